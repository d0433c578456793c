/* dtw_cli.c -- `sigfish dtw [OPTIONS] genome.fa reads.blow5` on the B200 path.
 *
 * Same options, defaults, validation rules and stderr summary as the reference's dtw_main()
 * (reference src/dtw_main.c:17-43 option table, 125-285 parsing + validation, 299-326 batch loop,
 * 331-345 summary).  Differences, all documented in INTEGRATION.md:
 *   - loading, record decoding and GPU submission run as pipeline stages on successive batches, with two
 *     batches in flight on the devices;
 *   - without -K / -B the batch is sized to two full waves of DTW tasks per GPU instead of 512 reads / 20 MB;
 *   - --gpus N / --gpu-first I choose the devices (default: all visible B200s); reads are sharded
 *     over them with the reference replicated;
 *   - --pore rna004 is accepted (the reference's validity test rejects it by mistake);
 *   - --profile-cpu / --accel are accepted and ignored: the stage timers are always printed.
 */
#include <errno.h>
#include <getopt.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "sfhost.h"

static struct option long_options[] = {
    {"slow5", required_argument, 0, 's'},    /* 0 vestigial */
    {"genome", required_argument, 0, 'g'},   /* 1 vestigial */
    {"threads", required_argument, 0, 't'},  /* 2 */
    {"batchsize", required_argument, 0, 'K'}, /* 3 */
    {"max-bytes", required_argument, 0, 'B'}, /* 4 */
    {"verbose", required_argument, 0, 'v'},  /* 5 */
    {"help", no_argument, 0, 'h'},           /* 6 */
    {"version", no_argument, 0, 'V'},        /* 7 */
    {"kmer-model", required_argument, 0, 0}, /* 8 */
    {"meth-model", required_argument, 0, 0}, /* 9 */
    {"output", required_argument, 0, 'o'},   /* 10 */
    {"window", required_argument, 0, 'w'},   /* 11 */
    {"rna", no_argument, 0, 0},              /* 12 */
    {"prefix", required_argument, 0, 'b'},   /* 13 */
    {"query-size", required_argument, 0, 'q'}, /* 14 */
    {"debug-break", required_argument, 0, 0}, /* 15 */
    {"dtw-std", no_argument, 0, 0},          /* 16 */
    {"invert", no_argument, 0, 0},           /* 17 */
    {"secondary", required_argument, 0, 0},  /* 18 */
    {"full-ref", no_argument, 0, 0},         /* 19 */
    {"from-end", no_argument, 0, 0},         /* 20 */
    {"profile-cpu", required_argument, 0, 0}, /* 21 */
    {"accel", required_argument, 0, 0},      /* 22 */
    {"sam", no_argument, 0, 'a'},            /* 23 */
    {"pore", required_argument, 0, 0},       /* 24 */
    {"gpus", required_argument, 0, 0},       /* 25 B200 */
    {"gpu-first", required_argument, 0, 0},  /* 26 B200 */
    {"device-decode", required_argument, 0, 0}, /* 27 B200 */
    {0, 0, 0, 0}};

static int64_t parse_num(const char *str)
{
    char *p;
    double x = strtod(str, &p);
    if (*p == 'G' || *p == 'g')
        x *= 1e9;
    else if (*p == 'M' || *p == 'm')
        x *= 1e6;
    else if (*p == 'K' || *p == 'k')
        x *= 1e3;
    return (int64_t)(x + .499);
}

static void yes_or_no(opt_t *opt, uint32_t flag, const char *name, const char *arg)
{
    if (!strcmp(arg, "yes") || !strcmp(arg, "y"))
        opt->flag |= flag;
    else if (!strcmp(arg, "no") || !strcmp(arg, "n"))
        opt->flag &= ~flag;
    else
        SF_WARNING("option '--%s' only accepts 'yes' or 'no'.", name);
}

static void print_help_msg(FILE *fp, const opt_t *opt)
{
    fprintf(fp, "Usage: sigfish dtw [OPTIONS] genome.fa reads.blow5\n");
    fprintf(fp, "\nbasic options:\n");
    fprintf(fp, "   -t INT                     number of host threads decoding records [%d]\n", opt->num_thread);
    fprintf(fp, "   -K INT                     batch size (max number of reads loaded at once) [auto: ~1e12 DTW cells per GPU, >= %d]\n", opt->batch_size);
    fprintf(fp, "   -B FLOAT[K/M/G]            max number of bytes loaded at once [auto, >= %.1fM]\n", opt->batch_size_bytes / (float)(1000 * 1000));
    fprintf(fp, "   -h                         help\n");
    fprintf(fp, "   -o FILE                    output to file [stdout]\n");
    fprintf(fp, "   --verbose INT              verbosity level [%d]\n", (int)opt->verbosity);
    fprintf(fp, "   --version                  print version\n");
    fprintf(fp, "   --pore STR                 set the pore chemistry (r9, r10 or rna004) [auto]\n");
    fprintf(fp, "   --gpus INT                 number of B200 GPUs to shard the reads over [all]\n");
    fprintf(fp, "   --gpu-first INT            ordinal of the first GPU to use [0]\n");
    fprintf(fp, "   --device-decode=yes|no     inflate / svb-zd decode BLOW5 records on the GPUs instead of the host threads [yes]\n");
    fprintf(fp, "\nadvanced options:\n");
    fprintf(fp, "   --kmer-model FILE          nucleotide k-mer model file (required: no built-in models in this build)\n");
    fprintf(fp, "   --rna                      the dataset is direct RNA\n");
    fprintf(fp, "   -q INT                     the number of events in query signal to align [%d]\n", opt->query_size);
    fprintf(fp, "   -p INT                     the number of events to trim at query signal start [%d]\n", opt->prefix_size);
    fprintf(fp, "   --debug-break INT          break after processing the specified no. of batches\n");
    fprintf(fp, "   --profile-cpu=yes|no       accepted for compatibility (stage timers are always printed)\n");
    fprintf(fp, "   --dtw-std                  use DTW standard instead of DTW subsequence\n");
    fprintf(fp, "   --invert                   reverse the reference events instead of query\n");
    fprintf(fp, "   --full-ref                 map to the full reference\n");
    fprintf(fp, "   --from-end                 Map the end portion of the query instead of the beginning\n");
    fprintf(fp, "   --sam                      Output in SAM format\n");
}

/* ---- host pipeline: loader thread -> decoder thread -> main thread ---- */
#define FEED_DBS 4

typedef struct {
    db_t *item[FEED_DBS + 1];
    int head, count;
    pthread_mutex_t mu;
    pthread_cond_t cv;
} queue_t;

static void queue_init(queue_t *q)
{
    memset(q, 0, sizeof *q);
    pthread_mutex_init(&q->mu, NULL);
    pthread_cond_init(&q->cv, NULL);
}

static void queue_push(queue_t *q, db_t *d)
{
    pthread_mutex_lock(&q->mu);
    q->item[(q->head + q->count) % (FEED_DBS + 1)] = d;
    q->count++;
    pthread_cond_signal(&q->cv);
    pthread_mutex_unlock(&q->mu);
}

/* oldest entry; NULL if the queue is empty and wait == 0 */
static db_t *queue_pop(queue_t *q, int wait)
{
    db_t *d = NULL;
    pthread_mutex_lock(&q->mu);
    while (q->count == 0 && wait)
        pthread_cond_wait(&q->cv, &q->mu);
    if (q->count > 0) {
        d = q->item[q->head];
        q->head = (q->head + 1) % (FEED_DBS + 1);
        q->count--;
    }
    pthread_mutex_unlock(&q->mu);
    return d;
}

typedef struct {
    core_t *core;
    const opt_t *opt;
    double realtime0;
    db_t *db[FEED_DBS];
    queue_t empty, loaded, parsed;
    pthread_t loader, decoder;
} feed_t;

static void *feed_loader(void *p)
{
    feed_t *f = (feed_t *)p;
    core_t *core = f->core;
    int32_t n_loaded = 0;
    for (;;) {
        db_t *d = queue_pop(&f->empty, 1);
        const ret_status_t st = load_db(core, d);
        fprintf(stderr, "[%s::%.3f*%.2f] %d Entries (%.1fM bytes) loaded\n", "dtw_main", sf_realtime() - f->realtime0,
                sf_cputime() / (sf_realtime() - f->realtime0), st.num_reads, st.num_bytes / (1000.0 * 1000.0));
        /* src/dtw_main.c:299-300, 322-325 */
        const int more = (st.num_reads >= core->opt.batch_size || st.num_bytes >= core->opt.batch_size_bytes) &&
                         f->opt->debug_break != n_loaded;
        n_loaded++;
        d->last_batch = !more;
        queue_push(&f->loaded, d);
        if (!more)
            return NULL;
    }
}

static void *feed_decoder(void *p)
{
    feed_t *f = (feed_t *)p;
    for (;;) {
        db_t *d = queue_pop(&f->loaded, 1);
        parse_db(f->core, d);
        const int last = d->last_batch; /* read before the batch is handed on */
        queue_push(&f->parsed, d);
        if (last)
            return NULL;
    }
}

static void feed_init(feed_t *f, core_t *core, const opt_t *opt, double realtime0)
{
    memset(f, 0, sizeof *f);
    f->core = core;
    f->opt = opt;
    f->realtime0 = realtime0;
    queue_init(&f->empty);
    queue_init(&f->loaded);
    queue_init(&f->parsed);
    for (int i = 0; i < FEED_DBS; i++) {
        f->db[i] = init_db(core);
        queue_push(&f->empty, f->db[i]);
    }
    if (pthread_create(&f->loader, NULL, feed_loader, f) || pthread_create(&f->decoder, NULL, feed_decoder, f)) {
        SF_FATAL("%s", "pthread_create failed");
    }
}

static void feed_destroy(feed_t *f)
{
    pthread_join(f->loader, NULL);
    pthread_join(f->decoder, NULL);
    for (int i = 0; i < FEED_DBS; i++)
        free_db(f->db[i]);
}

int dtw_main(int argc, char *argv[])
{
    const double realtime0 = sf_realtime();
    const char *optstring = "p:q:t:B:K:v:o:w:ahV";
    int longindex = 0;
    int c;
    FILE *fp_help = stderr;
    opt_t opt;
    init_opt(&opt);
    int k_set = 0, b_set = 0;

    while ((c = getopt_long(argc, argv, optstring, long_options, &longindex)) >= 0) {
        if (c == 'w') {
            opt.region_str = optarg;
        } else if (c == 'B') {
            opt.batch_size_bytes = parse_num(optarg);
            b_set = 1;
            if (opt.batch_size_bytes <= 0)
                SF_FATAL("%s", "Maximum number of bytes should be larger than 0.");
        } else if (c == 'K') {
            opt.batch_size = atoi(optarg);
            k_set = 1;
            if (opt.batch_size < 1)
                SF_FATAL("Batch size should larger than 0. You entered %d", opt.batch_size);
        } else if (c == 't') {
            opt.num_thread = atoi(optarg);
            if (opt.num_thread < 1)
                SF_FATAL("Number of threads should larger than 0. You entered %d", opt.num_thread);
        } else if (c == 'v') {
            opt.verbosity = (int8_t)atoi(optarg);
            sf_verbosity = opt.verbosity;
        } else if (c == 'V') {
            fprintf(stdout, "sigfish %s\n", SFHOST_VERSION);
            exit(EXIT_SUCCESS);
        } else if (c == 'h') {
            fp_help = stdout;
        } else if (c == 'p') {
            opt.prefix_size = atoi(optarg);
            if (opt.prefix_size < 0)
                SF_INFO("%s", "Autodetect query start.");
        } else if (c == 'q') {
            opt.query_size = atoi(optarg);
            if (opt.query_size < 0)
                SF_FATAL("Query size should larger than 0. You entered %d", opt.query_size);
        } else if (c == 0 && longindex == 8) {
            opt.model_file = optarg;
        } else if (c == 0 && longindex == 9) {
            opt.meth_model_file = optarg;
        } else if (c == 'o') {
            if (strcmp(optarg, "-") != 0 && freopen(optarg, "wb", stdout) == NULL)
                SF_FATAL("failed to write the output to file %s : %s", optarg, strerror(errno));
        } else if (c == 'a') {
            opt.flag |= SIGFISH_SAM;
        } else if (c == 0 && longindex == 12) {
            opt.flag |= SIGFISH_RNA;
        } else if (c == 0 && longindex == 15) {
            opt.debug_break = atoi(optarg);
        } else if (c == 0 && longindex == 16) {
            opt.flag |= SIGFISH_DTW;
        } else if (c == 0 && longindex == 17) {
            opt.flag |= SIGFISH_INV;
        } else if (c == 0 && longindex == 18) {
            yes_or_no(&opt, SIGFISH_SEC, "secondary", optarg);
        } else if (c == 0 && longindex == 19) {
            opt.flag |= SIGFISH_REF;
        } else if (c == 0 && longindex == 20) {
            opt.flag |= SIGFISH_END;
        } else if (c == 0 && longindex == 21) {
            yes_or_no(&opt, SIGFISH_PRF, "profile-cpu", optarg);
        } else if (c == 0 && longindex == 22) {
            yes_or_no(&opt, SIGFISH_ACC, "accel", optarg);
        } else if (c == 0 && longindex == 24) {
            opt.pore = optarg;
            if (strcmp(opt.pore, "r9") && strcmp(opt.pore, "r10") && strcmp(opt.pore, "rna004"))
                SF_FATAL("%s", "Pore model should be r9, r10 or rna004");
            if (!strcmp(opt.pore, "r10")) {
                opt.flag |= SIGFISH_R10;
                opt.pore_flag = OPT_PORE_R10;
            } else if (!strcmp(opt.pore, "rna004")) {
                opt.flag |= SIGFISH_RNA | SIGFISH_R10;
                opt.pore_flag = OPT_PORE_RNA004;
            }
        } else if (c == 0 && longindex == 25) {
            opt.num_gpus = atoi(optarg);
            if (opt.num_gpus < 1)
                SF_FATAL("Number of GPUs should larger than 0. You entered %d", opt.num_gpus);
        } else if (c == 0 && longindex == 26) {
            opt.first_gpu = atoi(optarg);
        } else if (c == 0 && longindex == 27) {
            if (!strcmp(optarg, "yes") || !strcmp(optarg, "y"))
                opt.device_decode = 1;
            else if (!strcmp(optarg, "no") || !strcmp(optarg, "n"))
                opt.device_decode = 0;
            else
                SF_FATAL("%s", "option '--device-decode' requires an argument 'yes' or 'no'");
        }
    }

    if (argc - optind != 2 || fp_help == stdout) {
        print_help_msg(fp_help, &opt);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    const char *fastafile = argv[optind];
    char *slow5file = argv[optind + 1];

    /* src/dtw_main.c:248-277 (the checks run before RNA auto-detection, as in the reference) */
    if (!(opt.flag & SIGFISH_RNA)) {
        if (opt.flag & SIGFISH_DTW)
            SF_FATAL("%s", "DTW is only available for RNA.");
        if (opt.flag & SIGFISH_INV)
            SF_FATAL("%s", "Inversion is only available for RNA.");
        if (opt.flag & SIGFISH_REF)
            SF_FATAL("%s", "--full-ref is only available for RNA.");
    }
    if (opt.prefix_size < 0) {
        if (!(opt.flag & SIGFISH_RNA))
            SF_FATAL("%s", "DNA does not support auto query start detection.");
        if (opt.flag & SIGFISH_INV)
            SF_FATAL("%s", "Inversion is not compatible with auto query start detection.");
        if (opt.flag & SIGFISH_END)
            SF_FATAL("%s", "Mapping from query end is not compatible with auto query start detection.");
    }

    core_t *core = init_core(fastafile, slow5file, opt, realtime0);
    /* Batch size: -K / -B are honoured when given (long references are cut into pieces, so any batch fills the
     * GPU).  Otherwise a batch is about 1e12 DTW cells per GPU (~0.12 s of device time: long enough
     * for the fixed costs of a batch -- launches, the tail of the last tasks, the host epilogue -- to stay small,
     * short enough for the GPUs to start early: on 8 GPUs a first batch of two waves per GPU kept them idle for
     * 0.33 s while it was read, allocated for and staged), at least half a wave and at most 65536 reads per GPU. */
    if (!k_set) {
        const int32_t wave = sfgpu_wave_reads(core->gpu[0]);
        const double cells_per_read = (double)core->opt.query_size * (double)sfgpu_ref_columns(core->gpu[0]);
        int64_t per_gpu = cells_per_read > 0 ? (int64_t)(1e12 / cells_per_read) : 0;
        if (per_gpu > 65536)
            per_gpu = 65536;
        if (per_gpu < wave / 2)
            per_gpu = wave / 2;
        if (per_gpu < 64)
            per_gpu = 64;
        const int64_t want = per_gpu * core->num_gpus;
        if (want > core->opt.batch_size)
            core->opt.batch_size = (int32_t)(want > 262144 ? 262144 : want);
    }
    if (!b_set) {
        const int64_t want = (int64_t)core->opt.batch_size * 16000;
        if (want > core->opt.batch_size_bytes)
            core->opt.batch_size_bytes = want;
    }
    if (sf_verbosity >= 3)
        fprintf(stderr, "[%s] batch size: %d reads / %.1fM bytes on %d GPU(s)\n", __func__, core->opt.batch_size,
                core->opt.batch_size_bytes / 1e6, core->num_gpus);

    /* Three host stages run on different batches at once: a loader thread reads raw records, a decoder thread
     * (fanning out to the -t workers) inflates them, and this thread packs / submits them to the GPUs and runs
     * the epilogue.  Two batches are in flight on the devices (one slot each): the tail of one overlaps the
     * head of the next.  Output stays in input order. */
    feed_t feed;
    feed_init(&feed, core, &opt, realtime0);
    if (core->opt.flag & SIGFISH_SAM)
        sam_hdr_wr(core->ref);
    db_t *flying[2] = {NULL, NULL};
    int in_flight = 0, eof = 0;
    while (!eof || in_flight) {
        if (!eof && in_flight < 2) {
            /* with a batch on the devices, only take the next one if it is ready: otherwise collect first */
            db_t *d = queue_pop(&feed.parsed, in_flight == 0);
            if (d) {
                eof = d->last_batch;
                const double t0 = sf_realtime();
                submit_db(core, d);
                core->process_db_time += sf_realtime() - t0;
                flying[in_flight++] = d;
                continue;
            }
        }
        db_t *d = flying[0];
        const double t0 = sf_realtime();
        collect_db(core, d);
        core->process_db_time += sf_realtime() - t0;
        fprintf(stderr, "[%s::%.3f*%.2f] %d Entries (%.1fM bytes) processed\n", __func__, sf_realtime() - realtime0,
                sf_cputime() / (sf_realtime() - realtime0), d->n_rec, d->sum_bytes / (1000.0 * 1000.0));
        output_db(core, d);
        free_db_tmp(d);
        flying[0] = flying[1];
        flying[1] = NULL;
        in_flight--;
        queue_push(&feed.empty, d);
    }
    feed_destroy(&feed);

    fprintf(stderr, "[%s] total entries: %ld\tprefix fail: %ld\tignored: %ld\ttoo short: %ld", __func__,
            (long)core->total_reads, (long)core->prefix_fail, (long)core->ignored, (long)core->too_short);
    fprintf(stderr, "\n[%s] total bytes: %.1f M", __func__, core->sum_bytes / (float)(1000 * 1000));
    fprintf(stderr, "\n[%s] Data loading time: %.3f sec", __func__, core->load_db_time);
    fprintf(stderr, "\n[%s] Data processing time: %.3f sec", __func__, core->process_db_time);
    fprintf(stderr, "\n[%s]     - Parse time (host): %.3f sec", __func__, core->parse_time);
    fprintf(stderr, "\n[%s]     - H2D time (GPU 0): %.3f sec", __func__, core->h2d_time);
    fprintf(stderr, "\n[%s]     - Events + normalise time (GPU 0): %.3f sec", __func__, core->event_time);
    fprintf(stderr, "\n[%s]     - DTW time (GPU 0): %.3f sec", __func__, core->dtw_time);
    if (core->dtw_time > 0)
        fprintf(stderr, "\n[%s]     - DTW cells: %.4g on %d GPU(s)", __func__, core->cells, core->num_gpus);
    fprintf(stderr, "\n[%s] Data output time: %.3f sec", __func__, core->output_time);
    fprintf(stderr, "\n");
    free_core(core, opt);
    return 0;
}
