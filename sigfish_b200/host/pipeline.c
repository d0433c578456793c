/* pipeline.c -- the batch pipeline of `sigfish dtw` around the B200 C-ABI.
 *
 * Mirrors (reference, paths relative to /root/reference):
 *   init_opt      src/sigfish.c:1122-1144     defaults K=512, B=20 MB, t=8, p=50, q=250
 *   init_core     src/sigfish.c:81-206        open reads, detect RNA / pore, load model, build reference
 *   init_db       src/sigfish.c:234-270
 *   load_db       src/sigfish.c:274-315       up to K records or B bytes of raw records
 *   process_db    src/sigfish.c:1018-1047     here: decode (host threads) -> pack -> GPUs -> epilogue
 *   dtw_single    src/sigfish.c:969-985       the part left on the host: strand flip, offset, MAPQ
 *   aln_to_str    src/sigfish.c:796-826 and paf_str 628-660
 *   output_db     src/sigfish.c:1051-1086     rows in input order, counters
 *   free_*        src/sigfish.c:208-231, 1089-1119
 */
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/time.h>

#include "refio.h"
#include "s5read.h"
#include "sfhost.h"

int8_t sf_verbosity = 4;

double sf_realtime(void)
{
    struct timeval tp;
    gettimeofday(&tp, NULL);
    return tp.tv_sec + tp.tv_usec * 1e-6;
}

double sf_cputime(void)
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

long sf_peakrss(void)
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return r.ru_maxrss * 1024;
}

void init_opt(opt_t *opt)
{
    memset(opt, 0, sizeof(opt_t));
    opt->batch_size = 512;
    opt->batch_size_bytes = 20 * 1000 * 1000;
    opt->num_thread = 8;
    opt->debug_break = -1;
    opt->prefix_size = 50;
    opt->query_size = 250;
    opt->verbosity = 4;
    opt->device_decode = 1;
}

/* src/sigfish.c:22-52 */
static int drna_detect(const sf_s5file_t *sf)
{
    const char *exp = sf_s5_hdr_get(sf, "experiment_type", 0);
    if (!exp) {
        SF_WARNING("%s", "experiment_type not found in SLOW5 header. Assuming genomic_dna");
        return 0;
    }
    int rna = 0;
    if (!strcmp(exp, "genomic_dna"))
        rna = 0;
    else if (!strcmp(exp, "rna"))
        rna = 1;
    else
        SF_WARNING("Unknown experiment type: %s. Assuming genomic_dna", exp);
    for (uint32_t g = 1; g < sf_s5_num_read_groups(sf); g++) {
        const char *cur = sf_s5_hdr_get(sf, "experiment_type", g);
        if (cur && strcmp(cur, exp))
            SF_WARNING("Experiment type mismatch: %s != %s in read group %d. Defaulted to %s", cur, exp, (int)g, exp);
    }
    return rna;
}

/* src/sigfish.c:54-80 */
static int8_t pore_detect(const sf_s5file_t *sf)
{
    const char *kit = sf_s5_hdr_get(sf, "sequencing_kit", 0);
    if (!kit) {
        SF_WARNING("%s", "sequencing_kit not found in SLOW5 header. Assuming R9.4.1");
        return OPT_PORE_R9;
    }
    int8_t pore = OPT_PORE_R9;
    if (strstr(kit, "114"))
        pore = OPT_PORE_R10;
    else if (strstr(kit, "rna004"))
        pore = OPT_PORE_RNA004;
    for (uint32_t g = 1; g < sf_s5_num_read_groups(sf); g++) {
        const char *cur = sf_s5_hdr_get(sf, "sequencing_kit", g);
        if (cur && strcmp(cur, kit))
            SF_WARNING("sequencing_kit type mismatch: %s != %s in read group %d. Defaulted to %s", cur, kit, (int)g, kit);
    }
    return pore;
}

typedef struct {
    core_t *core;
    int g, device;
    const opt_t *opt;
    const sf_fasta_t *fa;
    int write_ref; /* only one worker fills the shared ref_lengths / offsets arrays */
    int failed;
    char err[512];
} gpu_init_arg_t;

static void *gpu_init_worker(void *p)
{
    gpu_init_arg_t *a = (gpu_init_arg_t *)p;
    core_t *core = a->core;
    const opt_t *opt = a->opt;
    sfgpu_opt_t go;
    memset(&go, 0, sizeof go);
    go.device = a->device;
    go.flags = opt->flag & (SFGPU_RNA | SFGPU_DTW | SFGPU_INV | SFGPU_REF | SFGPU_END | SFGPU_SAM);
    go.query_size = opt->query_size;
    go.prefix_size = opt->prefix_size;
    go.kmer_size = (int32_t)core->kmer_size;
    go.n_slots = 2;
    go.pore = opt->pore_flag;
    if (sfgpu_create(&core->gpu[a->g], &go, core->level_mean) != SFGPU_OK) {
        snprintf(a->err, sizeof a->err, "%s", sfgpu_strerror(NULL));
        a->failed = 1;
        return NULL;
    }
    refsynth_t *ref = core->ref;
    if (sfgpu_set_ref(core->gpu[a->g], a->fa->num_ref, a->fa->bases, a->fa->off, a->write_ref ? ref->ref_lengths : NULL,
                      a->write_ref ? ref->ref_seq_lengths : NULL, a->write_ref ? ref->ref_st_offset : NULL) != SFGPU_OK) {
        snprintf(a->err, sizeof a->err, "%s", sfgpu_strerror(core->gpu[a->g]));
        a->failed = 1;
    }
    return NULL;
}

core_t *init_core(const char *fastafile, char *slow5file, opt_t opt, double realtime0)
{
    core_t *core = (core_t *)calloc(1, sizeof(core_t));
    char err[512];
    sf_verbosity = opt.verbosity;

    const double t_init0 = sf_realtime();
#define SF_STAGE(msg)                                                                                 \
    do {                                                                                              \
        if (sf_verbosity >= 5)                                                                        \
            fprintf(stderr, "[init_core::%.3f] %s\n", sf_realtime() - t_init0, msg);                  \
    } while (0)
    core->sf = sf_s5_open(slow5file, err, sizeof err);
    if (!core->sf) {
        SF_FATAL("Error opening SLOW5 file: %s", err);
    }
    if (drna_detect(core->sf)) {
        opt.flag |= SIGFISH_RNA;
        if (sf_verbosity >= 3)
            fprintf(stderr, "[INFO] %s: Detected RNA data. --rna was set automatically.\n", __func__); /* VERBOSE(), src/error.h:36 */
    }
    if (opt.pore == NULL) {
        const int8_t pore = pore_detect(core->sf);
        opt.pore_flag = pore;
        if (pore) {
            opt.flag |= SIGFISH_R10;
            if (sf_verbosity >= 3)
                fprintf(stderr, "[INFO] %s: Detected %s data. --pore %s was set automatically.\n", __func__,
                        pore == OPT_PORE_R10 ? "R10" : "RNA004", pore == OPT_PORE_R10 ? "r10" : "rna004");
        }
    }

    /* model (src/sigfish.c:143-164).  The built-in tables of the reference live in src/model.h, which is
     * not part of the reference mount, so there is nothing to fall back to. */
    if (!opt.model_file) {
        SF_FATAL("%s", "no k-mer model: the built-in pore models are not available in this build, pass --kmer-model FILE");
    }
    if (sf_model_read(opt.model_file, &core->level_mean, &core->kmer_size, err, sizeof err)) {
        SF_FATAL("%s", err);
    }

    SF_STAGE("model read");
    /* GPUs.  Initialising the CUDA driver costs about 0.6 s per visible GPU on a B200 box (measured:
     * tools/gpuinit_probe.cu, profiles/r02_cli_trace_*), whether the GPU is used or not, so when the user asked for
     * fewer GPUs than the box has, only those are made visible before the first CUDA call. */
    if (opt.num_gpus > 0) {
        const int first_req = opt.first_gpu > 0 ? opt.first_gpu : 0;
        const char *vis = getenv("CUDA_VISIBLE_DEVICES");
        char list[1024];
        size_t len = 0;
        int ok = 1;
        if (vis && *vis) { /* entries [first, first + n) of the list that is already in force */
            const char *p = vis;
            int idx = 0, taken = 0;
            while (*p && taken < opt.num_gpus) {
                const char *e = strchr(p, ',');
                const size_t l = e ? (size_t)(e - p) : strlen(p);
                if (idx >= first_req) {
                    if (len + l + 2 > sizeof list) { ok = 0; break; }
                    if (taken) list[len++] = ',';
                    memcpy(list + len, p, l);
                    len += l;
                    taken++;
                }
                idx++;
                if (!e) break;
                p = e + 1;
            }
            if (taken < opt.num_gpus) ok = 0; /* fewer entries than asked for: let the check below report it */
        } else {
            for (int g = 0; g < opt.num_gpus && ok; g++) {
                const int w = snprintf(list + len, sizeof list - len, "%s%d", g ? "," : "", first_req + g);
                if (w < 0 || (size_t)w >= sizeof list - len) ok = 0; else len += (size_t)w;
            }
        }
        if (ok && len > 0) {
            list[len] = 0;
            setenv("CUDA_VISIBLE_DEVICES", list, 1);
            opt.first_gpu = 0;
        }
    }
    int ndev = sfgpu_device_count();
    SF_STAGE("device count");
    if (ndev < 1) {
        SF_FATAL("%s", "no sm_100 (B200) GPU visible: this build has no CPU path");
    }
    int first = opt.first_gpu > 0 ? opt.first_gpu : 0;
    int want = opt.num_gpus > 0 ? opt.num_gpus : ndev - first;
    if (first + want > ndev || want < 1) {
        SF_FATAL("requested GPUs %d..%d but %d visible", first, first + want - 1, ndev);
    }
    if (want > SFHOST_MAX_GPUS)
        want = SFHOST_MAX_GPUS;
    core->num_gpus = want;

    /* reference (replaces gen_ref, src/sigfish.c:178): FASTA on the host, events on every GPU */
    sf_fasta_t fa;
    if (sf_fasta_read(fastafile, &fa, err, sizeof err)) {
        SF_FATAL("%s", err);
    }
    SF_STAGE("FASTA read");
    refsynth_t *ref = (refsynth_t *)calloc(1, sizeof(refsynth_t));
    ref->num_ref = fa.num_ref;
    ref->ref_names = fa.names;
    ref->ref_lengths = (int32_t *)calloc(fa.num_ref, sizeof(int32_t));
    ref->ref_seq_lengths = (int32_t *)calloc(fa.num_ref, sizeof(int32_t));
    ref->ref_st_offset = (int32_t *)calloc(fa.num_ref, sizeof(int32_t));
    core->ref = ref;
    int64_t bad = 0;
    for (int64_t i = 0; i < fa.off[fa.num_ref]; i++) {
        switch (fa.bases[i]) {
        case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': break;
        default: bad++;
        }
    }
    if (bad)
        SF_WARNING("%ld non-ACGT reference bases are treated as 'A' (reverse strand: 'T'), as in the reference", (long)bad);

    /* one context per GPU, each with the whole reference; built concurrently (CUDA context creation and the
     * serial-order z-score of a long contig take ~1 s per device) */
    {
        gpu_init_arg_t *ia = (gpu_init_arg_t *)calloc((size_t)core->num_gpus, sizeof(gpu_init_arg_t));
        pthread_t *tid = (pthread_t *)calloc((size_t)core->num_gpus, sizeof(pthread_t));
        for (int g = 0; g < core->num_gpus; g++) {
            ia[g].core = core;
            ia[g].g = g;
            ia[g].device = first + g;
            ia[g].opt = &opt;
            ia[g].fa = &fa;
            ia[g].write_ref = g == 0;
            if (core->num_gpus == 1)
                gpu_init_worker(&ia[g]);
            else if (pthread_create(&tid[g], NULL, gpu_init_worker, &ia[g])) {
                SF_FATAL("%s", "pthread_create failed");
            }
        }
        for (int g = 0; g < core->num_gpus; g++) {
            if (core->num_gpus > 1)
                pthread_join(tid[g], NULL);
            if (ia[g].failed) {
                SF_FATAL("GPU %d: %s", first + g, ia[g].err);
            }
        }
        free(ia);
        free(tid);
    }
    SF_STAGE("GPU contexts + resident reference");
    free(fa.bases);
    free(fa.off);
#undef SF_STAGE

    core->opt = opt;
    core->realtime0 = realtime0;
    /* BLOW5 records go to the GPUs as they are (zlib / uncompressed records, svb-zd / uncompressed signals) */
    core->device_decode = opt.device_decode && sf_s5_is_binary(core->sf) && sf_s5_record_press(core->sf) <= 1 &&
                          sf_s5_signal_press(core->sf) <= 1;
    return core;
}

void free_core(core_t *core, opt_t opt)
{
    (void)opt;
    for (int g = 0; g < core->num_gpus; g++)
        sfgpu_destroy(core->gpu[g]);
    if (core->ref) {
        for (int i = 0; i < core->ref->num_ref; i++)
            free(core->ref->ref_names[i]);
        free(core->ref->ref_names);
        free(core->ref->ref_lengths);
        free(core->ref->ref_seq_lengths);
        free(core->ref->ref_st_offset);
        free(core->ref);
    }
    free(core->level_mean);
    sf_s5_close(core->sf);
    free(core);
}

db_t *init_db(core_t *core)
{
    db_t *db = (db_t *)calloc(1, sizeof(db_t));
    const size_t n = (size_t)core->opt.batch_size;
    db->capacity_rec = core->opt.batch_size;
    db->mem_records = (char **)calloc(n, sizeof(char *));
    db->mem_bytes = (size_t *)calloc(n, sizeof(size_t));
    db->mem_cap = (size_t *)calloc(n, sizeof(size_t));
    db->rec = (sf_rec_t *)calloc(n, sizeof(sf_rec_t));
    db->res = (sfgpu_result_t *)calloc(n, sizeof(sfgpu_result_t));
    db->aln = (aln_t *)calloc(n, sizeof(aln_t));
    db->out = (char **)calloc(n, sizeof(char *));
    db->mem_views = sf_s5_is_mapped(core->sf);
    db->sig_pos = (int32_t *)calloc(n, sizeof(int32_t));
    db->sig_bytes = (int64_t *)calloc(n, sizeof(int64_t));
    db->rec_bytes = (int64_t *)calloc(n, sizeof(int64_t));
    db->sig_off = (int64_t *)calloc(n + 1, sizeof(int64_t));
    db->sig_ptr = (int16_t **)calloc(n, sizeof(int16_t *));
    db->dig = (float *)calloc(n, sizeof(float));
    db->off = (float *)calloc(n, sizeof(float));
    db->rng = (float *)calloc(n, sizeof(float));
    if (core->opt.flag & SIGFISH_SAM) {
        db->move_off = (int64_t *)calloc(n + 1, sizeof(int64_t));
        db->n_moves = (int32_t *)calloc(n, sizeof(int32_t));
        db->win_start = (uint64_t *)calloc(n * (size_t)core->opt.query_size, sizeof(uint64_t));
        db->win_len = (float *)calloc(n * (size_t)core->opt.query_size, sizeof(float));
    }
    return db;
}

ret_status_t load_db(core_t *core, db_t *db)
{
    const double t0 = sf_realtime();
    db->n_rec = 0;
    db->sum_bytes = 0;
    db->total_reads = 0;
    db->prefix_fail = 0;
    db->ignored = 0;
    db->too_short = 0;
    db->submitted = 0;
    ret_status_t status = {0, 0};
    while (db->n_rec < db->capacity_rec && db->sum_bytes < core->opt.batch_size_bytes) {
        const int i = db->n_rec;
        int64_t got;
        if (db->mem_views) { /* mapped BLOW5: the record stays where it is in the page cache */
            const char *view = NULL;
            got = sf_s5_get_next_view(core->sf, &view);
            db->mem_records[i] = (char *)view;
        } else {
            got = sf_s5_get_next_mem(core->sf, &db->mem_records[i], &db->mem_cap[i]);
        }
        if (got < 0) {
            SF_FATAL("Error reading from SLOW5 file: %s", sf_s5_error(core->sf));
        }
        if (got == 0)
            break;
        db->mem_bytes[i] = (size_t)got;
        db->n_rec++;
        db->total_reads++;
        db->sum_bytes += got;
    }
    if (db->mem_views && db->n_rec > 0) /* records of a batch are contiguous in the file */
        sf_s5_prefault(core->sf, db->mem_records[0],
                       (size_t)(db->mem_records[db->n_rec - 1] - db->mem_records[0]) + db->mem_bytes[db->n_rec - 1]);
    status.num_reads = db->n_rec;
    status.num_bytes = db->sum_bytes;
    core->load_db_time += sf_realtime() - t0;
    return status;
}

/* ---- record decoding on host threads (parse_single, src/sigfish.c:317-328) ---- */
typedef struct {
    core_t *core;
    db_t *db;
    int32_t *next;
    char *scratch;
    size_t scratch_cap;
    int failed;
    int heads_only;
} parse_arg_t;

static void *parse_worker(void *p)
{
    parse_arg_t *a = (parse_arg_t *)p;
    for (;;) {
        const int32_t i = __sync_fetch_and_add(a->next, 1);
        if (i >= a->db->n_rec)
            break;
        if (a->heads_only) {
            if (sf_s5_parse_head(a->core->sf, a->db->mem_records[i], a->db->mem_bytes[i], &a->db->rec[i], &a->db->sig_pos[i],
                                 &a->db->sig_bytes[i], &a->scratch, &a->scratch_cap)) {
                a->failed = i + 1; /* the whole batch is then decoded on the host */
                break;
            }
            continue;
        }
        if (sf_s5_parse(a->core->sf, a->db->mem_records[i], a->db->mem_bytes[i], &a->db->rec[i], &a->scratch,
                        &a->scratch_cap)) {
            a->failed = i + 1;
            break;
        }
    }
    free(a->scratch);
    a->scratch = NULL;
    return NULL;
}

/* heads_only: read id, scaling and sample count of every record, the device decodes the signals; a record whose
 * head cannot be read that way sends the whole batch through the full host decoder */
static int parse_db_mode(core_t *core, db_t *db, int heads_only)
{
    const double t0 = sf_realtime();
    int nt = core->opt.num_thread < 1 ? 1 : core->opt.num_thread;
    if (nt > db->n_rec)
        nt = db->n_rec > 0 ? db->n_rec : 1;
    int32_t next = 0;
    parse_arg_t *args = (parse_arg_t *)calloc((size_t)nt, sizeof(parse_arg_t));
    pthread_t *tid = (pthread_t *)calloc((size_t)nt, sizeof(pthread_t));
    for (int t = 0; t < nt; t++) {
        args[t].core = core;
        args[t].db = db;
        args[t].next = &next;
        args[t].heads_only = heads_only;
    }
    if (nt == 1) {
        parse_worker(&args[0]);
    } else {
        for (int t = 0; t < nt; t++)
            if (pthread_create(&tid[t], NULL, parse_worker, &args[t])) {
                SF_FATAL("%s", "pthread_create failed");
            }
        for (int t = 0; t < nt; t++)
            pthread_join(tid[t], NULL);
    }
    int failed = 0;
    for (int t = 0; t < nt; t++)
        if (args[t].failed) {
            if (!heads_only) {
                SF_FATAL("Error parsing the record %d of the batch", args[t].failed - 1);
            }
            failed = 1;
        }
    free(args);
    free(tid);
    core->parse_time += sf_realtime() - t0;
    return failed;
}

void parse_db(core_t *core, db_t *db)
{
    db->heads_only = 0;
    if (core->device_decode && db->n_rec > 0) {
        if (parse_db_mode(core, db, 1) == 0) {
            db->heads_only = 1;
            return;
        }
        core->decode_fallbacks++;
    }
    parse_db_mode(core, db, 0);
}

/* Splits n_rec reads into G contiguous ranges [begin[g], begin[g+1]) of about the same total weight (a
 * contiguous split keeps the output in input order with a plain concatenation). */
void sf_shard_ranges(int32_t n_rec, const int64_t *weight, int32_t G, int32_t *begin)
{
    int64_t total = 0;
    for (int i = 0; i < n_rec; i++)
        total += weight[i];
    int64_t acc = 0;
    int g = 0;
    begin[0] = 0;
    for (int i = 0; i < n_rec && g + 1 < G; i++) {
        acc += weight[i];
        while (g + 1 < G && acc * G >= total * (g + 1)) {
            g++;
            begin[g] = i + 1;
        }
    }
    while (g < G) {
        g++;
        begin[g] = n_rec;
    }
}

typedef struct {
    core_t *core;
    db_t *db;
    int g;
    int failed;
} gpu_submit_arg_t;

static void *gpu_submit_worker(void *p)
{
    gpu_submit_arg_t *a = (gpu_submit_arg_t *)p;
    db_t *db = a->db;
    const int b = db->shard_begin[a->g], e = db->shard_begin[a->g + 1];
    int rc;
    if (db->heads_only)
        rc = sfgpu_submit_records(a->core->gpu[a->g], db->slot, e - b, (const uint8_t *const *)db->mem_records + b, db->rec_bytes + b,
                                  sf_s5_record_press(a->core->sf), sf_s5_signal_press(a->core->sf), db->sig_pos + b,
                                  db->sig_bytes + b, db->sig_off + b, db->dig + b, db->off + b, db->rng + b);
    else
        rc = sfgpu_submit_reads(a->core->gpu[a->g], db->slot, e - b, (const int16_t *const *)db->sig_ptr + b, db->sig_off + b,
                                db->dig + b, db->off + b, db->rng + b);
    if (rc != SFGPU_OK)
        a->failed = 1;
    return NULL;
}

void submit_db(core_t *core, db_t *db)
{
    /* shards: contiguous read ranges of about the same estimated device work.  A read costs its DTW, which does not
     * depend on its length (query_size rows x every reference column: the unit here is one DTW cell), plus the
     * copy and event detection of its samples (~640 cell times per sample: 2 bytes over PCIe against 8 TCUPS).
     * Balancing samples alone, as the first version did, gives a GPU with a few long reads much less DTW work than
     * one with many short reads, and the batch waits for the slowest GPU. */
    const int G = core->num_gpus;
    {
        int64_t *w = (int64_t *)malloc(sizeof(int64_t) * (size_t)(db->n_rec > 0 ? db->n_rec : 1));
        const int64_t dtw = (int64_t)core->opt.query_size * sfgpu_ref_columns(core->gpu[0]);
        for (int i = 0; i < db->n_rec; i++)
            w[i] = db->rec[i].len_raw_signal > 0 ? dtw + 640 * (int64_t)db->rec[i].len_raw_signal : 1;
        sf_shard_ranges(db->n_rec, w, G, db->shard_begin);
        free(w);
    }
    int g;

    db->slot = core->next_slot;
    core->next_slot ^= 1;
    /* per-read pointers / lengths / scalings once for the whole batch; every shard is a sub-range */
    for (int i = 0; i < db->n_rec; i++) {
        const sf_rec_t *r = &db->rec[i];
        db->sig_ptr[i] = r->raw_signal;
        db->rec_bytes[i] = (int64_t)db->mem_bytes[i];
        db->sig_off[i] = (int64_t)r->len_raw_signal; /* used as the length array here */
        /* narrowed to float exactly as event_single() does (src/sigfish.c:335-337) */
        db->dig[i] = (float)r->digitisation;
        db->off[i] = (float)r->offset;
        db->rng[i] = (float)r->range;
    }
    /* one host thread per GPU copies its shard into that GPU's pinned staging buffer and launches */
    gpu_submit_arg_t sa[SFHOST_MAX_GPUS];
    pthread_t tid[SFHOST_MAX_GPUS];
    for (g = 0; g < G; g++) {
        sa[g].core = core;
        sa[g].db = db;
        sa[g].g = g;
        sa[g].failed = 0;
        if (G == 1)
            gpu_submit_worker(&sa[g]);
        else if (pthread_create(&tid[g], NULL, gpu_submit_worker, &sa[g])) {
            SF_FATAL("%s", "pthread_create failed");
        }
    }
    for (g = 0; g < G; g++) {
        if (G > 1)
            pthread_join(tid[g], NULL);
        if (sa[g].failed) {
            SF_FATAL("GPU %d: %s", g, sfgpu_strerror(core->gpu[g]));
        }
    }
    db->submitted = 1;
}

/* (int)round(x) as the x86-64 reference evaluates it (cvttsd2si: out of range -> INT_MIN) */
static int32_t to_int_like_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return (int32_t)0x80000000;
    return (int32_t)v;
}

static char *paf_str(const aln_t *aln, const char *read_id, const char *rname, uint64_t start_raw, uint64_t end_raw,
                     uint64_t query_size, uint64_t len_raw_signal, uint64_t rlength)
{
    /* src/sigfish.c:628-660 */
    const float block_len = (float)(aln->pos_end - aln->pos_st);
    const float residue = block_len - aln->score * block_len / (float)query_size;
    const size_t cap = strlen(read_id) + strlen(rname) + 256;
    char *s = (char *)malloc(cap);
    snprintf(s, cap, "%s\t%ld\t%ld\t%ld\t%c\t%s\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\td1:f:%.2f\td2:f:%.2f\n", read_id,
             (long)len_raw_signal, (long)start_raw, (long)end_raw, aln->d, rname, (int)rlength, aln->pos_st, aln->pos_end,
             to_int_like_x86(round((double)residue)), to_int_like_x86(round((double)block_len)), aln->mapq,
             (double)aln->score, (double)aln->score2);
    return s;
}

/* ---- SAM (src/sigfish.c:530-571 path_to_map, 663-768 r2qevent_map_to_ss, 770-794 sam_str) ---- */

typedef struct {
    int32_t start, stop; /* index_pair_t, src/sigfish.h:141-144 */
} index_pair_t;

typedef struct {
    char *s;
    size_t l, m;
} sbuf_t;

static void sbuf_printf(sbuf_t *b, const char *fmt, ...)
{
    for (;;) {
        va_list ap;
        va_start(ap, fmt);
        const int w = vsnprintf(b->s + b->l, b->m - b->l, fmt, ap);
        va_end(ap);
        if (w >= 0 && (size_t)w < b->m - b->l) {
            b->l += (size_t)w;
            return;
        }
        b->m = b->m * 2 + (size_t)(w > 0 ? w : 64);
        b->s = (char *)realloc(b->s, b->m);
    }
}

/* moves: the winner's path backwards from (qlen-1, pos_end), as sfgpu_collect_paths() returns it */
static char *sam_str(const aln_t *aln, const sfgpu_result_t *r, const uint8_t *moves, int32_t n_moves,
                     const uint64_t *ev_start, const float *ev_len, const char *read_id, const char *rname, int rna)
{
    const int32_t n_kmers = r->pos_end - r->pos_st + 1;
    index_pair_t *map = (index_pair_t *)malloc(sizeof(index_pair_t) * (size_t)n_kmers);
    for (int32_t i = 0; i < n_kmers; i++)
        map[i].start = map[i].stop = -1;
    /* forward order: undo the moves from the far end */
    const int32_t k = n_moves + 1;
    int32_t *px = (int32_t *)malloc(sizeof(int32_t) * (size_t)k), *py = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    {
        int32_t i = r->qlen - 1, j = r->pos_end;
        px[k - 1] = i;
        py[k - 1] = j;
        for (int32_t s = 0; s < n_moves; s++) {
            if (moves[s] == 0) { i--; j--; }
            else if (moves[s] == 1) { j--; }
            else { i--; }
            px[k - 2 - s] = i;
            py[k - 2 - s] = j;
        }
    }
    int32_t prev_q = -1;
    for (int32_t s = 0; s < k; s++) { /* path_to_map */
        const int32_t at = py[s] - py[0], qi = px[s];
        if (map[at].start == -1)
            map[at].start = qi;
        map[at].stop = qi;
        if (prev_q == qi)
            map[at].start = map[at].stop = -1;
        prev_q = qi;
    }
    free(px);
    free(py);

    /* r2qevent_map_to_ss: event indices relative to the window here (the reference adds qstart and indexes
     * the full event table; ev_start/ev_len hold the window only) */
    if (rna) {
        const int32_t end = map[n_kmers - 1].stop;
        for (int32_t i = 0; i < n_kmers; i++)
            if (map[i].start != -1) {
                map[i].start = end - map[i].start;
                map[i].stop = end - map[i].stop;
            }
        for (int32_t a = 0; a < n_kmers / 2; a++) {
            index_pair_t t = map[a];
            map[a] = map[n_kmers - 1 - a];
            map[n_kmers - 1 - a] = t;
        }
        for (int32_t i = 0; i < n_kmers; i++) {
            int32_t t = map[i].start;
            map[i].start = map[i].stop;
            map[i].stop = t;
        }
    }

    sbuf_t b;
    b.m = 4000;
    b.l = 0;
    b.s = (char *)malloc(b.m);
    const uint64_t qsize = (uint64_t)(r->qend - 1) - (uint64_t)r->qstart;
    sbuf_printf(&b, "%s\t%d\t%s\t%ld\t%d\t%ldM\t*\t0\t0\t*\t*\t", read_id, aln->d == '+' ? 0 : 16, rname,
                (long)aln->pos_st + 1, aln->mapq, (long)qsize);
    sbuf_printf(&b, "si:Z:%ld,%ld,%ld,%ld\tss:Z:", (long)r->start_raw, (long)r->end_raw,
                (long)(rna ? aln->pos_end : aln->pos_st), (long)(rna ? aln->pos_st : aln->pos_end));
    int64_t ci = 0, mi = 0, d = 0;
    int ff = 1;
    for (int32_t j = 0; j < n_kmers; j++) {
        if (map[j].start == -1) {
            if (!ff)
                d++;
            continue;
        }
        const int64_t s0 = (int64_t)ev_start[map[j].start];
        const int64_t s1 = (int64_t)ev_start[map[j].stop] + (int)ev_len[map[j].stop];
        ff = 0;
        if (d > 0) {
            sbuf_printf(&b, "%dD", (int)d);
            d = 0;
        }
        if (j == 0)
            ci = s0;
        ci += (mi = s0 - ci);
        if (mi)
            sbuf_printf(&b, "%dI", (int)mi);
        ci += (mi = s1 - s0);
        if (mi)
            sbuf_printf(&b, "%d,", (int)mi);
    }
    sbuf_printf(&b, "\n");
    free(map);
    return b.s;
}

/* ---- per-read epilogue of a collected batch ---- */
#define SF_EPI_CHUNK 512
typedef struct {
    core_t *core;
    db_t *db;
    int32_t *next;
    int64_t ignored, too_short, prefix_fail;
} epi_arg_t;

static void epilogue_read(core_t *core, db_t *db, int i, epi_arg_t *cnt)
{
    const refsynth_t *ref = core->ref;
    {
        const sfgpu_result_t *r = &db->res[i];
        db->out[i] = NULL;
        if (r->status & 1)
            cnt->ignored++;
        if (r->status & 2)
            cnt->too_short++;
        if (r->status & 16)
            cnt->prefix_fail++;
        /* no hit (only possible for degenerate queries, e.g. a constant signal whose z-score is NaN: the
         * reference's behaviour is undefined there): print nothing */
        if (db->rec[i].len_raw_signal == 0 || r->qlen <= 0 || r->rid < 0 || r->pos_st < 0 || r->pos_end < 0)
            return;
        /* src/sigfish.c:969-985 */
        aln_t *a = &db->aln[i];
        a->score = r->score;
        a->score2 = r->score2;
        a->rid = r->rid;
        a->d = r->strand ? '-' : '+';
        const int32_t rlen = ref->ref_lengths[r->rid];
        a->pos_st = r->strand ? rlen - r->pos_end : r->pos_st;
        a->pos_end = r->strand ? rlen - r->pos_st : r->pos_end;
        a->pos_st += ref->ref_st_offset[r->rid];
        a->pos_end += ref->ref_st_offset[r->rid];
        const float ratio = 500 * (a->score2 - a->score) / a->score;
        int32_t mq = to_int_like_x86(round((double)ratio));
        if (mq > 60)
            mq = 60;
        a->mapq = (uint8_t)mq;
        /* src/sigfish.c:800-807: query_size = (qend-1) - qstart */
        const uint64_t query_size = (uint64_t)(r->qend - 1) - (uint64_t)r->qstart;
        if (core->opt.flag & SIGFISH_SAM) {
            if (db->n_moves[i] < 0)
                return;
            const int q = core->opt.query_size;
            db->out[i] = sam_str(a, r, db->moves + db->move_off[i], db->n_moves[i], db->win_start + (size_t)i * q,
                                 db->win_len + (size_t)i * q, db->rec[i].read_id, ref->ref_names[r->rid],
                                 (core->opt.flag & SIGFISH_RNA) != 0);
        } else {
            db->out[i] = paf_str(a, db->rec[i].read_id, ref->ref_names[r->rid], r->start_raw, r->end_raw, query_size,
                                 db->rec[i].len_raw_signal, (uint64_t)ref->ref_seq_lengths[r->rid]);
        }
    }
}

static void *epilogue_worker(void *p)
{
    epi_arg_t *a = (epi_arg_t *)p;
    for (;;) {
        const int32_t b = __sync_fetch_and_add(a->next, SF_EPI_CHUNK);
        if (b >= a->db->n_rec)
            break;
        const int32_t e = b + SF_EPI_CHUNK < a->db->n_rec ? b + SF_EPI_CHUNK : a->db->n_rec;
        for (int32_t i = b; i < e; i++)
            epilogue_read(a->core, a->db, i, a);
    }
    return NULL;
}

void collect_db(core_t *core, db_t *db)
{
    if (!db->submitted)
        return;
    if (db->heads_only) {
        /* a record the device could not decode (malformed, or a stream its inflate rejects): the host decodes the
         * batch with its own reader -- which reports what is wrong with the record -- and submits the samples */
        int bad = 0;
        for (int g = 0; g < core->num_gpus; g++) {
            const int rc = sfgpu_collect(core->gpu[g], db->slot, db->res + db->shard_begin[g]);
            if (rc == SFGPU_EDECODE) {
                SF_WARNING("GPU %d: %s; decoding this batch on the host", g, sfgpu_strerror(core->gpu[g]));
                bad = 1;
            } else if (rc != SFGPU_OK) {
                SF_FATAL("GPU %d: %s", g, sfgpu_strerror(core->gpu[g]));
            }
        }
        if (bad) {
            core->decode_fallbacks++;
            db->heads_only = 0;
            parse_db_mode(core, db, 0);
            const int32_t keep = core->next_slot;
            core->next_slot = db->slot; /* same slot again; the rotation of the batches in flight is not disturbed */
            submit_db(core, db);
            core->next_slot = keep;
        }
    }
    for (int g = 0; g < core->num_gpus; g++) {
        const int b = db->shard_begin[g];
        if (sfgpu_collect(core->gpu[g], db->slot, db->res + b) != SFGPU_OK) {
            SF_FATAL("GPU %d: %s", g, sfgpu_strerror(core->gpu[g]));
        }
        if (core->opt.flag & SIGFISH_SAM) { /* the winners' warping paths + window event boundaries */
            const int e = db->shard_begin[g + 1];
            const int q = core->opt.query_size;
            db->move_off[b] = b == 0 ? 0 : db->move_off[b];
            for (int i = b; i < e; i++) {
                const sfgpu_result_t *r = &db->res[i];
                const int64_t need = r->qlen > 0 && r->rid >= 0 && r->pos_st >= 0 ? (int64_t)r->qlen + r->pos_end - r->pos_st : 0;
                db->move_off[i + 1] = db->move_off[i] + need;
            }
            const int64_t base = db->move_off[b];
            if ((size_t)db->move_off[e] > db->moves_cap) {
                db->moves_cap = (size_t)db->move_off[e] * 2 + 4096;
                db->moves = (uint8_t *)realloc(db->moves, db->moves_cap);
            }
            /* offsets relative to the shard */
            int64_t *rel = (int64_t *)malloc(sizeof(int64_t) * (size_t)(e - b + 1));
            for (int i = b; i <= e; i++)
                rel[i - b] = db->move_off[i] - base;
            if (sfgpu_collect_paths(core->gpu[g], db->slot, rel, db->moves + base, db->n_moves + b,
                                    db->win_start + (size_t)b * q, db->win_len + (size_t)b * q) != SFGPU_OK) {
                SF_FATAL("GPU %d: %s", g, sfgpu_strerror(core->gpu[g]));
            }
            free(rel);
        }
        sfgpu_timing_t t;
        if (sfgpu_timing(core->gpu[g], db->slot, &t) == SFGPU_OK) {
            if (g == 0) {
                core->h2d_time += t.h2d_ms * 1e-3;
                core->event_time += t.events_ms * 1e-3;
                core->dtw_time += (t.dtw_ms + t.trace_ms) * 1e-3;
                core->d2h_time += t.d2h_ms * 1e-3;
            }
            core->cells += t.cells;
        }
    }
    db->submitted = 0;
    /* the epilogue of the reads (strand flip, offsets, MAPQ, PAF / SAM text) on the -t worker threads, as the
     * reference does it inside work_db(); chunks of reads are claimed from a shared counter, db->out[] keeps the order */
    int nt = core->opt.num_thread < 1 ? 1 : core->opt.num_thread;
    if (nt > 32)
        nt = 32;
    if (db->n_rec < 2 * SF_EPI_CHUNK)
        nt = 1;
    int32_t next = 0;
    epi_arg_t args[32];
    pthread_t tid[32];
    for (int t = 0; t < nt; t++) {
        args[t].core = core;
        args[t].db = db;
        args[t].next = &next;
        args[t].ignored = args[t].too_short = args[t].prefix_fail = 0;
    }
    if (nt == 1) {
        epilogue_worker(&args[0]);
    } else {
        for (int t = 0; t < nt; t++)
            if (pthread_create(&tid[t], NULL, epilogue_worker, &args[t])) {
                SF_FATAL("%s", "pthread_create failed");
            }
        for (int t = 0; t < nt; t++)
            pthread_join(tid[t], NULL);
    }
    for (int t = 0; t < nt; t++) {
        db->ignored += args[t].ignored;
        db->too_short += args[t].too_short;
        db->prefix_fail += args[t].prefix_fail;
    }
}

void process_db(core_t *core, db_t *db)
{
    const double t0 = sf_realtime();
    parse_db(core, db);
    submit_db(core, db);
    collect_db(core, db);
    core->process_db_time += sf_realtime() - t0;
}

void output_db(core_t *core, db_t *db)
{
    const double t0 = sf_realtime();
    for (int i = 0; i < db->n_rec; i++)
        if (db->out[i])
            fputs(db->out[i], stdout);
    fflush(stdout);
    core->sum_bytes += db->sum_bytes;
    core->total_reads += db->total_reads;
    core->prefix_fail += db->prefix_fail;
    core->ignored += db->ignored;
    core->too_short += db->too_short;
    core->output_time += sf_realtime() - t0;
}

void free_db_tmp(db_t *db)
{
    for (int i = 0; i < db->n_rec; i++) {
        free(db->out[i]);
        db->out[i] = NULL;
    }
}

void free_db(db_t *db)
{
    for (int i = 0; i < db->capacity_rec; i++) {
        if (!db->mem_views)
            free(db->mem_records[i]);
        free(db->rec[i].read_id);
        free(db->rec[i].raw_signal);
    }
    free(db->mem_records); free(db->mem_bytes); free(db->mem_cap); free(db->rec); free(db->res); free(db->aln);
    free(db->out); free(db->sig_off); free(db->sig_ptr); free(db->dig); free(db->off); free(db->rng);
    free(db->move_off); free(db->n_moves); free(db->win_start); free(db->win_len); free(db->moves);
    free(db->sig_pos); free(db->sig_bytes); free(db->rec_bytes);
    free(db);
}

void sam_hdr_wr(const refsynth_t *ref)
{
    /* src/dtw_main.c:118-123: LN is the k-mer count, as in the reference */
    for (int i = 0; i < ref->num_ref; i++)
        printf("@SQ\tSN:%s\tLN:%ld\n", ref->ref_names[i], (long)ref->ref_lengths[i]);
    printf("@PG\tID:sigfish\tPN:sigfish\tVN:%s\n", SFHOST_VERSION_SHORT);
}
