/* sfhost.h -- host side of the B200 `dtw` path: the batch pipeline of the reference
 * (src/sigfish.h:274-306 -- init_opt / init_core / init_db / load_db / process_db / output_db /
 * free_db_tmp / free_db / free_core, same names, argument meaning and error behaviour) re-plumbed
 * around the C-ABI of include/sfgpu.h.
 *
 * What changed with respect to the reference (paths relative to /root/reference):
 *   - process_db() (src/sigfish.c:1018-1047) no longer fans reads out to pthreads (src/thread.c); it
 *     packs the batch, shards it over the GPUs and runs the device stages.  submit_db()/collect_db()
 *     expose the two halves so that the main loop can load batch n+1 while batch n is on the GPUs.
 *   - host threads (-t) are used only for decoding the BLOW5 records (parse_single, sigfish.c:317-328).
 *   - FASTA, k-mer model and SLOW5/BLOW5 reading are done by the small readers in this directory
 *     (the reference links kseq.h and slow5lib).
 *   - the built-in pore-model tables are absent (src/model.h is missing from the reference mount):
 *     a model file (--kmer-model) is required.
 * Errors: like the reference, fatal conditions print "[fn::ERROR] ..." to stderr and exit(EXIT_FAILURE).
 */
#ifndef SFHOST_H
#define SFHOST_H

#include <stdint.h>
#include <stdio.h>

#include "sfgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SFHOST_VERSION "0.1.0-b200 (sigfish 0.2.0 dtw path)"
#define SFHOST_VERSION_SHORT "0.2.0-b200"

/* option flags: same bits as the reference (src/sigfish.h:30-39) */
#define SIGFISH_RNA 0x001
#define SIGFISH_DTW 0x002
#define SIGFISH_INV 0x004
#define SIGFISH_SEC 0x008
#define SIGFISH_REF 0x010
#define SIGFISH_END 0x020
#define SIGFISH_PRF 0x040
#define SIGFISH_ACC 0x080
#define SIGFISH_SAM 0x100
#define SIGFISH_R10 0x200

#define OPT_PORE_R9 0
#define OPT_PORE_R10 1
#define OPT_PORE_RNA004 2

#define SFHOST_MAX_GPUS 16

/* user options: the fields of the reference's opt_t (src/sigfish.h:121-139) plus the GPU selection */
typedef struct {
    const char *model_file;
    const char *meth_model_file; /* accepted, unused (dead in the reference too) */
    uint32_t flag;
    int32_t batch_size;       /* -K */
    int64_t batch_size_bytes; /* -B */
    char *pore;
    int8_t pore_flag;
    int8_t device_decode; /* --device-decode=yes|no (default yes): decode BLOW5 records on the GPUs */
    int32_t num_thread; /* -t: host threads decoding records */
    int8_t verbosity;
    int32_t debug_break;
    char *region_str;
    int32_t prefix_size; /* -p */
    int32_t query_size;  /* -q */
    /* B200 additions */
    int32_t num_gpus;    /* --gpus N (0: all visible) */
    int32_t first_gpu;   /* --gpu-first ordinal */
} opt_t;

/* resident reference description (refsynth_t, src/sigfish.h:90-99, minus the event arrays, which
 * live in HBM) */
typedef struct {
    int32_t num_ref;
    char **ref_names;
    int32_t *ref_lengths;
    int32_t *ref_seq_lengths;
    int32_t *ref_st_offset;
} refsynth_t;

typedef struct {
    int32_t rid;
    int32_t pos_st;
    int32_t pos_end;
    float score;
    float score2;
    char d;
    uint8_t mapq;
} aln_t;

/* one decoded read (the fields of slow5_rec_t the path uses) */
typedef struct {
    char *read_id;
    double digitisation, offset, range, sampling_rate;
    uint64_t len_raw_signal;
    int16_t *raw_signal;
    size_t cap_signal; /* allocated samples (buffers are reused across batches like the reference's) */
} sf_rec_t;

struct sf_s5file;
typedef struct sf_s5file sf_s5file_t;

/* a batch (db_t, src/sigfish.h:161-197) */
typedef struct {
    int32_t n_rec;
    int32_t capacity_rec;
    char **mem_records;
    size_t *mem_bytes;
    size_t *mem_cap;
    int mem_views;      /* mem_records[] point into the mapped file (not owned, read only) */
    sf_rec_t *rec;
    sfgpu_result_t *res; /* device results of the batch */
    aln_t *aln;
    char **out;
    /* records decoded on the device: where the signal field lies in each decompressed record */
    int32_t *sig_pos;
    int64_t *sig_bytes;
    int64_t *rec_bytes;
    int heads_only;     /* parse_db read only the heads of this batch's records */
    /* packing scratch */
    int64_t *sig_off;   /* per-read sample counts */
    int16_t **sig_ptr;  /* per-read signal buffers */
    float *dig, *off, *rng;
    int32_t shard_begin[SFHOST_MAX_GPUS + 1];
    /* --sam: winners' paths and window event boundaries */
    int64_t *move_off;
    uint8_t *moves;
    size_t moves_cap;
    int32_t *n_moves;
    uint64_t *win_start;
    float *win_len;
    int32_t slot;        /* device slot this batch was submitted to */
    int submitted;
    int last_batch;      /* set by the loader: the file ended with this batch */
    /* stats */
    int64_t sum_bytes;
    int64_t total_reads;
    int64_t prefix_fail;
    int64_t ignored;
    int64_t too_short;
} db_t;

/* core (core_t, src/sigfish.h:202-244) */
typedef struct {
    sf_s5file_t *sf;
    float *level_mean; /* model[i].level_mean */
    uint32_t kmer_size;
    opt_t opt;
    double realtime0;
    double load_db_time, process_db_time, output_time;
    double parse_time, event_time, normalise_time, dtw_time; /* device stage times (CUDA events) */
    double h2d_time, d2h_time;
    int64_t sum_bytes, total_reads, prefix_fail, ignored, too_short;
    refsynth_t *ref;
    int32_t num_gpus;
    sfgpu_ctx *gpu[SFHOST_MAX_GPUS];
    int32_t next_slot;
    double cells; /* DTW cells computed so far */
    int device_decode;        /* BLOW5 records are inflated / svb-zd decoded on the GPUs (sfgpu_submit_records) */
    int64_t decode_fallbacks; /* batches the host had to decode after all */
} core_t;

typedef struct {
    int32_t num_reads;
    int64_t num_bytes;
} ret_status_t;

void init_opt(opt_t *opt);
core_t *init_core(const char *fastafile, char *slow5file, opt_t opt, double realtime0);
db_t *init_db(core_t *core);
ret_status_t load_db(core_t *core, db_t *db);
void process_db(core_t *core, db_t *db);
/* the three parts of process_db: decode the records on the host threads / pack + launch on the GPUs /
 * wait + per-read epilogue.  dtw_cli.c runs them as pipeline stages on different batches at once. */
void parse_db(core_t *core, db_t *db);
void submit_db(core_t *core, db_t *db);
void collect_db(core_t *core, db_t *db);
void output_db(core_t *core, db_t *db);
void free_db_tmp(db_t *db);
void free_db(db_t *db);
void free_core(core_t *core, opt_t opt);

/* contiguous read ranges per GPU with equal sample counts: begin[0..G] */
void sf_shard_ranges(int32_t n_rec, const int64_t *weight, int32_t G, int32_t *begin);

/* writes the @SQ header lines of --sam output (src/dtw_main.c:118-123) */
void sam_hdr_wr(const refsynth_t *ref);

/* `sigfish dtw` command line (src/dtw_main.c) */
int dtw_main(int argc, char *argv[]);

/* `sigfish eval truth.paf test.paf` (src/eval.c) */
int eval_main(int argc, char *argv[]);

/* helpers shared by the sources of this directory */
double sf_realtime(void);
double sf_cputime(void);
long sf_peakrss(void);
extern int8_t sf_verbosity;

#define SF_ERROR(fmt, ...) fprintf(stderr, "[%s::ERROR]\033[1;31m " fmt "\033[0m\n", __func__, __VA_ARGS__)
#define SF_WARNING(fmt, ...)                                                                       \
    do {                                                                                           \
        if (sf_verbosity >= 2)                                                                     \
            fprintf(stderr, "[%s::WARNING]\033[1;33m " fmt "\033[0m\n", __func__, __VA_ARGS__);     \
    } while (0)
#define SF_INFO(fmt, ...)                                                                          \
    do {                                                                                           \
        if (sf_verbosity >= 3)                                                                     \
            fprintf(stderr, "[%s::INFO]\033[1;34m " fmt "\033[0m\n", __func__, __VA_ARGS__);        \
    } while (0)
#define SF_FATAL(fmt, ...)                                                                         \
    do {                                                                                           \
        SF_ERROR(fmt, __VA_ARGS__);                                                                \
        exit(EXIT_FAILURE);                                                                        \
    } while (0)

#ifdef __cplusplus
}
#endif
#endif
