/* sfgpu_null.c -- a NULL device behind include/sfgpu.h, for profiling the HOST side of `sigfish-b200 dtw` only
 * (tools/hostfeed/run.py).  It is NOT a fallback and is never linked into the product: it aligns nothing.
 * sfgpu_submit_reads() copies the samples into a staging buffer (the cost of filling the pinned buffer) and
 * sfgpu_collect() returns made-up hits so that the PAF epilogue has something to print.  The numbers it yields
 * are the read rate at which the host code alone (BLOW5 load, record decode, sharding, staging copy, epilogue,
 * output) saturates, i.e. the ceiling the GPUs can be fed at. */
#include <stdlib.h>
#include <string.h>

#include "sfgpu.h"

#define NULL_SLOTS 4

struct sfgpu_ctx {
    sfgpu_opt_t opt;
    int32_t ref_len;
    int32_t n[NULL_SLOTS];
    int16_t *stage[NULL_SLOTS];
    int64_t cap[NULL_SLOTS];
};

int sfgpu_device_count(void)
{
    const char *e = getenv("HOSTFEED_GPUS");
    return e ? atoi(e) : 8;
}

int sfgpu_create(sfgpu_ctx **ctx, const sfgpu_opt_t *opt, const float *level_mean)
{
    (void)level_mean;
    sfgpu_ctx *c = calloc(1, sizeof *c);
    if (!c)
        return SFGPU_ELIMIT;
    c->opt = *opt;
    *ctx = c;
    return SFGPU_OK;
}

int sfgpu_set_ref(sfgpu_ctx *c, int32_t num_ref, const char *bases, const int64_t *base_off, int32_t *ref_lengths,
                  int32_t *ref_seq_lengths, int32_t *ref_st_offset)
{
    (void)bases;
    for (int32_t i = 0; i < num_ref; i++) {
        const int32_t len = (int32_t)(base_off[i + 1] - base_off[i]);
        if (ref_lengths) ref_lengths[i] = len + 1 - c->opt.kmer_size;
        if (ref_seq_lengths) ref_seq_lengths[i] = len;
        if (ref_st_offset) ref_st_offset[i] = 0;
        if (i == 0) c->ref_len = len + 1 - c->opt.kmer_size;
    }
    return SFGPU_OK;
}

/* records decoded on the device: the host's part is the copy of the compressed bytes into the staging buffer */
int sfgpu_submit_records(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const uint8_t *const *records, const int64_t *record_bytes,
                         int32_t record_press, int32_t signal_press, const int32_t *sig_pos, const int64_t *sig_bytes,
                         const int64_t *n_samples, const float *digitisation, const float *offset, const float *range)
{
    (void)record_press; (void)signal_press; (void)sig_pos; (void)sig_bytes; (void)n_samples;
    (void)digitisation; (void)offset; (void)range;
    if (slot < 0 || slot >= NULL_SLOTS)
        return SFGPU_EARG;
    int64_t tot = 0;
    for (int32_t i = 0; i < n_reads; i++)
        tot += (record_bytes[i] + 7) & ~7ll;
    tot = (tot + 1) / 2; /* the staging buffer is counted in int16 */
    if (tot > c->cap[slot]) {
        free(c->stage[slot]);
        c->cap[slot] = tot + tot / 4;
        c->stage[slot] = malloc(sizeof(int16_t) * c->cap[slot]);
        if (!c->stage[slot])
            return SFGPU_ELIMIT;
    }
    int64_t o = 0;
    for (int32_t i = 0; i < n_reads; i++) {
        memcpy((char *)c->stage[slot] + o, records[i], (size_t)record_bytes[i]);
        o += (record_bytes[i] + 7) & ~7ll;
    }
    c->n[slot] = n_reads;
    return SFGPU_OK;
}

int sfgpu_submit_reads(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const int16_t *const *signals,
                       const int64_t *n_samples, const float *digitisation, const float *offset, const float *range)
{
    (void)digitisation; (void)offset; (void)range;
    if (slot < 0 || slot >= NULL_SLOTS)
        return SFGPU_EARG;
    int64_t tot = 0;
    for (int32_t i = 0; i < n_reads; i++)
        tot += (n_samples[i] + 7) & ~7ll;
    if (tot > c->cap[slot]) {
        free(c->stage[slot]);
        c->cap[slot] = tot + tot / 4;
        c->stage[slot] = malloc(sizeof(int16_t) * c->cap[slot]);
        if (!c->stage[slot])
            return SFGPU_ELIMIT;
    }
    int64_t o = 0;
    for (int32_t i = 0; i < n_reads; i++) {
        memcpy(c->stage[slot] + o, signals[i], sizeof(int16_t) * n_samples[i]);
        o += (n_samples[i] + 7) & ~7ll;
    }
    c->n[slot] = n_reads;
    return SFGPU_OK;
}

int sfgpu_collect(sfgpu_ctx *c, int32_t slot, sfgpu_result_t *out)
{
    const int32_t span = c->ref_len > 1000 ? c->ref_len - 600 : 1;
    for (int32_t i = 0; i < c->n[slot]; i++) {
        sfgpu_result_t *r = &out[i];
        r->n_events = 600; r->qstart = c->opt.prefix_size > 0 ? c->opt.prefix_size : 50;
        r->qend = r->qstart + c->opt.query_size; r->qlen = c->opt.query_size; r->status = 0;
        r->start_raw = 500; r->end_raw = 3000;
        r->score = 30.5f + (float)(i % 7); r->score2 = r->score + 1.25f;
        r->rid = 0; r->strand = i & 1;
        r->pos_st = (int32_t)(((int64_t)i * 977) % span); r->pos_end = r->pos_st + 190;
    }
    return SFGPU_OK;
}

int sfgpu_collect_paths(sfgpu_ctx *c, int32_t slot, const int64_t *move_off, uint8_t *moves, int32_t *n_moves,
                        uint64_t *ev_start, float *ev_len)
{
    (void)c; (void)slot; (void)move_off; (void)moves; (void)n_moves; (void)ev_start; (void)ev_len;
    return SFGPU_EARG;
}

int sfgpu_timing(sfgpu_ctx *c, int32_t slot, sfgpu_timing_t *t)
{
    (void)c; (void)slot;
    memset(t, 0, sizeof *t);
    return SFGPU_OK;
}

int32_t sfgpu_wave_reads(const sfgpu_ctx *c) { (void)c; return 4144; }
int64_t sfgpu_ref_columns(const sfgpu_ctx *c) { (void)c; return 2 * 999992; }
const char *sfgpu_strerror(const sfgpu_ctx *c) { (void)c; return "null device"; }

void sfgpu_destroy(sfgpu_ctx *c)
{
    if (!c)
        return;
    for (int i = 0; i < NULL_SLOTS; i++)
        free(c->stage[i]);
    free(c);
}
