/* sfgpu_oracle.c -- TEST INFRASTRUCTURE ONLY: a test double of libsfgpu.so that answers include/sfgpu.h with the CPU
 * oracle (oracle/liboracle.so).
 *
 * Why: the product's host code (sigfish_b200/host/: readers, batch loop, sharding over contexts, decode fallback,
 * epilogue, PAF / SAM writers, threads) can then be run end to end WITHOUT a GPU and compared byte for byte with what
 * the unmodified reference binary printed (tests/golden/) -- tests/test_host_cpu_parity.py.  It is built by that test
 * into tests/mockdev/_build/ under the name libsfgpu.so, next to a private build of the host sources; it is never
 * linked into, shipped with or loaded by the product (sigfish_b200/, bench.py).  The real library is the CUDA one.
 *
 * Knobs (environment): MOCK_GPUS = devices reported (default 1); MOCK_REJECT_RECORDS = 1 makes sfgpu_collect()
 * answer SFGPU_EDECODE for batches submitted as records, which sends the host down its own-decoder fallback. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "oracle.h"
#include "sfgpu.h"

#define MOCK_SLOTS 4

typedef struct {
    int32_t n;
    int reject;
    sfgpu_result_t *res;
    /* --sam: per read the backward moves of the winner's path and the window's event boundaries */
    uint8_t **moves;
    int32_t *n_moves;
    uint64_t *ev_start;
    float *ev_len;
    double cells;
    int64_t samples;
} mock_slot;

struct sfgpu_ctx {
    sfgpu_opt_t opt;
    float *level_mean;
    orc_ref_t *ref;
    int64_t ref_columns;
    mock_slot slot[MOCK_SLOTS];
    char err[256];
};

static char g_err[256] = "";

static int fail(sfgpu_ctx *c, int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    if (c)
        snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}

int sfgpu_device_count(void)
{
    const char *e = getenv("MOCK_GPUS");
    return e ? atoi(e) : 1;
}

int sfgpu_create(sfgpu_ctx **ctx, const sfgpu_opt_t *opt, const float *level_mean)
{
    if (!ctx || !opt || !level_mean)
        return fail(NULL, SFGPU_EARG, "null argument");
    if (opt->query_size < 1 || opt->query_size > 1024)
        return fail(NULL, SFGPU_ELIMIT, "query_size out of range (1 .. 1024)");
    if (opt->kmer_size < 1 || opt->kmer_size > 12)
        return fail(NULL, SFGPU_EARG, "bad kmer_size");
    sfgpu_ctx *c = (sfgpu_ctx *)calloc(1, sizeof *c);
    c->opt = *opt;
    const size_t n = (size_t)1 << (2 * opt->kmer_size);
    c->level_mean = (float *)malloc(sizeof(float) * n);
    memcpy(c->level_mean, level_mean, sizeof(float) * n);
    *ctx = c;
    return SFGPU_OK;
}

static uint32_t oracle_flags(const sfgpu_ctx *c)
{
    uint32_t f = c->opt.flags & (SFGPU_RNA | SFGPU_DTW | SFGPU_INV | SFGPU_REF | SFGPU_END);
    if (c->opt.pore == 2)
        f |= ORC_RNA004;
    return f;
}

int sfgpu_set_ref(sfgpu_ctx *c, int32_t num_ref, const char *bases, const int64_t *base_off, int32_t *ref_lengths,
                  int32_t *ref_seq_lengths, int32_t *ref_st_offset)
{
    if (!c || num_ref < 1 || !bases || !base_off)
        return fail(c, SFGPU_EARG, "bad reference");
    const char **seqs = (const char **)malloc(sizeof(char *) * (size_t)num_ref);
    int32_t *lens = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_ref);
    for (int32_t i = 0; i < num_ref; i++) {
        seqs[i] = bases + base_off[i];
        lens[i] = (int32_t)(base_off[i + 1] - base_off[i]);
        if (base_off[i + 1] - base_off[i] < c->opt.kmer_size) { /* as libsfgpu.so does (sfgpu.cu: sfgpu_set_ref) */
            free(seqs);
            free(lens);
            return fail(c, SFGPU_EARG, "a contig is shorter than the k-mer size");
        }
    }
    if (c->ref)
        orc_ref_free(c->ref);
    c->ref = orc_ref_build(num_ref, seqs, lens, c->level_mean, c->opt.kmer_size, oracle_flags(c), c->opt.query_size);
    free(seqs);
    free(lens);
    if (!c->ref)
        return fail(c, SFGPU_EARG, "reference could not be built");
    c->ref_columns = 0;
    for (int32_t i = 0; i < num_ref; i++) {
        if (ref_lengths) ref_lengths[i] = c->ref->ref_lengths[i];
        if (ref_seq_lengths) ref_seq_lengths[i] = c->ref->ref_seq_lengths[i];
        if (ref_st_offset) ref_st_offset[i] = c->ref->ref_st_offset[i];
        c->ref_columns += (int64_t)c->ref->ref_lengths[i] * (c->ref->has_reverse ? 2 : 1);
    }
    return SFGPU_OK;
}

static void slot_clear(mock_slot *s)
{
    if (s->moves)
        for (int32_t i = 0; i < s->n; i++)
            free(s->moves[i]);
    free(s->moves);
    free(s->n_moves);
    free(s->ev_start);
    free(s->ev_len);
    free(s->res);
    memset(s, 0, sizeof *s);
}

static int slot_begin(sfgpu_ctx *c, int32_t slot, int32_t n)
{
    if (!c)
        return fail(NULL, SFGPU_EARG, "null context");
    if (!c->ref)
        return fail(c, SFGPU_ESTATE, "submit before sfgpu_set_ref");
    if (slot < 0 || slot >= MOCK_SLOTS || n < 0)
        return fail(c, SFGPU_EARG, "bad slot or read count");
    mock_slot *s = &c->slot[slot];
    slot_clear(s);
    s->n = n;
    s->res = (sfgpu_result_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(sfgpu_result_t));
    if (c->opt.flags & SFGPU_SAM) {
        const size_t q = (size_t)c->opt.query_size;
        s->moves = (uint8_t **)calloc((size_t)(n > 0 ? n : 1), sizeof(uint8_t *));
        s->n_moves = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        s->ev_start = (uint64_t *)calloc((size_t)(n > 0 ? n : 1) * q, sizeof(uint64_t));
        s->ev_len = (float *)calloc((size_t)(n > 0 ? n : 1) * q, sizeof(float));
    }
    return SFGPU_OK;
}

/* one read through the oracle -> the fields the device leaves in sfgpu_result_t */
static void map_one(sfgpu_ctx *c, mock_slot *s, int32_t i, const int16_t *raw, int64_t n, float dig, float off, float range)
{
    const uint32_t flags = oracle_flags(c);
    orc_hit_t h;
    orc_event_t *ev = NULL;
    orc_map_read_events(c->ref, raw, n, dig, off, range, flags, c->opt.query_size, c->opt.prefix_size, &h, &ev);
    sfgpu_result_t *r = &s->res[i];
    memset(r, 0, sizeof *r);
    r->rid = -1;
    r->pos_st = r->pos_end = -1;
    r->n_events = h.n_events > 0 ? h.n_events : 0;
    r->status = (h.status & 3) | ((h.status & 4) ? 16 : 0) | (n > 0 && h.n_events <= 0 ? 4 : 0);
    r->qstart = (int32_t)h.qstart;
    r->qend = (int32_t)h.qend;
    s->samples += n;
    if (s->n_moves)
        s->n_moves[i] = -1;
    if (h.mapped) {
        r->qlen = (int32_t)(h.qend - h.qstart);
        r->start_raw = h.start_raw;
        r->end_raw = h.end_raw;
        r->score = h.score;
        r->score2 = h.score2;
        r->rid = h.rid;
        r->strand = h.strand == '-';
        r->pos_st = h.raw_pos_st;
        r->pos_end = h.raw_pos_end;
        s->cells += (double)r->qlen * (double)c->ref_columns;
        if (s->moves && h.rid >= 0 && h.raw_pos_st >= 0) {
            /* the winner's path: rebuild its cost matrix as the reference's update_aln() does (sigfish.c:599-618) */
            const int qlen = r->qlen, rlen = c->ref->ref_lengths[h.rid];
            const int rna = (flags & ORC_RNA) != 0;
            float *query = (float *)malloc(sizeof(float) * (size_t)qlen);
            for (int j = 0; j < qlen; j++) {
                if (rna && !(flags & ORC_INV))
                    query[qlen - 1 - j] = ev[h.qstart + j].mean;
                else
                    query[j] = ev[h.qstart + j].mean;
            }
            const float *y = h.strand == '+' ? c->ref->forward[h.rid] : c->ref->reverse[h.rid];
            float *cost = (float *)malloc(sizeof(float) * (size_t)qlen * (size_t)rlen);
            if (flags & ORC_DTW)
                orc_std_dtw(query, y, qlen, rlen, cost);
            else
                orc_subsequence(query, y, qlen, rlen, cost);
            int32_t *px = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + rlen));
            int32_t *py = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + rlen));
            const int32_t k = orc_path_full(cost, qlen, rlen, h.raw_pos_end, px, py); /* start -> end */
            s->moves[i] = (uint8_t *)malloc((size_t)(k > 1 ? k - 1 : 1));
            for (int32_t t = k - 1, m = 0; t > 0; t--, m++) {
                const int di = px[t] - px[t - 1], dj = py[t] - py[t - 1];
                s->moves[i][m] = (uint8_t)(di && dj ? 0 : (dj ? 1 : 2));
            }
            s->n_moves[i] = k - 1;
            for (int j = 0; j < qlen; j++) {
                s->ev_start[(size_t)i * (size_t)c->opt.query_size + (size_t)j] = ev[h.qstart + j].start;
                s->ev_len[(size_t)i * (size_t)c->opt.query_size + (size_t)j] = ev[h.qstart + j].length;
            }
            free(query);
            free(cost);
            free(px);
            free(py);
        }
    }
    orc_free(ev);
}

int sfgpu_submit_reads(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const int16_t *const *signals, const int64_t *n_samples,
                       const float *digitisation, const float *offset, const float *range)
{
    const int rc = slot_begin(c, slot, n_reads);
    if (rc)
        return rc;
    for (int32_t i = 0; i < n_reads; i++)
        map_one(c, &c->slot[slot], i, signals[i], n_samples[i], digitisation[i], offset[i], range[i]);
    return SFGPU_OK;
}

int sfgpu_submit(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const int16_t *signals, const int64_t *sig_off,
                 const float *digitisation, const float *offset, const float *range)
{
    const int rc = slot_begin(c, slot, n_reads);
    if (rc)
        return rc;
    for (int32_t i = 0; i < n_reads; i++)
        map_one(c, &c->slot[slot], i, signals + sig_off[i], sig_off[i + 1] - sig_off[i], digitisation[i], offset[i], range[i]);
    return SFGPU_OK;
}

/* svb-zd: u32 count, 2-bit byte counts, zigzag deltas (slow5lib src/slow5_press.c:1085-1133); 0 or -1 */
static int svb_zd(const uint8_t *in, int64_t n_in, int16_t *out, int64_t n)
{
    if (n_in < 4)
        return n == 0 ? 0 : -1;
    uint32_t count;
    memcpy(&count, in, 4);
    if ((int64_t)count != n)
        return -1;
    const int64_t n_key = ((int64_t)count + 3) / 4;
    if (4 + n_key > n_in)
        return -1;
    const uint8_t *key = in + 4, *data = in + 4 + n_key, *end = in + n_in;
    int32_t prev = 0;
    for (int64_t i = 0; i < n; i++) {
        const int nb = ((key[i >> 2] >> (2 * (i & 3))) & 3) + 1;
        if (end - data < nb)
            return -1;
        uint32_t v = 0;
        for (int b = 0; b < nb; b++)
            v |= (uint32_t)data[b] << (8 * b);
        data += nb;
        prev += (int32_t)((v >> 1) ^ (0u - (v & 1u)));
        out[i] = (int16_t)prev;
    }
    return 0;
}

int sfgpu_submit_records(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const uint8_t *const *records, const int64_t *record_bytes,
                         int32_t record_press, int32_t signal_press, const int32_t *sig_pos, const int64_t *sig_bytes,
                         const int64_t *n_samples, const float *digitisation, const float *offset, const float *range)
{
    if (record_press < 0 || record_press > 1 || signal_press < 0 || signal_press > 1)
        return fail(c, SFGPU_ELIMIT, "compression method not decoded by the device");
    const int rc = slot_begin(c, slot, n_reads);
    if (rc)
        return rc;
    mock_slot *s = &c->slot[slot];
    const char *rej = getenv("MOCK_REJECT_RECORDS");
    if (rej && atoi(rej)) {
        s->reject = 1;
        return SFGPU_OK;
    }
    for (int32_t i = 0; i < n_reads && !s->reject; i++) {
        const int64_t need = (int64_t)sig_pos[i] + sig_bytes[i];
        const uint8_t *rec = records[i];
        uint8_t *plain = NULL;
        if (record_press) { /* the whole record, trailer checked, as the device does */
            uLongf cap = (uLongf)need + 4096;
            int zrc = Z_BUF_ERROR;
            while (zrc == Z_BUF_ERROR && cap < ((uLongf)1 << 31)) {
                free(plain);
                plain = (uint8_t *)malloc(cap);
                uLongf got = cap;
                zrc = uncompress(plain, &got, rec, (uLong)record_bytes[i]);
                if (zrc == Z_OK && (int64_t)got < need)
                    zrc = Z_DATA_ERROR;
                cap *= 4;
            }
            if (zrc != Z_OK) {
                free(plain);
                s->reject = 1;
                break;
            }
            rec = plain;
        } else if (need > record_bytes[i]) {
            s->reject = 1;
            break;
        }
        int16_t *sig = (int16_t *)malloc(sizeof(int16_t) * (size_t)(n_samples[i] > 0 ? n_samples[i] : 1));
        if (signal_press) {
            if (svb_zd(rec + sig_pos[i], sig_bytes[i], sig, n_samples[i]))
                s->reject = 1;
        } else if (sig_bytes[i] != 2 * n_samples[i]) {
            s->reject = 1;
        } else {
            memcpy(sig, rec + sig_pos[i], (size_t)sig_bytes[i]);
        }
        if (!s->reject)
            map_one(c, s, i, sig, n_samples[i], digitisation[i], offset[i], range[i]);
        free(sig);
        free(plain);
    }
    return SFGPU_OK;
}

int sfgpu_resubmit(sfgpu_ctx *c, int32_t slot)
{
    (void)slot;
    return fail(c, SFGPU_ESTATE, "test double: nothing is resident");
}

int sfgpu_collect(sfgpu_ctx *c, int32_t slot, sfgpu_result_t *out)
{
    if (!c || slot < 0 || slot >= MOCK_SLOTS)
        return fail(c, SFGPU_EARG, "bad slot");
    mock_slot *s = &c->slot[slot];
    if (s->reject) {
        s->reject = 0;
        return fail(c, SFGPU_EDECODE, "test double: a record of the batch was not decoded");
    }
    if (s->n > 0 && out)
        memcpy(out, s->res, sizeof(sfgpu_result_t) * (size_t)s->n);
    return SFGPU_OK;
}

int sfgpu_collect_paths(sfgpu_ctx *c, int32_t slot, const int64_t *move_off, uint8_t *moves, int32_t *n_moves,
                        uint64_t *ev_start, float *ev_len)
{
    if (!c || slot < 0 || slot >= MOCK_SLOTS || !(c->opt.flags & SFGPU_SAM))
        return fail(c, SFGPU_ESTATE, "paths need a context created with SFGPU_SAM");
    mock_slot *s = &c->slot[slot];
    const size_t q = (size_t)c->opt.query_size;
    for (int32_t i = 0; i < s->n; i++) {
        n_moves[i] = s->n_moves[i];
        if (s->n_moves[i] > 0) {
            if (move_off[i + 1] - move_off[i] < s->n_moves[i])
                return fail(c, SFGPU_EARG, "move buffer too small");
            memcpy(moves + move_off[i], s->moves[i], (size_t)s->n_moves[i]);
        }
    }
    memcpy(ev_start, s->ev_start, sizeof(uint64_t) * (size_t)s->n * q);
    memcpy(ev_len, s->ev_len, sizeof(float) * (size_t)s->n * q);
    return SFGPU_OK;
}

int sfgpu_timing(sfgpu_ctx *c, int32_t slot, sfgpu_timing_t *t)
{
    if (!c || slot < 0 || slot >= MOCK_SLOTS || !t)
        return SFGPU_EARG;
    memset(t, 0, sizeof *t);
    t->cells = c->slot[slot].cells;
    t->samples = c->slot[slot].samples;
    return SFGPU_OK;
}

void sfgpu_destroy(sfgpu_ctx *c)
{
    if (!c)
        return;
    for (int i = 0; i < MOCK_SLOTS; i++)
        slot_clear(&c->slot[i]);
    if (c->ref)
        orc_ref_free(c->ref);
    free(c->level_mean);
    free(c);
}

const char *sfgpu_strerror(const sfgpu_ctx *c) { return c ? c->err : g_err; }

/* the inspection entry points are the CUDA library's business */
int sfgpu_ref_events(sfgpu_ctx *c, int32_t rid, int32_t strand, float *out, int32_t cap)
{
    (void)rid; (void)strand; (void)out; (void)cap;
    return fail(c, SFGPU_ESTATE, "test double");
}
int64_t sfgpu_event_table(sfgpu_ctx *c, const int16_t *signal, int64_t n_samples, float digitisation, float offset, float range,
                          uint64_t *start, float *length, float *mean, int64_t cap)
{
    (void)signal; (void)n_samples; (void)digitisation; (void)offset; (void)range; (void)start; (void)length; (void)mean; (void)cap;
    return fail(c, SFGPU_ESTATE, "test double");
}
int64_t sfgpu_slot_signal(sfgpu_ctx *c, int32_t slot, int32_t read, int16_t *out, int64_t cap)
{
    (void)slot; (void)read; (void)out; (void)cap;
    return fail(c, SFGPU_ESTATE, "test double");
}
int sfgpu_query(sfgpu_ctx *c, int32_t slot, int32_t read, float *out, int32_t cap)
{
    (void)slot; (void)read; (void)out; (void)cap;
    return fail(c, SFGPU_ESTATE, "test double");
}
int sfgpu_set_ref_events(sfgpu_ctx *c, int32_t num_ref, int32_t has_reverse, const float *events, const int64_t *ev_off)
{
    (void)num_ref; (void)has_reverse; (void)events; (void)ev_off;
    return fail(c, SFGPU_ESTATE, "test double");
}
int sfgpu_submit_queries(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const float *queries, const int32_t *qlen)
{
    (void)slot; (void)n_reads; (void)queries; (void)qlen;
    return fail(c, SFGPU_ESTATE, "test double");
}
int32_t sfgpu_wave_reads(const sfgpu_ctx *c) { (void)c; return 64; }
int64_t sfgpu_ref_columns(const sfgpu_ctx *c) { return c ? c->ref_columns : 0; }
