/* oracle_map.c -- per-read driver restated on the CPU (test oracle; see oracle.h).
 *
 * Follows reference src/sigfish.c:
 *   normalise_single 424-505, init_aln 507-520, update_aln 575-626,
 *   paf_str 628-660, aln_to_str 796-826, dtw_single 828-985.
 */
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define KEEP 5 /* SECONDARY_CAP, sigfish.h:41 */

typedef struct {
    float score;
    int32_t rid;
    int32_t pos_st;
    int32_t pos_end;
    char strand;
} cand_t;

/* sigfish.c:424-505.  Window selection, then z-score of the event means inside it. */
int orc_window_normalise(orc_event_t *ev, int64_t *n_events, uint32_t flags, int32_t q, int32_t p,
                         int64_t *qstart, int64_t *qend, int32_t *status)
{
    int64_t n = *n_events;
    *status = 0;
    *qstart = 0;
    *qend = 0;
    if (n <= 0)
        return 0;
    int64_t lo, hi;
    if (!(flags & ORC_END)) { /* sigfish.c:435-462 (auto query start, p < 0, is not restated here) */
        lo = p;
        hi = lo + q;
        if (lo + 25 > n) {
            lo = hi = 0;
            n = 0;
            *status |= 1;
        } else if (hi > n) {
            hi = n;
            *status |= 2;
        }
    } else { /* sigfish.c:464-478 */
        lo = n - p - q;
        hi = n - p;
        if (lo < 0) {
            lo = 0;
            *status |= 2;
        }
        if (hi < 0) {
            hi = 0;
            n = 0;
            *status |= 1;
        }
    }
    *qstart = lo;
    *qend = hi;
    *n_events = n;

    /* sigfish.c:483-502 -- runs even for an ignored read (empty window) */
    const float cnt = (float)(hi - lo);
    float mean = 0.0f;
    for (int64_t j = lo; j < hi; j++)
        mean += ev[j].mean;
    mean /= cnt;
    float var = 0.0f;
    for (int64_t j = lo; j < hi; j++) {
        float d = ev[j].mean - mean;
        var += d * d;
    }
    var /= cnt;
    const float sd = (float)sqrt((double)var);
    for (int64_t j = lo; j < hi; j++)
        ev[j].mean = (ev[j].mean - mean) / sd;
    return n > 0;
}

/* sigfish.c:575-626.  `list` is ordered worst (index 0) to best (index KEEP-1).  A new score
 * is placed above every entry whose score is >= it, so of two equal scores the later one ranks
 * better; entry 0 falls off. */
static void keep_best(cand_t *list, float score, int32_t rid, int32_t pos, char strand,
                      const float *cost, int qlen, int rlen)
{
    int slot = 0;
    while (slot < KEEP && !(score > list[slot].score))
        slot++;
    if (slot == 0)
        return;
    for (int s = 0; s + 1 < slot; s++)
        list[s] = list[s + 1];
    cand_t *c = &list[slot - 1];
    c->score = score;
    c->rid = rid;
    c->pos_end = pos;
    c->strand = strand;
    c->pos_st = -1;
    if (pos < rlen) /* cdtw.c:107-108: path() refuses an out-of-range start column */
        c->pos_st = orc_path_start(cost, qlen, rlen, pos < 0 ? rlen - 1 : pos);
}

/* sigfish.c:891-901 / 938-948 -- cut the last row into blocks of qlen columns; each block
 * offers its first (strict <) minimum */
static void offer_blocks(cand_t *list, const float *cost, int qlen, int rlen, int32_t rid, char strand)
{
    const float *last = cost + (size_t)(qlen - 1) * rlen;
    for (int b = 0; b < rlen; b += qlen) {
        float best = INFINITY;
        int32_t at = -1;
        for (int c = b; c < b + qlen && c < rlen; c++) {
            if (last[c] < best) {
                best = last[c];
                at = c;
            }
        }
        /* with no finite cell the reference passes min_pos = -1, i.e. pos = -1 - (qlen-1)*rlen */
        int32_t pos = at >= 0 ? at : (int32_t)(-1 - (int64_t)(qlen - 1) * rlen);
        keep_best(list, best, rid, pos, strand, cost, qlen, rlen);
    }
}

/* sigfish.c:979: (int)round(...) of a double; x86-64 cvttsd2si yields INT_MIN when out of range */
static int32_t to_int_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0))
        return INT_MIN;
    return (int32_t)v;
}

/* sigfish.c:828-985 */
void orc_align(const orc_ref_t *ref, const orc_event_t *ev, int64_t qstart, int64_t qend,
               uint32_t flags, orc_hit_t *hit)
{
    const int rna = (flags & ORC_RNA) != 0;
    const int qlen = (int)(qend - qstart);

    cand_t list[KEEP];
    for (int s = 0; s < KEEP; s++) {
        list[s].score = INFINITY;
        list[s].rid = -1;
        list[s].pos_st = list[s].pos_end = -1;
        list[s].strand = 0;
    }

    float *query = (float *)malloc(sizeof(float) * (size_t)(qlen > 0 ? qlen : 1));
    for (int j = 0; j < qlen; j++) { /* 857-867: RNA reads are 3'->5', flip unless --invert */
        if (rna && !(flags & ORC_INV))
            query[qlen - 1 - j] = ev[qstart + j].mean;
        else
            query[j] = ev[qstart + j].mean;
    }

    for (int32_t r = 0; r < ref->num_ref; r++) {
        const int rlen = ref->ref_lengths[r];
        float *cost = (float *)malloc(sizeof(float) * (size_t)qlen * (size_t)rlen);
        if (!(flags & ORC_DTW)) {
            orc_subsequence(query, ref->forward[r], qlen, rlen, cost);
            offer_blocks(list, cost, qlen, rlen, r, '+');
        } else { /* 914-917: only the bottom-right cell competes */
            float s = orc_std_dtw(query, ref->forward[r], qlen, rlen, cost);
            keep_best(list, s, r, rlen - 1, '+', cost, qlen, rlen);
        }
        if (!rna) {
            orc_subsequence(query, ref->reverse[r], qlen, rlen, cost);
            offer_blocks(list, cost, qlen, rlen, r, '-');
        }
        free(cost);
    }
    free(query);

    /* 969-983 */
    const cand_t *w = &list[KEEP - 1];
    hit->score = w->score;
    hit->score2 = list[KEEP - 2].score;
    hit->rid = w->rid;
    hit->strand = w->strand;
    hit->raw_pos_st = w->pos_st;
    hit->raw_pos_end = w->pos_end;
    if (w->rid >= 0) {
        const int32_t rl = ref->ref_lengths[w->rid];
        hit->pos_st = w->strand == '+' ? w->pos_st : rl - w->pos_end;
        hit->pos_end = w->strand == '+' ? w->pos_end : rl - w->pos_st;
        hit->pos_st += ref->ref_st_offset[w->rid];
        hit->pos_end += ref->ref_st_offset[w->rid];
    } else {
        hit->pos_st = hit->pos_end = -1;
    }
    float ratio = 500 * (hit->score2 - hit->score) / hit->score;
    int32_t mq = to_int_x86(round((double)ratio));
    if (mq > 60)
        mq = 60;
    hit->mapq = (int32_t)(uint8_t)mq; /* aln_t.mapq is uint8_t (sigfish.h:153) */
}

static void map_read_core(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation,
                          float offset, float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit,
                          orc_event_t **ev_out);

void orc_map_read(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation,
                  float offset, float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit)
{
    map_read_core(ref, raw, n, digitisation, offset, range, flags, q, p, hit, NULL);
}

/* orc_map_read that also hands out the event table it worked on (normalised in place by the window step; caller
 * frees with orc_free): what a test double of the device needs to rebuild the winner's warping path */
void orc_map_read_events(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation, float offset,
                         float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit, orc_event_t **ev_out)
{
    map_read_core(ref, raw, n, digitisation, offset, range, flags, q, p, hit, ev_out);
}

int orc_map_read_sam(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation, float offset,
                     float range, uint32_t flags, int32_t q, int32_t p, const char *read_id,
                     const char *const *rnames, char *buf, size_t cap)
{
    orc_hit_t hit;
    orc_event_t *ev = NULL;
    map_read_core(ref, raw, n, digitisation, offset, range, flags, q, p, &hit, &ev);
    int len = 0;
    if (hit.mapped && hit.rid >= 0)
        len = orc_sam_line(ref, ev, flags, &hit, read_id, rnames[hit.rid], buf, cap);
    free(ev);
    return len;
}

static void map_read_core(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation,
                          float offset, float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit,
                          orc_event_t **ev_out)
{
    memset(hit, 0, sizeof(*hit));
    hit->rid = -1;
    if (n <= 0)
        return;
    orc_event_t *ev = NULL;
    int64_t nev = orc_detect_events(raw, n, digitisation, offset, range, (flags & ORC_RNA) != 0, &ev);
    hit->n_events = nev;
    if (nev <= 0) {
        free(ev);
        return;
    }
    int64_t lo, hi;
    int32_t st;
    int64_t nkeep = nev;
    int32_t fail = 0;
    if (p < 0 && !(flags & ORC_END)) { /* sigfish.c:438-447 + detect_query_start 380-422 */
        const int64_t pe = orc_polya_end_sample(raw, n, digitisation, offset, range, (flags & ORC_RNA004) != 0);
        int64_t first = -1;
        if (pe > 0) {
            int64_t e = 0;
            while (e < nev && ev[e].start < (uint64_t)pe)
                e++;
            first = e < nev ? e : -1;
        }
        if (first < 0) {
            fail = 4; /* prefix_fail: fall back to 50 events */
            first = 50;
        }
        p = (int32_t)first;
    }
    int go = orc_window_normalise(ev, &nkeep, flags, q, p, &lo, &hi, &st);
    st |= fail;
    hit->status = st;
    hit->qstart = lo;
    hit->qend = hi;
    if (go) {
        hit->mapped = 1;
        /* sigfish.c:800-807 */
        hit->start_raw = ev[lo].start;
        /* uint64 + float is evaluated in fp32 and converted back (C usual arithmetic conversions) */
        hit->end_raw = (uint64_t)((float)ev[hi - 1].start + ev[hi - 1].length);
        orc_align(ref, ev, lo, hi, flags, hit);
    }
    if (ev_out)
        *ev_out = ev;
    else
        free(ev);
}

/* sigfish.c:628-660 with the arguments aln_to_str passes (796-826) */
int orc_paf_line(char *buf, size_t cap, const orc_hit_t *hit, const char *read_id,
                 const char *rname, int32_t ref_seq_len, int64_t len_raw_signal)
{
    const uint64_t query_size = (uint64_t)(hit->qend - 1) - (uint64_t)hit->qstart;
    const float block = (float)(hit->pos_end - hit->pos_st);
    const float residue = block - hit->score * block / (float)query_size;
    return snprintf(buf, cap,
                    "%s\t%ld\t%ld\t%ld\t%c\t%s\t%d\t%d\t%d\t%d\t%d\t%d\ttp:A:P\td1:f:%.2f\td2:f:%.2f\n",
                    read_id, (long)len_raw_signal, (long)hit->start_raw, (long)hit->end_raw,
                    hit->strand, rname, ref_seq_len, hit->pos_st, hit->pos_end,
                    to_int_x86(round((double)residue)), to_int_x86(round((double)block)),
                    hit->mapq, (double)hit->score, (double)hit->score2);
}

/* test helper: dtw_single() (orc_align) on a bare array of already normalised event means */
void orc_align_means(const orc_ref_t *ref, const float *means, int32_t n, uint32_t flags, orc_hit_t *hit)
{
    orc_event_t *ev = (orc_event_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(orc_event_t));
    for (int32_t j = 0; j < n; j++)
        ev[j].mean = means[j];
    memset(hit, 0, sizeof(*hit));
    hit->rid = -1;
    hit->mapped = 1;
    hit->qstart = 0;
    hit->qend = n;
    orc_align(ref, ev, 0, n, flags, hit);
    free(ev);
}

/* ---- SAM output (reference src/sigfish.c:530-571 path_to_map, 663-768 r2qevent_map_to_ss, 770-794 sam_str) ---- */

typedef struct {
    int32_t first, last; /* query events aligned to one reference position, -1/-1: none */
} span_t;

/* sigfish.c:530-571: walk the warping path in forward order */
static span_t *spans_from_path(const int32_t *px, const int32_t *py, int32_t k, int32_t len)
{
    span_t *sp = (span_t *)malloc(sizeof(span_t) * (size_t)(len > 0 ? len : 1));
    for (int32_t i = 0; i < len; i++)
        sp[i].first = sp[i].last = -1;
    const int32_t origin = py[0];
    int32_t prev_q = -1;
    for (int32_t s = 0; s < k; s++) {
        const int32_t at = py[s] - origin, qi = px[s];
        if (sp[at].first == -1)
            sp[at].first = qi;
        sp[at].last = qi;
        if (prev_q == qi) /* same query event again (a horizontal step): this position gets nothing */
            sp[at].first = sp[at].last = -1;
        prev_q = qi;
    }
    return sp;
}

static int appendf(char *buf, size_t cap, size_t *len, const char *fmt, long v)
{
    int w = snprintf(buf + *len, *len < cap ? cap - *len : 0, fmt, v);
    if (w > 0)
        *len += (size_t)w;
    return w;
}

/* sam_str() for the winning hit of `hit` (as filled by orc_align on the same normalised events).
 * Returns the number of bytes written. */
int orc_sam_line(const orc_ref_t *ref, const orc_event_t *ev, uint32_t flags, const orc_hit_t *hit,
                 const char *read_id, const char *rname, char *buf, size_t cap)
{
    const int rna = (flags & ORC_RNA) != 0;
    const int64_t qstart = hit->qstart, qend = hit->qend;
    const int qlen = (int)(qend - qstart);
    const int rlen = ref->ref_lengths[hit->rid];
    /* rebuild the winner's cost matrix and its path */
    float *query = (float *)malloc(sizeof(float) * (size_t)qlen);
    for (int j = 0; j < qlen; j++) {
        if (rna && !(flags & ORC_INV))
            query[qlen - 1 - j] = ev[qstart + j].mean;
        else
            query[j] = ev[qstart + j].mean;
    }
    const float *y = hit->strand == '+' ? ref->forward[hit->rid] : ref->reverse[hit->rid];
    float *cost = (float *)malloc(sizeof(float) * (size_t)qlen * (size_t)rlen);
    if (flags & ORC_DTW)
        orc_std_dtw(query, y, qlen, rlen, cost);
    else
        orc_subsequence(query, y, qlen, rlen, cost);
    int32_t *px = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + rlen));
    int32_t *py = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + rlen));
    const int32_t k = orc_path_full(cost, qlen, rlen, hit->raw_pos_end, px, py);
    const int32_t n_pos = hit->raw_pos_end - hit->raw_pos_st + 1;
    span_t *sp = spans_from_path(px, py, k, n_pos);
    free(cost);
    free(query);
    free(px);
    free(py);

    /* sigfish.c:667-695: RNA queries were reversed, map indices back; then make them absolute */
    if (rna) {
        const int32_t last = sp[n_pos - 1].last;
        for (int32_t i = 0; i < n_pos; i++)
            if (sp[i].first != -1) {
                sp[i].first = last - sp[i].first;
                sp[i].last = last - sp[i].last;
            }
    }
    for (int32_t i = 0; i < n_pos; i++)
        if (sp[i].first != -1) {
            sp[i].first += (int32_t)qstart;
            sp[i].last += (int32_t)qstart;
        }
    /* sigfish.c:709-721: RNA reports positions back to front, each span swapped */
    if (rna) {
        for (int32_t a = 0; a < n_pos / 2; a++) {
            span_t t = sp[a];
            sp[a] = sp[n_pos - 1 - a];
            sp[n_pos - 1 - a] = t;
        }
        for (int32_t i = 0; i < n_pos; i++) {
            int32_t t = sp[i].first;
            sp[i].first = sp[i].last;
            sp[i].last = t;
        }
    }

    size_t len = 0;
    /* sigfish.c:770-794 */
    int w = snprintf(buf, cap, "%s\t%d\t%s\t%ld\t%d\t%ldM\t*\t0\t0\t*\t*\tsi:Z:%ld,%ld,%ld,%ld\tss:Z:", read_id,
                     hit->strand == '+' ? 0 : 16, rname, (long)hit->pos_st + 1, hit->mapq,
                     (long)((qend - 1) - qstart), (long)hit->start_raw, (long)hit->end_raw,
                     (long)(rna ? hit->pos_end : hit->pos_st), (long)(rna ? hit->pos_st : hit->pos_end));
    if (w > 0)
        len = (size_t)w;
    /* sigfish.c:723-762: run-length string over the raw signal: "<n>," match, "<n>D" skipped reference
     * positions, "<n>I" skipped samples */
    int64_t cursor = 0, gap = 0;
    int first_seen = 0;
    for (int32_t j = 0; j < n_pos; j++) {
        if (sp[j].first == -1) {
            if (first_seen)
                gap++;
            continue;
        }
        const int64_t s0 = (int64_t)ev[sp[j].first].start;
        const int64_t s1 = (int64_t)ev[sp[j].last].start + (int)ev[sp[j].last].length;
        first_seen = 1;
        if (gap > 0) {
            appendf(buf, cap, &len, "%ldD", (long)gap);
            gap = 0;
        }
        if (j == 0)
            cursor = s0;
        int64_t mi = s0 - cursor;
        cursor += mi;
        if (mi)
            appendf(buf, cap, &len, "%ldI", (long)(int)mi);
        mi = s1 - s0;
        cursor += mi;
        if (mi)
            appendf(buf, cap, &len, "%ld,", (long)(int)mi);
    }
    if (len + 1 < cap) {
        buf[len++] = '\n';
        buf[len] = 0;
    }
    free(sp);
    return (int)len;
}
