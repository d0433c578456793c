/* oracle_events.c -- event detection restated on the CPU (test oracle; see oracle.h).
 *
 * Follows reference src/events.c:297-577 (scrappie segmentation as vendored by
 * sigfish) and the pA conversion of src/sigfish.c:334-347.  The MAD trimming of
 * events.c:99-269 is not restated: getevents() discards its result (events.c:567),
 * so it cannot influence the output.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* sigfish.c:344-347 -- all fp32: one division, then (raw + offset) * unit per sample */
void orc_to_picoamps(const int16_t *raw, int64_t n, float digitisation, float offset,
                     float range, float *pa)
{
    const float unit = range / digitisation;
    for (int64_t s = 0; s < n; s++) {
        float shifted = (float)raw[s] + offset;
        pa[s] = shifted * unit;
    }
}

/* events.c:297-307 -- running fp64 sums; the square is an fp32 product (then widened) */
void orc_prefix_sums(const float *x, int64_t n, double *sum, double *sumsq)
{
    double acc = 0.0, acc2 = 0.0;
    sum[0] = 0.0;
    sumsq[0] = 0.0;
    for (int64_t s = 0; s < n; s++) {
        float sq = x[s] * x[s];
        acc = acc + (double)x[s];
        acc2 = acc2 + (double)sq;
        sum[s + 1] = acc;
        sumsq[s + 1] = acc2;
    }
}

/* events.c:319-368 -- two-window t-statistic.  The mixed fp32/fp64 evaluation is
 * spelled out step by step. */
void orc_tstat(const double *sum, const double *sumsq, int64_t n, int64_t w, float *t)
{
    memset(t, 0, sizeof(float) * (size_t)n);
    if (n < 2 * w || w < 2) /* events.c:332-334 */
        return;
    const float wf = (float)w;
    for (int64_t i = w; i <= n - w; i++) {
        /* left window [i-w, i): kept in fp64; no subtraction on the very first one */
        double lsum = sum[i], lsq = sumsq[i];
        if (i > w) {
            lsum -= sum[i - w];
            lsq -= sumsq[i - w];
        }
        /* right window [i, i+w): narrowed to fp32 immediately */
        float rsum = (float)(sum[i + w] - sum[i]);
        float rsq = (float)(sumsq[i + w] - sumsq[i]);

        float lmean = (float)(lsum / (double)wf);
        float rmean = rsum / wf;

        float lmean2 = lmean * lmean; /* fp32 products */
        float rmean2 = rmean * rmean;
        float rq = rsq / wf;          /* fp32 quotient */
        double acc = lsq / (double)wf;
        acc = acc - (double)lmean2;
        acc = acc + (double)rq;
        acc = acc - (double)rmean2;
        float var = (float)acc;
        var = fmaxf(var, FLT_MIN);

        float dm = rmean - lmean;
        float vq = var / wf;
        t[i] = (float)(fabs((double)dm) / sqrt((double)vq));
    }
}

/* one of the two coupled peak finders of events.c:273-285 */
typedef struct {
    const float *sig;
    float threshold;
    uint64_t window;
    uint64_t masked_to;
    int peak_pos; /* -1: no maximum recorded yet */
    float peak_val;
    int valid;
} finder_t;

static void finder_reset(finder_t *f)
{
    f->peak_pos = -1;
    f->peak_val = FLT_MAX;
    f->valid = 0;
}

/* events.c:375-447.  Parameters from events.c:47-58. */
int64_t orc_peaks(const float *t1, const float *t2, int64_t n, int rna, uint64_t *peaks)
{
    finder_t f[2];
    f[0].sig = t1;
    f[1].sig = t2;
    f[0].threshold = rna ? 2.5f : 1.4f;
    f[1].threshold = 9.0f;
    f[0].window = rna ? 7 : 3;
    f[1].window = rna ? 14 : 6;
    const float height = rna ? 1.0f : 0.2f;
    for (int d = 0; d < 2; d++) {
        f[d].masked_to = 0;
        finder_reset(&f[d]);
    }

    int64_t count = 0;
    for (uint64_t i = 0; i < (uint64_t)n; i++) {
        for (int d = 0; d < 2; d++) { /* short finder first, then long: order matters */
            finder_t *me = &f[d];
            if (me->masked_to >= i)
                continue;
            const float v = me->sig[i];
            if (me->peak_pos < 0) {
                /* hunting: track the running minimum until the signal climbs `height` above it */
                if (v < me->peak_val) {
                    me->peak_val = v;
                } else if (v - me->peak_val > height) {
                    me->peak_val = v;
                    me->peak_pos = (int)i;
                }
                continue;
            }
            /* inside a candidate peak */
            if (v > me->peak_val) {
                me->peak_val = v;
                me->peak_pos = (int)i;
            }
            if (d == 0 && me->peak_val > me->threshold) {
                /* a short-window peak that will fire silences the long finder */
                f[1].masked_to = (uint64_t)me->peak_pos + me->window;
                finder_reset(&f[1]);
            }
            if (me->peak_val - v > height && me->peak_val > me->threshold)
                me->valid = 1;
            if (me->valid && (i - (uint64_t)me->peak_pos) > me->window / 2) {
                peaks[count++] = (uint64_t)me->peak_pos;
                me->peak_pos = -1;
                me->peak_val = v;
                me->valid = 0;
            }
        }
    }
    return count;
}

/* events.c:461-477 */
static orc_event_t make_event(uint64_t lo, uint64_t hi, const double *sum, const double *sumsq)
{
    orc_event_t e;
    e.start = lo;
    e.length = (float)(hi - lo); /* unsigned wrap if the peak list is ever non-monotone */
    float dsum = (float)(sum[hi] - sum[lo]);
    e.mean = dsum / e.length;
    float dsq = (float)(sumsq[hi] - sumsq[lo]);
    float var = dsq / e.length - e.mean * e.mean;
    e.stdv = sqrtf(fmaxf(var, 0.0f));
    return e;
}

/* events.c:510-554 + 479-508 */
int64_t orc_detect_events(const int16_t *raw, int64_t n, float digitisation, float offset,
                          float range, int rna, orc_event_t **out)
{
    *out = NULL;
    if (n <= 0)
        return 0;
    float *pa = (float *)malloc(sizeof(float) * (size_t)n);
    double *sum = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *sumsq = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    float *t1 = (float *)malloc(sizeof(float) * (size_t)n);
    float *t2 = (float *)malloc(sizeof(float) * (size_t)n);
    uint64_t *peaks = (uint64_t *)calloc((size_t)n, sizeof(uint64_t));

    orc_to_picoamps(raw, n, digitisation, offset, range, pa);
    orc_prefix_sums(pa, n, sum, sumsq);
    orc_tstat(sum, sumsq, n, rna ? 7 : 3, t1);
    orc_tstat(sum, sumsq, n, rna ? 14 : 6, t2);
    int64_t npk = orc_peaks(t1, t2, n, rna, peaks);

    int64_t nev = -1;
    if (npk >= 1) { /* with no peak the reference reads peaks[-1] (events.c:504) */
        nev = npk + 1; /* every emitted position lies in [1, n-1] => all are counted (484-489) */
        orc_event_t *ev = (orc_event_t *)calloc((size_t)nev, sizeof(orc_event_t));
        for (int64_t e = 0; e < nev; e++) {
            uint64_t lo = e == 0 ? 0 : peaks[e - 1];
            uint64_t hi = e == nev - 1 ? (uint64_t)n : peaks[e];
            ev[e] = make_event(lo, hi, sum, sumsq);
        }
        *out = ev;
    }
    free(peaks);
    free(t2);
    free(t1);
    free(sumsq);
    free(sum);
    free(pa);
    return nev;
}

void orc_free(void *p) { free(p); }
