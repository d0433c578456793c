/* oracle_jnn.c -- automatic query start (-p < 0) restated on the CPU (test oracle; see oracle.h).
 *
 * Follows reference src/jnn.c:21-60 (rolling mean), 62-79 (outlier clamp), 100-177 (jnnv2 adaptor
 * finder), 191-279 (jnn_core), 354-376 (find_polya), src/jnn.h:30-99 (parameters), src/stat.h:17-44
 * (meanf / stdvf) and src/sigfish.c:380-422 (detect_query_start).
 * All sums are fp32 and sequential, as in the reference.
 */
#include <math.h>
#include <stdlib.h>

#include "oracle.h"

/* stat.h:17-24 */
static float mean_f32(const float *x, int n)
{
    float acc = 0;
    for (int i = 0; i < n; i++)
        acc += x[i];
    return acc / n;
}

/* stat.h:36-44 */
static float stdv_f32(const float *x, int n)
{
    const float m = mean_f32(x, n);
    float acc = 0;
    for (int i = 0; i < n; i++)
        acc += (x[i] - m) * (x[i] - m);
    return sqrtf(acc / n);
}

static float clamp_adc(float v) /* jnn.c:17-18, 62-96: values outside [0, 1200] are clipped */
{
    if (v > 1200)
        return 1200;
    if (v < 0)
        return 0;
    return v;
}

/* jnn.c:100-177: first dip of the windowed mean below mean - scale*stdv that lasts lo..hi samples.
 * Returns 0 and (a,b) on success; a = b = 0 when nothing qualifies, -1 when the read is too short. */
static void adaptor_span(const int16_t *raw, int64_t n, int rna004, int64_t *a_out, int64_t *b_out)
{
    const float scale = rna004 ? 0.7f : 0.5f;
    const int merge_gap = 1500, win = 2000, hi_len = 200000, lo_len = rna004 ? 500 : 2000;
    if (!(n > win)) {
        *a_out = -1;
        *b_out = -1;
        return;
    }
    const int m = (int)n - win;
    float *roll = (float *)malloc(sizeof(float) * (size_t)m);
    float run = 0.0f;
    for (int i = 0; i < win; i++)
        run += clamp_adc((float)raw[i]);
    roll[0] = run / win;
    for (int i = 1; i < m; i++) { /* jnn.c:43-48: slide by subtracting the leaving and adding the entering sample */
        run -= clamp_adc((float)raw[i - 1]);
        run += clamp_adc((float)raw[i + win - 1]);
        roll[i] = run / win;
    }
    const float mu = mean_f32(roll, m);
    const float sd = stdv_f32(roll, m);
    const float floor_v = mu - sd * scale;

    int64_t cap = 1000, cnt = 0;
    int64_t *sx = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap), *sy = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    int inside = 0, st = 0, en = 0;
    for (int j = 0; j < m; j++) {
        const float v = roll[j];
        if (v < floor_v && !inside) {
            st = j;
            inside = 1;
        } else if (v < floor_v) {
            en = j;
        } else if (v > floor_v && inside) {
            if (cnt && st - sy[cnt - 1] < merge_gap) {
                sy[cnt - 1] = en;
            } else {
                if (cnt >= cap) {
                    cap *= 2;
                    sx = (int64_t *)realloc(sx, sizeof(int64_t) * (size_t)cap);
                    sy = (int64_t *)realloc(sy, sizeof(int64_t) * (size_t)cap);
                }
                sx[cnt] = st;
                sy[cnt] = en;
                cnt++;
            }
            st = 0;
            en = 0;
            inside = 0;
        }
    }
    *a_out = 0;
    *b_out = 0;
    for (int64_t s = 0; s < cnt; s++) {
        const int64_t len = sy[s] - sx[s];
        if (len > hi_len || len < lo_len)
            continue;
        *a_out = sx[s] + win / 2 - 1;
        *b_out = sy[s] + win / 2 - 1;
        break;
    }
    free(roll);
    free(sx);
    free(sy);
}

/* jnn.c:191-279 with the poly-A parameters of jnn.h:56-77 (std_scale -1: fixed band), reduced to the
 * first segment, which is all find_polya() returns (jnn.c:354-376).  Returns its end or -1. */
static int64_t first_plateau_end(const float *pa, int64_t n, float top, float bot)
{
    const int win = 250, max_err = 30, merge_gap = 200;
    const float stall = 1.0f;
    int corr = 50;
    int on = 0, err = 0, run_err = 0, c = 0;
    int64_t st = 0;
    /* the reference collects all segments and merges neighbours closer than merge_gap; the first
     * segment's end can still grow by merging, so the scan cannot stop at the first closure */
    int64_t first_x = -1, first_y = -1;
    int n_seg = 0;
    int64_t last_y = 0;
    for (int64_t i = 0; i < n; i++) {
        const float a = clamp_adc(pa[i]);
        if (a < top && a > bot) {
            if (!on) {
                st = i;
                on = 1;
            }
            c++;
            corr++;
            if (run_err)
                run_err = 0;
            if (c >= win && c >= corr && !(c % corr))
                err--;
        } else {
            if (on && err < max_err) {
                c++;
                err++;
                run_err++;
                if (c >= win && c >= corr && !(c % corr))
                    err--;
            } else if (on && (c >= win || (!n_seg && c >= win * stall))) {
                const int64_t en = i - run_err;
                on = 0;
                if (n_seg && st - last_y < merge_gap) {
                    last_y = en;
                    if (n_seg == 1)
                        first_y = en;
                } else {
                    if (n_seg == 0) {
                        first_x = st;
                        first_y = en;
                    }
                    last_y = en;
                    n_seg++;
                }
                c = 0;
                err = 0;
                run_err = 0;
            } else if (on) {
                on = 0;
                c = 0;
                err = 0;
                run_err = 0;
            }
        }
    }
    (void)first_x;
    return n_seg > 0 ? first_y : -1;
}

/* sigfish.c:380-422 up to the point where the event table is consulted: the raw-sample index at which
 * the poly-A tail ends (events starting at or after it form the query), or -1 */
int64_t orc_polya_end_sample(const int16_t *raw, int64_t n, float digitisation, float offset, float range,
                             int rna004)
{
    int64_t ax, ay;
    adaptor_span(raw, n, rna004, &ax, &ay);
    if (!(ay > 0))
        return -1;
    float *pa = (float *)malloc(sizeof(float) * (size_t)n);
    orc_to_picoamps(raw, n, digitisation, offset, range, pa);
    const float level = mean_f32(pa + ax, (int)(ay - ax));
    const int64_t e = first_plateau_end(pa + ay, n - ay, level + 30 + 20, level + 30 - 20);
    free(pa);
    if (!(e > 0))
        return -1;
    return e + ay;
}
