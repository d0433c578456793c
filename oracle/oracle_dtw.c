/* oracle_dtw.c -- the DTW recurrences and the warp-path walk restated on the CPU
 * (test oracle; see oracle.h).  Follows reference src/cdtw.c (mlpy):
 *   min3 25-36, std_dtw 69-94, path 98-167, subsequence 171-189,
 *   subsequence_path 192-227.
 * Full cost matrix, row-major with stride m, exactly like the reference, so the
 * walk below can test the same fp32 equalities.
 */
#include <math.h>
#include <stdlib.h>

#include "oracle.h"

/* cdtw.c:25-36 -- first argument wins ties; NaN never replaces the running value */
static inline float least3(float a, float b, float c)
{
    float v = a;
    if (b < v)
        v = b;
    if (c < v)
        v = c;
    return v;
}

static inline float gap(float a, float b) { return (float)fabs((double)(a - b)); }

/* cdtw.c:171-189 -- free start along the reference: row 0 carries no accumulated cost */
void orc_subsequence(const float *x, const float *y, int n, int m, float *cost)
{
    for (int j = 0; j < m; j++)
        cost[j] = gap(x[0], y[j]);
    for (int i = 1; i < n; i++) {
        float *row = cost + (size_t)i * m;
        const float *above = row - m;
        row[0] = gap(x[i], y[0]) + above[0];
        for (int j = 1; j < m; j++)
            row[j] = gap(x[i], y[j]) + least3(above[j], above[j - 1], row[j - 1]);
    }
}

/* cdtw.c:69-94 with squared == 0 -- global alignment: row 0 accumulates too */
float orc_std_dtw(const float *x, const float *y, int n, int m, float *cost)
{
    cost[0] = gap(x[0], y[0]);
    for (int j = 1; j < m; j++)
        cost[j] = gap(x[0], y[j]) + cost[j - 1];
    for (int i = 1; i < n; i++) {
        float *row = cost + (size_t)i * m;
        const float *above = row - m;
        row[0] = gap(x[i], y[0]) + above[0];
        for (int j = 1; j < m; j++)
            row[j] = gap(x[i], y[j]) + least3(above[j], above[j - 1], row[j - 1]);
    }
    return cost[(size_t)n * m - 1];
}

/* One backward step of cdtw.c:129-147: diagonal if it equals the minimum, else left, else up. */
static inline void step_back(const float *cost, int m, int *i, int *j)
{
    if (*i == 0) {
        (*j)--;
    } else if (*j == 0) {
        (*i)--;
    } else {
        const float up = cost[(size_t)(*i - 1) * m + *j];
        const float dg = cost[(size_t)(*i - 1) * m + (*j - 1)];
        const float lf = cost[(size_t)(*i) * m + (*j - 1)];
        const float best = least3(up, dg, lf);
        if (dg == best) {
            (*i)--;
            (*j)--;
        } else if (lf == best) {
            (*j)--;
        } else {
            (*i)--;
        }
    }
}

/* cdtw.c:98-167 walks from (n-1, end_col) all the way to (0,0); cdtw.c:192-227 then drops
 * the leading run of row-0 cells except the last one.  The first retained column is thus the
 * column at which the backward walk first reaches row 0 (0 if it reaches column 0 first). */
int32_t orc_path_start(const float *cost, int n, int m, int end_col)
{
    int i = n - 1, j = end_col;
    while (i > 0)
        step_back(cost, m, &i, &j);
    return j;
}

int32_t orc_path_full(const float *cost, int n, int m, int end_col, int32_t *px, int32_t *py)
{
    int i = n - 1, j = end_col;
    int32_t k = 0;
    px[k] = i;
    py[k] = j;
    k++;
    while (i > 0) {
        step_back(cost, m, &i, &j);
        px[k] = i;
        py[k] = j;
        k++;
    }
    for (int32_t a = 0, b = k - 1; a < b; a++, b--) {
        int32_t t = px[a]; px[a] = px[b]; px[b] = t;
        t = py[a]; py[a] = py[b]; py[b] = t;
    }
    return k;
}
