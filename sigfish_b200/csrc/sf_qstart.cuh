// sf_qstart.cuh -- automatic query start (-p < 0, direct RNA): where does the poly-A tail end?
//
// Replaces (reference, paths relative to /root/reference):
//   src/jnn.c:21-60     rolling_window        (windowed mean, slid by subtract / add in fp32)
//   src/jnn.c:62-96     rm_outlier(f)         (clip to [0, 1200])
//   src/jnn.c:100-177   jnnv2 / find_adaptor  (first 2000..200000-sample dip of the windowed mean)
//   src/jnn.c:191-279   jnn_core, 354-376 find_polya (first in-band plateau after the adaptor)
//   src/stat.h:17-44    meanf / stdvf
//   src/sigfish.c:380-405 detect_query_start up to polya.y (the event lookup is done by the event kernel)
// Every quantity is an fp32 running sum in sample order, so the whole chain is sequential: one thread
// per read, all reads of the batch in flight at once (the rolling mean is recomputed in each of its
// three passes instead of being stored).
#pragma once
#include <cuda_runtime.h>
#include "sf_types.cuh"

struct sf_qs_args {
    const int16_t *signal;
    const int64_t *sig_off;
    const int64_t *sig_len;
    const float *digitisation, *offset, *range;
    int32_t n_reads;
    int32_t rna004;        // pore_flag == OPT_PORE_RNA004: jnn.h:91-97 instead of 83-89
    int64_t *polya_end;    // [n_reads] raw-sample index, -1: not found
};

__device__ __forceinline__ float sf_clip_adc(float v) { return v > 1200.0f ? 1200.0f : (v < 0.0f ? 0.0f : v); }

__global__ void sf_qstart_kernel(const sf_qs_args a)
{
    const int read = blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= a.n_reads)
        return;
    const long long n = a.sig_len[read];
    const int16_t *raw = a.signal + a.sig_off[read];
    long long result = -1;
    const int win = 2000;
    if (n > win) {
        const int m = (int)n - win;
        const float wf = (float)win;
        // ---- pass 1: mean of the windowed means (stat.h:17-24 over jnn.c:38-48) ----
        float run = 0.0f;
        for (int i = 0; i < win; i++)
            run = __fadd_rn(run, sf_clip_adc((float)raw[i]));
        const float run0 = run;
        float acc = __fdiv_rn(run, wf);
        for (int i = 1; i < m; i++) {
            run = __fsub_rn(run, sf_clip_adc((float)raw[i - 1]));
            run = __fadd_rn(run, sf_clip_adc((float)raw[i + win - 1]));
            acc = __fadd_rn(acc, __fdiv_rn(run, wf));
        }
        const float mu = __fdiv_rn(acc, (float)m);
        // ---- pass 2: population stdv (stat.h:36-44) ----
        run = run0;
        float d = __fsub_rn(__fdiv_rn(run, wf), mu);
        acc = __fmul_rn(d, d);
        for (int i = 1; i < m; i++) {
            run = __fsub_rn(run, sf_clip_adc((float)raw[i - 1]));
            run = __fadd_rn(run, sf_clip_adc((float)raw[i + win - 1]));
            d = __fsub_rn(__fdiv_rn(run, wf), mu);
            acc = __fadd_rn(acc, __fmul_rn(d, d));
        }
        const float sd = __fsqrt_rn(__fdiv_rn(acc, (float)m));
        const float scale = a.rna004 ? 0.7f : 0.5f;
        const float floor_v = __fsub_rn(mu, __fmul_rn(sd, scale));
        const int lo_len = a.rna004 ? 500 : 2000, hi_len = 200000, merge_gap = 1500;
        // ---- pass 3: dips below the floor; neighbours closer than merge_gap are one segment; the first
        //      segment of admissible length wins (jnn.c:126-163).  Only the open last segment is kept. ----
        int inside = 0, st = 0, en = 0;
        int have = 0, sx = 0, sy = 0;       // last segment
        long long ax = 0, ay = 0;
        int found = 0;
        run = run0;
        for (int j = 0; j < m && !found; j++) {
            if (j > 0) {
                run = __fsub_rn(run, sf_clip_adc((float)raw[j - 1]));
                run = __fadd_rn(run, sf_clip_adc((float)raw[j + win - 1]));
            }
            const float v = __fdiv_rn(run, wf);
            if (v < floor_v && !inside) {
                st = j;
                inside = 1;
            } else if (v < floor_v) {
                en = j;
            } else if (v > floor_v && inside) {
                if (have && st - sy < merge_gap) {
                    sy = en;
                } else {
                    if (have) { // the previous segment is final now
                        const int len = sy - sx;
                        if (!(len > hi_len) && !(len < lo_len)) {
                            ax = sx + win / 2 - 1;
                            ay = sy + win / 2 - 1;
                            found = 1;
                        }
                    }
                    sx = st;
                    sy = en;
                    have = 1;
                }
                st = 0;
                en = 0;
                inside = 0;
            }
        }
        if (!found && have) {
            const int len = sy - sx;
            if (!(len > hi_len) && !(len < lo_len)) {
                ax = sx + win / 2 - 1;
                ay = sy + win / 2 - 1;
                found = 1;
            }
        }
        if (found && ay > 0) {
            // ---- adaptor level in pA (sigfish.c:388), then the poly-A plateau (jnn.c:191-279 with the
            //      fixed band of jnn.h:56-77: window 250, corrector 50, error 30, merge distance 200) ----
            const float unit = __fdiv_rn(a.range[read], a.digitisation[read]);
            const float offs = a.offset[read];
            float lv = 0.0f;
            for (long long i = ax; i < ay; i++)
                lv = __fadd_rn(lv, __fmul_rn(__fadd_rn((float)raw[i], offs), unit));
            lv = __fdiv_rn(lv, (float)(int)(ay - ax));
            const float top = __fadd_rn(__fadd_rn(lv, 30.0f), 20.0f);
            const float bot = __fsub_rn(__fadd_rn(lv, 30.0f), 20.0f);
            const int pwin = 250, max_err = 30, gap = 200;
            int corr = 50, on = 0, err = 0, run_err = 0, c = 0, n_seg = 0;
            long long pst = 0, last_y = 0, first_y = -1;
            const long long rem = n - ay;
            for (long long i = 0; i < rem && n_seg < 2; i++) {
                const float v = sf_clip_adc(__fmul_rn(__fadd_rn((float)raw[ay + i], offs), unit));
                if (v < top && v > bot) {
                    if (!on) {
                        pst = i;
                        on = 1;
                    }
                    c++;
                    corr++;
                    run_err = 0;
                    if (c >= pwin && c >= corr && !(c % corr))
                        err--;
                } else if (on && err < max_err) {
                    c++;
                    err++;
                    run_err++;
                    if (c >= pwin && c >= corr && !(c % corr))
                        err--;
                } else if (on && (c >= pwin || (!n_seg && (float)c >= (float)pwin * 1.0f))) {
                    const long long pen = i - run_err;
                    on = 0;
                    if (n_seg && pst - last_y < gap) {
                        last_y = pen;
                        if (n_seg == 1)
                            first_y = pen;
                    } else {
                        if (n_seg == 0)
                            first_y = pen;
                        last_y = pen;
                        n_seg++; // a second, unmerged segment freezes the first one: the scan can stop
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
            if (n_seg > 0 && first_y > 0)
                result = first_y + ay;
        }
    }
    a.polya_end[read] = result;
}
