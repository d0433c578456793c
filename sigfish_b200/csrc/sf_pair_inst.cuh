// sf_pair_inst.cuh -- the instantiations of sf_dtw_pair_kernel, spread over several translation units so that they
// compile in parallel (one kernel per query size 64 < q <= 256: R2 = ceil(q / 16) rows per lane, RQ = (q - 1) % R2 the
// register of the last query row; with the redo variants ~300 kernels).  sfgpu.cu reaches them through sf_pair_run().
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include "sf_dtw.cuh"

enum { SF_PAIR_OCCUPANCY = 0, SF_PAIR_LAUNCH = 1, SF_PAIR_LAUNCH_FIX = 2 };

struct sf_pair_op {
    int what;               // SF_PAIR_*
    const sf_dtw_args *args;
    int grid;
    size_t smem;
    cudaStream_t stream;
    int blocks_per_sm;      // out: SF_PAIR_OCCUPANCY
    cudaError_t err;        // out
};

// each returns false when (r2, rq, std_dtw) is not instantiated in its translation unit
bool sf_pair_run_r16(int r2, int rq, bool std_dtw, sf_pair_op &op);     // R2 = 16, both recurrences
bool sf_pair_run_r5_8(int r2, int rq, bool std_dtw, sf_pair_op &op);    // R2 = 5 .. 8, subsequence DTW
bool sf_pair_run_r9_12(int r2, int rq, bool std_dtw, sf_pair_op &op);   // R2 = 9 .. 12, subsequence DTW
bool sf_pair_run_r13_15(int r2, int rq, bool std_dtw, sf_pair_op &op);  // R2 = 13 .. 15, subsequence DTW

inline bool sf_pair_run(int r2, int rq, bool std_dtw, sf_pair_op &op)
{
    op.err = cudaErrorInvalidValue;
    op.blocks_per_sm = 0;
    return sf_pair_run_r16(r2, rq, std_dtw, op) || sf_pair_run_r5_8(r2, rq, std_dtw, op) || sf_pair_run_r9_12(r2, rq, std_dtw, op) ||
           sf_pair_run_r13_15(r2, rq, std_dtw, op);
}
// is the pair layout built for this query size and recurrence?
inline bool sf_pair_exists(int r2, bool std_dtw) { return r2 == 16 || (!std_dtw && r2 >= 5 && r2 <= 15); }

#ifdef SF_PAIR_INST_IMPL
template <int R2, int RQ, bool STD> void sf_pair_do(sf_pair_op &op)
{
    switch (op.what) {
    case SF_PAIR_OCCUPANCY:
        op.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&op.blocks_per_sm, sf_dtw_pair_kernel<R2, STD, RQ, false>, SF_DTW_THREADS, op.smem);
        break;
    case SF_PAIR_LAUNCH: {
        // launched behind sf_dtw_score_kernel in the same stream, allowed to start once that kernel's blocks are resident
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)op.grid);
        cfg.blockDim = dim3(SF_DTW_THREADS);
        cfg.dynamicSmemBytes = op.smem;
        cfg.stream = op.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        op.err = cudaLaunchKernelEx(&cfg, sf_dtw_pair_kernel<R2, STD, RQ, false>, *op.args);
        break;
    }
    case SF_PAIR_LAUNCH_FIX: // redo pass of the pieces whose warm front did not verify; subsequence DTW only
        if constexpr (!STD) {
            sf_dtw_pair_kernel<R2, false, RQ, true><<<op.grid, SF_DTW_THREADS, op.smem, op.stream>>>(*op.args);
            op.err = cudaGetLastError();
        }
        break;
    }
}

template <int R2, bool STD, int... RQs> bool sf_pair_rq(int rq, sf_pair_op &op, std::integer_sequence<int, RQs...>)
{
    return ((rq == RQs && (sf_pair_do<R2, RQs, STD>(op), true)) || ...);
}
// every register the last query row can sit in: RQ = 0 .. R2 - 1
template <int R2, bool STD> bool sf_pair_rows(int r2, int rq, sf_pair_op &op)
{
    return r2 == R2 && sf_pair_rq<R2, STD>(rq, op, std::make_integer_sequence<int, R2>{});
}
#endif
