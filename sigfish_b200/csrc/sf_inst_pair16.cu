// sf_dtw_pair_kernel<16, STD, RQ>: 240 < q <= 256, and every 128 < q <= 256 of --dtw-std
#define SF_PAIR_INST_IMPL
#include "sf_pair_inst.cuh"
bool sf_pair_run_r16(int r2, int rq, bool std_dtw, sf_pair_op &op)
{
    return std_dtw ? sf_pair_rows<16, true>(r2, rq, op) : sf_pair_rows<16, false>(r2, rq, op);
}
