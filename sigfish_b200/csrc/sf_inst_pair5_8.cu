// sf_dtw_pair_kernel<5..8, false, RQ>: 64 < q <= 128
#define SF_PAIR_INST_IMPL
#include "sf_pair_inst.cuh"
bool sf_pair_run_r5_8(int r2, int rq, bool std_dtw, sf_pair_op &op)
{
    if (std_dtw)
        return false;
    return sf_pair_rows<5, false>(r2, rq, op) || sf_pair_rows<6, false>(r2, rq, op) || sf_pair_rows<7, false>(r2, rq, op) ||
           sf_pair_rows<8, false>(r2, rq, op);
}
