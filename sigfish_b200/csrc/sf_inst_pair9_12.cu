// sf_dtw_pair_kernel<9..12, false, RQ>: 128 < q <= 192
#define SF_PAIR_INST_IMPL
#include "sf_pair_inst.cuh"
bool sf_pair_run_r9_12(int r2, int rq, bool std_dtw, sf_pair_op &op)
{
    if (std_dtw)
        return false;
    return sf_pair_rows<9, false>(r2, rq, op) || sf_pair_rows<10, false>(r2, rq, op) || sf_pair_rows<11, false>(r2, rq, op) ||
           sf_pair_rows<12, false>(r2, rq, op);
}
