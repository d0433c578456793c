// sf_dtw_pair_kernel<13..15, false, RQ>: 192 < q <= 240
#define SF_PAIR_INST_IMPL
#include "sf_pair_inst.cuh"
bool sf_pair_run_r13_15(int r2, int rq, bool std_dtw, sf_pair_op &op)
{
    if (std_dtw)
        return false;
    return sf_pair_rows<13, false>(r2, rq, op) || sf_pair_rows<14, false>(r2, rq, op) || sf_pair_rows<15, false>(r2, rq, op);
}
