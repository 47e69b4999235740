// rc_pat36.cu — the <RC_PAT_MAXL> instantiation of the pattern kernel (rc_pat_kernel.cuh) and its launcher.
#include "rc_pat_kernel.cuh"

cudaError_t rc_launch_pat36(dim3 grid, size_t smem, cudaStream_t st, const RcPlanDev& pl, int e, int binBase, int ringRows,
                            const RcModelDev* models, const RcModelEpochDev* mep, const RcSEvDev* sevs, const double* ktab,
                            const double* aShort, const double* wpool, const double* bPrev, double* bCur, long long bStride,
                            double* dAcc, double* pOut) {
    rc_propagate_pat_kernel<RC_PAT_MAXL, 1><<<grid, RC_PBLOCK, smem, st>>>(pl, e, binBase, ringRows, models, mep, sevs, ktab, aShort, wpool, bPrev, bCur,
                                                               bStride, dAcc, pOut);
    return cudaGetLastError();
}
