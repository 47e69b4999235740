// rc_kernels.h — launch wrappers of the sm_100a kernels (rc_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "rc_types.h"

size_t rc_propagate_smem_bytes(int cap, int nnzw, int P, int rowLen);

cudaError_t rc_launch_tables(const RcModelDev* models, const RcSizeStepDev* sizes, const int* sizeOff, const RcMaskStepDev* masks,
                             const int* maskOff, const double* dts, double* tab, int P, int A1, int rowsMax, int B, cudaStream_t st);
cudaError_t rc_launch_propagate(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs,
                                const double* tab, const double* wpool, double* pOut, int b0, int nB,
                                unsigned char* scratch, size_t scratchBytes, cudaStream_t st);
size_t rc_big_scratch_bytes(int cap, int nnzw);
cudaError_t rc_launch_propagate_fast(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs,
                                     const double* tab, const double* wpool, double* pOut, int b0, int nB,
                                     cudaStream_t st);
size_t rc_fast_smem_bytes(int block);
cudaError_t rc_launch_cum(int nEpochs, int P, int A1, const RcModelDev* models, const RcModelEpochDev* mep, const double* tab,
                          double* cum, int b0, int nB, cudaStream_t st);
cudaError_t rc_launch_traj(const RcPlanDev& pl, int e, int nThr, const RcModelDev* models, const RcModelEpochDev* mep,
                           const RcSEvDev* sevs, const double* tab, const double* cum, const double* dts, double* ktab,
                           double* aEnd, double* aShort, int b0, int nB, cudaStream_t st);
// the three instantiations of the pattern kernel live in translation units of their own (rc_pat12.cu, rc_pat25.cu, rc_pat36.cu)
#define RC_PAT_LAUNCH_ARGS dim3 grid, size_t smem, cudaStream_t st, const RcPlanDev& pl, int e, int binBase, int ringRows, \
    const RcModelDev* models, const RcModelEpochDev* mep, const RcSEvDev* sevs, const double* ktab, const double* aShort, \
    const double* wpool, const double* bPrev, double* bCur, long long bStride, double* dAcc, double* pOut
cudaError_t rc_launch_pat12(RC_PAT_LAUNCH_ARGS);
cudaError_t rc_launch_pat25(RC_PAT_LAUNCH_ARGS);
cudaError_t rc_launch_pat36(RC_PAT_LAUNCH_ARGS);
cudaError_t rc_launch_propagate_keyed(const RcPlanDev& pl, int e, int binBase, int nBins, int rowMode, int maxFetch, int maxB, int zSlices, const RcModelDev* models,
                                      const RcModelEpochDev* mep, const RcSEvDev* sevs, const double* ktab,
                                      const double* aShort, const double* wpool, const double* bPrev, double* bCur,
                                      long long bStride, double* dAcc, double* pOut, int b0, int nB, cudaStream_t st);
cudaError_t rc_launch_reduce(const double* pOut, const long long* counts, const int* patGlobal,
                             const RcModelDev* models, int nPat, int nPatGlobal, int B, double* partial,
                             double* dProbs, void* scratch, cudaStream_t st);
size_t rc_reduce_scratch_bytes(int B);
cudaError_t rc_launch_finalize(const double* partial, int B, double other, double siteRed, double* logl,
                               cudaStream_t st);
cudaError_t rc_launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t st);
