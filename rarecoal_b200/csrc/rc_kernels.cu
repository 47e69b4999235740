// rc_kernels.cu — sm_100a FP64 kernels for the rarecoal likelihood hot path.
//
//   rc_tables_kernel     per (model, grid step): coalescence tables shared by all patterns
//   rc_propagate_kernel  persistent time loop: updateA / updateB / updateD / Join / Split / shortcut
//                        (Core.hs:72-288) for a bin of patterns of one model per block
//   rc_reduce_kernel     getProb epilogue reduction: sum cnt*log p, sum p per model (MaxUtils.hs:85-91)
//   rc_finalize_kernel   siteRed * (ll + other * log(1 - sp))
//   rc_fp64_peak_kernel  DFMA-rate microbenchmark (roofline denominator)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rc_kernels.h"

// ------------------------------------------------------------------------------------------------
// Per-step tables.  Formulas keep the reference's operation order (Core.hs:240,270,279).
__global__ void rc_tables_kernel(const RcModelDev* __restrict__ models, const RcSegDev* __restrict__ segs,
                                 const int* __restrict__ segOff, const double* __restrict__ dts,
                                 double* __restrict__ tab, int P, int A1) {
    const int b = blockIdx.y;
    const int s = blockIdx.x;
    const RcModelDev& M = models[b];
    if (s >= M.rowsAvail) return;
    // last segment with sBegin <= s
    int lo = segOff[b], hi = segOff[b + 1] - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (segs[mid].sBegin <= s) lo = mid; else hi = mid - 1;
    }
    const RcSegDev& sg = segs[lo];
    const double dt = dts[s];
    const int rowLen = rc_row_len(P, A1);
    double* row = tab + M.tabOff + (long long)s * rowLen;
    const int nE = P * A1;
    for (int i = threadIdx.x; i < rowLen; i += blockDim.x) {
        double v;
        if (i < 2 * nE) {
            int ii = i < nE ? i : i - nE;
            int k = ii / A1, j = ii - k * A1;
            bool frozen = (sg.frozenMask >> k) & 1u;
            double xx = (double)j, pp = sg.N[k];
            if (frozen) v = 0.0;
            else if (i < nE) v = 1.0 - exp(-(xx * (xx + 1.0) / 2.0 * (1.0 / pp) * dt));
            else v = xx * (xx - 1.0) / 2.0 * (1.0 / pp);
        } else {
            int r = i - 2 * nE;
            if (r < P) v = exp(-(0.5 * dt / sg.N[r]));
            else if (r < 2 * P) v = ((sg.frozenMask >> (r - P)) & 1u) ? 0.0 : 1.0 / sg.N[r - P];
            else if (r < 3 * P) v = sg.N[r - 2 * P];
            else if (r == 3 * P) v = dt;
            else v = (double)sg.frozenMask;
        }
        row[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// chooseCont n k (Utils.hs:213-215)
__device__ __forceinline__ double rc_choose_cont(double n, int k) {
    double p = 1.0;
    for (int j = 1; j <= k; ++j) p = p * ((n + 1.0 - (double)j) / (double)j);
    return p;
}

__device__ __forceinline__ int rc_find_owner(const int* off, int np, int slot) {
    int lo = 0, hi = np - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= slot) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct RcSmem {
    double *bA, *bB, *dacc, *a, *r, *dpat, *tab;
    uint32_t *words, *meta;
    int *patOff, *patOff2, *patId, *patDone;
    unsigned short* owner;
};

__device__ __forceinline__ RcSmem rc_carve(unsigned char* base, int cap, int nnzw, int P, int rowLen) {
    RcSmem S;
    double* d = reinterpret_cast<double*>(base);
    S.bA = d; d += cap;
    S.bB = d; d += cap;
    S.dacc = d; d += cap;
    S.a = d; d += RC_NPB * P;
    S.r = d; d += RC_NPB * P;
    S.dpat = d; d += RC_NPB;
    S.tab = d; d += rowLen;
    uint32_t* u = reinterpret_cast<uint32_t*>(d);
    S.words = u; u += (size_t)cap * nnzw;
    S.meta = u; u += cap;
    int* ip = reinterpret_cast<int*>(u);
    S.patOff = ip; ip += RC_NPB + 1;
    S.patOff2 = ip; ip += RC_NPB + 1;
    S.patId = ip; ip += RC_NPB;
    S.patDone = ip; ip += RC_NPB;
    S.owner = reinterpret_cast<unsigned short*>(ip);
    return S;
}

size_t rc_propagate_smem_bytes(int cap, int nnzw, int P, int rowLen) {
    size_t n = 0;
    n += (size_t)(3 * cap + 2 * RC_NPB * P + RC_NPB + rowLen) * sizeof(double);
    n += ((size_t)cap * nnzw + cap) * sizeof(uint32_t);
    n += (size_t)(2 * (RC_NPB + 1) + 2 * RC_NPB) * sizeof(int);
    n += (size_t)cap * sizeof(unsigned short);
    return (n + 15) & ~(size_t)15;
}

// One block = one model x one bin of patterns.  Slots (live states) of all patterns of the bin are
// packed side by side in shared memory; every thread advances its slots through the whole time grid.
__global__ void __launch_bounds__(RC_BLOCK)
rc_propagate_kernel(RcPlanDev pl, const RcModelDev* __restrict__ models, const RcSEvDev* __restrict__ sevs,
                    const double* __restrict__ tab, const double* __restrict__ wpool,
                    double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int bin = blockIdx.x, b = blockIdx.y;
    const RcModelDev& M = models[b];
    if (M.nSteps < 0) return;   // model failed validation on the host
    const int P = pl.P, A1 = pl.A1, nnzw = pl.nnzw, nPat = pl.nPat;
    const int rowLen = rc_row_len(P, A1);
    RcSmem S = rc_carve(smemRaw, pl.cap, nnzw, P, rowLen);
    const double* E2 = S.tab;
    const double* CN = S.tab + P * A1;
    const double* EA = S.tab + 2 * P * A1;
    const double* invN = EA + P;
    const int dtIdx = 2 * P * A1 + 3 * P;

    const int pat0 = pl.binPatOff[bin];
    const int np = pl.binPatOff[bin + 1] - pat0;

    // ---- init (makeInitCoalState, Core.hs:58-70) ----
    if (tid < np) {
        S.patId[tid] = pl.binPats[pat0 + tid];
        S.patDone[tid] = 0;
        S.dpat[tid] = 0.0;
    }
    __syncthreads();
    for (int i = tid; i < np * P; i += nthr) {
        int p = i / P, k = i - p * P;
        S.a[i] = (double)pl.aInit[(size_t)S.patId[p] * P + k];
        S.r[i] = 0.0;
    }
    if (tid == 0) {
        int run = 0;
        for (int p = 0; p < np; ++p) {
            S.patOff[p] = run;
            run += pl.slotOff[S.patId[p] + 1] - pl.slotOff[S.patId[p]];
        }
        S.patOff[np] = run;
    }
    __syncthreads();
    int nSlots = S.patOff[np];
    double* bc = S.bA;
    double* bn = S.bB;
    for (int slot = tid; slot < nSlots; slot += nthr) {
        int p = rc_find_owner(S.patOff, np, slot);
        int local = slot - S.patOff[p];
        int pid = S.patId[p];
        long long g = pl.epochBase[0] + pl.slotOff[pid] + local;
        for (int d = 0; d < nnzw; ++d) S.words[slot * nnzw + d] = pl.words[g * nnzw + d];
        S.meta[slot] = pl.meta[g];
        S.owner[slot] = (unsigned short)p;
        S.dacc[slot] = 0.0;
        // the seed state m is the last one fillUp emits (all components at their maximum)
        bc[slot] = (slot + 1 == S.patOff[p + 1]) ? 1.0 : 0.0;
    }
    __syncthreads();

    int epoch = 0, nextEv = 0;
    const RcSEvDev* sev = sevs + M.sevOff;
    const int nSteps = M.nSteps;

    for (int s = 0; s < nSteps; ++s) {
        // ---- shortcut (canUseShortcut / propagateStateShortcut, Core.hs:80-118) ----
        if (s == M.sShort) {
            int mine = 1;
            if (tid < np) {
                int pid = S.patId[tid];
                if (pl.shortcutOK[pid]) {
                    const double* aP = S.a + tid * P;
                    double nrA = 0.0;
                    int popIndex = -1;
                    for (int k = 0; k < P; ++k) {
                        nrA = nrA + aP[k];
                        if (popIndex < 0 && aP[k] > 0.0) popIndex = k;
                    }
                    const double* rowN = tab + M.tabOff + (long long)s * rowLen + 2 * P * A1 + 2 * P;
                    double popSize = rowN[popIndex < 0 ? 0 : popIndex];
                    int base = S.patOff[tid], n = S.patOff[tid + 1] - base;
                    double acc = S.dpat[tid], add = 0.0;
                    for (int i = 0; i < n; ++i) {
                        uint32_t mt = S.meta[base + i];
                        if (mt & RC_META_UNIT) acc = acc + S.dacc[base + i];
                        int nd = (int)(mt >> 16);
                        double withComb = 2.0 * popSize / (double)nd;
                        add = add + bc[base + i] * (withComb / rc_choose_cont(nrA + (double)nd, nd));
                    }
                    S.dpat[tid] = acc + add;
                    S.patDone[tid] = 2;
                } else mine = 0;
            }
            if (__syncthreads_and(mine)) break;
        }
        // ---- structural events firing before this step's update (singleStep, Core.hs:120-135) ----
        while (nextEv < M.nSEv && sev[nextEv].step == s) {
            const RcSEvDev ev = sev[nextEv];
            if (tid < np) {
                int base = S.patOff[tid], n = S.patOff[tid + 1] - base;
                double acc = S.dpat[tid];
                for (int i = 0; i < n; ++i)
                    if (S.meta[base + i] & RC_META_UNIT) acc = acc + S.dacc[base + i];
                S.dpat[tid] = acc;
                double* aP = S.a + tid * P;
                double al = aP[ev.l], ak = aP[ev.k];
                if (ev.type == RC_EV_JOIN_D) {          // popJoinA, Core.hs:165-170
                    aP[ev.l] = 0.0;
                    aP[ev.k] = al + ak;
                } else {                                // popSplitA, Core.hs:195-200
                    aP[ev.l] = (1.0 - ev.m) * al;
                    aP[ev.k] = ev.m * al + ak;
                }
            }
            const int* so2 = pl.slotOff + (size_t)(epoch + 1) * (nPat + 1);
            if (tid == 0) {
                int run = 0;
                for (int p = 0; p < np; ++p) {
                    S.patOff2[p] = run;
                    run += so2[S.patId[p] + 1] - so2[S.patId[p]];
                }
                S.patOff2[np] = run;
            }
            __syncthreads();
            const int nNew = S.patOff2[np];
            const long long eb = pl.epochBase[epoch + 1];
            const int* trOff = pl.trOff + pl.trBase[epoch];
            const uint32_t* trEnt = pl.trEnt + pl.trEntBase[epoch];
            const double* W = wpool + M.wPool + ev.wOff;
            // popJoinB / popSplitB as a gather (Core.hs:172-221)
            for (int slot = tid; slot < nNew; slot += nthr) {
                int p = rc_find_owner(S.patOff2, np, slot);
                int local = slot - S.patOff2[p];
                int gl = so2[S.patId[p]] + local;
                int e0 = trOff[gl], e1 = trOff[gl + 1];
                const int oldBase = S.patOff[p];
                double acc = 0.0;
                for (int q = e0; q < e1; ++q) {
                    uint32_t en = trEnt[q];
                    uint32_t widx = en >> 16;
                    double old = bc[oldBase + (int)(en & 0xFFFFu)];
                    acc = (widx == RC_W_ONE) ? acc + old : acc + old * W[widx];
                }
                bn[slot] = acc;
                long long g = eb + gl;
                for (int d = 0; d < nnzw; ++d) S.words[slot * nnzw + d] = pl.words[g * nnzw + d];
                S.meta[slot] = pl.meta[g];
                S.owner[slot] = (unsigned short)p;
                S.dacc[slot] = 0.0;
            }
            __syncthreads();
            if (tid <= np) S.patOff[tid] = S.patOff2[tid];
            double* t = bc; bc = bn; bn = t;
            nSlots = nNew;
            ++epoch;
            ++nextEv;
            __syncthreads();
        }
        // ---- tables of this step ----
        const double* row = tab + M.tabOff + (long long)s * rowLen;
        for (int i = tid; i < rowLen; i += nthr) S.tab[i] = row[i];
        __syncthreads();
        const double dt = S.tab[dtIdx];
        if (dt > 0.0) {
            // ---- updateA (Core.hs:229-240) ----
            const unsigned fmask = (unsigned)S.tab[dtIdx + 1];
            for (int i = tid; i < np * P; i += nthr) {
                int p = i / P, k = i - p * P;
                if (S.patDone[p]) continue;
                double a = S.a[i];
                if (!((fmask >> k) & 1u) && a >= 1.0) {
                    a = 1.0 / (1.0 + ((1.0 / a) - 1.0) * EA[k]);
                    S.a[i] = a;
                }
                S.r[i] = a * invN[k];
            }
            __syncthreads();
            // ---- updateB + updateD (Core.hs:242-288) ----
            for (int slot = tid; slot < nSlots; slot += nthr) {
                const int p = S.owner[slot];
                if (S.patDone[p]) continue;
                const uint32_t mt = S.meta[slot];
                const int nnz = (int)(mt & 0xFFu);
                const int base = S.patOff[p];
                const double* rP = S.r + p * P;
                double t1 = 0.0, t2 = 0.0;
                for (int d = 0; d < nnz; ++d) {
                    const uint32_t w = S.words[slot * nnzw + d];
                    const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                    const uint32_t up = w >> 16;
                    t1 = t1 + (CN[k * A1 + x] + (double)x * rP[k]);
                    if (up != RC_NO_UP) t2 = t2 + bc[base + (int)up] * E2[k * A1 + x];
                }
                const double bnew = bc[slot] * exp(-(t1 * dt)) + t2;
                bn[slot] = bnew;
                if (mt & RC_META_UNIT) S.dacc[slot] += dt * bnew;
            }
            __syncthreads();
            double* t = bc; bc = bn; bn = t;
        }
    }
    __syncthreads();
    // ---- getProb epilogue (Core.hs:49-54) ----
    if (tid < np) {
        int pid = S.patId[tid];
        double d = S.dpat[tid];
        if (S.patDone[tid] != 2) {
            int base = S.patOff[tid], n = S.patOff[tid + 1] - base;
            for (int i = 0; i < n; ++i)
                if (S.meta[base + i] & RC_META_UNIT) d = d + S.dacc[base + i];
        }
        double discFac = 1.0;
        for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? M.disc[k] : 1.0);
        pOut[(size_t)b * nPat + pid] = d * M.theta * pl.combFac[pid] * discFac;
    }
}

// ------------------------------------------------------------------------------------------------
// One block per model; fixed summation order (thread-strided partials, then a shared-memory tree),
// so results are reproducible run to run.
__global__ void __launch_bounds__(1024)
rc_reduce_kernel(const double* __restrict__ pOut, const long long* __restrict__ counts,
                 const int* __restrict__ patGlobal, const RcModelDev* __restrict__ models, int nPat,
                 int nPatGlobal, double* __restrict__ partial, double* __restrict__ dProbs) {
    __shared__ double sl[1024], sp[1024];
    const int b = blockIdx.x, tid = threadIdx.x;   // b = position in the plan-sorted model array
    const bool bad = models[b].nSteps < 0;
    const int orig = models[b].pad;                // caller's model index
    double ll = 0.0, ps = 0.0;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = tid; i < nPat; i += blockDim.x) {
        double p = bad ? nanv : pOut[(size_t)b * nPat + i];
        ll += log(p) * (double)counts[i];
        ps += p;
        if (dProbs) dProbs[(size_t)orig * nPatGlobal + patGlobal[i]] = p;
    }
    sl[tid] = ll; sp[tid] = ps;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if (tid < w) { sl[tid] += sl[tid + w]; sp[tid] += sp[tid + w]; }
        __syncthreads();
    }
    if (tid == 0) { partial[2 * orig] = sl[0]; partial[2 * orig + 1] = sp[0]; }
}

__global__ void rc_finalize_kernel(const double* __restrict__ partial, int B, double other, double siteRed,
                                   double* __restrict__ logl) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) logl[b] = siteRed * (partial[2 * b] + other * log(1.0 - partial[2 * b + 1]));
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rc_fp64_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
cudaError_t rc_launch_tables(const RcModelDev* models, const RcSegDev* segs, const int* segOff, const double* dts,
                             double* tab, int P, int A1, int rowsMax, int B, cudaStream_t st) {
    if (rowsMax <= 0 || B <= 0) return cudaSuccess;
    dim3 grid((unsigned)rowsMax, (unsigned)B);
    rc_tables_kernel<<<grid, 128, 0, st>>>(models, segs, segOff, dts, tab, P, A1);
    return cudaGetLastError();
}

cudaError_t rc_launch_propagate(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs,
                                const double* tab, const double* wpool, double* pOut, int b0, int nB,
                                cudaStream_t st) {
    if (pl.nBins <= 0 || nB <= 0) return cudaSuccess;
    size_t smem = rc_propagate_smem_bytes(pl.cap, pl.nnzw, pl.P, rc_row_len(pl.P, pl.A1));
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(rc_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid((unsigned)pl.nBins, (unsigned)nB);
    rc_propagate_kernel<<<grid, RC_BLOCK, smem, st>>>(pl, models + b0, sevs, tab, wpool, pOut + (size_t)b0 * pl.nPat);
    return cudaGetLastError();
}

cudaError_t rc_launch_reduce(const double* pOut, const long long* counts, const int* patGlobal,
                             const RcModelDev* models, int nPat, int nPatGlobal, int B, double* partial,
                             double* dProbs, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    rc_reduce_kernel<<<B, 1024, 0, st>>>(pOut, counts, patGlobal, models, nPat, nPatGlobal, partial, dProbs);
    return cudaGetLastError();
}

cudaError_t rc_launch_finalize(const double* partial, int B, double other, double siteRed, double* logl,
                               cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    rc_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(partial, B, other, siteRed, logl);
    return cudaGetLastError();
}

cudaError_t rc_launch_fp64_peak(double* out, int blocks, int iters, cudaStream_t st) {
    rc_fp64_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 1.0);
    return cudaGetLastError();
}
