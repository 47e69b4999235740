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
#include <stddef.h>
#include <stdint.h>

#include "rc_kernels.h"
#include "rc_dev.cuh"
#include <cstdlib>

#ifndef RC_RED_SLICES
#define RC_RED_SLICES 32       // blocks per model in rc_reduce_kernel
#endif
#define RC_RED_BLOCK 256
#ifndef RC_TRAJ_FILL
#define RC_TRAJ_FILL 151552     // threads a trajectory launch aims for: 148 SMs x 8 blocks of 128
#endif
#ifndef RC_TRAJ_ZMAX
#define RC_TRAJ_ZMAX 16
#endif

// ------------------------------------------------------------------------------------------------
// Per-step tables.  Formulas keep the reference's operation order (Core.hs:240,270,279).
__global__ void rc_tables_kernel(const RcModelDev* __restrict__ models, const RcSizeStepDev* __restrict__ sizes,
                                 const int* __restrict__ sizeOff, const RcMaskStepDev* __restrict__ masks,
                                 const int* __restrict__ maskOff, const double* __restrict__ dts,
                                 double* __restrict__ tab, int P, int A1) {
    const int b = blockIdx.y;
    const int s = blockIdx.x;
    const RcModelDev& M = models[b];
    if (s >= M.rowsAvail) return;
    __shared__ double Ns[RC_MAX_P];
    __shared__ unsigned maskS;
    // last entry with sBegin <= s: thread k for branch k's sizes, thread P for the frozen mask
    if ((int)threadIdx.x <= P) {
        const bool isMask = (int)threadIdx.x == P;
        int lo = isMask ? maskOff[b] : sizeOff[b * P + threadIdx.x];
        int hi = (isMask ? maskOff[b + 1] : sizeOff[b * P + threadIdx.x + 1]) - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((isMask ? masks[mid].sBegin : sizes[mid].sBegin) <= s) lo = mid; else hi = mid - 1;
        }
        if (isMask) maskS = masks[lo].frozenMask;
        else Ns[threadIdx.x] = sizes[lo].N;
    }
    __syncthreads();
    const unsigned frozenMask = maskS;
    const double dt = dts[s];
    const int rowLen = rc_row_len(P, A1);
    double* row = tab + M.tabOff + (long long)s * rowLen;
    const int nE = P * A1;
    for (int i = threadIdx.x; i < rowLen; i += blockDim.x) {
        double v = 0.0;
        if (i == 0) v = dt;
        else if (i == 1) v = (double)frozenMask;
        else if (i < RC_ROW_PAIRS + 2 * nE) {
            const int q = i - RC_ROW_PAIRS;
            const int pi = q >> 1, k = pi / A1, j = pi - k * A1;
            const bool frozen = (frozenMask >> k) & 1u;
            const double xx = (double)j, pp = Ns[k];
            if (j == 0) {
                if (q & 1) v = frozen ? 0.0 : 1.0 / pp;
                else v = (frozen || !(dt > 0.0)) ? 1.0 : exp(-(0.5 * dt / pp));
            } else if (frozen) v = 0.0;
            else if (q & 1) v = 1.0 - exp(-(xx * (xx + 1.0) / 2.0 * (1.0 / pp) * dt));
            else v = xx * (xx - 1.0) / 2.0 * (1.0 / pp);
        } else if (i < RC_ROW_PAIRS + 2 * nE + P) v = Ns[i - RC_ROW_PAIRS - 2 * nE];
        row[i] = v;
    }
}

struct RcSmem {
    double *bA, *bB, *dacc, *a, *r, *dpat, *tab;
    uint32_t *words, *meta;
    int *patOff, *patOff2, *patId, *patDone;
    unsigned short* owner;
};

// big: per-slot arrays of a block whose live sets do not fit shared memory (global scratch, rc_big_scratch_bytes per block);
// null: everything in shared memory
__device__ __forceinline__ RcSmem rc_carve(unsigned char* base, int cap, int nnzw, int P, int rowLen, unsigned char* big) {
    RcSmem S;
    double* d = reinterpret_cast<double*>(base);
    if (big) {
        double* g = reinterpret_cast<double*>(big);
        S.bA = g; g += cap;
        S.bB = g; g += cap;
        S.dacc = g; g += cap;
        uint32_t* gu = reinterpret_cast<uint32_t*>(g);
        S.words = gu; gu += (size_t)cap * nnzw;
        S.meta = gu; gu += cap;
        S.owner = reinterpret_cast<unsigned short*>(gu);
        cap = 0;
    } else {
        S.bA = d; d += cap;
        S.bB = d; d += cap;
        S.dacc = d; d += cap;
    }
    S.a = d; d += RC_NPB * P;
    S.r = d; d += RC_NPB * P;
    S.dpat = d; d += RC_NPB;
    S.tab = d; d += rowLen;
    uint32_t* u = reinterpret_cast<uint32_t*>(d);
    if (!big) { S.words = u; u += (size_t)cap * nnzw; S.meta = u; u += cap; }
    int* ip = reinterpret_cast<int*>(u);
    S.patOff = ip; ip += RC_NPB + 1;
    S.patOff2 = ip; ip += RC_NPB + 1;
    S.patId = ip; ip += RC_NPB;
    S.patDone = ip; ip += RC_NPB;
    if (!big) S.owner = reinterpret_cast<unsigned short*>(ip);
    return S;
}

size_t rc_big_scratch_bytes(int cap, int nnzw) {
    size_t n = (size_t)3 * cap * sizeof(double) + ((size_t)cap * nnzw + cap) * sizeof(uint32_t) + (size_t)cap * sizeof(unsigned short);
    return (n + 255) & ~(size_t)255;
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
                    double* __restrict__ pOut, unsigned char* scratch, long long scratchPerBlock) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int bin = blockIdx.x, b = blockIdx.y;
    const RcModelDev& M = models[b];
    if (M.nSteps < 0) return;   // model failed validation on the host
    if (M.mode == 1 && bin < pl.nBinsKB) return;   // these patterns run in row mode on the trajectory-key tier for this model
    const int P = pl.P, A1 = pl.A1, nnzw = pl.nnzw, nPat = pl.nPat, cap = pl.cap;
    const int rowLen = rc_row_len(P, A1);
    // live sets beyond the shared-memory tier (pl.bigGlobal): b, the structure words and the accumulators of the block live in
    // its slice of a global scratch buffer (same code, L1 / L2 instead of shared memory; the reference's dense vectors take any
    // size, Core.hs:58-70)
    RcSmem S = rc_carve(smemRaw, pl.cap, nnzw, P, rowLen,
                        scratch ? scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)scratchPerBlock : nullptr);
    const double* PR = S.tab + RC_ROW_PAIRS;   // pairs: (k,0) = {EA, invN}, (k,j) = {CN, E2}
    const int dtIdx = 0;

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
        for (int d = 0; d < nnzw; ++d) S.words[d * cap + slot] = pl.words[g * nnzw + d];
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
                    const double* rowN = tab + M.tabOff + (long long)s * rowLen + rc_row_n(P, A1);
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
                for (int d = 0; d < nnzw; ++d) S.words[d * cap + slot] = pl.words[g * nnzw + d];
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
            for (int i = tid; i < np * P; i += nthr) {
                int p = i / P, k = i - p * P;
                if (S.patDone[p]) continue;
                double a = S.a[i];
                const double ea = PR[2 * k * A1], invN = PR[2 * k * A1 + 1];
                if (ea != 1.0 && a >= 1.0) {         // EA == 1 encodes a frozen branch
                    a = 1.0 / (1.0 + ((1.0 / a) - 1.0) * ea);
                    S.a[i] = a;
                }
                S.r[i] = a * invN;
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
                    const uint32_t w = S.words[d * cap + slot];
                    const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                    const uint32_t up = w >> 16;
                    t1 = t1 + (PR[2 * (k * A1 + x)] + (double)x * rP[k]);
                    if (up != RC_NO_UP) t2 = t2 + bc[base + (int)up] * PR[2 * (k * A1 + x) + 1];
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
// Fast tier: one live state per thread, structure in registers, ONE barrier per grid step.
//
//   * every thread owns one slot (live state x of one pattern): b(x), its decay factor and, per non-zero
//     component k, two ABSOLUTE shared-memory addresses resolved once per structural epoch from the plan: the
//     16-byte pair {rate row, 1-exp table entry} and the neighbour b(x+e_k) (an always-zero slot if that state
//     is not live);
//   * thread (p,k) additionally owns the ancestral lineage count a_k of pattern p (updateA, Core.hs:229-240),
//     carried as u = 1/a - 1 so that propagateA is one multiply and one division, and publishes ONE GRID STEP
//     AHEAD of the states the pairs
//         RE[p][k][j] = { j(j-1)/2 (1/N_k) + j a_k (1/N_k)   (k-th term of t1, Core.hs:270),
//                         1 - exp(-j(j+1)/2 dt/N_k)          (factor of b(x+e_k) in t2, Core.hs:279) }
//     so a state issues one LDS.128, one add, one LDS.64 and one FMA per non-zero component;
//   * exp(-t1 dt) of consecutive steps differs by exp(-delta) with |delta| ~ 1e-4: the factor is carried
//     multiplicatively with a 4-term series (truncation < 1e-17) and re-anchored with a true exp() after
//     every event, every 128 steps and whenever |delta| >= 2^-10;
//   * the time loop is unrolled by two so that all buffer parities are immediates; table rows are
//     prefetched two steps ahead into a two-slot ring.
// BLOCK threads per block (64, 128 or 256): a block holds whole patterns, so BLOCK bounds the live set of a
// pattern in this tier; small blocks keep the per-step barrier between few warps.
template <int BLOCK>
struct RcFastSmem {                        // every member at a compile-time offset: accesses are base + immediate
    static constexpr int BSTRIDE = BLOCK + 8;      // doubles per b buffer; entry BLOCK is the always-zero neighbour
    static constexpr int RCAP = 2 * BLOCK;         // pair rows per parity
    double bbuf[2 * BSTRIDE];
    double RE[2 * 2 * RCAP];               // [parity][row]{rate, e2}; row RCAP-1 = {0,0} padding, RCAP-2 = {dt,0}
    double tabS[2 * RC_ROWCAP];            // two-slot ring of table rows
    double dslot[BLOCK];
    double aS[BLOCK];
    double dpat[RC_NPB];
    uint32_t metaS[BLOCK];
    int patOff[RC_NPB + 1];
    int patOff2[RC_NPB + 1];
    int patR[RC_NPB + 1];
    int patId[RC_NPB];
    int patDone[RC_NPB];
};

size_t rc_fast_smem_bytes(int block) {
    size_t n = block == 64 ? sizeof(RcFastSmem<64>) : block == 128 ? sizeof(RcFastSmem<128>) : sizeof(RcFastSmem<256>);
    return (n + 15) & ~(size_t)15;
}

// Per-thread state of the fast kernel (all in registers; every function below is force-inlined).
template <int NNZ>
struct RcFastRegs {
    unsigned aRE[NNZ], aB[NNZ];   // absolute shared addresses at parity 0
    int nw;                       // components to visit (warp maximum)
    bool unit, slotOn, anchor;
    double bv, D, zprev, dacc;
    unsigned myB;                 // absolute address of this thread's b slot at parity 0
    // A-task (thread tid <-> pattern tid / P of the bin, branch tid % P)
    bool taskOn, aProp;
    double av, uv;                // a (when !aProp) or u = 1/a - 1 (when aProp, i.e. a >= 1)
    unsigned tTab;                // absolute address of pair (tk, 0) = {EA, invN} in ring slot 0
    unsigned tRE;                 // absolute address of this task's pairs at parity 0
    int taskM;                    // number of pairs (largest x_k over the pattern's live states)
};

template <int BLOCK>
struct RcFastK {
    typedef RcFastSmem<BLOCK> SM;
    static constexpr unsigned B_PAR = SM::BSTRIDE * 8u;        // byte distance between the two b buffers
    static constexpr unsigned RE_PAR = SM::RCAP * 16u;         // ... the two pair buffers
    static constexpr unsigned T_PAR = RC_ROWCAP * 8u;          // ... the two ring slots
    static constexpr unsigned RE_DT = (SM::RCAP - 2) * 16u;    // pair holding {dt, 0}
    static constexpr unsigned RE_ZERO = (SM::RCAP - 1) * 16u;  // pair holding {0, 0}
    static constexpr unsigned O_B = (unsigned)offsetof(SM, bbuf);
    static constexpr unsigned O_RE = (unsigned)offsetof(SM, RE);
    static constexpr unsigned O_T = (unsigned)offsetof(SM, tabS);
    static_assert(O_RE % 16 == 0 && O_T % 16 == 0, "16-byte alignment of pair tables");
};

// updateA for one grid step from ring slot TPH, pairs written at parity RPH.
template <int NNZ, int BLOCK, int TPH, int RPH>
__device__ __forceinline__ void rc_fast_taskA(RcFastRegs<NNZ>& R) {
    typedef RcFastK<BLOCK> K;
    if (R.taskOn) {
        double ea, invN;
        rc_lds128(R.tTab + TPH * K::T_PAR, ea, invN);
        double a = R.av;
        if (R.aProp) {
            R.uv = R.uv * ea;                               // EA == 1 when frozen or dt <= 0
            a = 1.0 / (1.0 + R.uv);                         // propagateA, Core.hs:239-240
        }
        const double r = a * invN;
        unsigned src = R.tTab + TPH * K::T_PAR + 16u, dst = R.tRE + RPH * K::RE_PAR;
        double jj = 1.0;
        for (int j = 0; j < R.taskM; ++j, jj += 1.0, src += 16u, dst += 16u) {
            double cn, e2;
            rc_lds128(src, cn, e2);
            rc_sts128(dst, fma(jj, r, cn), e2);
        }
    }
}

template <int NNZ, int BLOCK, int N, int PH>
__device__ __forceinline__ void rc_fast_dims(const RcFastRegs<NNZ>& R, double& t1, double& t2) {
    typedef RcFastK<BLOCK> K;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        double rate, e2;
        rc_lds128(R.aRE[d] + PH * K::RE_PAR, rate, e2);
        t1 += rate;
        t2 = fma(rc_lds64(R.aB[d] + PH * K::B_PAR), e2, t2);
    }
}

// One grid step with compile-time buffer parity PH (b, RE and the table ring all alternate).
template <int NNZ, int BLOCK, int PH>
__device__ __forceinline__ void rc_fast_step(RcFastRegs<NNZ>& R, const unsigned sBase, const double* __restrict__ pf,
                                             const int rowLen, const bool havePre, const bool doA, const int s) {
    typedef RcFastK<BLOCK> K;
    constexpr int QH = PH ^ 1;
    constexpr int NPRE = (RC_ROWCAP + BLOCK - 1) / BLOCK;
    // table row of step s+2: issued now, stored at the end (its ring slot held row s, which is dead)
    double pre[NPRE];
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
        pre[i] = 0.0;
        if (havePre && (int)threadIdx.x + i * BLOCK < rowLen) pre[i] = __ldg(pf + i * BLOCK);
    }
    const double dt = rc_lds64(sBase + K::O_RE + PH * K::RE_PAR + K::RE_DT);
    // ---- updateB + updateD for this thread's state (Core.hs:242-288) ----
    if (R.slotOn) {
        if (dt > 0.0) {
            double t1 = 0.0, t2 = 0.0;
            switch (R.nw) {
                case 1: rc_fast_dims<NNZ, BLOCK, 1, PH>(R, t1, t2); break;
                case 2: rc_fast_dims<NNZ, BLOCK, 2, PH>(R, t1, t2); break;
                case 3: rc_fast_dims<NNZ, BLOCK, (NNZ >= 3 ? 3 : NNZ), PH>(R, t1, t2); break;
                case 4: rc_fast_dims<NNZ, BLOCK, (NNZ >= 4 ? 4 : NNZ), PH>(R, t1, t2); break;
                case 5: rc_fast_dims<NNZ, BLOCK, (NNZ >= 5 ? 5 : NNZ), PH>(R, t1, t2); break;
                case 6: rc_fast_dims<NNZ, BLOCK, (NNZ >= 6 ? 6 : NNZ), PH>(R, t1, t2); break;
                case 7: rc_fast_dims<NNZ, BLOCK, (NNZ >= 7 ? 7 : NNZ), PH>(R, t1, t2); break;
                case 8: rc_fast_dims<NNZ, BLOCK, (NNZ >= 8 ? 8 : NNZ), PH>(R, t1, t2); break;
                default: break;
            }
            const double z = t1 * dt;
            const double del = z - R.zprev;
            R.zprev = z;
            if (!R.anchor && fabs(del) < 0.0009765625 && (s & 127)) {
                // exp(-del) - 1 = del * (-1 + del/2 - del^2/6 + del^3/24)
                double q = fma(del, 1.0 / 24.0, -1.0 / 6.0);
                q = fma(q, del, 0.5);
                q = fma(q, del, -1.0);
                R.D = fma(R.D, del * q, R.D);
            } else {
                R.D = exp(-z);
                R.anchor = false;
            }
            R.bv = fma(R.bv, R.D, t2);
            if (R.unit) R.dacc = fma(dt, R.bv, R.dacc);
        }
        rc_sts64(R.myB + QH * K::B_PAR, R.bv);
    }
    // ---- updateA one step ahead and the pairs of step s+1 (from ring slot QH, to parity QH) ----
    if (doA) {
        if (threadIdx.x == 0)
            rc_sts128(sBase + K::O_RE + QH * K::RE_PAR + K::RE_DT, rc_lds64(sBase + K::O_T + QH * K::T_PAR), 0.0);
        rc_fast_taskA<NNZ, BLOCK, QH, QH>(R);
    }
#pragma unroll
    for (int i = 0; i < NPRE; ++i)
        if (havePre && (int)threadIdx.x + i * BLOCK < rowLen)
            rc_sts64(sBase + K::O_T + PH * K::T_PAR + (threadIdx.x + i * BLOCK) * 8u, pre[i]);
    __syncthreads();
}

template <int NNZ, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK)
rc_propagate_fast_kernel(RcPlanDev pl, const RcModelDev* __restrict__ models, const RcSEvDev* __restrict__ sevs,
                         const double* __restrict__ tab, const double* __restrict__ wpool,
                         double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x;
    const int bin = blockIdx.x, b = blockIdx.y;
    if (models[b].nSteps < 0 || models[b].mode != 0) return;   // failed validation / handled by the keyed kernel
    const int P = pl.P, A1 = pl.A1;
    const int rowLen = rc_row_len(P, A1);
    typedef RcFastSmem<BLOCK> SM;
    typedef RcFastK<BLOCK> K;
    SM& S = *reinterpret_cast<SM*>(smemRaw);
    const unsigned sBase = rc_smem_base(smemRaw);
    const int pat0 = pl.fBinPatOff[bin];
    const int np = pl.fBinPatOff[bin + 1] - pat0;
    const int rowsAvail = models[b].rowsAvail, nSteps = models[b].nSteps, sShort = models[b].sShort;
    const int nSEv = models[b].nSEv;
    const double* rows = tab + models[b].tabOff;
    const RcSEvDev* sev = sevs + models[b].sevOff;

    const int tp = tid / P, tk = tid - tp * P;
    const bool isTask = tp < np;

    RcFastRegs<NNZ> R;
    const unsigned zeroRE = sBase + K::O_RE + K::RE_ZERO;
    const unsigned zeroB = sBase + K::O_B + BLOCK * 8u;
#pragma unroll
    for (int d = 0; d < NNZ; ++d) { R.aRE[d] = zeroRE; R.aB[d] = zeroB; }
    R.nw = 0; R.unit = false; R.slotOn = false; R.anchor = true;
    R.bv = 0.0; R.D = 1.0; R.zprev = 0.0; R.dacc = 0.0;
    R.myB = sBase + K::O_B + (unsigned)tid * 8u;
    R.taskOn = false; R.aProp = false; R.av = 0.0; R.uv = 0.0;
    R.tTab = sBase + K::O_T + (unsigned)(RC_ROW_PAIRS + 2 * tk * A1) * 8u;
    R.tRE = zeroRE; R.taskM = 0;

    if (tid < np) {
        S.patId[tid] = pl.fBinPats[pat0 + tid];
        S.patDone[tid] = 0;
        S.dpat[tid] = 0.0;
    }
    if (tid < 4) S.RE[(tid >> 1) * 2 * SM::RCAP + 2 * (SM::RCAP - 1) + (tid & 1)] = 0.0;    // {0,0} padding pairs
    if (tid < 2) S.bbuf[tid * SM::BSTRIDE + BLOCK] = 0.0;                               // zero neighbours
    __syncthreads();
    if (isTask) R.av = (double)pl.aInit[(size_t)S.patId[tp] * P + tk];   // makeInitCoalState, Core.hs:60

    int epoch = 0, nextEv = 0;
    int nextEvStep = nSEv > 0 ? sev[0].step : -1;
    int s = 0;
    int cp = 0;               // parity of the buffer that holds the current b when the slow path is entered
    bool init = true;

    for (;;) {
        // ================= slow path: shortcut test, structural events, epoch set-up =================
        // On exit everything is normalised to parity 0: b in bbuf[0], pairs and dt of step s at parity 0,
        // table rows s and s+1 in ring slots 0 and 1.
        {
            const int nnzw = pl.nnzw, nPat = pl.nPat;
            for (int r = 0; r < 2 && s + r < rowsAvail; ++r)
                for (int i = tid; i < rowLen; i += BLOCK) S.tabS[r * RC_ROWCAP + i] = rows[(long long)(s + r) * rowLen + i];
            if (isTask) S.aS[tid] = R.aProp ? 1.0 / (1.0 + R.uv) : R.av;
            S.dslot[tid] = R.dacc;
            R.dacc = 0.0;
            __syncthreads();
            if (tid < np && !S.patDone[tid] && !init) {       // updateD contributions so far (Core.hs:282-288)
                const int base = S.patOff[tid], cnt = S.patOff[tid + 1] - base;
                double acc = S.dpat[tid];
                for (int i = 0; i < cnt; ++i)
                    if (S.metaS[base + i] & RC_META_UNIT) acc = acc + S.dslot[base + i];
                S.dpat[tid] = acc;
            }
            if (s == sShort && !init) {                       // canUseShortcut / propagateStateShortcut, Core.hs:80-118
                int ok = 1;
                if (tid < np && !S.patDone[tid]) {
                    if (pl.shortcutOK[S.patId[tid]]) {
                        const double* aP = S.aS + tid * P;
                        double nrA = 0.0;
                        int popIndex = -1;
                        for (int k = 0; k < P; ++k) {
                            nrA = nrA + aP[k];
                            if (popIndex < 0 && aP[k] > 0.0) popIndex = k;
                        }
                        const double popSize = S.tabS[rc_row_n(P, A1) + (popIndex < 0 ? 0 : popIndex)];
                        const int base = S.patOff[tid], cnt = S.patOff[tid + 1] - base;
                        const double* bc = S.bbuf + cp * SM::BSTRIDE + base;
                        double add = 0.0;
                        for (int i = 0; i < cnt; ++i) {
                            const int nd = (int)(S.metaS[base + i] >> 16);
                            const double withComb = 2.0 * popSize / (double)nd;
                            add = add + bc[i] * (withComb / rc_choose_cont(nrA + (double)nd, nd));
                        }
                        S.dpat[tid] = S.dpat[tid] + add;
                        S.patDone[tid] = 2;
                    } else ok = 0;
                }
                if (__syncthreads_and(ok)) break;
            }
            // the initial state is handled as a pseudo-event that builds epoch 0 without a transition
            bool first = init;
            bool rebuilt = false;
            while (first || (nextEv < nSEv && nextEvStep == s)) {
                const double* W = wpool + models[b].wPool;
                if (!first) {
                    const RcSEvDev ev = sev[nextEv];
                    if (tid < np && !S.patDone[tid]) {
                        double* aP = S.aS + tid * P;
                        const double al = aP[ev.l], ak = aP[ev.k];
                        if (ev.type == RC_EV_JOIN_D) {            // popJoinA, Core.hs:165-170
                            aP[ev.l] = 0.0;
                            aP[ev.k] = al + ak;
                        } else {                                  // popSplitA, Core.hs:195-200
                            aP[ev.l] = (1.0 - ev.m) * al;
                            aP[ev.k] = ev.m * al + ak;
                        }
                    }
                    W += ev.wOff;
                    ++epoch;
                    ++nextEv;
                    nextEvStep = nextEv < nSEv ? sev[nextEv].step : -1;
                }
                // ---- per-thread structure of `epoch`; b is carried through the event's transition table
                //      (popJoinB/popSplitB as a gather, Core.hs:172-221)
                const int* so = pl.slotOff + (size_t)epoch * (nPat + 1);
                if (tid < np) {
                    int pid = S.patId[tid];
                    S.patOff2[tid] = so[pid + 1] - so[pid];
                    const uint8_t* mx = pl.maxX + ((size_t)epoch * nPat + pid) * P;
                    int r = 0;
                    for (int k = 0; k < P; ++k) r += mx[k];
                    S.patR[tid] = r;
                }
                __syncthreads();
                if (tid == 0) {
                    int run = 0, rr = 0;
                    for (int p = 0; p < np; ++p) {
                        int sz = S.patOff2[p]; S.patOff2[p] = run; run += sz;
                        int q = S.patR[p]; S.patR[p] = rr; rr += q;
                    }
                    S.patOff2[np] = run;
                    S.patR[np] = rr;
                }
                __syncthreads();
                if (isTask) {
                    const uint8_t* mx = pl.maxX + ((size_t)epoch * nPat + S.patId[tp]) * P;
                    int off = S.patR[tp];
                    for (int k = 0; k < tk; ++k) off += mx[k];
                    R.tRE = sBase + K::O_RE + (unsigned)off * 16u;
                    R.taskM = mx[tk];
                }
                const int nNew = S.patOff2[np];
                R.slotOn = tid < nNew;
                int n = 0;
                R.unit = false; R.anchor = true; R.bv = 0.0;
#pragma unroll
                for (int d = 0; d < NNZ; ++d) { R.aRE[d] = zeroRE; R.aB[d] = zeroB; }
                if (R.slotOn) {
                    const int p = rc_find_owner(S.patOff2, np, tid);
                    const int local = tid - S.patOff2[p];
                    const int pid = S.patId[p];
                    const int gl = so[pid] + local;
                    const long long g = pl.epochBase[epoch] + gl;
                    if (S.patDone[p]) R.slotOn = false;
                    if (!first) {
                        const int* trOff = pl.trOff + pl.trBase[epoch - 1];
                        const uint32_t* trEnt = pl.trEnt + pl.trEntBase[epoch - 1];
                        const double* bo = S.bbuf + cp * SM::BSTRIDE + S.patOff[p];
                        const int e0 = trOff[gl], e1 = trOff[gl + 1];
                        double acc = 0.0;
                        for (int q = e0; q < e1; ++q) {
                            const uint32_t en = trEnt[q];
                            const uint32_t widx = en >> 16;
                            const double old = bo[en & 0xFFFFu];
                            acc = (widx == RC_W_ONE) ? acc + old : acc + old * W[widx];
                        }
                        R.bv = acc;
                    } else {
                        // the seed state m is the last one fillUp emits (all components at their maximum)
                        R.bv = (tid + 1 == S.patOff2[p + 1]) ? 1.0 : 0.0;
                    }
                    const uint32_t mt = pl.meta[g];
                    S.metaS[tid] = mt;
                    n = (int)(mt & 0xFFu);
                    R.unit = (mt & RC_META_UNIT) != 0;
                    const uint8_t* mx = pl.maxX + ((size_t)epoch * nPat + pid) * P;
                    const int rBase = S.patR[p], bBase = S.patOff2[p];
#pragma unroll
                    for (int d = 0; d < NNZ; ++d) {
                        if (d < n) {
                            const uint32_t w = pl.words[g * nnzw + d];
                            const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                            const uint32_t up = w >> 16;
                            int off = rBase;
                            for (int kk = 0; kk < k; ++kk) off += mx[kk];
                            R.aRE[d] = sBase + K::O_RE + (unsigned)(off + x - 1) * 16u;
                            if (up != RC_NO_UP)
                                R.aB[d] = sBase + K::O_B + (unsigned)(bBase + (int)up) * 8u;
                        }
                    }
                }
                R.nw = __reduce_max_sync(0xFFFFFFFFu, n);
                __syncthreads();                      // every gather has read the old layout
                S.bbuf[tid] = R.bv;                   // new layout lives at parity 0
                cp = 0;
                if (tid <= np) S.patOff[tid] = S.patOff2[tid];
                first = false;
                rebuilt = true;
                __syncthreads();
            }
            init = false;
            if (!rebuilt) {                           // no event: only move b to parity 0 and refresh activity
                S.bbuf[tid] = R.bv;
                if (R.slotOn) {
                    const int p = rc_find_owner(S.patOff, np, tid);
                    if (S.patDone[p]) R.slotOn = false;
                }
            }
            // ---- a (possibly changed by events) back into the task registers; updateA of step s in place
            if (isTask) {
                R.av = S.aS[tid];
                R.aProp = R.av >= 1.0;
                if (R.aProp) R.uv = 1.0 / R.av - 1.0;
                R.taskOn = !S.patDone[tp];
            }
            if (s >= nSteps) break;
            if (tid == 0) { S.RE[2 * (SM::RCAP - 2)] = S.tabS[0]; S.RE[2 * (SM::RCAP - 2) + 1] = 0.0; }
            rc_fast_taskA<NNZ, BLOCK, 0, 0>(R);
            __syncthreads();
        }
        // ================================ fast path: two grid steps per trip ================================
        // next step index at which the slow path must run again
        int nextStop = nSteps;
        if (sShort > s && sShort < nextStop) nextStop = sShort;
        if (nextEvStep > s && nextEvStep < nextStop) nextStop = nextEvStep;
        const int pfStop = min(nextStop, rowsAvail - 1);       // last row worth prefetching
        const double* pf = rows + (long long)(s + 2) * rowLen + tid;
        for (;;) {
            {
                const bool bnd = s + 1 == nextStop;
                rc_fast_step<NNZ, BLOCK, 0>(R, sBase, pf, rowLen, s + 2 <= pfStop, !bnd, s);
                ++s;
                pf += rowLen;
                if (bnd) { cp = 1; break; }
            }
            {
                const bool bnd = s + 1 == nextStop;
                rc_fast_step<NNZ, BLOCK, 1>(R, sBase, pf, rowLen, s + 2 <= pfStop, !bnd, s);
                ++s;
                pf += rowLen;
                if (bnd) { cp = 0; break; }
            }
        }
        if (s >= nSteps && s != sShort) break;
    }
    // ---- getProb epilogue (Core.hs:49-54) ----
    __syncthreads();
    S.dslot[tid] = R.dacc;
    __syncthreads();
    if (tid < np) {
        const int pid = S.patId[tid];
        double d = S.dpat[tid];
        if (S.patDone[tid] != 2) {
            const int base = S.patOff[tid], cnt = S.patOff[tid + 1] - base;
            for (int i = 0; i < cnt; ++i)
                if (S.metaS[base + i] & RC_META_UNIT) d = d + S.dslot[base + i];
        }
        double discFac = 1.0;
        for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? models[b].disc[k] : 1.0);
        pOut[(size_t)b * pl.nPat + pid] = d * models[b].theta * pl.combFac[pid] * discFac;
    }
}

// ------------------------------------------------------------------------------------------------
// Keyed fast tier.
//
// Trajectories.  updateA (Core.hs:239-240) is a' = 1 / (1 + (1/a - 1) EA) with EA = exp(-dt/2N_k): with u = 1/a - 1 the
// recursion is u' = u EA, so inside a structural epoch
//     a(s) = 1 / (1 + u0 C_k(s)),   u0 = 1/a0 - 1,   C_k(s) = prod of EA_k over the epoch's steps up to and including s
// and C does not depend on the key, only on (model, branch).  rc_cum_kernel writes C (one warp scan per model, epoch and
// branch over the table rows rc_tables_kernel has written: EA is exactly 1 where the reference performs no update, i.e.
// dt <= 0 or a frozen branch); rc_traj_kernel then has NO sequential chain inside an epoch: every (key, allele count, step)
// is independent, and a launch cuts the epoch's steps into slices (grid.z) so that even a single model fills the device.
// Only the epochs follow each other (a0 comes from the parents' values at the end of the previous epoch through
// popJoinA / popSplitA, Core.hs:165-170,195-200, and u0 is formed from it with the reference's expression).  Against the
// literal recursion, a is not rounded and inverted again between steps: the values differ by a few 1e-16 relative (the first
// step of an epoch not at all), three orders of magnitude inside the 1e-10 the probabilities are tested to.  a < 1 is never
// propagated (Core.hs:239), and 1 / (1 + u) with -1 < u <= 0 cannot fall below 1, so "a >= 1" holds for the whole epoch or
// not at all.
// With the literal recursion a single model spent 0.26 of its 0.42 ms on the GPU in this chain (eight kernels of a few warps,
// each step two dependent divisions and an exp).
//
// For every step the kernel writes the pair
//     { G = exp(-(j(j-1)/2 (1/N_k) + j a_k (1/N_k)) dt),   E2 = 1 - exp(-j(j+1)/2 (1/N_k) dt) }
// i.e. the k-th factor of exp(-t1 dt) for a state with x_k = j (Core.hs:259,270) and the coalescence factor of
// b(x+e_k) (Core.hs:279).  All patterns that share the key read the same pairs.
__global__ void __launch_bounds__(32 * RC_MAX_P)
rc_cum_kernel(int nEpochs, const RcModelDev* __restrict__ models, const RcModelEpochDev* __restrict__ mep,
              const double* __restrict__ tab, double* __restrict__ cum, int P, int A1) {
    const int b = blockIdx.y, e = blockIdx.x;
    const RcModelDev& M = models[b];
    if (M.nSteps < 0 || M.mode != 1 || e >= nEpochs) return;
    const RcModelEpochDev me = mep[M.mepOff + e];
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= P) return;
    const int rowLen = rc_row_len(P, A1);
    const double* ea = tab + M.tabOff + RC_ROW_PAIRS + 2 * (k * A1);
    double* out = cum + (M.tabOff / rowLen) * P + k;
    double carry = 1.0;
    for (int s0 = me.beg; s0 < me.end; s0 += 32) {
        const int s = s0 + lane;
        double v = s < me.end ? ea[(long long)s * rowLen] : 1.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double t = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v *= t;
        }
        v *= carry;
        if (s < me.end) out[(long long)s * P] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
}

#ifndef RC_TCH
#define RC_TCH 1
#endif
__global__ void __launch_bounds__(128)
rc_traj_kernel(RcPlanDev pl, int e, const RcModelDev* __restrict__ models, const RcModelEpochDev* __restrict__ mep,
               const RcSEvDev* __restrict__ sevs, const double* __restrict__ tab, const double* __restrict__ cum,
               const double* __restrict__ dts, double* __restrict__ ktab, double* __restrict__ aEnd,
               double* __restrict__ aShort) {
    const int b = blockIdx.y;
    const RcModelDev& M = models[b];
    if (M.nSteps < 0 || M.mode != 1) return;
    const RcEpochDev E = pl.ep[e];
    const RcModelEpochDev me = mep[M.mepOff + e];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nst = me.end - me.beg;
    const bool lastEpoch = e + 1 == pl.nEpochs;
    // this slice of the epoch's steps
    const int per = (nst + (int)gridDim.z - 1) / (int)gridDim.z;
    const int sBeg = me.beg + (int)blockIdx.z * per, sEnd = min(me.end, sBeg + per);
    // pair 0 of every step row: {1, dt}: its first half doubles as the neutral decay factor of the pattern kernel's unused
    // unit-branch slots; a step with dt <= 0 performs no update (Core.hs:130)
    if (blockIdx.z == 0)
        for (int i = t; i < nst; i += gridDim.x * blockDim.x) {
            const double dt = dts[me.beg + i];
            double* p0 = ktab + (me.tabBase + (long long)i * E.rowPairs) * 2;
            p0[0] = 1.0;
            p0[1] = dt > 0.0 ? dt : 0.0;
        }
    if (t >= E.nThr) return;
    const int key = pl.thrKey[E.thrOff + t];
    const int kk = E.keyOff + key;
    const int k = pl.keyBranch[kk];
    const int j = t - pl.keyThrBase[kk] + 1;
    const bool hasPair = j <= (int)pl.keyMaxJ[kk];
    // a key of even length is stored with an odd stride (rc_pair_stride): the thread of its last pair also writes the padding
    // pair, so that a step row is written in whole 32-byte sectors
    const bool padLast = j == (int)pl.keyMaxJ[kk] && rc_pair_stride(j + pl.zeroRow) != j + pl.zeroRow;
    // a plan with a Split: the pair of x = 0, {1, 0}, in front of the key's pairs (pattern mode on bounding boxes, rc_plan.cpp)
    const bool zeroFirst = pl.zeroRow && j == 1;
    const bool special = j == 1 && blockIdx.z == 0;          // writes the key's a at the end of the epoch / at the shortcut step
    if (sBeg >= sEnd && !special) return;
    // ---- starting value ----
    double a0;
    {
        const int op = pl.keyOp[kk];
        const double* aPrev = e > 0 ? aEnd + M.aEndOff + pl.ep[e - 1].keyOff : aEnd;
        const int pa = pl.keyPa[kk], pb = pl.keyPb[kk];
        if (op == RC_KOP_INIT) a0 = (double)pl.keyInit[kk];                // makeInitCoalState, Core.hs:60
        else if (op == RC_KOP_COPY) a0 = aPrev[pa];
        else if (op == RC_KOP_JOIN) a0 = aPrev[pb] + aPrev[pa];            // popJoinA: al + ak
        else if (op == RC_KOP_ZERO) a0 = 0.0;
        else {
            const double m = sevs[M.sevOff + e - 1].m;
            if (op == RC_KOP_SPLIT_K) a0 = m * aPrev[pb] + aPrev[pa];      // popSplitA: m * al + ak
            else a0 = (1.0 - m) * aPrev[pa];                               // (1 - m) * al
        }
    }
    const bool act = a0 >= 1.0;
    const double u0 = act ? (1.0 / a0) - 1.0 : 0.0;
    const int rowLen = rc_row_len(pl.P, pl.A1);
    const double* cm = cum + (M.tabOff / rowLen) * pl.P + k;
    if (special) {
        // a after the step before s (the value the step s starts from)
        const double aLast = (act && nst > 0) ? 1.0 / (1.0 + u0 * cm[(long long)(me.end - 1) * pl.P]) : a0;
        aEnd[M.aEndOff + kk] = aLast;
        if (lastEpoch && M.sShort >= me.beg && M.sShort <= me.end)
            aShort[M.aShortOff + key] = (act && M.sShort > me.beg) ? 1.0 / (1.0 + u0 * cm[(long long)(M.sShort - 1) * pl.P]) : a0;
    }
    if (!hasPair && !padLast && !zeroFirst) return;
    // The key-independent factors of a step — 1/N_k, CN = j(j-1)/2N_k (Core.hs:270) and E2 (Core.hs:279) — come from the
    // per-step table rc_tables_kernel has written for this model (rc_types.h, table row layout); only the a-dependent
    // exponential is left per (key, j, step).
    const double xx = (double)j;
    const long long stride = (long long)E.rowPairs * 2;
    double* out = ktab + (me.tabBase * 2) + (long long)(pl.keyPairBase[kk] + j - 1) * 2 + (long long)(sBeg - me.beg) * stride;
    const double* row = tab + M.tabOff + (long long)sBeg * rowLen + RC_ROW_PAIRS + 2 * (k * pl.A1);
    const int jOff = hasPair ? 2 * j : 0;
    cm += (long long)sBeg * pl.P;
    // The steps are independent: the table entries of RC_TCH steps are loaded first (one L2 round trip for all of them),
    // then the steps are computed without a branch around the arithmetic (a step without update — dt <= 0 or a frozen
    // branch — selects {1, 0} afterwards).
    for (int s0 = sBeg; s0 < sEnd; s0 += RC_TCH) {
        double2 p0[RC_TCH], pj[RC_TCH];
        double dt[RC_TCH], c[RC_TCH];
#pragma unroll
        for (int q = 0; q < RC_TCH; ++q) {
            const long long o = min(s0 + q, sEnd - 1) - sBeg;
            p0[q] = *reinterpret_cast<const double2*>(row + o * rowLen);          // {EA, 1/N}; 1/N == 0: frozen
            pj[q] = *reinterpret_cast<const double2*>(row + o * rowLen + jOff);   // {CN, E2}
            dt[q] = dts[sBeg + o];
            c[q] = cm[o * pl.P];
        }
#pragma unroll
        for (int q = 0; q < RC_TCH; ++q) {
            if (s0 + q < sEnd) {
                const bool upd = dt[q] > 0.0 && p0[q].y != 0.0;
                const double a = act ? 1.0 / (1.0 + u0 * c[q]) : a0;               // propagateA, Core.hs:239-240
                const double t1k = pj[q].x + xx * a * p0[q].y;                     // Core.hs:270 (new a)
                const double G = exp(-(t1k * dt[q]));
                double* o = out + (long long)(s0 + q - sBeg) * stride;
                if (hasPair) *reinterpret_cast<double2*>(o) = upd ? make_double2(G, pj[q].y) : make_double2(1.0, 0.0);
                if (padLast) *reinterpret_cast<double2*>(o + 2) = make_double2(1.0, 0.0);
                if (zeroFirst) *reinterpret_cast<double2*>(o - 2) = make_double2(1.0, 0.0);
            }
        }
    }
}

// rc_propagate_keyed_kernel — the state side.  One live state per thread; per non-zero component one LDS.128 of the
// pair {G, E2}, one multiply into the decay factor, one LDS.64 of the neighbour and one FMA:
//     b'(x) = b(x) * prod_k G_k(x_k) + sum_k b(x + e_k) E2_k(x_k)            (Core.hs:259)
// The pairs of the bin's keys are staged from the trajectory table into an RC_KDEPTH-deep shared-memory ring with
// cp.async (one 16-byte pair row per lane, RC_KDEPTH-1 grid steps ahead); one barrier per grid step; the time loop is
// specialised per number of components and unrolled by the ring depth so that every buffer offset is an immediate.
#define RC_KDEPTH 8
struct RcKeySmem {
    static constexpr int BSTRIDE = RC_BLOCK + 8;       // entry RC_BLOCK of each b buffer is the always-zero neighbour
    double bbuf[2 * BSTRIDE];
    double RE[RC_KDEPTH * 2 * RC_KROWS];               // [slot][row]{G, E2}; row 0 = {dt, 0}; last row = {1, 0}
    double dslot[RC_BLOCK];
    double dpat[RC_NPB];
    uint32_t metaS[RC_BLOCK];
    int patOff[RC_NPB + 1];
    int patId[RC_NPB];
    int patDone[RC_NPB];
};
struct RcKeyK {
    static constexpr unsigned B_PAR = RcKeySmem::BSTRIDE * 8u;
    static constexpr unsigned RE_SLOT = RC_KROWS * 16u;
    static constexpr unsigned RE_PAD = (RC_KROWS - 1) * 16u;
    static constexpr unsigned O_B = (unsigned)offsetof(RcKeySmem, bbuf);
    static constexpr unsigned O_RE = (unsigned)offsetof(RcKeySmem, RE);
    static_assert(O_RE % 16 == 0, "16-byte alignment of the pair table");
};

template <int NNZ>
struct RcKeyRegs {
    unsigned aRE[NNZ], aB[NNZ];   // absolute shared addresses (ring slot 0 / b parity 0)
    int nw;
    bool unit, slotOn, fetchOn;
    double bv, dacc;
    const double* fPtr;           // this lane's pair in the step row to be copied next
};

// One grid step of the keyed kernel; ring slot U (compile time), b parity U & 1.
template <int NNZ, int N, int U>
__device__ __forceinline__ void rc_key_step(RcKeyRegs<NNZ>& R, const unsigned sBase, const unsigned myRE, const unsigned myB,
                                            const long long strideD, const unsigned fetchPred, const bool warpOn) {
    typedef RcKeyK K;
    constexpr int PH = U & 1, QH = PH ^ 1;
    constexpr unsigned slot = (unsigned)U * K::RE_SLOT;
    // pairs of step s + DEPTH - 1 -> ring slot (U - 1) mod DEPTH (read one step ago).  Steps past the end of the epoch
    // read rows that are never used (the table carries RC_KDEPTH rows of slack).
    rc_cp_async16_if(myRE + (unsigned)((U + RC_KDEPTH - 1) % RC_KDEPTH) * K::RE_SLOT, R.fPtr, fetchPred);
    rc_cp_commit();
    R.fPtr += strideD;
    if (warpOn) {      // warp-uniform; lanes without a state run on the padding pair / zero neighbour and keep b = 0
        double g, e2;
        rc_lds_pair(R.aRE[0] + slot, g, e2);
        double D = g;
        double t2 = rc_lds64(R.aB[0] + PH * K::B_PAR) * e2;
#pragma unroll
        for (int d = 1; d < N; ++d) {
            rc_lds_pair(R.aRE[d] + slot, g, e2);
            D *= g;
            t2 = fma(rc_lds64(R.aB[d] + PH * K::B_PAR), e2, t2);
        }
        R.bv = fma(R.bv, D, t2);                                               // updateB, Core.hs:259
        rc_sts64(myB + QH * K::B_PAR, R.bv);
        if (R.unit) R.dacc = fma(rc_lds64(sBase + K::O_RE + slot + 8u), R.bv, R.dacc);   // updateD, Core.hs:282-288
    }
    rc_cp_wait<RC_KDEPTH - 2>();       // the pairs of the next step have landed
    __syncthreads();
}

// Grid steps s .. nextStop-1 for a warp whose states have at most N non-zero components.  Returns the parity of the
// b buffer that holds the result.
template <int NNZ, int N>
__device__ __forceinline__ int rc_key_run(RcKeyRegs<NNZ>& R, const unsigned sBase, const long long strideD, int& s,
                                          const int nextStop) {
    typedef RcKeyK K;
    const unsigned myRE = sBase + K::O_RE + threadIdx.x * 16u;
    const unsigned myB = sBase + K::O_B + threadIdx.x * 8u;
    const unsigned fetchPred = R.fetchOn ? 1u : 0u;
    const bool warpOn = __any_sync(0xFFFFFFFFu, R.slotOn);
    const int s0 = s;
    // whole ring turns without exit tests
    while (s + RC_KDEPTH <= nextStop) {
        rc_key_step<NNZ, N, 0>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 1>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 2>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 3>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 4>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 5>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 6>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        rc_key_step<NNZ, N, 7>(R, sBase, myRE, myB, strideD, fetchPred, warpOn);
        s += RC_KDEPTH;
    }
    static_assert(RC_KDEPTH == 8, "the unrolled ring turn above assumes a depth of 8");
    // remainder (< DEPTH steps)
    if (s < nextStop) { rc_key_step<NNZ, N, 0>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 1>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 2>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 3>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 4>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 5>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    if (s < nextStop) { rc_key_step<NNZ, N, 6>(R, sBase, myRE, myB, strideD, fetchPred, warpOn); ++s; }
    return (s - s0) & 1;
}

#ifndef RC_KEYED_MINB
#define RC_KEYED_MINB 4
#endif
// One launch per structural epoch `e`: block = (bin of that epoch, model).  b enters through the Join/Split
// transition table of the event that opened the epoch (popJoinB / popSplitB as a gather, Core.hs:172-221) from the
// previous launch's carry buffer and leaves through this launch's; the per-pattern updateD sums travel in dAcc.
// The last epoch also does the shortcut (Core.hs:80-118) and the getProb epilogue (Core.hs:49-54).
template <int NNZ>
__global__ void __launch_bounds__(RC_BLOCK, RC_KEYED_MINB)
rc_propagate_keyed_kernel(RcPlanDev pl, int e, int binBase, const RcModelDev* __restrict__ models,
                          const RcModelEpochDev* __restrict__ mep, const RcSEvDev* __restrict__ sevs,
                          const double* __restrict__ ktab, const double* __restrict__ aShort,
                          const double* __restrict__ wpool, const double* __restrict__ bPrev, double* __restrict__ bCur,
                          long long bStride, double* __restrict__ dAcc, double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    typedef RcKeySmem SM;
    typedef RcKeyK K;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    if (models[b].nSteps < 0 || models[b].mode != 1) return;
    const int P = pl.P, nPat = pl.nPat;
    SM& S = *reinterpret_cast<SM*>(smemRaw);
    const unsigned sBase = rc_smem_base(smemRaw);
    const int binIdx = binBase + (int)blockIdx.x;
    // bin image (rc_types.h): header -> pattern / lane records -> transition entries
    const int4 hdrA = reinterpret_cast<const int4*>(pl.kBinHdr)[(RC_HDR_N / 4) * binIdx];
    const int4 hdrB = reinterpret_cast<const int4*>(pl.kBinHdr)[(RC_HDR_N / 4) * binIdx + 1];
    const int pat0 = hdrA.x, np = hdrA.y, nSlots = hdrA.z;
    const bool isLast = e + 1 == pl.nEpochs;
    const int sShort = models[b].sShort;
    const RcModelEpochDev mE = mep[models[b].mepOff + e];

    RcKeyRegs<NNZ> R;
    const unsigned padRE = sBase + K::O_RE + K::RE_PAD;
    const unsigned zeroB = sBase + K::O_B + RC_BLOCK * 8u;
    R.nw = 1; R.unit = false; R.fetchOn = false;
    R.bv = 0.0; R.dacc = 0.0;
    R.fPtr = ktab;
    R.slotOn = tid < nSlots;

    // ---- set-up ----
    // start the pair ring right away: its latency overlaps the rest of the set-up
    const int rowPairs = pl.ep[e].rowPairs;
    const long long strideD = (long long)rowPairs * 2;
    R.fetchOn = tid < hdrB.y;
    const int fOff = R.fetchOn ? pl.kFetchRows[hdrB.x + tid] : 0;
    int s = mE.beg;
    bool ringStarted = false;
    if (s < mE.end) {
        R.fPtr = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs + fOff) * 2;
#pragma unroll
        for (int u = 0; u < RC_KDEPTH - 1; ++u) {
            rc_cp_async16_if(sBase + K::O_RE + (unsigned)u * K::RE_SLOT + tid * 16u, R.fPtr, R.fetchOn ? 1u : 0u);
            rc_cp_commit();
            R.fPtr += strideD;
        }
        ringStarted = true;
    }
    if (tid < np) {
        const int2 rec = reinterpret_cast<const int2*>(pl.kPatRec)[pat0 + tid];
        S.patId[tid] = rec.x;
        S.patOff[tid] = rec.y & 0xFFFF;
        S.patDone[tid] = 0;
        S.dpat[tid] = e == 0 ? 0.0 : dAcc[(size_t)b * nPat + rec.x];
    }
    if (tid == 0) S.patOff[np] = nSlots;
    if (tid < 2 * RC_KDEPTH)                                                   // {1,0} padding pair of every ring slot
        S.RE[(tid >> 1) * 2 * RC_KROWS + 2 * (RC_KROWS - 1) + (tid & 1)] = (tid & 1) ? 0.0 : 1.0;
    if (tid < 2) S.bbuf[tid * SM::BSTRIDE + RC_BLOCK] = 0.0;                   // zero neighbours
    int myLocal = 0;
    {
        int n = 0;
        uint32_t mt = 0u;
        uint4 la = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u), lb = make_uint4(0u, 0u, 0u, 0u);
        if (R.slotOn) {
            const uint4* lane = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + 2 * (size_t)tid;
            la = lane[0];
            lb = lane[1];
            mt = lb.x;
            myLocal = (int)lb.y;
            n = (int)(mt & 0xFFu);
            R.unit = (mt & RC_META_UNIT) != 0;
            if (lb.z == RC_IMG_SEED) R.bv = 1.0;                                // makeInitCoalState, Core.hs:66
            else {
                // popJoinB / popSplitB as a gather over the previous epoch's live states of this pattern
                const int cnt = (int)(lb.w >> RC_IMG_CNT_SHIFT);
                if (cnt) {
                    const double* W = wpool + models[b].wPool + sevs[models[b].sevOff + e - 1].wOff;
                    const double* bo = bPrev + (size_t)b * bStride + (lb.w & ((1u << RC_IMG_CNT_SHIFT) - 1u));
                    R.bv = rc_gather(pl.trEnt + lb.z, 0, cnt, bo, W);
                }
            }
        }
        S.metaS[tid] = mt;
#pragma unroll
        for (int d = 0; d < NNZ; ++d) {
            const uint32_t rw = d < 4 ? la.x : la.y, uw = d < 4 ? la.z : la.w;
            const unsigned row = (rw >> (8 * (d & 3))) & 0xFFu, up = (uw >> (8 * (d & 3))) & 0xFFu;
            R.aRE[d] = sBase + K::O_RE + row * 16u;
            R.aB[d] = (!R.slotOn || up == (unsigned)tid) ? zeroB : sBase + K::O_B + up * 8u;
        }
        R.nw = max(1, __reduce_max_sync(0xFFFFFFFFu, n));
    }
    S.bbuf[tid] = R.bv;
    __syncthreads();

    int cp = 0;
    bool finished = false;
    // at most two runs: up to the end of the epoch (or to the shortcut step in the last epoch), then — only if some
    // pattern of the bin cannot take the shortcut — on to the end of the grid
    for (int phase = 0; phase < 2 && !finished; ++phase) {
        int nextStop = mE.end;
#ifdef RC_PROBE_NOLOOP
        nextStop = s;      // timing probe only (tools/setup_probe.py): set-up and hand-over without the grid steps
#endif
        if (isLast && sShort >= s && phase == 0) nextStop = min(nextStop, sShort);
        if (s < nextStop) {
            if (!ringStarted) {
                R.fPtr = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs + fOff) * 2;
#pragma unroll
                for (int u = 0; u < RC_KDEPTH - 1; ++u) {
                    rc_cp_async16_if(sBase + K::O_RE + (unsigned)u * K::RE_SLOT + tid * 16u, R.fPtr, R.fetchOn ? 1u : 0u);
                    rc_cp_commit();
                    R.fPtr += strideD;
                }
            }
            ringStarted = false;
            rc_cp_wait<RC_KDEPTH - 2>();
            __syncthreads();
            switch (R.nw) {
                case 1: cp = rc_key_run<NNZ, 1>(R, sBase, strideD, s, nextStop); break;
                case 2: cp = rc_key_run<NNZ, 2>(R, sBase, strideD, s, nextStop); break;
                case 3: cp = rc_key_run<NNZ, (NNZ >= 3 ? 3 : NNZ)>(R, sBase, strideD, s, nextStop); break;
                case 4: cp = rc_key_run<NNZ, (NNZ >= 4 ? 4 : NNZ)>(R, sBase, strideD, s, nextStop); break;
                case 5: cp = rc_key_run<NNZ, (NNZ >= 5 ? 5 : NNZ)>(R, sBase, strideD, s, nextStop); break;
                case 6: cp = rc_key_run<NNZ, (NNZ >= 6 ? 6 : NNZ)>(R, sBase, strideD, s, nextStop); break;
                case 7: cp = rc_key_run<NNZ, (NNZ >= 7 ? 7 : NNZ)>(R, sBase, strideD, s, nextStop); break;
                default: cp = rc_key_run<NNZ, NNZ>(R, sBase, strideD, s, nextStop); break;
            }
            rc_cp_wait<0>();
        }
        // ---- updateD contributions of the run (Core.hs:282-288) ----
        S.dslot[tid] = R.dacc;
        R.dacc = 0.0;
        __syncthreads();
        if (tid < np && !S.patDone[tid]) {
            const int base = S.patOff[tid], cnt = S.patOff[tid + 1] - base;
            double acc = S.dpat[tid];
            for (int i = 0; i < cnt; ++i)
                if (S.metaS[base + i] & RC_META_UNIT) acc = acc + S.dslot[base + i];
            S.dpat[tid] = acc;
        }
        if (!(isLast && phase == 0 && s == sShort && s < models[b].nSteps)) { finished = true; break; }
        // ---- canUseShortcut / propagateStateShortcut (Core.hs:80-118): the event queue is empty from here on ----
        int ok = 1;
        if (tid < np) {
            const int pid = S.patId[tid];
            if (pl.shortcutOK[pid]) {
                const double* aK = aShort + models[b].aShortOff;
                const int* kid = pl.lastKeyId + (size_t)pid * P;
                double nrA = 0.0;
                int popIndex = -1;
                for (int k = 0; k < P; ++k) {
                    const double ak = aK[kid[k]];
                    nrA = nrA + ak;
                    if (popIndex < 0 && ak > 0.0) popIndex = k;
                }
                const double popSize = models[b].Nshort[popIndex < 0 ? 0 : popIndex];
                const int base = S.patOff[tid], cnt = S.patOff[tid + 1] - base;
                const double* bc = S.bbuf + cp * SM::BSTRIDE + base;
                double add = 0.0;
                for (int i = 0; i < cnt; ++i) {
                    const int nd = (int)(S.metaS[base + i] >> 16);
                    const double withComb = 2.0 * popSize / (double)nd;
                    add = add + bc[i] * (withComb / rc_choose_cont(nrA + (double)nd, nd));
                }
                S.dpat[tid] = S.dpat[tid] + add;
                S.patDone[tid] = 2;
            } else ok = 0;
        }
        if (__syncthreads_and(ok)) { finished = true; break; }
        // some pattern continues to the end of the grid: park the finished ones, move b to parity 0
        if (R.slotOn) {
            const int p = rc_find_owner(S.patOff, np, tid);
            if (S.patDone[p]) {
                R.slotOn = false;
                R.unit = false;
#pragma unroll
                for (int d = 0; d < NNZ; ++d) { R.aRE[d] = padRE; R.aB[d] = zeroB; }
            }
        }
        S.bbuf[tid] = R.bv;
        cp = 0;
        __syncthreads();
    }
    if (!isLast) {
        // ---- hand b and the updateD sums to the next epoch's launch ----
        if (tid < nSlots) bCur[(size_t)b * bStride + myLocal] = R.bv;
        if (tid < np) dAcc[(size_t)b * nPat + S.patId[tid]] = S.dpat[tid];
        return;
    }
    // ---- getProb epilogue (Core.hs:49-54) ----
    if (tid < np) {
        const int pid = S.patId[tid];
        const double d = S.dpat[tid];
        double discFac = 1.0;
        for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? models[b].disc[k] : 1.0);
        pOut[(size_t)b * nPat + pid] = d * models[b].theta * pl.combFac[pid] * discFac;
    }
}

// ------------------------------------------------------------------------------------------------
// Row mode of the keyed tier (epochs whose live sets are boxes with few non-zero branches): one thread owns the
// row of states x_r = 1..M along the widest branch r of its pattern, the other coordinates fixed.  b of the row lives
// in registers; the neighbour along r is the next register, the {G,E2} pairs of the other branches are loaded once per
// row, and only the neighbouring rows and the row branch's own pairs come from shared memory:
//     b'(j) = b(j) * [prod_{d != r} G_d(x_d)] * G_r(j) + sum_{d != r} b_{row+e_d}(j) * E2_d(x_d) + b(j+1) * E2_r(j)
// (Core.hs:259).  Rows are stored row-major with an odd stride (M | 1), so a warp reading one column of consecutive
// rows touches distinct banks.
#define RC_RDEPTH 4
#ifndef RC_ROW_GU
#define RC_ROW_GU 2
#endif
#ifndef RC_ROW_SPLIT_O
#define RC_ROW_SPLIT_O 0
#endif
#ifndef RC_ROW_SPLIT_R
#define RC_ROW_SPLIT_R 0
#endif
struct RcRowSmem {
    static constexpr int BPAR = RC_ROW_BCAP + 16;      // doubles per parity; the last 16 are the always-zero row
    double bbuf[2 * BPAR];
    double RE[RC_RDEPTH * 2 * RC_KROWS];               // [slot][row]{G, E2}; row 0 = {dt, 0}; last row = {1, 0}
    double dslot[RC_BLOCK];
    double dpat[RC_NPBR];
    int patRowOff[RC_NPBR + 1];
    int patId[RC_NPBR];
    int patDone[RC_NPBR];
};
struct RcRowK {
    static constexpr unsigned B_PAR = RcRowSmem::BPAR * 8u;
    static constexpr unsigned RE_SLOT = RC_KROWS * 16u;
    static constexpr unsigned RE_PAD = (RC_KROWS - 1) * 16u;
    static constexpr unsigned O_B = (unsigned)offsetof(RcRowSmem, bbuf);
    static constexpr unsigned O_RE = (unsigned)offsetof(RcRowSmem, RE);
    static constexpr unsigned ZERO_B = RC_ROW_BCAP * 8u;
    static_assert(O_RE % 16 == 0, "16-byte alignment of the pair table");
};

template <int NOCAP>
struct RcRowRegs {
    double b[RC_ROW_MAXM];
    unsigned aRE[NOCAP], aB[NOCAP];   // pair of (branch d, x_d) in ring slot 0 / neighbour row (parity 0)
    unsigned aREr, myB;                        // pair of (row branch, j = 1) in ring slot 0 / own row (parity 0)
    int M, NO;
    bool unit, rowOn, fetchOn;
    double dacc;
    const double* fPtr;
};

// grid steps s .. nextStop-1 for a warp whose rows have at most NO other branches and length at most M
// EX: every lane of the warp owns a row of exactly M states — no per-state predicates
template <int NO, int M, bool EX, int NOCAP>
__device__ __forceinline__ int rc_row_run(RcRowRegs<NOCAP>& R, const unsigned sBase, const long long strideD, int& s,
                                          const int nextStop) {
    typedef RcRowK K;
    const unsigned myRE = sBase + K::O_RE + threadIdx.x * 16u;
    const unsigned fetchPred = R.fetchOn ? 1u : 0u;
    const bool warpOn = __any_sync(0xFFFFFFFFu, R.rowOn);
    unsigned slot = 0, prev = (RC_RDEPTH - 1) * K::RE_SLOT;    // ring slot of step s / of step s-1
    unsigned po = 0, qo = K::B_PAR;                            // parity offsets of the b buffer read / written
    const int s0 = s;
    for (; s < nextStop; ++s) {
        rc_cp_async16_if(myRE + prev, R.fPtr, fetchPred);      // pairs of step s + DEPTH - 1 -> slot of step s - 1
        rc_cp_commit();
        R.fPtr += strideD;
        if (warpOn) {
            double D = 1.0, e2o[NO > 0 ? NO : 1];
#pragma unroll
            for (int d = 0; d < NO; ++d) {
                double g;
                rc_lds_pair<RC_ROW_SPLIT_O>(R.aRE[d] + slot, g, e2o[d]);
                D *= g;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const bool v = EX || j < R.M;
                double gj = 1.0, e2j = 0.0;
                if (v) rc_lds_pair<RC_ROW_SPLIT_R>(R.aREr + slot + (unsigned)j * 16u, gj, e2j);
                double t2 = (j + 1 < M) ? R.b[j + 1] * e2j : 0.0;             // old b(j+1): not yet overwritten
#pragma unroll
                for (int d = 0; d < NO; ++d) {
                    // (predicating off the loads of rows without a neighbour along d, instead of pointing them at the zero
                    // row, was measured: fewer bank conflicts, but 5 % slower for the extra predicate arithmetic)
                    double nb = 0.0;
                    if (v) nb = rc_lds64(R.aB[d] + po + (unsigned)j * 8u);
                    t2 = fma(nb, e2o[d], t2);
                }
                R.b[j] = fma(R.b[j], D * gj, t2);                            // updateB, Core.hs:259
                if (v) rc_sts64(R.myB + qo + (unsigned)j * 8u, R.b[j]);
            }
            if (R.unit) R.dacc = fma(rc_lds64(sBase + K::O_RE + slot + 8u), R.b[0], R.dacc);   // updateD, Core.hs:282-288
        }
        rc_cp_wait<RC_RDEPTH - 2>();
        __syncthreads();
        prev = slot;
        slot = slot + K::RE_SLOT == RC_RDEPTH * K::RE_SLOT ? 0u : slot + K::RE_SLOT;
        const unsigned t = po; po = qo; qo = t;
    }
    return (s - s0) & 1;
}

// Row lengths that occur with NO other branches: any up to RC_ROW_MAXM for NO <= 3, else NO + M <= RC_ROW_SUM (planner rule).
// A length the planner never produces maps onto the longest variant that exists, so no dead variants are compiled.
template <int NO, int M>
struct RcRowLen {
    static constexpr int MAXM = NO <= 3 ? RC_ROW_MAXM : (RC_ROW_SUM - NO < 1 ? 1 : RC_ROW_SUM - NO);
    static constexpr int value = M <= MAXM ? M : MAXM;
};
template <int NO, int NOCAP>
__device__ __forceinline__ int rc_row_dispatch_m(RcRowRegs<NOCAP>& R, const unsigned sBase, const long long strideD, int& s,
                                                 const int nextStop, const int Mw) {
    // bins hold rows of one length (rc_plan.cpp, cost classes), so most warps take the predicate-free variant
    if (__all_sync(0xFFFFFFFFu, R.rowOn && R.M == Mw)) {
        switch (Mw) {
            case 1: return rc_row_run<NO, 1, true>(R, sBase, strideD, s, nextStop);
            case 2: return rc_row_run<NO, RcRowLen<NO, 2>::value, true>(R, sBase, strideD, s, nextStop);
            case 3: return rc_row_run<NO, RcRowLen<NO, 3>::value, true>(R, sBase, strideD, s, nextStop);
            case 4: return rc_row_run<NO, RcRowLen<NO, 4>::value, true>(R, sBase, strideD, s, nextStop);
            case 5: return rc_row_run<NO, RcRowLen<NO, 5>::value, true>(R, sBase, strideD, s, nextStop);
            case 6: return rc_row_run<NO, RcRowLen<NO, 6>::value, true>(R, sBase, strideD, s, nextStop);
            case 7: return rc_row_run<NO, RcRowLen<NO, 7>::value, true>(R, sBase, strideD, s, nextStop);
            case 8: return rc_row_run<NO, RcRowLen<NO, 8>::value, true>(R, sBase, strideD, s, nextStop);
            case 9: return rc_row_run<NO, RcRowLen<NO, 9>::value, true>(R, sBase, strideD, s, nextStop);
            default: return rc_row_run<NO, RcRowLen<NO, RC_ROW_MAXM>::value, true>(R, sBase, strideD, s, nextStop);
        }
    }
    switch (Mw) {
        case 1: return rc_row_run<NO, 1, false>(R, sBase, strideD, s, nextStop);
        case 2: return rc_row_run<NO, RcRowLen<NO, 2>::value, false>(R, sBase, strideD, s, nextStop);
        case 3: return rc_row_run<NO, RcRowLen<NO, 3>::value, false>(R, sBase, strideD, s, nextStop);
        case 4: return rc_row_run<NO, RcRowLen<NO, 4>::value, false>(R, sBase, strideD, s, nextStop);
        case 5: return rc_row_run<NO, RcRowLen<NO, 5>::value, false>(R, sBase, strideD, s, nextStop);
        case 6: return rc_row_run<NO, RcRowLen<NO, 6>::value, false>(R, sBase, strideD, s, nextStop);
        case 7: case 8: return rc_row_run<NO, RcRowLen<NO, 8>::value, false>(R, sBase, strideD, s, nextStop);
        default: return rc_row_run<NO, RcRowLen<NO, RC_ROW_MAXM>::value, false>(R, sBase, strideD, s, nextStop);
    }
}

#ifndef RC_ROWS_MINB
#define RC_ROWS_MINB 3
#endif
// NOCAP = 3 for epochs whose rows all have at most 3 other branches (fewer registers, smaller code), else RC_ROW_NO
template <int NOCAP>
__global__ void __launch_bounds__(RC_BLOCK, RC_ROWS_MINB)
rc_propagate_rows_kernel(RcPlanDev pl, int e, int binBase, const RcModelDev* __restrict__ models,
                         const RcModelEpochDev* __restrict__ mep, const RcSEvDev* __restrict__ sevs,
                         const double* __restrict__ ktab, const double* __restrict__ aShort,
                         const double* __restrict__ wpool, const double* __restrict__ bPrev, double* __restrict__ bCur,
                         long long bStride, double* __restrict__ dAcc, double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    typedef RcRowSmem SM;
    typedef RcRowK K;
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    if (models[b].nSteps < 0 || models[b].mode != 1) return;
    const int P = pl.P, nPat = pl.nPat;
    SM& S = *reinterpret_cast<SM*>(smemRaw);
    const unsigned sBase = rc_smem_base(smemRaw);
    const int binIdx = binBase + (int)blockIdx.x;
    // bin image (rc_types.h): header -> pattern / row / element records -> transition entries
    const int4 hdrA = reinterpret_cast<const int4*>(pl.kBinHdr)[(RC_HDR_N / 4) * binIdx];
    const int4 hdrB = reinterpret_cast<const int4*>(pl.kBinHdr)[(RC_HDR_N / 4) * binIdx + 1];
    const int pat0 = hdrA.x, np = hdrA.y, nRows = hdrA.z, nB = hdrA.w;
    const bool isLast = e + 1 == pl.nEpochs;
    const int sShort = models[b].sShort;
    const RcModelEpochDev mE = mep[models[b].mepOff + e];

    RcRowRegs<NOCAP> R;
#pragma unroll
    for (int j = 0; j < RC_ROW_MAXM; ++j) R.b[j] = 0.0;
    const unsigned padRE = sBase + K::O_RE + K::RE_PAD;
    const unsigned zeroB = sBase + K::O_B + K::ZERO_B;
#pragma unroll
    for (int d = 0; d < NOCAP; ++d) { R.aRE[d] = padRE; R.aB[d] = zeroB; }
    R.aREr = padRE; R.myB = zeroB; R.M = 0; R.NO = 0;
    R.unit = false; R.fetchOn = false; R.dacc = 0.0;
    R.fPtr = ktab;
    R.rowOn = tid < nRows;

    // ---- set-up ----
    // start the pair ring right away: its latency overlaps the rest of the set-up
    const int rowPairs = pl.ep[e].rowPairs;
    const long long strideD = (long long)rowPairs * 2;
    R.fetchOn = tid < hdrB.y;
    const int fOff = R.fetchOn ? pl.kFetchRows[hdrB.x + tid] : 0;
    int s = mE.beg;
    bool ringStarted = false;
    if (s < mE.end) {
        R.fPtr = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs + fOff) * 2;
#pragma unroll
        for (int u = 0; u < RC_RDEPTH - 1; ++u) {
            rc_cp_async16_if(sBase + K::O_RE + (unsigned)u * K::RE_SLOT + tid * 16u, R.fPtr, R.fetchOn ? 1u : 0u);
            rc_cp_commit();
            R.fPtr += strideD;
        }
        ringStarted = true;
    }
    // the row's record; its loads are in flight while the block gathers b
    uint4 la = make_uint4(0u, 0u, 0u, 0u), lb = make_uint4(0u, 0u, 0u, 0u), lc = make_uint4(0u, 0u, 0u, 0u);
    if (R.rowOn) {
        const uint4* lane = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + 3 * (size_t)tid;
        la = lane[0];
        lb = lane[1];
        lc = lane[2];
    }
    if (tid < np) {
        const int2 rec = reinterpret_cast<const int2*>(pl.kPatRec)[pat0 + tid];
        S.patId[tid] = rec.x;
        S.patRowOff[tid] = rec.y & 0xFFFF;
        S.patDone[tid] = 0;
        S.dpat[tid] = e == 0 ? 0.0 : dAcc[(size_t)b * nPat + rec.x];
    }
    if (tid == 0) S.patRowOff[np] = nRows;
    if (tid < 2 * RC_RDEPTH)                                                   // {1,0} padding pair of every ring slot
        S.RE[(tid >> 1) * 2 * RC_KROWS + 2 * (RC_KROWS - 1) + (tid & 1)] = (tid & 1) ? 0.0 : 1.0;
    if (tid < 32) S.bbuf[(tid >> 4) * SM::BPAR + RC_ROW_BCAP + (tid & 15)] = 0.0;   // zero rows
    // ---- b of every state of the bin into shared memory (parity 0), all threads cooperating: the seed state (epoch 0)
    //      or a gather through the event's transition table from the previous launch's carry buffer ----
    {
        const uint2* el = reinterpret_cast<const uint2*>(pl.kElem) + hdrB.w;
        const double* W = e > 0 ? wpool + models[b].wPool + sevs[models[b].sevOff + e - 1].wOff : nullptr;
        const double* bp = bPrev + (size_t)b * bStride;
        constexpr int GU = RC_ROW_GU;
        for (int i = tid; i < nB; i += GU * RC_BLOCK) {
            uint2 g[GU];
            double v[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) g[u] = i + u * RC_BLOCK < nB ? el[i + u * RC_BLOCK] : make_uint2(0u, 0u);
            rc_gather_batch<GU, 4>(g, pl.trEnt, bp, W, v);
#pragma unroll
            for (int u = 0; u < GU; ++u)
                if (i + u * RC_BLOCK < nB) S.bbuf[i + u * RC_BLOCK] = v[u];
        }
        // the other parity starts as zeros: a row may be longer than its neighbour row (the staircase a Split leaves), and
        // the elements the neighbour does not own — states that are not live — must read as b = 0 in both parities
        for (int i = tid; i < nB; i += RC_BLOCK) S.bbuf[SM::BPAR + i] = 0.0;
    }
    __syncthreads();
    const uint16_t* myRefs = pl.rowRefAll;
    int myPid = 0, myOut = 0;
    if (R.rowOn) {
        R.M = (int)(la.z & 0xFFu);
        R.NO = (int)((la.z >> 8) & 0xFFu);
        R.unit = R.NO == 0;                                     // x = e_r is the j = 1 state of a single-branch pattern
        R.aREr = sBase + K::O_RE + (la.y >> 24) * 16u;
        R.myB = sBase + K::O_B + (la.w & 0xFFFFu) * 8u;
        const unsigned rowsW[2] = {la.x, la.y};
        const unsigned offW[4] = {la.w, lb.x, lb.y, lb.z};       // 8 x u16: own row, neighbour rows
#pragma unroll
        for (int d = 0; d < NOCAP; ++d) {
            R.aRE[d] = sBase + K::O_RE + ((rowsW[d >> 2] >> (8 * (d & 3))) & 0xFFu) * 16u;
            R.aB[d] = sBase + K::O_B + ((offW[(d + 1) >> 1] >> (16 * ((d + 1) & 1))) & 0xFFFFu) * 8u;
        }
        myOut = (int)lb.w;
        myRefs = pl.rowRefAll + (size_t)lc.x * RC_ROW_MAXM;
        myPid = (int)lc.y;
        // b of the row from shared memory (filled cooperatively above)
#pragma unroll
        for (int j = 0; j < RC_ROW_MAXM; ++j)
            if (j < R.M) R.b[j] = rc_lds64(R.myB + (unsigned)j * 8u);
    }
    const int NOw = __reduce_max_sync(0xFFFFFFFFu, R.NO);
    const int Mw = max(1, __reduce_max_sync(0xFFFFFFFFu, R.M));
    __syncthreads();

    int cp = 0;
    bool finished = false;
    for (int phase = 0; phase < 2 && !finished; ++phase) {
        int nextStop = mE.end;
#ifdef RC_PROBE_NOLOOP
        nextStop = s;      // timing probe only (tools/setup_probe.py): set-up and hand-over without the grid steps
#endif
        if (isLast && sShort >= s && phase == 0) nextStop = min(nextStop, sShort);
        if (s < nextStop) {
            if (!ringStarted) {
                R.fPtr = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs + fOff) * 2;
#pragma unroll
                for (int u = 0; u < RC_RDEPTH - 1; ++u) {
                    rc_cp_async16_if(sBase + K::O_RE + (unsigned)u * K::RE_SLOT + tid * 16u, R.fPtr, R.fetchOn ? 1u : 0u);
                    rc_cp_commit();
                    R.fPtr += strideD;
                }
            }
            ringStarted = false;
            rc_cp_wait<RC_RDEPTH - 2>();
            __syncthreads();
            int par;
            static_assert(RC_ROW_NO == 7, "dispatch over the number of other branches");
            if constexpr (NOCAP <= 3) {
                switch (NOw) {
                    case 0: par = rc_row_dispatch_m<0>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 1: par = rc_row_dispatch_m<1>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 2: par = rc_row_dispatch_m<2>(R, sBase, strideD, s, nextStop, Mw); break;
                    default: par = rc_row_dispatch_m<3>(R, sBase, strideD, s, nextStop, Mw); break;
                }
            } else {
                // a warp at the border of two cost classes can combine many other branches (one lane) with a long row
                // (another lane): no specialised variant exists for that (RcRowLen), the general one covers it
                if (NOw > 3 && NOw + Mw > RC_ROW_SUM)
                    par = rc_row_run<RC_ROW_NO, RC_ROW_MAXM, false>(R, sBase, strideD, s, nextStop);
                else switch (NOw) {
                    case 0: par = rc_row_dispatch_m<0>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 1: par = rc_row_dispatch_m<1>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 2: par = rc_row_dispatch_m<2>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 3: par = rc_row_dispatch_m<3>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 4: par = rc_row_dispatch_m<4>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 5: par = rc_row_dispatch_m<5>(R, sBase, strideD, s, nextStop, Mw); break;
                    case 6: par = rc_row_dispatch_m<6>(R, sBase, strideD, s, nextStop, Mw); break;
                    default: par = rc_row_dispatch_m<7>(R, sBase, strideD, s, nextStop, Mw); break;
                }
            }
            cp ^= par;
            rc_cp_wait<0>();
        }
        // ---- updateD contributions of the run ----
        S.dslot[tid] = R.dacc;
        R.dacc = 0.0;
        __syncthreads();
        if (tid < np && !S.patDone[tid]) {
            const int r0 = S.patRowOff[tid], cnt = S.patRowOff[tid + 1] - r0;
            double acc = S.dpat[tid];
            for (int i = 0; i < cnt; ++i) acc = acc + S.dslot[r0 + i];       // non-unit rows carry 0
            S.dpat[tid] = acc;
        }
        if (!(isLast && phase == 0 && s == sShort && s < models[b].nSteps)) { finished = true; break; }
        // ---- canUseShortcut / propagateStateShortcut (Core.hs:80-118) ----
        // a pattern that passes has single-branch states only: one row, |x| = j
        __syncthreads();
        if (R.rowOn && pl.shortcutOK[myPid]) {
            const double* aK = aShort + models[b].aShortOff;
            const int* kid = pl.lastKeyId + (size_t)myPid * P;
            double nrA = 0.0;
            int popIndex = -1;
            for (int k = 0; k < P; ++k) {
                const double ak = aK[kid[k]];
                nrA = nrA + ak;
                if (popIndex < 0 && ak > 0.0) popIndex = k;
            }
            const double popSize = models[b].Nshort[popIndex < 0 ? 0 : popIndex];
            // sum in the reference's live-list order (fillUp order = ascending j for a single branch)
            // chooseCont (nrA + nd) nd = chooseCont (nrA + nd - 1) (nd - 1) * (nrA + nd) / nd  (Utils.hs:213-215 as a recurrence
            // along the row; the factors are the same, multiplied in ascending instead of descending order)
            double add = 0.0, cc = 1.0;
#pragma unroll
            for (int j = 0; j < RC_ROW_MAXM; ++j) {
                if (j < R.M) {
                    const double nd = (double)(j + 1);
                    cc = cc * ((nrA + nd) / nd);
                    add = add + R.b[j] * (2.0 * popSize / nd / cc);
                }
            }
            S.dslot[tid] = add;
        } else S.dslot[tid] = 0.0;
        __syncthreads();
        int ok = 1;
        if (tid < np) {
            if (pl.shortcutOK[S.patId[tid]]) {
                // single-branch states only: one row per live branch (several after a Split)
                double acc = S.dpat[tid];
                for (int r = S.patRowOff[tid]; r < S.patRowOff[tid + 1]; ++r) acc = acc + S.dslot[r];
                S.dpat[tid] = acc;
                S.patDone[tid] = 2;
            } else ok = 0;
        }
        if (__syncthreads_and(ok)) { finished = true; break; }
        // some pattern continues to the end of the grid: park the finished rows
        if (R.rowOn && pl.shortcutOK[myPid]) {
            R.rowOn = false;
            R.unit = false;
            R.M = 0;
        }
        // b of the continuing rows must sit at parity 0 for the next run
        if (cp) {
#pragma unroll
            for (int j = 0; j < RC_ROW_MAXM; ++j)
                if (j < R.M) rc_sts64(R.myB + (unsigned)j * 8u, R.b[j]);
            cp = 0;
        }
        __syncthreads();
    }
    if (!isLast) {
        if (tid < nRows) {
            double* out = bCur + (size_t)b * bStride + myOut;
#pragma unroll
            for (int j = 0; j < RC_ROW_MAXM; ++j)
                if (j < R.M) out[myRefs[j]] = R.b[j];
        }
        if (tid < np) dAcc[(size_t)b * nPat + S.patId[tid]] = S.dpat[tid];
        return;
    }
    if (tid < np) {
        const int pid = S.patId[tid];
        const double d = S.dpat[tid];
        double discFac = 1.0;
        for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? models[b].disc[k] : 1.0);
        pOut[(size_t)b * nPat + pid] = d * models[b].theta * pl.combFac[pid] * discFac;
    }
}

// ------------------------------------------------------------------------------------------------
// RC_RED_SLICES blocks per model, each over a fixed contiguous slice of the patterns (thread-strided partials, then a
// shared-memory tree); the block that finishes last adds the slices' sums in slice order.  The order of every addition is a
// function of the number of patterns only: results are reproducible run to run and do not depend on the batch.  (One block
// per model took 37 us on C3 — on 44 of 148 SMs for the bench batch, and a tenth of the span of a single likelihood.)
__global__ void __launch_bounds__(RC_RED_BLOCK)
rc_reduce_kernel(const double* __restrict__ pOut, const long long* __restrict__ counts,
                 const int* __restrict__ patGlobal, const RcModelDev* __restrict__ models, int nPat,
                 int nPatGlobal, double* __restrict__ partial, double* __restrict__ dProbs,
                 double* __restrict__ slices, unsigned* __restrict__ arrived) {
    __shared__ double sl[RC_RED_BLOCK], sp[RC_RED_BLOCK];
    const int b = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;   // b = position in the plan-sorted model array
    const bool bad = models[b].nSteps < 0;
    const int orig = models[b].pad;                // caller's model index
    const int per = (nPat + RC_RED_SLICES - 1) / RC_RED_SLICES;
    const int i0 = r * per, i1 = min(nPat, i0 + per);
    double ll = 0.0, ps = 0.0;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = i0 + tid; i < i1; i += RC_RED_BLOCK) {
        double p = bad ? nanv : pOut[(size_t)b * nPat + i];
        ll += log(p) * (double)counts[i];
        ps += p;
        if (dProbs) dProbs[(size_t)orig * nPatGlobal + patGlobal[i]] = p;
    }
    sl[tid] = ll; sp[tid] = ps;
    __syncthreads();
    for (int w = RC_RED_BLOCK >> 1; w > 0; w >>= 1) {
        if (tid < w) { sl[tid] += sl[tid + w]; sp[tid] += sp[tid + w]; }
        __syncthreads();
    }
    if (tid == 0) {
        double* mine = slices + ((size_t)b * RC_RED_SLICES + r) * 2;
        __stcg(mine, sl[0]);
        __stcg(mine + 1, sp[0]);
        __threadfence();
        if (atomicAdd(&arrived[b], 1u) == RC_RED_SLICES - 1) {
            __threadfence();
            double tl = 0.0, tp = 0.0;
            const double* all = slices + (size_t)b * RC_RED_SLICES * 2;
            for (int q = 0; q < RC_RED_SLICES; ++q) { tl += __ldcg(all + 2 * q); tp += __ldcg(all + 2 * q + 1); }
            partial[2 * orig] = tl;
            partial[2 * orig + 1] = tp;
            arrived[b] = 0;                        // ready for the next call
        }
    }
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
//
// cudaFuncSetAttribute applies to the CURRENT device and the library may drive several devices from several threads
// (rc_create_multi): the opt-in to large dynamic shared memory is tracked per (kernel, device).
#include <map>
#include <mutex>
#include <utility>
static cudaError_t rc_ensure_smem(const void* kernel, size_t smem, bool carveout = false) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = done[std::make_pair(kernel, dev)];
    if (smem <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (carveout) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
    }
    have = smem;
    return cudaSuccess;
}
cudaError_t rc_launch_tables(const RcModelDev* models, const RcSizeStepDev* sizes, const int* sizeOff, const RcMaskStepDev* masks,
                             const int* maskOff, const double* dts, double* tab, int P, int A1, int rowsMax, int B, cudaStream_t st) {
    if (rowsMax <= 0 || B <= 0) return cudaSuccess;
    dim3 grid((unsigned)rowsMax, (unsigned)B);
    rc_tables_kernel<<<grid, 128, 0, st>>>(models, sizes, sizeOff, masks, maskOff, dts, tab, P, A1);
    return cudaGetLastError();
}

cudaError_t rc_launch_propagate(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs,
                                const double* tab, const double* wpool, double* pOut, int b0, int nB,
                                unsigned char* scratch, size_t scratchBytes, cudaStream_t st) {
    if (pl.nBins <= 0 || nB <= 0) return cudaSuccess;
    size_t smem = rc_propagate_smem_bytes(pl.bigGlobal ? 0 : pl.cap, pl.nnzw, pl.P, rc_row_len(pl.P, pl.A1));
    cudaError_t e = rc_ensure_smem((const void*)rc_propagate_kernel, smem);
    if (e != cudaSuccess) return e;
    if (!pl.bigGlobal) {
        dim3 grid((unsigned)pl.nBins, (unsigned)nB);
        rc_propagate_kernel<<<grid, RC_BLOCK, smem, st>>>(pl, models + b0, sevs, tab, wpool, pOut + (size_t)b0 * pl.nPat, nullptr, 0);
        return cudaGetLastError();
    }
    // global-memory tier: as many models per launch as the scratch buffer holds
    const size_t per = rc_big_scratch_bytes(pl.cap, pl.nnzw);
    const int slice = (int)std::min<size_t>((size_t)nB, scratchBytes / (per * (size_t)pl.nBins));
    if (slice < 1) return cudaErrorMemoryAllocation;
    for (int m0 = 0; m0 < nB; m0 += slice) {
        dim3 grid((unsigned)pl.nBins, (unsigned)std::min(slice, nB - m0));
        rc_propagate_kernel<<<grid, RC_BLOCK, smem, st>>>(pl, models + b0 + m0, sevs, tab, wpool, pOut + (size_t)(b0 + m0) * pl.nPat,
                                                         scratch, (long long)per);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

template <int NNZ, int BLOCK>
static cudaError_t rc_launch_fast_t(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs, const double* tab,
                                    const double* wpool, double* pOut, int nB, cudaStream_t st) {
    const size_t smem = rc_fast_smem_bytes(BLOCK);
    cudaError_t e = rc_ensure_smem((const void*)rc_propagate_fast_kernel<NNZ, BLOCK>, smem, true);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)pl.nFastBins, (unsigned)nB);
    rc_propagate_fast_kernel<NNZ, BLOCK><<<grid, BLOCK, smem, st>>>(pl, models, sevs, tab, wpool, pOut);
    return cudaGetLastError();
}

cudaError_t rc_launch_propagate_fast(const RcPlanDev& pl, const RcModelDev* models, const RcSEvDev* sevs,
                                     const double* tab, const double* wpool, double* pOut, int b0, int nB,
                                     cudaStream_t st) {
    if (pl.nFastBins <= 0 || nB <= 0) return cudaSuccess;
    const RcModelDev* m = models + b0;
    double* out = pOut + (size_t)b0 * pl.nPat;
    const bool n4 = pl.nnzw <= 4;
    switch (pl.fastBlock) {
        case 64: return n4 ? rc_launch_fast_t<4, 64>(pl, m, sevs, tab, wpool, out, nB, st) : rc_launch_fast_t<8, 64>(pl, m, sevs, tab, wpool, out, nB, st);
        case 128: return n4 ? rc_launch_fast_t<4, 128>(pl, m, sevs, tab, wpool, out, nB, st) : rc_launch_fast_t<8, 128>(pl, m, sevs, tab, wpool, out, nB, st);
        default: return n4 ? rc_launch_fast_t<4, 256>(pl, m, sevs, tab, wpool, out, nB, st) : rc_launch_fast_t<8, 256>(pl, m, sevs, tab, wpool, out, nB, st);
    }
}

cudaError_t rc_launch_cum(int nEpochs, int P, int A1, const RcModelDev* models, const RcModelEpochDev* mep, const double* tab,
                          double* cum, int b0, int nB, cudaStream_t st) {
    if (nB <= 0 || nEpochs <= 0) return cudaSuccess;
    dim3 grid((unsigned)nEpochs, (unsigned)nB);
    rc_cum_kernel<<<grid, 32 * P, 0, st>>>(nEpochs, models + b0, mep, tab, cum, P, A1);
    return cudaGetLastError();
}

cudaError_t rc_launch_traj(const RcPlanDev& pl, int e, int nThr, const RcModelDev* models, const RcModelEpochDev* mep,
                           const RcSEvDev* sevs, const double* tab, const double* cum, const double* dts, double* ktab,
                           double* aEnd, double* aShort, int b0, int nB, cudaStream_t st) {
    if (nB <= 0) return cudaSuccess;
    // slices of the epoch's steps: enough threads to fill the device when the batch is small (the count depends on the batch
    // and the plan only, so a captured launch graph stays valid when the event times move)
    const long long thr = (long long)((nThr + 127) / 128) * 128 * nB;
    static const long long fill = std::getenv("RC_TRAJ_FILL") ? std::atoll(std::getenv("RC_TRAJ_FILL")) : RC_TRAJ_FILL;
    static const int zMax = std::getenv("RC_TRAJ_ZMAX") ? std::atoi(std::getenv("RC_TRAJ_ZMAX")) : RC_TRAJ_ZMAX;
    int z = (int)((fill + thr - 1) / thr);
    z = std::max(1, std::min(z, zMax));
    dim3 grid((unsigned)((nThr + 127) / 128), (unsigned)nB, (unsigned)z);
    rc_traj_kernel<<<grid, 128, 0, st>>>(pl, e, models + b0, mep, sevs, tab, cum, dts, ktab, aEnd, aShort);
    return cudaGetLastError();
}

cudaError_t rc_launch_propagate_keyed(const RcPlanDev& pl, int e, int binBase, int nBins, int rowMode, int maxFetch, int maxB, int zSlices, const RcModelDev* models,
                                      const RcModelEpochDev* mep, const RcSEvDev* sevs, const double* ktab,
                                      const double* aShort, const double* wpool, const double* bPrev, double* bCur,
                                      long long bStride, double* dAcc, double* pOut, int b0, int nB, cudaStream_t st) {
    if (nBins <= 0 || nB <= 0) return cudaSuccess;
    const int eArg = e;                 // RC_E_FUSE (pattern kernel only)
    e &= ~RC_E_FUSE;
    const size_t smem = (sizeof(RcKeySmem) + 15) & ~(size_t)15;
    dim3 grid((unsigned)nBins, (unsigned)nB);
    double* out = pOut + (size_t)b0 * pl.nPat;
    double* dac = dAcc + (size_t)b0 * pl.nPat;
    const double* bp = bPrev + (size_t)b0 * bStride;
    double* bc = bCur + (size_t)b0 * bStride;
    if (rowMode >= RC_MODE_PAT) {
        // ring of RC_PDEPTH slots sized for the group's largest fetch list, aliased with the gathered b of the largest bin
        const int ringRows = ((maxFetch + RC_PBLOCK - 1) / RC_PBLOCK) * RC_PBLOCK;
        const size_t psmem = (std::max((size_t)RC_PDEPTH * ringRows * 16, (size_t)maxB * 8) + 15) & ~(size_t)15;
        grid.z = (unsigned)std::max(1, zSlices);      // > 1: bins of the <RC_PAT_MAXT> group (RC_PAT12_NP patterns per thread there) in a <RC_PAT_MAXS> launch
        if (rowMode == RC_MODE_PAT) return rc_launch_pat12(grid, psmem, st, pl, eArg, binBase, ringRows, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
        if (rowMode == RC_MODE_PAT + 1) return rc_launch_pat25(grid, psmem, st, pl, eArg, binBase, ringRows, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
        return rc_launch_pat36(grid, psmem, st, pl, eArg, binBase, ringRows, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
    }
    if (rowMode) {
        const size_t rsmem = (sizeof(RcRowSmem) + 15) & ~(size_t)15;
        if (rowMode == 1) {
            cudaError_t err = rc_ensure_smem((const void*)rc_propagate_rows_kernel<3>, rsmem);
            if (err != cudaSuccess) return err;
            rc_propagate_rows_kernel<3><<<grid, RC_BLOCK, rsmem, st>>>(pl, e, binBase, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
        } else {
            cudaError_t err = rc_ensure_smem((const void*)rc_propagate_rows_kernel<RC_ROW_NO>, rsmem);
            if (err != cudaSuccess) return err;
            rc_propagate_rows_kernel<RC_ROW_NO><<<grid, RC_BLOCK, rsmem, st>>>(pl, e, binBase, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
        }
        return cudaGetLastError();
    }
    if (pl.nnzw <= 4) {
        cudaError_t err = rc_ensure_smem((const void*)rc_propagate_keyed_kernel<4>, smem);
        if (err != cudaSuccess) return err;
        rc_propagate_keyed_kernel<4><<<grid, RC_BLOCK, smem, st>>>(pl, e, binBase, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
    } else {
        cudaError_t err = rc_ensure_smem((const void*)rc_propagate_keyed_kernel<8>, smem);
        if (err != cudaSuccess) return err;
        rc_propagate_keyed_kernel<8><<<grid, RC_BLOCK, smem, st>>>(pl, e, binBase, models + b0, mep, sevs, ktab, aShort, wpool, bp, bc, bStride, dac, out);
    }
    return cudaGetLastError();
}

cudaError_t rc_launch_reduce(const double* pOut, const long long* counts, const int* patGlobal,
                             const RcModelDev* models, int nPat, int nPatGlobal, int B, double* partial,
                             double* dProbs, void* scratch, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    // scratch (rc_reduce_scratch_bytes(B), zeroed once): B arrival counters, then the slices' sums
    unsigned* arrived = static_cast<unsigned*>(scratch);
    double* slices = reinterpret_cast<double*>(static_cast<unsigned char*>(scratch) + (((size_t)B * 4 + 255) & ~(size_t)255));
    rc_reduce_kernel<<<dim3((unsigned)B, RC_RED_SLICES), RC_RED_BLOCK, 0, st>>>(pOut, counts, patGlobal, models, nPat, nPatGlobal, partial, dProbs,
                                                                              slices, arrived);
    return cudaGetLastError();
}

size_t rc_reduce_scratch_bytes(int B) {
    return (((size_t)std::max(B, 1) * 4 + 255) & ~(size_t)255) + (size_t)std::max(B, 1) * RC_RED_SLICES * 2 * sizeof(double);
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
