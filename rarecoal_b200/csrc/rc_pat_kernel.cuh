// rc_pat_kernel.cuh — rc_propagate_pat_kernel<MAXS>; one translation unit per instantiation (rc_pat12.cu, rc_pat25.cu,
// rc_pat36.cu) so that the three compile side by side.
#pragma once
#include "rc_dev.cuh"

// ------------------------------------------------------------------------------------------------
// Pattern mode of the keyed tier: one thread owns a whole pattern whose live set is a box of at most MAXS states (two
// instantiations: RC_PAT_MAXS, and RC_PAT_MAXL with fewer resident blocks for the larger boxes) —
// the row branch A, up to four more branches with m >= 2 (B..E; the shape is a template parameter) and any number of
// branches with m = 1, which only contribute a decay factor.  b lives in registers, so every neighbour of a state is a
// register too and shared memory only carries the {G,E2} pairs:
//     b'(x) = b(x) prod_k G_k(x_k) + sum_{k in A..D} b(x + e_k) E2_k(x_k)                          (Core.hs:259)
// with sum_k m_k pair loads per pattern and step instead of one per state and branch.  Blocks are small (RC_PBLOCK threads,
// several resident per SM) and hold patterns of one shape only (rc_plan.cpp), so the update is fully unrolled without
// predicates; every thread fetches up to four pair rows per grid step into the ring.
#ifndef RC_PAT_GU
#define RC_PAT_GU 4
#endif
#ifndef RC_PDEPTH
#define RC_PDEPTH 4
#endif
#define RC_PFETCH (RC_KROWS / RC_PBLOCK)
// RC_PAT_TMA 0 (default): every thread copies up to RC_PFETCH pairs per grid step with cp.async.cg.  1: the ring is filled by
// bulk copies (cp.async.bulk, one per segment of the bin and grid step, issued by the first lanes) that complete on one mbarrier
// per slot.  Measured on C3 (44 models): 11.01 ms against 9.03 ms — UBLKCP takes its operands from uniform registers, so the
// per-lane copies become a serial loop over the active lanes (ELECT, 5 x R2UR, UBLKCP per segment: profiles/
// r02_tma_ring_waterfall.sass.txt), and a bin has 10 - 60 segments; the waiting on the slot's mbarrier adds a dependent ~90 cycles
// per step.  Kept for reproduction: python tools/build_variant.py tma -DRC_PAT_TMA=1.
#ifndef RC_PAT_TMA
#define RC_PAT_TMA 0
#endif
__device__ __forceinline__ void rc_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rc_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rc_mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void rc_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Shared memory of a pattern-mode block (dynamic, sized per bin group by the host): the pair ring, RC_PDEPTH slots of
// `ringRows` pairs — [slot][row]{G, E2}, row 0 = {dt, 0} — and, aliased onto it, the b values of the bin gathered before the
// ring starts (pattern after pattern).
template <int MAXS>
struct RcPatRegs {
    double b[MAXS];
    unsigned long long rows64;    // pair rows (bytes): A, then the shaped branches, then the m = 1 branches
    int nU;                       // number of m = 1 branches
    const double* dptr;           // direct bins (no ring): the pattern's first pair in the step row to be read next, else null
    int dPairBase;                // ... and its pair index inside the step row (pair 0 of the row is {dt, 0})
    bool direct;                  // block-uniform
    bool anyU;                    // block-uniform: some pattern of the bin has a branch with m = 1
    int pf;                       // block-uniform: pair rows of the bin in units of RC_PBLOCK, rounded up to 1, 2 or RC_PFETCH
    bool on, unit;
    double dacc;
#if RC_PAT_TMA
    const double* rowPtr;         // step row of the trajectory table the next fill reads (block-uniform)
    unsigned segSrc, segDL;       // this lane's segment: first pair in a step row; first ring row | pair rows << 16 (0: none)
    int nSeg;                     // segments of the bin (lanes beyond the 64th read theirs from shared memory)
    unsigned stepBytes;           // bytes all segments bring in per grid step
    unsigned mb0, segS;           // shared-memory addresses: the RC_PDEPTH mbarriers, the segment table
    unsigned k;                   // grid steps consumed so far: slot k % RC_PDEPTH, mbarrier phase parity (k / RC_PDEPTH) & 1
#else
    const double* fp[RC_PFETCH];  // this thread's pairs in the step row to be copied next
    unsigned fmask;               // bit q: the thread fetches pair row q * RC_PBLOCK + tid
#endif
};

template <int MAXS>
__device__ __forceinline__ unsigned rc_pat_row(const RcPatRegs<MAXS>& R, int q) { return (unsigned)(R.rows64 >> (8 * q)) & 0xFFu; }

#if RC_PAT_TMA
// one grid step's pairs into ring slot `idx`: thread 0 arms the slot's mbarrier with the byte count, the first lanes issue one
// bulk copy per segment
template <int MAXS>
__device__ __forceinline__ void rc_pat_fill(RcPatRegs<MAXS>& R, const unsigned sBase, const unsigned idx, const unsigned RE_SLOT, const long long strideD) {
    const unsigned bar = R.mb0 + idx * 8u, dst0 = sBase + idx * RE_SLOT;
    if (threadIdx.x == 0) rc_mbar_expect_tx(bar, R.stepBytes);
    if (R.segDL) rc_bulk_g2s(dst0 + (R.segDL & 0xFFFFu) * 16u, R.rowPtr + 2 * (size_t)R.segSrc, (R.segDL >> 16) * 16u, bar);
    if (R.nSeg > RC_PBLOCK) {
        for (int i = RC_PBLOCK + (int)threadIdx.x; i < R.nSeg; i += RC_PBLOCK) {
            unsigned sg, dl;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sg), "=r"(dl) : "r"(R.segS + (unsigned)i * 8u));
            rc_bulk_g2s(dst0 + (dl & 0xFFFFu) * 16u, R.rowPtr + 2 * (size_t)sg, (dl >> 16) * 16u, bar);
        }
    }
    R.rowPtr += strideD;
}
#endif

template <int MA, int MB, int MC, int MD, int ME, bool HASU, int PF, int MAXS>
__device__ __forceinline__ void rc_pat_run(RcPatRegs<MAXS>& R, const unsigned sBase, const long long strideD, int& s, const int nextStop,
                                           unsigned& slot, const unsigned RE_SLOT) {
    constexpr int MBp = MB ? MB : 1, MCp = MC ? MC : 1, MDp = MD ? MD : 1, MEp = ME ? ME : 1;
    constexpr int SA = MBp * MCp * MDp * MEp, SB = MCp * MDp * MEp, SC = MDp * MEp, SD = MEp;
    constexpr int NS = (MB ? 1 : 0) + (MC ? 1 : 0) + (MD ? 1 : 0) + (ME ? 1 : 0);
    static_assert(MA * SA <= MAXS, "shape");
    if constexpr (NS == 0) {
        // Direct bin (every pattern of the group has one live branch and its own trajectory key — the root epoch of a tree):
        // nothing is shared between threads, the pairs come straight from the table, no ring, no barrier.
        if (R.direct) {
            for (; s < nextStop; ++s) {
                if (R.on) {
                    const double2* row = reinterpret_cast<const double2*>(R.dptr);
#pragma unroll
                    for (int i = 0; i < MA; ++i) {
                        const double2 p = row[i];
                        const double t2 = (i + 1 < MA) ? R.b[i + 1] * p.y : 0.0;
                        R.b[i] = fma(R.b[i], p.x, t2);                                      // updateB, Core.hs:259
                    }
                    if (R.unit) R.dacc = fma(*(R.dptr - 2 * (long long)R.dPairBase + 1), R.b[0], R.dacc);   // updateD, Core.hs:282-288
                    R.dptr += strideD;
                }
            }
            return;
        }
    }
#if !RC_PAT_TMA
    const unsigned myRE = sBase + threadIdx.x * 16u;
#endif
    const unsigned long long unitRows = R.rows64 >> (8 * (1 + NS));      // pair rows of the m = 1 branches, one byte each
    const unsigned aA = sBase + rc_pat_row(R, 0) * 16u, aB = sBase + rc_pat_row(R, 1) * 16u, aC = sBase + rc_pat_row(R, 2) * 16u,
                   aD = sBase + rc_pat_row(R, 3) * 16u, aE = sBase + rc_pat_row(R, 4) * 16u;
    for (; s < nextStop; ++s) {
#if RC_PAT_TMA
        rc_pat_fill(R, sBase, (R.k + RC_PDEPTH - 1) & (RC_PDEPTH - 1), RE_SLOT, strideD);      // the slot of step s - 1 is free again
        const unsigned slot = (R.k & (RC_PDEPTH - 1)) * RE_SLOT;
#else
        const unsigned prev = slot == 0 ? (RC_PDEPTH - 1) * RE_SLOT : slot - RE_SLOT;      // slot of step s - 1: free again
        // (PF: the bin has at most PF * RC_PBLOCK pair rows — a loop of its own per ring size, chosen once per block)
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            rc_cp_async16_if(myRE + prev + (unsigned)q * RC_PBLOCK * 16u, R.fp[q], (R.fmask >> q) & 1u);
            R.fp[q] += strideD;
        }
        rc_cp_commit();       // (skipping the copies beyond the group's ring size with uniform branches measured slower: 8.58 -> 9.16 ms)
#endif
        if (R.on) {
            double gB[MBp], eB[MBp], gC[MCp], eC[MCp], gD[MDp], eD[MDp], gE[MEp], eE[MEp];
            if (MB) {
#pragma unroll
                for (int j = 0; j < MBp; ++j) rc_lds128(aB + slot + (unsigned)j * 16u, gB[j], eB[j]);
            }
            if (MC) {
#pragma unroll
                for (int j = 0; j < MCp; ++j) rc_lds128(aC + slot + (unsigned)j * 16u, gC[j], eC[j]);
            }
            if (MD) {
#pragma unroll
                for (int j = 0; j < MDp; ++j) rc_lds128(aD + slot + (unsigned)j * 16u, gD[j], eD[j]);
            }
            if (ME) {
#pragma unroll
                for (int j = 0; j < MEp; ++j) rc_lds128(aE + slot + (unsigned)j * 16u, gE[j], eE[j]);
            }
            // branches with m = 1 contribute G_k(1) only.  Every lane multiplies all 7 - NS slots: a spare slot points at pair
            // row 0, whose first half is 1.0 — no predicates, no constants to materialise, a balanced product
            // A bin without any such branch (HASU false: the last epochs of a tree) runs a loop of its own without these loads
            // and products (a uniform branch around them inside one loop measured slower: 10.21 against 10.13 ms)
            double D1 = 1.0;
            if constexpr (HASU) {
                double g[7];
#pragma unroll
                for (int u = 0; u < 7 - NS; ++u) g[u] = rc_lds64(sBase + slot + ((unsigned)(unitRows >> (8 * u)) & 0xFFu) * 16u);
                D1 = ((g[0] * g[1]) * (g[2] * (NS < 4 ? g[3] : 1.0))) * ((NS < 3 ? g[4] : 1.0) * (NS < 2 ? g[5] : 1.0) * (NS < 1 ? g[6] : 1.0));
            }
#pragma unroll
            for (int i = 0; i < MA; ++i) {
                double gA, eA;
                rc_lds128(aA + slot + (unsigned)i * 16u, gA, eA);
                const double dA = HASU ? D1 * gA : gA;
#pragma unroll
                for (int j = 0; j < MBp; ++j) {
                    const double dB = MB ? dA * gB[j] : dA;
#pragma unroll
                    for (int k = 0; k < MCp; ++k) {
                        const double dC = MC ? dB * gC[k] : dB;
#pragma unroll
                        for (int l = 0; l < MDp; ++l) {
                            const double dD = MD ? dC * gD[l] : dC;
#pragma unroll
                            for (int m = 0; m < MEp; ++m) {
                                const double dE = ME ? dD * gE[m] : dD;
                                const int x = i * SA + j * SB + k * SC + l * SD + m;
                                double t2 = (i + 1 < MA) ? R.b[x + SA] * eA : 0.0;      // old values: not yet overwritten
                                if (MB && j + 1 < MB) t2 = fma(R.b[x + SB], eB[j], t2);
                                if (MC && k + 1 < MC) t2 = fma(R.b[x + SC], eC[k], t2);
                                if (MD && l + 1 < MD) t2 = fma(R.b[x + SD], eD[l], t2);
                                if (ME && m + 1 < ME) t2 = fma(R.b[x + 1], eE[m], t2);
                                R.b[x] = fma(R.b[x], dE, t2);                           // updateB, Core.hs:259
                            }
                        }
                    }
                }
            }
            if (R.unit) R.dacc = fma(rc_lds64(sBase + slot + 8u), R.b[0], R.dacc);      // updateD, Core.hs:282-288
        }
#if RC_PAT_TMA
        ++R.k;
        rc_mbar_wait(R.mb0 + (R.k & (RC_PDEPTH - 1)) * 8u, (R.k / RC_PDEPTH) & 1u);      // the pairs of the next step have landed
        __syncthreads();       // ... and every thread is done with the slot the next fill overwrites
#else
        rc_cp_wait<RC_PDEPTH - 2>();
        __syncthreads();       // (probe: with a warp-level sync here instead — results undefined — C3 ran 3.8 % faster: the barrier is not the
                               // bound; two grid steps per barrier, the step body instantiated twice per round: 7.59 -> 10.1 ms and a 10-minute build)
        slot = slot + RE_SLOT == RC_PDEPTH * RE_SLOT ? 0u : slot + RE_SLOT;
#endif
    }
}

// shapes the planner admits (pat_shape in rc_plan.cpp): MA >= MB >= MC >= MD >= ME, each of B..E absent (0) or >= 2, product
// of the present ones <= MAXS; <RC_PAT_MAXS> also holds the shapes of <RC_PAT_MAXT> (small batches run both groups in one
// launch, see rc_api.cu), <RC_PAT_MAXL> only those the others do not hold
template <int MB, int MC, int MD, int ME, int MAXS>
__device__ __forceinline__ void rc_pat_dispatch_a(RcPatRegs<MAXS>& R, const unsigned sBase, const long long strideD, int& s, const int nextStop,
                                                  unsigned& slot, const unsigned reSlot, const int MA) {
    constexpr int SA = (MB ? MB : 1) * (MC ? MC : 1) * (MD ? MD : 1) * (ME ? ME : 1);
#define RC_PAT_CASE(A)                                                                                         \
    case A:                                                                                                    \
        if constexpr (A >= MB && A * SA <= MAXS && (MAXS <= RC_PAT_MAXS || A * SA > RC_PAT_MAXS)) {            \
            if (R.anyU) {                                                                                      \
                if (R.pf == 1) rc_pat_run<A, MB, MC, MD, ME, true, 1>(R, sBase, strideD, s, nextStop, slot, reSlot);           \
                else if (R.pf == 2) rc_pat_run<A, MB, MC, MD, ME, true, 2>(R, sBase, strideD, s, nextStop, slot, reSlot);      \
                else rc_pat_run<A, MB, MC, MD, ME, true, RC_PFETCH>(R, sBase, strideD, s, nextStop, slot, reSlot);             \
            } else {                                                                                           \
                if (R.pf == 1) rc_pat_run<A, MB, MC, MD, ME, false, 1>(R, sBase, strideD, s, nextStop, slot, reSlot);          \
                else if (R.pf == 2) rc_pat_run<A, MB, MC, MD, ME, false, 2>(R, sBase, strideD, s, nextStop, slot, reSlot);     \
                else rc_pat_run<A, MB, MC, MD, ME, false, RC_PFETCH>(R, sBase, strideD, s, nextStop, slot, reSlot);            \
            }                                                                                                  \
        }                                                                                                      \
        break;
    switch (MA) {
        RC_PAT_CASE(1) RC_PAT_CASE(2) RC_PAT_CASE(3) RC_PAT_CASE(4) RC_PAT_CASE(5)
        RC_PAT_CASE(6) RC_PAT_CASE(7) RC_PAT_CASE(8) RC_PAT_CASE(9) RC_PAT_CASE(10)
        default: break;
    }
#undef RC_PAT_CASE
    static_assert(RC_ROW_MAXM == 10, "cases above");
}

#ifndef RC_PAT_MINB
#define RC_PAT_MINB 6      // measured on C3 with the final launch scheme: 5 -> 10.19 ms, 6 -> 10.07, 7 -> 10.22, 8 -> 10.35
#endif
#ifndef RC_PATL_MINB
#define RC_PATL_MINB 6
#endif
template <int MAXS>
#ifndef RC_PATT_MINB
#define RC_PATT_MINB 8      // measured on C3: 8 -> 11.0 ms, 10 -> 11.2, 12 -> 11.6 (without this instantiation 11.5)
#endif
__global__ void __launch_bounds__(RC_PBLOCK, MAXS <= RC_PAT_MAXT ? RC_PATT_MINB : MAXS <= RC_PAT_MAXS ? RC_PAT_MINB : RC_PATL_MINB)
rc_propagate_pat_kernel(RcPlanDev pl, int eArg, int binBase, int ringRows, const RcModelDev* __restrict__ models, const RcModelEpochDev* __restrict__ mep,
                        const RcSEvDev* __restrict__ sevs, const double* __restrict__ ktab, const double* __restrict__ aShort,
                        const double* __restrict__ wpool, const double* __restrict__ bPrev, double* __restrict__ bCur,
                        long long bStride, double* __restrict__ dAcc, double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    // RC_E_FUSE: the models of this launch execute no grid step in the last epoch (the shortcut applies at its first step) and
    // every pattern passes it, so this launch — the last but one epoch — finishes the patterns itself (hand-over below)
    const int e = eArg & ~RC_E_FUSE;
    const bool fuseLast = (eArg & RC_E_FUSE) != 0;
    // every load of the first level (model header, bin header, epoch record) is issued before the first branch: a warp issues
    // in order, so a branch on one of them would keep the others from starting until it has arrived
    const int mSteps = models[b].nSteps, mMode = models[b].mode, mSShort = models[b].sShort, mMepOff = models[b].mepOff;
    const int binIdx0 = binBase + (int)blockIdx.x;
    const int4 hdrA = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0);
    const int4 hdrB = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0 + 1);
    const int4 hdrC = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0 + 2);      // x first segment, y segments, z bytes per step, w shape
    const int hdrDir = __ldg(pl.kBinHdr + RC_HDR_N * binIdx0 + 12);                                            // the bin in the lane-ordered region
    const RcEpochDev epd = pl.ep[e];
    if (mSteps < 0 || mMode != 1) return;
    const int P = pl.P, nPat = pl.nPat;
    double* bb = reinterpret_cast<double*>(smemRaw);
    const unsigned sBase = rc_smem_base(smemRaw);
    const unsigned reSlot = (unsigned)ringRows * 16u;
    const int np = hdrA.y, nB = RC_HDR_NB(hdrA.w);
    const bool dirIn = (hdrA.w & RC_HDR_DIRIN) != 0;      // b arrives in lane order (the previous epoch's lanes applied the Join)
    const bool seedIn = (hdrA.w & RC_HDR_SEED) != 0;      // first epoch: b = 1 at the pattern itself, the lane's last state
    if (np == 0) return;              // the unused header between two bin groups launched together
    const bool isLast = e + 1 == pl.nEpochs;
    const int sShort = mSShort;
    const RcModelEpochDev mE = mep[mMepOff + e];

    RcPatRegs<MAXS> R;
#pragma unroll
    for (int x = 0; x < MAXS; ++x) R.b[x] = 0.0;
    R.on = tid < np;
    R.unit = false;
    R.dacc = 0.0;
    const int rowPairs = epd.rowPairs;
    const long long strideD = (long long)rowPairs * 2;
    int s = mE.beg;
    unsigned slot = 0;
#if RC_PAT_TMA
    // ---- pair ring: one bulk copy per segment of the bin and grid step, completing on the slot's mbarrier ----
    __shared__ __align__(8) unsigned long long mbarS[RC_PDEPTH];
    __shared__ uint2 segTab[RC_KROWS];
    R.mb0 = rc_smem_base(mbarS);
    R.segS = rc_smem_base(segTab);
    R.nSeg = hdrC.y;
    R.stepBytes = (unsigned)hdrC.z;
    R.segSrc = 0u;
    R.segDL = 0u;
    R.k = 0u;
    {
        const uint2* segs = reinterpret_cast<const uint2*>(pl.kSegs) + hdrC.x;
        if (tid < R.nSeg) { const uint2 sg = segs[tid]; R.segSrc = sg.x; R.segDL = sg.y; }
        for (int i = RC_PBLOCK + tid; i < R.nSeg; i += RC_PBLOCK) segTab[i] = segs[i];
        if (tid == 0) {
#pragma unroll
            for (int u = 0; u < RC_PDEPTH; ++u) rc_mbar_init(R.mb0 + (unsigned)u * 8u, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    // (re)start: the RC_PDEPTH - 1 slots after the last one consumed; the generic-proxy accesses to these bytes (the gathered b)
    // are ordered before the bulk copies by the fence
    auto ring_start = [&]() {
        R.rowPtr = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs) * 2;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int u = 0; u < RC_PDEPTH - 1; ++u) rc_pat_fill(R, sBase, (R.k + (unsigned)u) & (RC_PDEPTH - 1), reSlot, strideD);
    };
    // all fills under way have landed (the loop waited for the first of them already); they count as consumed
    auto ring_drain = [&]() {
#pragma unroll
        for (unsigned u = 1; u < RC_PDEPTH - 1; ++u) rc_mbar_wait(R.mb0 + ((R.k + u) & (RC_PDEPTH - 1)) * 8u, ((R.k + u) / RC_PDEPTH) & 1u);
        R.k += RC_PDEPTH - 1;
    };
#else
    // ---- pair ring: every thread copies the pair rows tid, tid + RC_PBLOCK, ... of the bin ----
    R.fmask = 0u;
    R.pf = hdrB.y <= RC_PBLOCK ? 1 : hdrB.y <= 2 * RC_PBLOCK ? 2 : RC_PFETCH;
    int fOff[RC_PFETCH];
#pragma unroll
    for (int q = 0; q < RC_PFETCH; ++q) {
        const int r = q * RC_PBLOCK + tid;
        fOff[q] = 0;
        if (r < hdrB.y) { fOff[q] = pl.kFetchRows[hdrB.x + r]; R.fmask |= 1u << q; }
    }
    auto ring_start = [&]() {
#pragma unroll
        for (int q = 0; q < RC_PFETCH; ++q) R.fp[q] = ktab + (mE.tabBase + (long long)(s - mE.beg) * rowPairs + fOff[q]) * 2;
#pragma unroll
        for (int u = 0; u < RC_PDEPTH - 1; ++u) {
#pragma unroll
            for (int q = 0; q < RC_PFETCH; ++q) {
                rc_cp_async16_if(sBase + (unsigned)u * reSlot + (unsigned)(q * RC_PBLOCK + tid) * 16u, R.fp[q], (R.fmask >> q) & 1u);
                R.fp[q] += strideD;
            }
            rc_cp_commit();
        }
        slot = 0;
    };
#endif
    // ---- the pattern's record; b of every state of the bin gathered cooperatively into shared memory (the ring starts
    //      afterwards: it lives in the same bytes; starting it before the gather, in bytes of its own, keeps the fetch
    //      pointers live across the gather and measured slower — spills, 9.96 -> 10.26 ms on C3) ----
    const unsigned shapeB = (unsigned)hdrC.w & 0xFFFFFu;      // MA | MB << 4 | MC << 8 | MD << 12 | ME << 16: one shape per bin
    int n = (int)(shapeB & 0xFu);
#pragma unroll
    for (int q = 1; q < 5; ++q) n *= max(1, (int)((shapeB >> (4 * q)) & 0xFu));
    if (dirIn) {
        // direct hand-over: the bin's b lies [state][lane] in the lane-ordered region — every lane reads its own states at
        // stride RC_PBLOCK, a warp 256 contiguous bytes per state, straight into registers; nothing but the header is needed
        const double* src = bPrev + (size_t)b * bStride + pl.dirOff + hdrDir + tid;
#pragma unroll
        for (int x = 0; x < MAXS; ++x)
            if (x < n) R.b[x] = __ldcg(src + x * RC_PBLOCK);
    }
    uint4 la = make_uint4(0u, 0u, 0u, 0u), lb = la, lc = la, ld = la;
    if (R.on) {
        const uint4* lane = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + (RC_PLANE_N / 4) * (size_t)tid;
        la = lane[0];
        lb = lane[1];
        lc = lane[2];
        ld = lane[3];
    }
    {
        const uint2* el = reinterpret_cast<const uint2*>(pl.kElem) + hdrB.w;
        const double* W = e > 0 ? wpool + models[b].wPool + sevs[models[b].sevOff + e - 1].wOff : nullptr;
        const double* bp = bPrev + (size_t)b * bStride;
        if (seedIn || dirIn) {
        } else if (e > 0 && sevs[models[b].sevOff + e - 1].type == RC_EV_JOIN_D) {
            // a Join: all weights are 1, twice as many states per batch
            constexpr int GU = 2 * RC_PAT_GU;
            for (int i = tid; i < nB; i += GU * RC_PBLOCK) {
                uint2 g[GU];
                double v[GU];
#pragma unroll
                for (int u = 0; u < GU; ++u) g[u] = i + u * RC_PBLOCK < nB ? el[i + u * RC_PBLOCK] : make_uint2(0u, 0u);
                rc_gather_batch<GU, 4, false>(g, pl.trEnt, bp, W, v);
#pragma unroll
                for (int u = 0; u < GU; ++u)
                    if (i + u * RC_PBLOCK < nB) bb[i + u * RC_PBLOCK] = v[u];
            }
        } else {
            constexpr int GU = RC_PAT_GU;
            for (int i = tid; i < nB; i += GU * RC_PBLOCK) {
                uint2 g[GU];
                double v[GU];
#pragma unroll
                for (int u = 0; u < GU; ++u) g[u] = i + u * RC_PBLOCK < nB ? el[i + u * RC_PBLOCK] : make_uint2(0u, 0u);
                rc_gather_batch<GU, 4>(g, pl.trEnt, bp, W, v);
#pragma unroll
                for (int u = 0; u < GU; ++u)
                    if (i + u * RC_PBLOCK < nB) bb[i + u * RC_PBLOCK] = v[u];
            }
        }
    }
    const int pid = (int)lb.x;
    double dp = (R.on && e > 0) ? dAcc[(size_t)b * nPat + pid] : 0.0;
    R.rows64 = (unsigned long long)la.x | ((unsigned long long)la.y << 32);
    // direct bin: no pair rows to fetch; the pattern reads its own key's pairs from the table
    R.dPairBase = (int)ld.w;
    R.direct = hdrB.y == 0;
    R.dptr = ktab + (mE.tabBase + (long long)R.dPairBase) * 2;
    if (seedIn) {
#pragma unroll
        for (int x = 0; x < MAXS; ++x) R.b[x] = x == n - 1 ? 1.0 : 0.0;       // Core.hs:60-66
    } else if (!dirIn) {
        __syncthreads();
        if (R.on) {
            const double* mine = bb + lb.y;
#pragma unroll
            for (int x = 0; x < MAXS; ++x)
                if (x < n) R.b[x] = mine[x];
        }
        __syncthreads();          // the ring starts in the same bytes
    }
    R.nU = (int)((la.z >> 20) & 0xFu);
    R.anyU = __syncthreads_or(R.on && R.nU > 0) != 0;

    R.unit = R.on && (la.z & RC_PLANE_UNIT) != 0;            // x = e_A, the lane's first state, is a state of the pattern (updateD)
    const int MAw = (int)(shapeB & 0xFu);

    bool finished = false, parked = false;
    for (int phase = 0; phase < 2 && !finished; ++phase) {
        int nextStop = mE.end;
#ifdef RC_PROBE_NOLOOP
        nextStop = s;
#endif
        if (isLast && sShort >= s && phase == 0) nextStop = min(nextStop, sShort);
        if (s < nextStop) {
            ring_start();
#if RC_PAT_TMA
            rc_mbar_wait(R.mb0 + (R.k & (RC_PDEPTH - 1)) * 8u, (R.k / RC_PDEPTH) & 1u);
#else
            rc_cp_wait<RC_PDEPTH - 2>();
#endif
            __syncthreads();
#define RC_PAT_SHAPE(code, B, C, D, E) \
    case code: rc_pat_dispatch_a<B, C, D, E>(R, sBase, strideD, s, nextStop, slot, reSlot, MAw); break;
            switch (shapeB >> 4) {          // MB | MC << 4 | MD << 8 | ME << 12
                RC_PAT_SHAPE(0x0000, 0, 0, 0, 0) RC_PAT_SHAPE(0x0002, 2, 0, 0, 0) RC_PAT_SHAPE(0x0003, 3, 0, 0, 0)
                RC_PAT_SHAPE(0x0004, 4, 0, 0, 0) RC_PAT_SHAPE(0x0005, 5, 0, 0, 0) RC_PAT_SHAPE(0x0006, 6, 0, 0, 0)
                RC_PAT_SHAPE(0x0022, 2, 2, 0, 0) RC_PAT_SHAPE(0x0023, 3, 2, 0, 0) RC_PAT_SHAPE(0x0024, 4, 2, 0, 0)
                RC_PAT_SHAPE(0x0033, 3, 3, 0, 0) RC_PAT_SHAPE(0x0222, 2, 2, 2, 0) RC_PAT_SHAPE(0x0223, 3, 2, 2, 0)
                RC_PAT_SHAPE(0x2222, 2, 2, 2, 2)
                default: break;             // no other shape has a product <= RC_PAT_MAXL
            }
#undef RC_PAT_SHAPE
            static_assert(RC_PAT_MAXS == 25 && RC_PAT_MAXL == 36, "shape list above");
#if RC_PAT_TMA
            ring_drain();
#else
            rc_cp_wait<0>();
#endif
            __syncthreads();
        }
        if (!parked) dp = dp + R.dacc;                                           // updateD contributions of the run
        R.dacc = 0.0;
        if (!(isLast && phase == 0 && s == sShort && s < models[b].nSteps)) { finished = true; break; }
        // ---- canUseShortcut / propagateStateShortcut (Core.hs:80-118): a pattern that passes has single-branch states only ----
        int ok = 1;
        if (R.on) {
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
                // chooseCont (nrA + nd) nd as a recurrence along the states (Utils.hs:213-215, ascending order of factors)
                double add = 0.0, cc = 1.0;
#pragma unroll
                for (int x = 0; x < RC_ROW_MAXM; ++x) {
                    if (x < n) {
                        const double nd = (double)(x + 1);
                        cc = cc * ((nrA + nd) / nd);
                        add = add + R.b[x] * (2.0 * popSize / nd / cc);
                    }
                }
                dp = dp + add;
                R.on = false;          // parked: b stays, nothing more is accumulated
                R.unit = false;
                parked = true;
            } else ok = 0;
        }
        if (__syncthreads_and(ok)) { finished = true; break; }
    }
    if (tid >= np) return;
    if (!isLast) {
        // ---- hand b and the updateD sums to the next epoch's launch ----
        // (the lane record is read again here instead of being kept through the grid-step loop: twelve registers less in the loop)
        {
            const uint4* lane = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + (RC_PLANE_N / 4) * (size_t)tid;
            la = __ldcg(lane);
            lb = __ldcg(lane + 1);
            lc = __ldcg(lane + 2);
            ld = __ldcg(lane + 3);
        }
        const unsigned refW[9] = {lb.z, lb.w, lc.x, lc.y, lc.z, lc.w, ld.x, ld.y, ld.z};
        if (la.z & RC_PLANE_DIROUT) {
            // direct hand-over: the lane applies the Join itself, new_b[J(x)] += b(x) (Core.hs:172-188), summing in its own
            // [state][lane] column of shared memory (the ring is drained), and writes the pattern's new states where its lane
            // of the next epoch reads them
            const int nOut = (int)((la.z >> 24) & 0x3Fu);
            double* col = bb + tid;
            for (int y = 0; y < nOut; ++y) col[y * RC_PBLOCK] = 0.0;
#pragma unroll
            for (int x = 0; x < MAXS; ++x)
                if (x < n) {
                    const unsigned y = (refW[x >> 2] >> (8 * (x & 3))) & 0xFFu;      // 0xFF: a place of the bounding box without a state
                    if (y != 0xFFu) {
                        double* t = col + y * RC_PBLOCK;
                        *t = *t + R.b[x];
                    }
                }
            if (fuseLast) {
                // The last epoch would run its one grid step (the step at which the last Join fires), then the shortcut (Core.hs:
                // 80-118) on these values — its patterns have one live branch, state y holding y + 1 lineages — and end with the
                // getProb epilogue (Core.hs:49-54): done here, the last epoch's launches (one block set-up per bin and model for a
                // handful of flops) are not issued at all
                {
                    const RcModelEpochDev mL = mep[mMepOff + e + 1];
                    const int rowPairsL = pl.ep[e + 1].rowPairs;
                    const double* row = ktab + (mL.tabBase + (long long)pl.lastPairBase[pid]) * 2;
                    double dlast = 0.0;
                    for (int s2 = mL.beg; s2 < mL.end; ++s2, row += 2 * (long long)rowPairsL) {
                        for (int y = 0; y < nOut; ++y) {
                            const double2 pr = *reinterpret_cast<const double2*>(row + 2 * y);
                            const double t2 = y + 1 < nOut ? col[(y + 1) * RC_PBLOCK] * pr.y : 0.0;
                            col[y * RC_PBLOCK] = fma(col[y * RC_PBLOCK], pr.x, t2);                       // updateB, Core.hs:259
                        }
                        dlast = fma(*(row - 2 * (long long)pl.lastPairBase[pid] + 1), col[0], dlast);     // updateD, Core.hs:282-288
                    }
                    dp = dp + dlast;
                }
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
                double add = 0.0, cc = 1.0;
                for (int y = 0; y < nOut; ++y) {
                    const double nd = (double)(y + 1);
                    cc = cc * ((nrA + nd) / nd);
                    add = add + col[y * RC_PBLOCK] * (2.0 * popSize / nd / cc);
                }
                dp = dp + add;
                double discFac = 1.0;
                for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? models[b].disc[k] : 1.0);
                pOut[(size_t)b * nPat + pid] = dp * models[b].theta * pl.combFac[pid] * discFac;
                return;
            }
            double* out = bCur + (size_t)b * bStride + pl.dirOff + la.w;
            for (int y = 0; y < nOut; ++y) out[y * RC_PBLOCK] = col[y * RC_PBLOCK];
        } else {
            double* out = bCur + (size_t)b * bStride + la.w;
#pragma unroll
            for (int x = 0; x < MAXS; ++x)
                if (x < n) {
                    const unsigned ref = (refW[x >> 2] >> (8 * (x & 3))) & 0xFFu;
                    if (ref != 0xFFu) out[ref] = R.b[x];
                }
        }
        dAcc[(size_t)b * nPat + pid] = dp;
        return;
    }
    double discFac = 1.0;
    for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid * P + k] ? models[b].disc[k] : 1.0);
    pOut[(size_t)b * nPat + pid] = dp * models[b].theta * pl.combFac[pid] * discFac;
}

