// rc_pat_kernel.cuh — rc_propagate_pat_kernel<MAXS, NP>; one translation unit per instantiation (rc_pat12.cu, rc_pat25.cu,
// rc_pat36.cu) so that the three compile side by side.
#pragma once
#include "rc_dev.cuh"

// ------------------------------------------------------------------------------------------------
// Pattern mode of the keyed tier: one thread owns NP whole patterns whose live set is a box of at most MAXS states (three
// instantiations by MAXS; the smallest runs NP = RC_PAT12_NP patterns per thread) —
// the row branch A, up to four more branches with m >= 2 (B..E; the shape is a template parameter) and any number of
// branches with m = 1, which only contribute a decay factor.  b lives in registers, so every neighbour of a state is a
// register too and shared memory only carries the {G,E2} pairs:
//     b'(x) = b(x) prod_k G_k(x_k) + sum_{k in A..D} b(x + e_k) E2_k(x_k)                          (Core.hs:259)
// with sum_k m_k pair loads per pattern and step instead of one per state and branch.  Blocks are small (RC_PBLOCK threads,
// several resident per SM) and hold patterns of one shape only (rc_plan.cpp), so the update is fully unrolled without
// predicates; every thread fetches up to four pair rows per grid step into the ring.
//
// NP = 2 (-DRC_PAT12_NP=2, boxes of at most RC_PAT_MAXT states): what a thread does per grid step besides the update — ring
// copies, barrier, slot arithmetic, about 40 instructions — is as much as the update of a 12-state pattern itself; two patterns
// of the bin per thread (lanes t and t + RC_PBLOCK) share it.  Measured slower (rc_types.h), so NP = 1 everywhere by default.
//
// The bulk-copy (cp.async.bulk + mbarrier) variant of the ring measured in round 2 (11.0 against 9.03 ms on C3; evidence in
// profiles/r02_tma_ring_waterfall.sass.txt) lived here until commit 6f245f2 (-DRC_PAT_TMA=1).
#ifndef RC_PAT_GU
#define RC_PAT_GU 4
#endif
#define RC_PFETCH (RC_KROWS / RC_PBLOCK)
// Shared memory of a pattern-mode block (dynamic, sized per bin group by the host): the pair ring, RC_PDEPTH slots of
// `ringRows` pairs — [slot][row]{G, E2}, row 0 = {1, dt} — and, aliased onto it, the b values of the bin gathered before the
// ring starts (pattern after pattern) and the columns of the hand-over.
template <int MAXS, int NP>
struct RcPatRegs {
    double b[NP][MAXS];
    unsigned long long rows64[NP];   // pair rows (bytes): A, then the shaped branches, then the m = 1 branches
    const double* dptr[NP];          // direct bins (no ring): the pattern's first pair in the step row to be read next
    int dPairBase[NP];               // ... and its pair index inside the step row (pair 0 of the row is {1, dt})
    bool direct;                     // block-uniform
    bool anyU;                       // block-uniform: some pattern of the bin has a branch with m = 1
    int pf;                          // block-uniform: pair rows of the bin in units of RC_PBLOCK, rounded up to 1, 2 or RC_PFETCH
    bool on[NP], unit[NP];
    double dacc[NP];
    const double* fp[RC_PFETCH];     // this thread's pairs in the step row to be copied next
    unsigned fmask;                  // bit q: the thread fetches pair row q * RC_PBLOCK + tid
};

template <int MA, int MB, int MC, int MD, int ME, bool HASU, int PF, int MAXS, int NP>
__device__ __forceinline__ void rc_pat_run(RcPatRegs<MAXS, NP>& R, const unsigned sBase, const long long strideD, int& s, const int nextStop,
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
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    if (R.on[p]) {
                        const double2* row = reinterpret_cast<const double2*>(R.dptr[p]);
#pragma unroll
                        for (int i = 0; i < MA; ++i) {
                            const double2 pr = row[i];
                            const double t2 = (i + 1 < MA) ? R.b[p][i + 1] * pr.y : 0.0;
                            R.b[p][i] = fma(R.b[p][i], pr.x, t2);                                      // updateB, Core.hs:259
                        }
                        if (R.unit[p]) R.dacc[p] = fma(*(R.dptr[p] - 2 * (long long)R.dPairBase[p] + 1), R.b[p][0], R.dacc[p]);   // updateD, Core.hs:282-288
                        R.dptr[p] += strideD;
                    }
                }
            }
            return;
        }
    }
    const unsigned myRE = sBase + threadIdx.x * 16u;
    unsigned long long unitRows[NP];      // pair rows of the m = 1 branches, one byte each
    unsigned aA[NP], aB[NP], aC[NP], aD[NP], aE[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        unitRows[p] = R.rows64[p] >> (8 * (1 + NS));
        aA[p] = sBase + ((unsigned)(R.rows64[p]) & 0xFFu) * 16u;
        aB[p] = sBase + ((unsigned)(R.rows64[p] >> 8) & 0xFFu) * 16u;
        aC[p] = sBase + ((unsigned)(R.rows64[p] >> 16) & 0xFFu) * 16u;
        aD[p] = sBase + ((unsigned)(R.rows64[p] >> 24) & 0xFFu) * 16u;
        aE[p] = sBase + ((unsigned)(R.rows64[p] >> 32) & 0xFFu) * 16u;
    }
    for (; s < nextStop; ++s) {
        const unsigned prev = slot == 0 ? (RC_PDEPTH - 1) * RE_SLOT : slot - RE_SLOT;      // slot of step s - 1: free again
        // (PF: the bin has at most PF * RC_PBLOCK pair rows — a loop of its own per ring size, chosen once per block)
#pragma unroll
        for (int q = 0; q < PF; ++q) {
            rc_cp_async16_if(myRE + prev + (unsigned)q * RC_PBLOCK * 16u, R.fp[q], (R.fmask >> q) & 1u);
            R.fp[q] += strideD;
        }
        rc_cp_commit();       // (skipping the copies beyond the group's ring size with uniform branches measured slower: 8.58 -> 9.16 ms)
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (R.on[p]) {
                double gB[MBp], eB[MBp], gC[MCp], eC[MCp], gD[MDp], eD[MDp], gE[MEp], eE[MEp];
                if (MB) {
#pragma unroll
                    for (int j = 0; j < MBp; ++j) rc_lds128(aB[p] + slot + (unsigned)j * 16u, gB[j], eB[j]);
                }
                if (MC) {
#pragma unroll
                    for (int j = 0; j < MCp; ++j) rc_lds128(aC[p] + slot + (unsigned)j * 16u, gC[j], eC[j]);
                }
                if (MD) {
#pragma unroll
                    for (int j = 0; j < MDp; ++j) rc_lds128(aD[p] + slot + (unsigned)j * 16u, gD[j], eD[j]);
                }
                if (ME) {
#pragma unroll
                    for (int j = 0; j < MEp; ++j) rc_lds128(aE[p] + slot + (unsigned)j * 16u, gE[j], eE[j]);
                }
                // branches with m = 1 contribute G_k(1) only.  Every lane multiplies all 7 - NS slots: a spare slot points at pair
                // row 0, whose first half is 1.0 — no predicates, no constants to materialise, a balanced product
                // A bin without any such branch (HASU false: the last epochs of a tree) runs a loop of its own without these loads
                // and products (a uniform branch around them inside one loop measured slower: 10.21 against 10.13 ms; loops of their
                // own for at most 2 / 4 / all such slots, by the most unit branches of a pattern of the bin: 7.65 against 7.61 ms)
                double D1 = 1.0;
                if constexpr (HASU) {
                    double g[7];
#pragma unroll
                    for (int u = 0; u < 7 - NS; ++u) g[u] = rc_lds64(sBase + slot + ((unsigned)(unitRows[p] >> (8 * u)) & 0xFFu) * 16u);
                    D1 = ((g[0] * g[1]) * (g[2] * (NS < 4 ? g[3] : 1.0))) * ((NS < 3 ? g[4] : 1.0) * (NS < 2 ? g[5] : 1.0) * (NS < 1 ? g[6] : 1.0));
                }
#pragma unroll
                for (int i = 0; i < MA; ++i) {
                    double gA, eA;
                    rc_lds128(aA[p] + slot + (unsigned)i * 16u, gA, eA);
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
                                    double t2 = (i + 1 < MA) ? R.b[p][x + SA] * eA : 0.0;      // old values: not yet overwritten
                                    if (MB && j + 1 < MB) t2 = fma(R.b[p][x + SB], eB[j], t2);
                                    if (MC && k + 1 < MC) t2 = fma(R.b[p][x + SC], eC[k], t2);
                                    if (MD && l + 1 < MD) t2 = fma(R.b[p][x + SD], eD[l], t2);
                                    if (ME && m + 1 < ME) t2 = fma(R.b[p][x + 1], eE[m], t2);
                                    R.b[p][x] = fma(R.b[p][x], dE, t2);                        // updateB, Core.hs:259
                                }
                            }
                        }
                    }
                }
                if (R.unit[p]) R.dacc[p] = fma(rc_lds64(sBase + slot + 8u), R.b[p][0], R.dacc[p]);      // updateD, Core.hs:282-288
            }
        }
        rc_cp_wait<RC_PDEPTH - 2>();
        __syncthreads();       // (probe: with a warp-level sync here instead — results undefined — C3 ran 3.8 % faster: the barrier is not the
                               // bound; two grid steps per barrier, the step body instantiated twice per round: 7.59 -> 10.1 ms and a 10-minute build)
        slot = slot + RE_SLOT == RC_PDEPTH * RE_SLOT ? 0u : slot + RE_SLOT;
    }
}

// shapes the planner admits (pat_shape in rc_plan.cpp): MA >= MB >= MC >= MD >= ME, each of B..E absent (0) or >= 2, product
// of the present ones <= MAXS; <RC_PAT_MAXS> also holds the shapes of <RC_PAT_MAXT> (small batches run both groups in one
// launch, see rc_api.cu), <RC_PAT_MAXL> only those the others do not hold
template <int MB, int MC, int MD, int ME, int MAXS, int NP>
__device__ __forceinline__ void rc_pat_dispatch_a(RcPatRegs<MAXS, NP>& R, const unsigned sBase, const long long strideD, int& s, const int nextStop,
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
#ifndef RC_PATT_MINB
#define RC_PATT_MINB (RC_PAT12_NP > 1 ? 6 : 8)      // one pattern per thread: 8 -> 11.0 ms, 10 -> 11.2, 12 -> 11.6 (round 1)
#endif
// A bin holds up to RC_PBLOCK * NP patterns, lane t + p * RC_PBLOCK being pattern p of thread t.  blockIdx.z selects a slice of
// RC_PBLOCK * NP lanes: 0 but for the launches that run the bins of the <RC_PAT_MAXT> group (RC_PAT12_NP patterns per thread
// there) with the <RC_PAT_MAXS> kernel, one pattern per thread (small batches: both groups in one launch).
template <int MAXS, int NP>
__global__ void __launch_bounds__(RC_PBLOCK, MAXS <= RC_PAT_MAXT ? RC_PATT_MINB : MAXS <= RC_PAT_MAXS ? RC_PAT_MINB : RC_PATL_MINB)
rc_propagate_pat_kernel(RcPlanDev pl, int eArg, int binBase, int ringRows, const RcModelDev* __restrict__ models, const RcModelEpochDev* __restrict__ mep,
                        const RcSEvDev* __restrict__ sevs, const double* __restrict__ ktab, const double* __restrict__ aShort,
                        const double* __restrict__ wpool, const double* __restrict__ bPrev, double* __restrict__ bCur,
                        long long bStride, double* __restrict__ dAcc, double* __restrict__ pOut) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    // RC_E_FUSE: the models of this launch execute the last epoch's one grid step right before the shortcut and every pattern
    // passes it, so this launch — the last but one epoch — finishes the patterns itself (hand-over below)
    const int e = eArg & ~RC_E_FUSE;
    const bool fuseLast = (eArg & RC_E_FUSE) != 0;
    // every load of the first level (model header, bin header, epoch record) is issued before the first branch: a warp issues
    // in order, so a branch on one of them would keep the others from starting until it has arrived
    const int mSteps = models[b].nSteps, mMode = models[b].mode, mSShort = models[b].sShort, mMepOff = models[b].mepOff;
    const int binIdx0 = binBase + (int)blockIdx.x;
    const int4 hdrA = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0);
    const int4 hdrB = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0 + 1);
    const int4 hdrC = __ldg(reinterpret_cast<const int4*>(pl.kBinHdr) + (RC_HDR_N / 4) * binIdx0 + 2);      // w: shape of the bin
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
    const int lane0 = (int)blockIdx.z * (RC_PBLOCK * NP) + tid;       // this thread's lanes: lane0 + p * RC_PBLOCK
    if (np <= (int)blockIdx.z * (RC_PBLOCK * NP)) return;            // also the unused header between two bin groups launched together (np = 0)
    const bool isLast = e + 1 == pl.nEpochs;
    const int sShort = mSShort;
    const RcModelEpochDev mE = mep[mMepOff + e];

    RcPatRegs<MAXS, NP> R;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int x = 0; x < MAXS; ++x) R.b[p][x] = 0.0;
        R.on[p] = lane0 + p * RC_PBLOCK < np;
        R.unit[p] = false;
        R.dacc[p] = 0.0;
    }
    const int rowPairs = epd.rowPairs;
    const long long strideD = (long long)rowPairs * 2;
    int s = mE.beg;
    unsigned slot = 0;
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
    // ---- b of the thread's patterns.  (The ring starts afterwards: it lives in the bytes the gather uses; starting it before
    //      the gather, in bytes of its own, keeps the fetch pointers live across the gather and measured slower — spills,
    //      9.96 -> 10.26 ms on C3) ----
    const unsigned shapeB = (unsigned)hdrC.w & 0xFFFFFu;      // MA | MB << 4 | MC << 8 | MD << 12 | ME << 16: one shape per bin
    int n = (int)(shapeB & 0xFu);
#pragma unroll
    for (int q = 1; q < 5; ++q) n *= max(1, (int)((shapeB >> (4 * q)) & 0xFu));
    if (dirIn) {
        // direct hand-over: the bin's b lies [state][lane] in the lane-ordered region (RC_DIR_LANES lanes per state) — every
        // lane reads its own states, a warp 256 contiguous bytes per state, straight into registers; nothing but the header is
        // needed
        const double* src = bPrev + (size_t)b * bStride + pl.dirOff + hdrDir + lane0;
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int x = 0; x < MAXS; ++x)
                if (x < n) R.b[p][x] = __ldcg(src + x * RC_DIR_LANES + p * RC_PBLOCK);
    }
    uint4 la[NP], lb[NP], lc[NP], ld[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        la[p] = make_uint4(0u, 0u, 0u, 0u);
        lb[p] = la[p];
        lc[p] = la[p];
        ld[p] = la[p];
        if (R.on[p]) {
            const uint4* lane = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + (RC_PLANE_N / 4) * (size_t)(lane0 + p * RC_PBLOCK);
            la[p] = lane[0];
            lb[p] = lane[1];
            lc[p] = lane[2];
            ld[p] = lane[3];
        }
    }
    if (!seedIn && !dirIn) {
        // gather through the element records and the transition entries, cooperatively, into shared memory (all lanes of the
        // bin: a slice of a bin — blockIdx.z — gathers the whole bin too)
        const uint2* el = reinterpret_cast<const uint2*>(pl.kElem) + hdrB.w;
        const double* W = e > 0 ? wpool + models[b].wPool + sevs[models[b].sevOff + e - 1].wOff : nullptr;
        const double* bp = bPrev + (size_t)b * bStride;
        if (e > 0 && sevs[models[b].sevOff + e - 1].type == RC_EV_JOIN_D) {
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
    int pid[NP];
    double dp[NP];
    bool anyUnitBranch = false;
    R.direct = hdrB.y == 0;                 // direct bin: no pair rows to fetch; a pattern reads its own key's pairs from the table
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        pid[p] = (int)lb[p].x;
        dp[p] = (R.on[p] && e > 0) ? dAcc[(size_t)b * nPat + pid[p]] : 0.0;
        R.rows64[p] = (unsigned long long)la[p].x | ((unsigned long long)la[p].y << 32);
        R.dPairBase[p] = (int)ld[p].w;
        R.dptr[p] = ktab + (mE.tabBase + (long long)R.dPairBase[p]) * 2;
        anyUnitBranch = anyUnitBranch || (R.on[p] && ((la[p].z >> 20) & 0xFu) > 0);
        R.unit[p] = R.on[p] && (la[p].z & RC_PLANE_UNIT) != 0;       // x = e_A, the lane's first state, is a state of the pattern (updateD)
    }
    if (seedIn) {
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int x = 0; x < MAXS; ++x) R.b[p][x] = x == n - 1 ? 1.0 : 0.0;       // Core.hs:60-66
    } else if (!dirIn) {
        __syncthreads();
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (R.on[p]) {
                const double* mine = bb + lb[p].y;
#pragma unroll
                for (int x = 0; x < MAXS; ++x)
                    if (x < n) R.b[p][x] = mine[x];
            }
        __syncthreads();          // the ring starts in the same bytes
    }
    R.anyU = __syncthreads_or(anyUnitBranch) != 0;
    const int MAw = (int)(shapeB & 0xFu);

    bool finished = false, parked[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) parked[p] = false;
    for (int phase = 0; phase < 2 && !finished; ++phase) {
        int nextStop = mE.end;
#ifdef RC_PROBE_NOLOOP
        nextStop = s;
#endif
        if (isLast && sShort >= s && phase == 0) nextStop = min(nextStop, sShort);
        if (s < nextStop) {
            ring_start();
            rc_cp_wait<RC_PDEPTH - 2>();
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
            rc_cp_wait<0>();
            __syncthreads();
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (!parked[p]) dp[p] = dp[p] + R.dacc[p];                           // updateD contributions of the run
            R.dacc[p] = 0.0;
        }
        if (!(isLast && phase == 0 && s == sShort && s < models[b].nSteps)) { finished = true; break; }
        // ---- canUseShortcut / propagateStateShortcut (Core.hs:80-118): a pattern that passes has single-branch states only ----
        int ok = 1;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (R.on[p]) {
                if (pl.shortcutOK[pid[p]]) {
                    const double* aK = aShort + models[b].aShortOff;
                    const int* kid = pl.lastKeyId + (size_t)pid[p] * P;
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
                    for (int x = 0; x < (RC_ROW_MAXM < MAXS ? RC_ROW_MAXM : MAXS); ++x) {
                        if (x < n) {
                            const double nd = (double)(x + 1);
                            cc = cc * ((nrA + nd) / nd);
                            add = add + R.b[p][x] * (2.0 * popSize / nd / cc);
                        }
                    }
                    dp[p] = dp[p] + add;
                    R.on[p] = false;          // parked: b stays, nothing more is accumulated
                    R.unit[p] = false;
                    parked[p] = true;
                } else ok = 0;
            }
        }
        if (__syncthreads_and(ok)) { finished = true; break; }
    }
    if (!isLast) {
        // ---- hand b and the updateD sums to the next epoch's launch ----
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int lane = lane0 + p * RC_PBLOCK;
            if (lane >= np) continue;
            // (the lane record is read again here instead of being kept through the grid-step loop: twelve registers less in the loop)
            {
                const uint4* lr = reinterpret_cast<const uint4*>(pl.kLane) + (size_t)hdrB.z + (RC_PLANE_N / 4) * (size_t)lane;
                la[p] = __ldcg(lr);
                lb[p] = __ldcg(lr + 1);
                lc[p] = __ldcg(lr + 2);
                ld[p] = __ldcg(lr + 3);
            }
            const unsigned refW[9] = {lb[p].z, lb[p].w, lc[p].x, lc[p].y, lc[p].z, lc[p].w, ld[p].x, ld[p].y, ld[p].z};
            if (la[p].z & RC_PLANE_DIROUT) {
                // direct hand-over: the lane applies the Join itself, new_b[J(x)] += b(x) (Core.hs:172-188), summing in its own
                // [state][lane] column of shared memory (the ring is drained), and writes the pattern's new states where its lane
                // of the next epoch reads them
                const int nOut = (int)((la[p].z >> 24) & 0x3Fu);
                constexpr int CS = RC_PBLOCK * NP;
                double* col = bb + p * RC_PBLOCK + tid;
                for (int y = 0; y < nOut; ++y) col[y * CS] = 0.0;
#pragma unroll
                for (int x = 0; x < MAXS; ++x)
                    if (x < n) {
                        const unsigned y = (refW[x >> 2] >> (8 * (x & 3))) & 0xFFu;      // 0xFF: a place of the bounding box without a state
                        if (y != 0xFFu) {
                            double* t = col + y * CS;
                            *t = *t + R.b[p][x];
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
                        const double* row = ktab + (mL.tabBase + (long long)pl.lastPairBase[pid[p]]) * 2;
                        double dlast = 0.0;
                        for (int s2 = mL.beg; s2 < mL.end; ++s2, row += 2 * (long long)rowPairsL) {
                            for (int y = 0; y < nOut; ++y) {
                                const double2 pr = *reinterpret_cast<const double2*>(row + 2 * y);
                                const double t2 = y + 1 < nOut ? col[(y + 1) * CS] * pr.y : 0.0;
                                col[y * CS] = fma(col[y * CS], pr.x, t2);                                     // updateB, Core.hs:259
                            }
                            dlast = fma(*(row - 2 * (long long)pl.lastPairBase[pid[p]] + 1), col[0], dlast);      // updateD, Core.hs:282-288
                        }
                        dp[p] = dp[p] + dlast;
                    }
                    const double* aK = aShort + models[b].aShortOff;
                    const int* kid = pl.lastKeyId + (size_t)pid[p] * P;
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
                        add = add + col[y * CS] * (2.0 * popSize / nd / cc);
                    }
                    dp[p] = dp[p] + add;
                    double discFac = 1.0;
                    for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid[p] * P + k] ? models[b].disc[k] : 1.0);
                    pOut[(size_t)b * nPat + pid[p]] = dp[p] * models[b].theta * pl.combFac[pid[p]] * discFac;
                    continue;
                }
                double* out = bCur + (size_t)b * bStride + pl.dirOff + la[p].w;
                for (int y = 0; y < nOut; ++y) out[y * RC_DIR_LANES] = col[y * CS];
            } else {
                double* out = bCur + (size_t)b * bStride + la[p].w;
#pragma unroll
                for (int x = 0; x < MAXS; ++x)
                    if (x < n) {
                        const unsigned ref = (refW[x >> 2] >> (8 * (x & 3))) & 0xFFu;
                        if (ref != 0xFFu) out[ref] = R.b[p][x];
                    }
            }
            dAcc[(size_t)b * nPat + pid[p]] = dp[p];
        }
        return;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (lane0 + p * RC_PBLOCK >= np) continue;
        double discFac = 1.0;
        for (int k = 0; k < P; ++k) discFac = discFac * (pl.support[(size_t)pid[p] * P + k] ? models[b].disc[k] : 1.0);
        pOut[(size_t)b * nPat + pid[p]] = dp[p] * models[b].theta * pl.combFac[pid[p]] * discFac;
    }
}
