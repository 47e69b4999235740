// rc_dev.cuh — device helpers shared by the kernel translation units (rc_kernels.cu, rc_pat*.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "rc_kernels.h"

#ifndef RC_PDEPTH
#define RC_PDEPTH 4      // slots of the pattern kernel's pair ring (rc_pat_kernel.cuh; the launcher sizes its shared memory)
#endif

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

// Exclusive prefix sum of the values held by threads tid < n (n <= RC_BLOCK) into out[0..n] (out[n] = total).
// Must be called by all threads of the block; `warpTot` is scratch for RC_BLOCK/32 ints.
__device__ __forceinline__ void rc_block_scan(int v, int n, int* out, int* warpTot) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int x = tid < n ? v : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warpTot[wid] = x;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < wid; ++w) base += warpTot[w];
    if (tid < n) out[tid + 1] = base + x;
    if (tid == 0) out[0] = 0;
    __syncthreads();
}

// new_b = sum over the transition entries [e0, e1) of old_b[src] * w, with several loads in flight
__device__ __forceinline__ double rc_gather(const uint32_t* __restrict__ trEnt, int e0, int e1, const double* __restrict__ bo,
                                            const double* __restrict__ W) {
    double acc = 0.0;
    int q = e0;
    for (; q + 4 <= e1; q += 4) {
        uint32_t en[4];
        double old[4], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = trEnt[q + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            old[u] = bo[en[u] & 0xFFFFu];
            w[u] = (en[u] >> 16) == RC_W_ONE ? 1.0 : W[en[u] >> 16];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = (en[u] >> 16) == RC_W_ONE ? acc + old[u] : acc + old[u] * w[u];
    }
    for (; q < e1; ++q) {
        const uint32_t en = trEnt[q];
        const uint32_t widx = en >> 16;
        const double old = bo[en & 0xFFFFu];
        acc = (widx == RC_W_ONE) ? acc + old : acc + old * W[widx];
    }
    return acc;
}

// The same gather for U states of a bin at once (records of rc_types.h, "gather record"): all entry words first, then all
// carried b values and weights, then the sums in entry order — three dependent loads per batch instead of per state.
// HASW = false: every weight is 1 (a Join, Core.hs:172-189) — no weight loads and half the registers, so twice the states.
template <int U, int E, bool HASW = true>
__device__ __forceinline__ void rc_gather_batch(const uint2 (&g)[U], const uint32_t* __restrict__ trEnt, const double* __restrict__ bp,
                                                const double* __restrict__ W, double (&v)[U]) {
    int cnt[U], maxCnt = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const bool seed = g[u].x == RC_IMG_SEED;                              // makeInitCoalState, Core.hs:66
        cnt[u] = seed ? 0 : (int)(g[u].y >> RC_IMG_CNT_SHIFT);
        v[u] = seed ? 1.0 : 0.0;
        maxCnt = max(maxCnt, cnt[u]);
    }
    for (int base = 0; base < maxCnt; base += E) {
        uint32_t en[U][E];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < E; ++q) en[u][q] = base + q < cnt[u] ? trEnt[(size_t)g[u].x + base + q] : (RC_W_ONE << 16);
        double old[U][E], w[HASW ? U : 1][HASW ? E : 1];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const bool on = base + q < cnt[u];
                const uint32_t widx = en[u][q] >> 16;
                old[u][q] = on ? bp[(size_t)(g[u].y & ((1u << RC_IMG_CNT_SHIFT) - 1u)) + (en[u][q] & 0xFFFFu)] : 0.0;
                if (HASW) w[u][q] = (on && widx != RC_W_ONE) ? W[widx] : 1.0;
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < E; ++q)
                if (base + q < cnt[u]) {
                    if (HASW) v[u] = (en[u][q] >> 16) == RC_W_ONE ? v[u] + old[u][q] : v[u] + old[u][q] * w[u][q];
                    else v[u] = v[u] + old[u][q];
                }
    }
}


// Shared-window address of the dynamic shared memory, made opaque to the compiler: left as a plain cvta it is "cheap to
// rematerialise", and ptxas then re-derives it (S2R SR_CgaCtaId + MOV + LEA) inside the grid-step loops instead of keeping
// it in a register (profiles/r01_pat_kernel_loop.sass.txt).
__device__ __forceinline__ unsigned rc_smem_base(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("mov.u32 %0, %0;" : "+r"(a));
    return a;
}
__device__ __forceinline__ double rc_lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void rc_lds128(unsigned addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
// {G, E2} pair of the keyed kernels.  A 128-bit shared load is served quarter-warp by quarter-warp (about 3 wavefronts per
// warp here) even when the lanes read a handful of distinct rows.  Two 64-bit loads (RC_PAIR_SPLIT=1) were measured 10 %
// slower on C3 (41.1 vs 37.2 ms per 44 models), so the pair stays one LDS.128.
#ifndef RC_PAIR_SPLIT
#define RC_PAIR_SPLIT 0
#endif
template <int SPLIT = RC_PAIR_SPLIT>
__device__ __forceinline__ void rc_lds_pair(unsigned addr, double& a, double& b);
template <>
__device__ __forceinline__ void rc_lds_pair<0>(unsigned addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
template <>
__device__ __forceinline__ void rc_lds_pair<1>(unsigned addr, double& a, double& b) {
    // volatile: ptxas would fuse the two loads back into one LDS.128
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(a) : "r"(addr));
    asm volatile("ld.volatile.shared.f64 %0, [%1+8];" : "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void rc_sts64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void rc_sts128(unsigned addr, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}


// the pairs are read once per block: no L1 allocation
#ifndef RC_CP_MODE
#define RC_CP_MODE "cg"
#endif
__device__ __forceinline__ void rc_cp_async16(unsigned dst, const double* src) {
    asm volatile("cp.async." RC_CP_MODE ".shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// predicated form: no branch around the copy
__device__ __forceinline__ void rc_cp_async16_if(unsigned dst, const double* src, unsigned pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async." RC_CP_MODE ".shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"(pred)
                 : "memory");
}
__device__ __forceinline__ void rc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

