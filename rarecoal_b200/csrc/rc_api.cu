// rc_api.cu — context, model resolution and the C ABI of librarecoal_b200.so (include/rarecoal_b200.h).
//
// Host side of the path, mirroring the reference:
//   validateModel            StateSpace.hs:163-197        -> validate_model()
//   makeInitModelState       Core.hs:290-298 (stable sort) -> resolve_model()
//   singleStep event firing  Core.hs:120-135 (strict t < nextTime; events do not move the clock)
//                                                          -> events are resolved to grid-step indices
//   computeFrequencySpectrum MaxUtils.hs:93-119            -> rc_eval (one launch for all patterns)
//   computeLogLikelihoodFromSpec MaxUtils.hs:85-91         -> rc_reduce_kernel + rc_finalize_kernel
// There is no CPU fallback: without a usable CUDA device rc_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/rarecoal_b200.h"
#include "rc_kernels.h"
#include "rc_multi.h"
#include "rc_nccl.h"
#include "rc_plan.h"

namespace {

#define CK(call)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return RC_ERR_CUDA;                                                             \
        }                                                                                   \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct PlanEntry {
    rc::PlanHost h;
    RcPlanDev d;
    std::vector<void*> allocs;
    struct Pending { const void* src; size_t bytes, off; const void** out; };
    std::vector<Pending> pending;        // arrays of the plan waiting for finish_upload (one device arena)
    size_t arenaBytes = 0;
    std::vector<RcEpochDev> eps;
    bool allShortcut = false;
    size_t allocBytes = 0, bytes = 0;    // device bytes uploaded so far / accounted in the cache
    long long lastUse = 0;               // serial of the last eval call that used the plan (LRU)
    void release() {
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        allocBytes = 0;
        pending.clear();
        arenaBytes = 0;
    }
};

// piecewise-constant sizes per branch and frozen masks per model, by firing step (layouts of rc_types.h)
struct SizeStep {
    int sBegin;
    double N;
};
struct MaskStep {
    int sBegin;
    unsigned mask;
};

struct ModelHost {
    int status = RC_OK;
    std::string err;
    std::vector<std::vector<SizeStep>> sizes;   // [P]
    std::vector<MaskStep> masks;
    double size_at(int k, int s) const {
        const std::vector<SizeStep>& v = sizes[(size_t)k];
        size_t i = 0, lo = 0, hi = v.size() - 1;
        while (lo < hi) { size_t mid = (lo + hi + 1) / 2; if (v[mid].sBegin <= s) lo = mid; else hi = mid - 1; }
        i = lo;
        return v[i].N;
    }
    unsigned mask_at(int s) const {
        size_t i = 0;
        while (i + 1 < masks.size() && masks[i + 1].sBegin <= s) ++i;
        return masks[i].mask;
    }
    std::vector<rc::StructEv> sev;
    std::vector<int> sevStep;
    std::vector<double> sevM;
    std::vector<rc::GapInfo> gaps;
    bool hasSplit = false;
    int sShort = -1;
    std::string sig;
    PlanEntry* plan = nullptr;
};

}  // namespace

struct rc_statespace {
    rc::StateSpace* ss;
};

#ifndef RC_NCHAIN
#define RC_NCHAIN 2
#endif
struct rc_ctx {
    int device = 0;
    int smemOptin = 0;
    int smCount = 0;
    std::string err;
    // problem
    int P = 0, maxAf = 0, nPatGlobal = 0;
    std::vector<int32_t> patG, nVec;
    std::vector<int64_t> cntG;
    int64_t totalSites = 0;
    long long sumCounts = 0;
    bool compErr = false;
    std::string compErrMsg;
    int rank = 0, world = 1;
    // shard
    int nPat = 0;
    std::vector<int32_t> patS;
    std::vector<int> patGlobalIdx;
    DevBuf dAInit, dCombFac, dSupport, dPatGlobal, dCounts;
    // grid
    std::vector<double> times, dts;
    DevBuf dDts;
    // plans
    std::map<std::string, PlanEntry*> plans;
    size_t planBytes = 0;                // device memory held by cached plans
    long long evalSerial = 0;            // incremented per eval_chunk
    int optPlanCacheMax = 64;            // "plan_cache_max"
    double optPlanCacheMB = 49152.0;     // "plan_cache_mb"
    // counters since rc_create (rc_stats reports the per-call deltas)
    long long planMisses = 0, planEvictions = 0, planMissesAtCall = 0, planEvictionsAtCall = 0;
    double planBuildMs = 0.0, planBuildMsAtCall = 0.0;
    // eval scratch
    DevBuf dBlob, dTab, dPOut, dPartial, dLogl, dProbs, dKTab, dAEnd, dAShort, dBCarry, dDAcc, dBigScratch, dCum, dRed;
    int optBigSmemKB = 0;             // > 0: shared-memory budget of the big tier (tests: forces the global-memory mode)
    // options (rc_set_option)
    int optKeyed = 1;                 // fast-tier patterns through the keyed kernels when the table fits
    int optPattern = 1;               // boxes of few states run one pattern per thread (else rows / states)
    int optFuseLast = 1;              // a last epoch without grid steps is finished by the launches of the last but one epoch
    int optPatternEmbed = 0;          // ... and live sets that are not boxes on their bounding boxes (rc_plan.cpp, pat_box)
    double optKtabModelMB = 1024.0;   // largest trajectory table of one model
    double optKtabTotalMB = 12288.0;  // largest trajectory table memory of one call
    PinBuf hBlob, hOut, hPlanStage;
    cudaStream_t stream = nullptr;
    cudaEvent_t evt[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // side stream of the trajectory kernels: rc_traj_kernel(e) depends only on rc_traj_kernel(e-1), so the chain runs beside
    // the propagation launches, each of which waits for the trajectory of its own epoch
    cudaStream_t trajStream = nullptr;
    cudaEvent_t evFork = nullptr;
    // the bin groups of one epoch touch disjoint patterns: their launches run side by side (group streams), so the few long
    // blocks of the large-pattern group do not leave the GPU idle in their tail; joined before the next epoch
    // ... and a large batch is cut into RC_NCHAIN model ranges whose epoch chains run side by side, so that the tail of one
    // range's launch is filled by the other's
    cudaStream_t chainStream[RC_NCHAIN] = {};                 // [0] unused: chain 0 runs on the caller's stream
    cudaStream_t grpStream[RC_NCHAIN][RC_NGROUP - 1] = {};
    cudaEvent_t evEpoch[RC_NCHAIN] = {}, evGrp[RC_NCHAIN][RC_NGROUP - 1] = {}, evChain[RC_NCHAIN] = {};
    std::vector<cudaEvent_t> evTraj;
    bool trajTimed = false;
    int chunkHint = 256;              // models per chunk; adapted to the trajectory-table size after every call
    bool statsPending = false;
    bool blobInFlight = false;
    bool firstChunk = true;
    // launch graphs by launch signature ("graphs" option); dropped whenever a plan or a buffer they point to goes away
    int optGraphs = 1;
    std::map<std::string, cudaGraphExec_t> graphs;
    std::map<std::string, int> graphLaunches;
    std::set<std::string> graphSeen;
    void clear_graphs() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
        graphLaunches.clear();
        graphSeen.clear();
    }
    rc_stats stats;
    rc_ctx* child = nullptr;
    // multi-GPU: this context is one rank of an NCCL communicator (rc_comm_init), or the front of a single-process
    // multi-device context (rc_create_multi, rc_multi.cpp) whose calls are forwarded
    rc::NcclComm comm = nullptr;
    int commRanks = 1, commRank = 0;
    RcMulti* multi = nullptr;

    void clear_plans() {
        clear_graphs();
        for (auto& kv : plans) {
            kv.second->release();
            delete kv.second;
        }
        plans.clear();
        planBytes = 0;
    }
};

namespace {

// ---- validateModel (StateSpace.hs:163-197) -------------------------------------------------------
int validate_model(int P, const double* disc, const std::vector<rc_event>& sorted, std::string& err) {
    for (const rc_event& e : sorted)
        if (e.t < 0) { err = "Negative event times"; return RC_ERR_MODEL; }
    for (int k = 0; k < P; ++k)
        if (disc[k] <= 0 || disc[k] > 1) { err = "illegal discovery Rate"; return RC_ERR_MODEL; }
    const int n = (int)sorted.size();
    auto bad_idx = [&](int i) { return i < 0 || i >= P; };
    char buf[160];
    for (int i = 0; i < n; ++i) {
        const rc_event& e = sorted[i];
        if (e.type == RC_EV_JOIN) {
            if (bad_idx(e.k) || bad_idx(e.l)) { err = "illegal branch indices in event Join"; return RC_ERR_MODEL; }
            bool reuse = false;   // branch l must not be referenced by any later event
            for (int j = i + 1; j < n && !reuse; ++j) {
                const rc_event& r = sorted[j];
                reuse = r.k == e.l || ((r.type == RC_EV_JOIN || r.type == RC_EV_SPLIT) && r.l == e.l);
            }
            if (e.k == e.l || reuse) {
                std::snprintf(buf, sizeof buf, "Illegal join from %d to %d at time %g", e.l, e.k, e.t);
                err = buf;
                return RC_ERR_MODEL;
            }
        } else if (e.type == RC_EV_POPSIZE) {
            if (bad_idx(e.k)) { err = "illegal branch indices in event SetPopSize"; return RC_ERR_MODEL; }
            if (e.v <= 0) { std::snprintf(buf, sizeof buf, "Illegal population size: %g", e.v); err = buf; return RC_ERR_MODEL; }
        } else if (e.type == RC_EV_SPLIT) {
            if (bad_idx(e.k) || bad_idx(e.l)) { err = "illegal branch indices in event Split"; return RC_ERR_MODEL; }
            if (e.v < 0.0 || e.v > 1.0) { std::snprintf(buf, sizeof buf, "Illegal split rate%g", e.v); err = buf; return RC_ERR_MODEL; }
            // The reference does not check k /= l for Split; its state update would then leave the
            // state space (Core.hs:213) and abort in stateToId.  Reported as a model error here.
            if (e.k == e.l) { err = "Illegal split of a branch into itself"; return RC_ERR_MODEL; }
        } else if (e.type == RC_EV_FREEZE) {
            if (bad_idx(e.k)) { err = "illegal branch indices in event SetFreeze"; return RC_ERR_MODEL; }
        } else if (e.type == RC_EV_GROWTH) {
            if (bad_idx(e.k)) { err = "illegal branch indices in event SetGrowthRate"; return RC_ERR_MODEL; }
            if (!std::isfinite(e.v)) { err = "Illegal growth rate"; return RC_ERR_MODEL; }
        } else {
            err = "unknown event type";
            return RC_ERR_MODEL;
        }
    }
    return RC_OK;
}

std::vector<rc_event> sort_events(const rc_event* ev, int n) {
    std::vector<rc_event> v(ev, ev + n);
    // sortBy time is stable (Core.hs:292-294): template order survives for equal times
    std::stable_sort(v.begin(), v.end(), [](const rc_event& a, const rc_event& b) { return a.t < b.t; });
    return v;
}

// SetGrowthRate is sugar (include/rarecoal_b200.h, RC_EV_GROWTH): from its time t0 on, the size of branch k follows
// N(t) = N(t0) * exp(-r (t - t0)) (t runs backwards in time, so r > 0 is a population that grew towards the present) until
// the next SetPopSize / SetGrowthRate on k or the Join that removes k.  It is rewritten into the reference's own events —
// one `SetPopSize k N(t_s)` at every grid point t0 < t_s < tStop (so grid step s+1 runs with the size at its recent end)
// — before anything else looks at the model; N(t0) is the size set by earlier events (1.0 initially, Core.hs:296).
// `sorted` must be stably sorted by time; so is the result (generated events after the explicit ones of the same time).
std::vector<rc_event> desugar_growth(const std::vector<double>& times, const std::vector<rc_event>& sorted) {
    bool any = false;
    for (const rc_event& e : sorted) any = any || e.type == RC_EV_GROWTH;
    if (!any) return sorted;
    struct G { bool on = false; double t0 = 0, n0 = 1, r = 0; };
    std::vector<G> g(RC_MAX_P);
    std::vector<double> size(RC_MAX_P, 1.0);
    std::vector<rc_event> out, gen;
    auto valid = [](int k) { return k >= 0 && k < RC_MAX_P; };
    auto at = [](const G& s, double t) { return s.n0 * std::exp(-(s.r * (t - s.t0))); };
    // per-step sizes of a growth phase of branch k that ends (exclusively) at tStop
    auto flush = [&](int k, double tStop) {
        G& s = g[(size_t)k];
        if (!s.on) return;
        auto it = std::upper_bound(times.begin(), times.end(), s.t0);
        for (; it != times.end() && *it < tStop; ++it) {
            rc_event e;
            e.t = *it; e.type = RC_EV_POPSIZE; e.k = k; e.l = 0; e.v = at(s, *it);
            gen.push_back(e);
        }
        size[(size_t)k] = at(s, tStop);
        s.on = false;
    };
    for (const rc_event& e : sorted) {
        if (e.type == RC_EV_GROWTH) {
            if (!valid(e.k)) continue;                       // rejected by validateModel
            flush(e.k, e.t);                                 // N is continuous across a change of rate
            G& s = g[(size_t)e.k];
            if (e.v == 0.0) {                                // rate 0: a plain size from here on
                rc_event p;
                p.t = e.t; p.type = RC_EV_POPSIZE; p.k = e.k; p.l = 0; p.v = size[(size_t)e.k];
                gen.push_back(p);
            } else { s.on = true; s.t0 = e.t; s.n0 = size[(size_t)e.k]; s.r = e.v; }
            continue;
        }
        if (e.type == RC_EV_POPSIZE && valid(e.k)) { flush(e.k, e.t); size[(size_t)e.k] = e.v; }
        if (e.type == RC_EV_JOIN && valid(e.l)) flush(e.l, e.t);   // the joined-away branch ends here
        out.push_back(e);
    }
    const double inf = std::numeric_limits<double>::infinity();
    for (int k = 0; k < RC_MAX_P; ++k) flush(k, inf);
    out.insert(out.end(), gen.begin(), gen.end());
    std::stable_sort(out.begin(), out.end(), [](const rc_event& a, const rc_event& b) { return a.t < b.t; });
    return out;
}

// GHC's (^): square-and-multiply with the accumulator order of GHC.Real.powImpl / powImplAcc, so the
// binomial factor m^s (1-m)^(x_l-s) (Core.hs:216) rounds as in the reference.
double ghc_pow(double x, int n) {
    if (n == 0) return 1.0;
    while ((n & 1) == 0) { x = x * x; n >>= 1; }
    if (n == 1) return x;
    double acc = x;
    x = x * x;
    n >>= 1;
    while (true) {
        if (n & 1) {
            if (n == 1) return x * acc;
            acc = x * acc;
        }
        x = x * x;
        n >>= 1;
    }
}

// Resolves one ModelSpec against the grid: event firing steps, (popSize, freeze) segments,
// structural events and the structural signature.
void resolve_model(const rc_ctx* ctx, const rc_event* ev, int nEv, const double* disc, int noShortcut, ModelHost& mh) {
    const int P = ctx->P, T = (int)ctx->times.size();
    std::vector<rc_event> sorted = sort_events(ev, nEv);
    mh.status = validate_model(P, disc, sorted, mh.err);
    if (mh.status != RC_OK) return;
    sorted = desugar_growth(ctx->times, sorted);
    mh.sizes.assign((size_t)P, std::vector<SizeStep>(1, SizeStep{0, 1.0}));     // Core.hs:296-297
    mh.masks.assign(1, MaskStep{0, 0u});
    bool allFire = true;
    int lastFire = -1;
    for (const rc_event& e : sorted) {
        // fires in the first step whose nextTime > t (Core.hs:125)
        int f = (int)(std::upper_bound(ctx->times.begin(), ctx->times.end(), e.t) - ctx->times.begin());
        if (f >= T) { allFire = false; continue; }
        lastFire = f;
        if (e.type == RC_EV_POPSIZE) {
            // the last SetPopSize that fires before a step's update is the size of that step
            std::vector<SizeStep>& v = mh.sizes[(size_t)e.k];
            if (v.back().sBegin == f) v.back().N = e.v;
            else v.push_back(SizeStep{f, e.v});
        } else if (e.type == RC_EV_FREEZE) {
            if (mh.masks.back().sBegin != f) mh.masks.push_back(MaskStep{f, mh.masks.back().mask});
            unsigned& m = mh.masks.back().mask;
            if (e.v != 0.0) m |= (1u << e.k);
            else m &= ~(1u << e.k);
        } else {
            rc::StructEv se;
            se.type = e.type;
            se.k = e.k;
            se.l = e.l;
            se.mclass = 2;
            if (e.type == RC_EV_SPLIT) {
                mh.hasSplit = true;
                se.mclass = e.v == 0.0 ? 0 : (e.v == 1.0 ? 1 : 2);
            }
            mh.sev.push_back(se);
            mh.sevStep.push_back(f);
            mh.sevM.push_back(e.v);
        }
    }
    // the shortcut test needs an empty event queue (Core.hs:88-89)
    mh.sShort = -1;
    if (!noShortcut && allFire) {
        int s = lastFire + 1;
        if (s < T) mh.sShort = s;
    }
    // signature
    std::string& sig = mh.sig;
    sig.reserve(8 * mh.sev.size() + 8);
    for (const rc::StructEv& se : mh.sev) {
        sig.push_back((char)se.type);
        sig.push_back((char)se.k);
        sig.push_back((char)se.l);
        sig.push_back((char)se.mclass);
    }
    if (mh.hasSplit) {
        mh.gaps.resize(mh.sev.size());
        int prev = 0;
        for (size_t g = 0; g < mh.sev.size(); ++g) {
            int cnt = std::min(mh.sevStep[g] - prev, ctx->maxAf);
            sig.push_back('|');
            sig.push_back((char)cnt);
            for (int s = 0; s < cnt; ++s) {
                unsigned m = mh.mask_at(prev + s);
                mh.gaps[g].masks.push_back(m);
                sig.append(reinterpret_cast<const char*>(&m), sizeof m);
            }
            prev = mh.sevStep[g];
        }
    }
}

// Arrays of a plan are laid out in one device allocation: upload() reserves a 256-byte aligned range, finish_upload()
// allocates the arena, gathers the arrays into a pinned staging buffer with a few host threads and sends them as one copy
// (a C3 plan is 340 MB: array-by-array copies from pageable memory took 0.1-0.4 s, a third of the planner's own time).
template <typename T>
int upload(rc_ctx*, PlanEntry* pe, const std::vector<T>& v, const T** out) {
    PlanEntry::Pending q;
    q.src = v.data();
    q.bytes = v.size() * sizeof(T);
    q.off = pe->arenaBytes;
    q.out = reinterpret_cast<const void**>(out);
    pe->arenaBytes += (std::max<size_t>(q.bytes, 1) + 255) & ~(size_t)255;
    pe->pending.push_back(q);
    return RC_OK;
}

int finish_upload(rc_ctx* ctx, PlanEntry* pe) {
    void* dev = nullptr;
    const size_t total = std::max<size_t>(pe->arenaBytes, 256);
    CK(cudaMalloc(&dev, total));
    pe->allocs.push_back(dev);
    pe->allocBytes += total;
    for (const auto& q : pe->pending) *q.out = static_cast<const char*>(dev) + q.off;
    if (ctx->hPlanStage.ensure(total) == cudaSuccess) {
        char* stage = static_cast<char*>(ctx->hPlanStage.p);
        // tasks of at most 4 MB
        struct Task { const char* src; char* dst; size_t n; };
        std::vector<Task> tasks;
        for (const auto& q : pe->pending)
            for (size_t o = 0; o < q.bytes; o += (size_t)4 << 20)
                tasks.push_back({static_cast<const char*>(q.src) + o, stage + q.off + o, std::min(q.bytes - o, (size_t)4 << 20)});
        std::atomic<size_t> next(0);
        auto worker = [&]() {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= tasks.size()) break;
                std::memcpy(tasks[i].dst, tasks[i].src, tasks[i].n);
            }
        };
        int nT = (int)std::thread::hardware_concurrency();
        nT = std::max(1, std::min(nT, 8));
        if (total < ((size_t)8 << 20)) nT = 1;
        std::vector<std::thread> pool;
        for (int t = 1; t < nT; ++t) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
        CK(cudaMemcpy(dev, stage, total, cudaMemcpyHostToDevice));
    } else {
        cudaGetLastError();          // no pinned memory to be had: array by array from pageable memory
        for (const auto& q : pe->pending)
            if (q.bytes) CK(cudaMemcpy(static_cast<char*>(dev) + q.off, q.src, q.bytes, cudaMemcpyHostToDevice));
    }
    pe->pending.clear();
    pe->eps.clear();
    return RC_OK;
}

// Device copy of a built plan.  On failure the caller releases what was uploaded so far.
int upload_plan(rc_ctx* ctx, PlanEntry* pe) {
    RcPlanDev& d = pe->d;
    std::memset(&d, 0, sizeof d);
    const rc::PlanHost& h = pe->h;
    d.P = h.P; d.maxAf = h.maxAf; d.A1 = h.maxAf + 1; d.nPat = h.nPat; d.nEpochs = h.nEpochs; d.nnzw = h.nnzw;
    d.nBins = h.nBins; d.nBinsKB = h.nBinsKB; d.cap = h.cap; d.bigGlobal = h.bigGlobal; d.zeroRow = h.zeroRow; d.nFastBins = h.nFastBins; d.fastBlock = h.fastBlock;
    int rc;
    if ((rc = upload(ctx, pe, h.fBinPatOff, &d.fBinPatOff))) return rc;
    if ((rc = upload(ctx, pe, h.fBinPats, &d.fBinPats))) return rc;
    if ((rc = upload(ctx, pe, h.maxX, &d.maxX))) return rc;
    {
        std::vector<RcEpochDev>& eps = pe->eps;
        eps.assign(h.ep.size(), RcEpochDev());
        for (size_t e = 0; e < h.ep.size(); ++e) {
            eps[e].nKeys = h.ep[e].nKeys; eps[e].nThr = h.ep[e].nThr; eps[e].rowPairs = h.ep[e].rowPairs;
            eps[e].keyOff = h.ep[e].keyOff; eps[e].thrOff = h.ep[e].thrOff; eps[e].elBase = h.ep[e].elBase;
        }
        if ((rc = upload(ctx, pe, eps, &d.ep))) return rc;
    }
    d.totalKeys = h.totalKeys;
    d.nKeysLast = h.ep.empty() ? 0 : h.ep.back().nKeys;
    if ((rc = upload(ctx, pe, h.keyBranch, &d.keyBranch))) return rc;
    if ((rc = upload(ctx, pe, h.keyMaxJ, &d.keyMaxJ))) return rc;
    if ((rc = upload(ctx, pe, h.keyOp, &d.keyOp))) return rc;
    if ((rc = upload(ctx, pe, h.keyPa, &d.keyPa))) return rc;
    if ((rc = upload(ctx, pe, h.keyPb, &d.keyPb))) return rc;
    if ((rc = upload(ctx, pe, h.keyInit, &d.keyInit))) return rc;
    if ((rc = upload(ctx, pe, h.keyPairBase, &d.keyPairBase))) return rc;
    if ((rc = upload(ctx, pe, h.keyThrBase, &d.keyThrBase))) return rc;
    if ((rc = upload(ctx, pe, h.thrKey, &d.thrKey))) return rc;
    if ((rc = upload(ctx, pe, h.lastKeyId, &d.lastKeyId))) return rc;
    if ((rc = upload(ctx, pe, h.lastPairBase, &d.lastPairBase))) return rc;
    if ((rc = upload(ctx, pe, h.kBinBase, &d.kBinBase))) return rc;
    if ((rc = upload(ctx, pe, h.nKBins, &d.nKBins))) return rc;
    if ((rc = upload(ctx, pe, h.kBinPatOff, &d.kBinPatOff))) return rc;
    if ((rc = upload(ctx, pe, h.kBinPats, &d.kBinPats))) return rc;
    if ((rc = upload(ctx, pe, h.kFetchOff, &d.kFetchOff))) return rc;
    if ((rc = upload(ctx, pe, h.kFetchRows, &d.kFetchRows))) return rc;
    if ((rc = upload(ctx, pe, h.kRowBase, &d.kRowBase))) return rc;
    if ((rc = upload(ctx, pe, h.kBinHdr, &d.kBinHdr))) return rc;
    if ((rc = upload(ctx, pe, h.kPatRec, &d.kPatRec))) return rc;
    if ((rc = upload(ctx, pe, h.kLane, &d.kLane))) return rc;
    if ((rc = upload(ctx, pe, h.kElem, &d.kElem))) return rc;
    if ((rc = upload(ctx, pe, h.kSegs, &d.kSegs))) return rc;
    d.maxEpochSlots = h.maxEpochSlots;
    d.dirOff = h.dirOff;
    if ((rc = upload(ctx, pe, h.epochMode, &d.epochMode))) return rc;
    if ((rc = upload(ctx, pe, h.rowOff, &d.rowOff))) return rc;
    if ((rc = upload(ctx, pe, h.rowEpochBase, &d.rowEpochBase))) return rc;
    if ((rc = upload(ctx, pe, h.rowRefAll, &d.rowRefAll))) return rc;
    if ((rc = upload(ctx, pe, h.rowWordsAll, &d.rowWordsAll))) return rc;
    if ((rc = upload(ctx, pe, h.patRow, &d.patRow))) return rc;
    if ((rc = upload(ctx, pe, h.binPatOff, &d.binPatOff))) return rc;
    if ((rc = upload(ctx, pe, h.binPats, &d.binPats))) return rc;
    if ((rc = upload(ctx, pe, h.slotOff, &d.slotOff))) return rc;
    if ((rc = upload(ctx, pe, h.epochBase, &d.epochBase))) return rc;
    if ((rc = upload(ctx, pe, h.words, &d.words))) return rc;
    if ((rc = upload(ctx, pe, h.meta, &d.meta))) return rc;
    if ((rc = upload(ctx, pe, h.trBase, &d.trBase))) return rc;
    if ((rc = upload(ctx, pe, h.trOff, &d.trOff))) return rc;
    if ((rc = upload(ctx, pe, h.trEntBase, &d.trEntBase))) return rc;
    if ((rc = upload(ctx, pe, h.trEnt, &d.trEnt))) return rc;
    if ((rc = upload(ctx, pe, h.shortcutOK, &d.shortcutOK))) return rc;
    d.aInit = (const int32_t*)ctx->dAInit.p;
    d.combFac = (const double*)ctx->dCombFac.p;
    d.support = (const uint8_t*)ctx->dSupport.p;
    d.patGlobal = (const int32_t*)ctx->dPatGlobal.p;
    pe->allShortcut = true;
    for (uint8_t f : h.shortcutOK) pe->allShortcut = pe->allShortcut && f;
    if ((rc = finish_upload(ctx, pe))) return rc;
    // the host keeps the small per-epoch / per-group arrays it launches from; the bulk now lives on the device only
    rc::PlanHost& hm = pe->h;
    auto drop = [](auto& v) { std::decay_t<decltype(v)>().swap(v); };
    drop(hm.words); drop(hm.meta); drop(hm.trOff); drop(hm.trEnt); drop(hm.slotOff); drop(hm.kLane); drop(hm.kElem);
    drop(hm.kPatRec); drop(hm.kRowBase); drop(hm.kFetchRows); drop(hm.kBinHdr); drop(hm.rowRefAll); drop(hm.rowWordsAll);
    drop(hm.patRow); drop(hm.thrKey);
    return RC_OK;
}

// Plan cache: least-recently-used eviction (an optimiser that crosses event orders keeps its working set; the plans of
// the batch being resolved are never evicted).  Limits: "plan_cache_max" plans and "plan_cache_mb" MiB of device memory.
void evict_plans(rc_ctx* ctx, size_t needBytes) {
    auto over = [&]() {
        return (int)ctx->plans.size() >= ctx->optPlanCacheMax ||
               (double)(ctx->planBytes + needBytes) > ctx->optPlanCacheMB * 1048576.0;
    };
    bool synced = false;
    while (!ctx->plans.empty() && over()) {
        auto victim = ctx->plans.end();
        for (auto it = ctx->plans.begin(); it != ctx->plans.end(); ++it)
            if (it->second->lastUse < ctx->evalSerial && (victim == ctx->plans.end() || it->second->lastUse < victim->second->lastUse))
                victim = it;
        if (victim == ctx->plans.end()) break;          // everything left belongs to the current batch
        if (!synced) { cudaDeviceSynchronize(); synced = true; ctx->clear_graphs(); }   // launches of earlier calls may still read the plan
        ctx->planBytes -= victim->second->bytes;
        victim->second->release();
        delete victim->second;
        ctx->plans.erase(victim);
        ++ctx->planEvictions;
    }
}

int get_plan(rc_ctx* ctx, ModelHost& mh, bool* hit) {
    auto it = ctx->plans.find(mh.sig);
    if (it != ctx->plans.end()) {
        mh.plan = it->second;
        mh.plan->lastUse = ctx->evalSerial;
        return RC_OK;
    }
    *hit = false;
    const auto t0 = std::chrono::steady_clock::now();
    PlanEntry* pe = new PlanEntry();
    std::string e = rc::build_plan(ctx->P, ctx->maxAf, ctx->nPat, ctx->patS.data(), ctx->nVec.data(), mh.sev, mh.gaps,
                                   mh.hasSplit, ctx->optBigSmemKB > 0 ? std::min(ctx->smemOptin, ctx->optBigSmemKB * 1024) : ctx->smemOptin, pe->h,
                                   ctx->optPattern ? (ctx->optPatternEmbed ? 3 : 1) : 0);
    if (!e.empty()) {
        delete pe;
        ctx->err = e;
        return RC_ERR_UNSUPPORTED;
    }
    const auto t1 = std::chrono::steady_clock::now();
    evict_plans(ctx, 0);
    int rc = upload_plan(ctx, pe);
    if (rc != RC_OK && cudaPeekAtLastError() == cudaErrorMemoryAllocation) {
        // out of device memory: drop every plan no model of this batch uses and try once more
        cudaGetLastError();
        pe->release();
        const int keepMax = ctx->optPlanCacheMax;
        ctx->optPlanCacheMax = 1;
        evict_plans(ctx, 0);
        ctx->optPlanCacheMax = keepMax;
        rc = upload_plan(ctx, pe);
    }
    if (rc != RC_OK) {
        pe->release();
        delete pe;
        return rc;
    }
    pe->bytes = pe->allocBytes;
    ctx->planBytes += pe->bytes;
    pe->lastUse = ctx->evalSerial;
    ctx->plans[mh.sig] = pe;
    mh.plan = pe;
    ++ctx->planMisses;
    if (std::getenv("RC_PLAN_TIMING"))
        std::fprintf(stderr, "[rc plan] built in %.1f ms, uploaded %.1f MB in %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                     (double)pe->bytes / 1048576.0, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    ctx->planBuildMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return RC_OK;
}

size_t align16(size_t n) { return (n + 15) & ~(size_t)15; }

// Executed state-steps (one unit = one iteration of Core.hs:248): live states x update
// steps, epoch by epoch; patterns that take the shortcut stop at sShort, the others run to T.
long long count_state_steps(const ModelHost& mh, const rc::PlanHost& h, int T) {
    long long W = 0;
    int prev = 0;
    for (int e = 0; e < h.nEpochs; ++e) {
        bool last = e + 1 == h.nEpochs;
        int endFull = last ? T : mh.sevStep[(size_t)e];
        int endShort = last ? (mh.sShort >= 0 ? mh.sShort : T) : endFull;
        W += h.epochSlotsShort[(size_t)e] * (long long)(endShort - prev) +
             (h.epochSlots[(size_t)e] - h.epochSlotsShort[(size_t)e]) * (long long)(endFull - prev);
        prev = endFull;
    }
    return W;
}

// FP64 instructions executed by the propagation kernels of the keyed tier (rc_plan.h, epochFp64), counted like the
// state-steps above.
long long count_fp64_instr(const ModelHost& mh, const rc::PlanHost& h, int T) {
    if (h.epochFp64.empty()) return 0;
    long long W = 0;
    int prev = 0;
    for (int e = 0; e < h.nEpochs; ++e) {
        bool last = e + 1 == h.nEpochs;
        int endFull = last ? T : mh.sevStep[(size_t)e];
        int endShort = last ? (mh.sShort >= 0 ? mh.sShort : T) : endFull;
        W += h.epochFp64Short[(size_t)e] * (long long)(endShort - prev) +
             (h.epochFp64[(size_t)e] - h.epochFp64Short[(size_t)e]) * (long long)(endFull - prev);
        prev = endFull;
    }
    return W;
}

// The whole evaluation up to the per-model partial sums (left on the device).
int eval_chunk(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
               const int32_t* noShortcut /*[B]*/, double* dPartial, double* dProbs, cudaStream_t st, int32_t* outStatus, long long* maxPairsOut) {
    ++ctx->evalSerial;
    const auto tHost0 = std::chrono::steady_clock::now();
    ctx->planMissesAtCall = ctx->planMisses;
    ctx->planEvictionsAtCall = ctx->planEvictions;
    ctx->planBuildMsAtCall = ctx->planBuildMs;
    const int P = ctx->P, A1 = ctx->maxAf + 1, T = (int)ctx->times.size();
    const int rowLen = rc_row_len(P, A1);
    std::memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->statsPending = false;

    std::vector<ModelHost> mhs((size_t)B);
    bool hit = true;
    // models are resolved against the grid independently of each other: a large batch (mcmc proposals, possibly with
    // hundreds of per-step size changes each) is spread over a few host threads; the plan cache is then consulted serially
    {
        auto resolve_range = [&](int b0, int b1) {
            for (int b = b0; b < b1; ++b) {
                ModelHost& mh = mhs[(size_t)b];
                if (ctx->compErr) {
                    mh.status = RC_ERR_COMP;
                    mh.err = ctx->compErrMsg;
                } else resolve_model(ctx, ev + evOff[b], evOff[b + 1] - evOff[b], disc + (size_t)b * P, noShortcut[b], mh);
            }
        };
        long long nEvTotal = (long long)evOff[B] - evOff[0];
        for (int q = evOff[0]; q < evOff[B]; ++q) nEvTotal += ev[q].type == RC_EV_GROWTH ? 512 : 0;   // rewritten into per-step events
        int nThr = (int)std::min<long long>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())), nEvTotal / 2048);
        nThr = std::min(nThr, B / 8);
        if (nThr <= 1) resolve_range(0, B);
        else {
            std::vector<std::thread> pool;
            for (int t = 1; t < nThr; ++t) pool.emplace_back(resolve_range, (int)((long long)B * t / nThr), (int)((long long)B * (t + 1) / nThr));
            resolve_range(0, B / nThr);
            for (auto& th : pool) th.join();
        }
    }
    for (int b = 0; b < B; ++b) {
        ModelHost& mh = mhs[(size_t)b];
        if (mh.status == RC_OK && ctx->nPat > 0) {
            int rc = get_plan(ctx, mh, &hit);
            if (rc != RC_OK) return rc;
        }
        if (outStatus) outStatus[b] = mh.status;
        if (mh.status != RC_OK && ctx->err.empty()) ctx->err = mh.err;
    }
    // order models by plan so that each plan is one launch
    std::vector<int> order((size_t)B);
    for (int b = 0; b < B; ++b) order[(size_t)b] = b;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return mhs[(size_t)x].plan < mhs[(size_t)y].plan; });

    // ---- pack the per-call blob: headers | structural events | segment offsets | segments | weights
    size_t nSev = 0, nSize = 0, nMask = 0, nW = 0, nMep = 0;
    for (auto& mh : mhs) {
        if (mh.status != RC_OK) continue;
        nMep += mh.sev.size() + 1;
        nSev += mh.sev.size();
        for (auto& v : mh.sizes) nSize += v.size();
        nMask += mh.masks.size();
        for (auto& se : mh.sev) if (se.type == RC_EV_SPLIT) nW += (size_t)A1 * A1;
    }
    // Layout: everything whose size follows from (B, structures) first, the per-step size lists — whose length moves with
    // the parameters under SetGrowthRate — last: the device addresses the kernels are launched with then stay the same from
    // call to call and a captured launch graph (below) can be replayed.
    size_t oModels = 0;
    size_t oSev = align16(oModels + (size_t)B * sizeof(RcModelDev));
    size_t oMep = align16(oSev + std::max<size_t>(nSev, 1) * sizeof(RcSEvDev));
    size_t oW = align16(oMep + std::max<size_t>(nMep, 1) * sizeof(RcModelEpochDev));
    size_t oSizeOff = align16(oW + std::max<size_t>(nW, 1) * sizeof(double));
    size_t oMaskOff = align16(oSizeOff + ((size_t)B * P + 1) * sizeof(int));
    size_t oMasks = align16(oMaskOff + (size_t)(B + 1) * sizeof(int));
    size_t oSizes = align16(oMasks + std::max<size_t>(nMask, 1) * sizeof(RcMaskStepDev));
    size_t blobBytes = align16(oSizes + std::max<size_t>(nSize, 1) * sizeof(RcSizeStepDev));
    // the pinned staging buffer is reused: wait until the previous call's upload has left it
    if (ctx->blobInFlight) {
        CK(cudaEventSynchronize(ctx->evt[5]));
        ctx->blobInFlight = false;
    }
    if (blobBytes > ctx->dBlob.cap) CK(cudaStreamSynchronize(st));   // kernels of the previous call may still read it
    CK(ctx->hBlob.ensure(blobBytes));
    CK(ctx->dBlob.ensure(blobBytes));
    unsigned char* hb = (unsigned char*)ctx->hBlob.p;
    std::memset(hb, 0, blobBytes);
    RcModelDev* hM = (RcModelDev*)(hb + oModels);
    RcSEvDev* hSev = (RcSEvDev*)(hb + oSev);
    int* hSizeOff = (int*)(hb + oSizeOff);
    int* hMaskOff = (int*)(hb + oMaskOff);
    RcSizeStepDev* hSizes = (RcSizeStepDev*)(hb + oSizes);
    RcMaskStepDev* hMasks = (RcMaskStepDev*)(hb + oMasks);
    double* hW = (double*)(hb + oW);
    RcModelEpochDev* hMep = (RcModelEpochDev*)(hb + oMep);
    size_t mepPos = 0;
    long long ktabPairs = 0, aEndDoubles = 0, aShortDoubles = 0, maxRowPairs = 0, carrySlots = 0;
    const long long ktabModelMax = (long long)(ctx->optKtabModelMB * 1048576.0 / 16.0);
    const long long ktabTotalMax = (long long)(ctx->optKtabTotalMB * 1048576.0 / 16.0);

    long long tabDoubles = 0;
    int rowsMax = 0;
    size_t sevPos = 0, sizePos = 0, maskPos = 0, wPos = 0;
    long long W = 0, slots0 = 0, Wfp = 0;
    const long long miss0 = ctx->planMissesAtCall, evict0 = ctx->planEvictionsAtCall;
    const double build0 = ctx->planBuildMsAtCall;
    for (int pos = 0; pos < B; ++pos) {
        const int b = order[(size_t)pos];
        ModelHost& mh = mhs[(size_t)b];
        RcModelDev& M = hM[pos];
        std::memset(&M, 0, sizeof M);      // the staging buffer is reused: a rejected model must not keep a previous call's
                                           // table offsets (rc_tables_kernel would write its rows over another model's)
        hMaskOff[pos] = (int)maskPos;
        for (int k = 0; k < P; ++k) hSizeOff[(size_t)pos * P + k] = (int)sizePos;   // overwritten below for a valid model
        M.pad = b;   // original model index: rows of partial / dProbs
        if (mh.status != RC_OK || ctx->nPat == 0) {
            M.nSteps = mh.status != RC_OK ? -1 : 0;
            continue;
        }
        const PlanEntry* pe = mh.plan;
        M.nSteps = T;
        M.sShort = mh.sShort;
        M.nSEv = (int)mh.sev.size();
        M.sevOff = (int)sevPos;
        const rc::PlanHost& hp = pe->h;
        // ---- keyed mode: per-epoch step ranges and the model's slice of the trajectory table ----
        {
            const int finalEnd = (mh.sShort >= 0 && pe->allShortcut) ? mh.sShort : T;
            long long pairs = 0;
            M.mepOff = (int)mepPos;
            for (int e = 0; e < hp.nEpochs; ++e) {
                RcModelEpochDev& me = hMep[mepPos++];
                me.beg = e == 0 ? 0 : mh.sevStep[(size_t)e - 1];
                me.end = e + 1 < hp.nEpochs ? mh.sevStep[(size_t)e] : finalEnd;
                if (me.end < me.beg) me.end = me.beg;
                me.tabBase = ktabPairs + pairs;
                pairs += (long long)(me.end - me.beg) * hp.ep[(size_t)e].rowPairs;
                maxRowPairs = std::max<long long>(maxRowPairs, hp.ep[(size_t)e].rowPairs);
            }
            M.mode = (ctx->optKeyed && hp.nFastBins > 0 && pairs <= ktabModelMax && ktabPairs + pairs <= ktabTotalMax) ? 1 : 0;
            if (M.mode == 1) {
                // only models that own a slice of the trajectory table bound the chunk size
                if (maxPairsOut && pairs > *maxPairsOut) *maxPairsOut = pairs;
                ktabPairs += pairs;
                M.aEndOff = aEndDoubles;
                M.aShortOff = aShortDoubles;
                aEndDoubles += hp.totalKeys;
                aShortDoubles += hp.ep.back().nKeys;
                carrySlots = std::max<long long>(carrySlots, hp.maxEpochSlots);
            }
            if (mh.sShort >= 0) {
                for (int k = 0; k < P; ++k) M.Nshort[k] = mh.size_at(k, mh.sShort);
            }
        }
        // per-step table rows: read by the self-contained tiers and by the trajectory kernel of the keyed tier
        M.rowsAvail = (mh.sShort >= 0 && pe->allShortcut) ? mh.sShort + 1 : T;
        M.tabOff = tabDoubles;
        M.wPool = 0;
        M.theta = theta[b];
        for (int k = 0; k < P; ++k) M.disc[k] = disc[(size_t)b * P + k];
        tabDoubles += (long long)M.rowsAvail * rowLen;
        rowsMax = std::max(rowsMax, M.rowsAvail);
        for (size_t g = 0; g < mh.sev.size(); ++g) {
            RcSEvDev& d = hSev[sevPos++];
            d.step = mh.sevStep[g];
            d.type = mh.sev[g].type;
            d.k = mh.sev[g].k;
            d.l = mh.sev[g].l;
            d.m = mh.sevM[g];
            d.wOff = 0;
            if (d.type == RC_EV_SPLIT) {
                d.wOff = (int)wPos;
                const double m = d.m;
                for (int xl = 0; xl < A1; ++xl)
                    for (int s = 0; s <= xl; ++s)   // Core.hs:216
                        hW[wPos + (size_t)xl * A1 + s] = ghc_pow(m, s) * ghc_pow(1.0 - m, xl - s) * rc::choose(xl, s);
                wPos += (size_t)A1 * A1;
            }
        }
        for (int k = 0; k < P; ++k) {
            hSizeOff[(size_t)pos * P + k] = (int)sizePos;
            for (const SizeStep& z : mh.sizes[(size_t)k]) {
                RcSizeStepDev& d = hSizes[sizePos++];
                d.sBegin = z.sBegin;
                d.pad = 0;
                d.N = z.N;
            }
        }
        for (const MaskStep& z : mh.masks) {
            RcMaskStepDev& d = hMasks[maskPos++];
            d.sBegin = z.sBegin;
            d.frozenMask = z.mask;
        }
        const rc::PlanHost& h = pe->h;
        W += count_state_steps(mh, h, T);
        if (M.mode == 1) Wfp += count_fp64_instr(mh, h, T);
        slots0 += h.epochSlots[0];
        ctx->stats.n_epochs = h.nEpochs;
        ctx->stats.n_steps = (mh.sShort >= 0 && pe->allShortcut) ? mh.sShort + 1 : T;
    }
    hSizeOff[(size_t)B * P] = (int)sizePos;
    hMaskOff[B] = (int)maskPos;
    ctx->stats.state_steps = W;
    ctx->stats.slots = slots0;
    ctx->stats.plan_cache_hit = hit ? 1 : 0;
    ctx->stats.fp64_instr = Wfp;
    ctx->stats.ktab_mb = (double)ktabPairs * 16.0 / 1048576.0;
    ctx->stats.chunks = 1;
    for (int pos = 0; pos < B; ++pos) ctx->stats.models_selfcontained += hM[pos].nSteps > 0 && hM[pos].mode == 0;
    ctx->stats.plan_misses = (int32_t)(ctx->planMisses - miss0);
    ctx->stats.plan_evictions = (int32_t)(ctx->planEvictions - evict0);
    ctx->stats.plan_build_ms = ctx->planBuildMs - build0;

    {
        size_t needTab = (size_t)std::max<long long>(tabDoubles, 1) * sizeof(double);
        size_t needOut = (size_t)B * std::max(ctx->nPat, 1) * sizeof(double);
        // the keyed kernel prefetches up to 8 step rows past the end of an epoch: keep that much slack
        size_t needK = (size_t)(std::max<long long>(ktabPairs, 1) + 8 * maxRowPairs + 64) * 16;
        size_t needA = (size_t)std::max<long long>(aEndDoubles, 1) * sizeof(double);
        // cumulative products of exp(-dt/2N) per (table row, branch): rc_cum_kernel
        size_t needCum = (size_t)std::max<long long>(tabDoubles / rowLen, 1) * P * sizeof(double);
        size_t needS = (size_t)std::max<long long>(aShortDoubles, 1) * sizeof(double);
        // b carried from one epoch's launch to the next (two buffers) and the per-pattern updateD accumulators
        size_t needC = (size_t)std::max<long long>(2 * carrySlots * B, 1) * sizeof(double);
        // global-memory mode of the big tier: one slice per block, at most 4 GiB (the launch is cut into model ranges beyond that)
        size_t needBig = 0;
        for (int pos = 0; pos < B;) {
            const PlanEntry* pe = mhs[(size_t)order[(size_t)pos]].plan;
            int end = pos + 1;
            while (end < B && mhs[(size_t)order[(size_t)end]].plan == pe) ++end;
            if (pe && pe->d.bigGlobal && pe->d.nBins > 0) {
                const size_t per = rc_big_scratch_bytes(pe->d.cap, pe->d.nnzw) * (size_t)pe->d.nBins;
                needBig = std::max(needBig, std::max(per, std::min(per * (size_t)(end - pos), (size_t)4 << 30)));
            }
            pos = end;
        }
        if (needTab > ctx->dTab.cap || needOut > ctx->dPOut.cap || needK > ctx->dKTab.cap || needA > ctx->dAEnd.cap ||
            needS > ctx->dAShort.cap || needC > ctx->dBCarry.cap || needOut > ctx->dDAcc.cap)
            CK(cudaStreamSynchronize(st));
        CK(ctx->dBCarry.ensure(needC));
        if (needBig > ctx->dBigScratch.cap) { CK(cudaStreamSynchronize(st)); CK(ctx->dBigScratch.ensure(needBig)); }
        CK(ctx->dDAcc.ensure(needOut));
        CK(ctx->dTab.ensure(needTab));
        CK(ctx->dPOut.ensure(needOut));
        CK(ctx->dKTab.ensure(needK));
        CK(ctx->dAEnd.ensure(needA));
        CK(ctx->dCum.ensure(needCum));
        {
            // slice sums and arrival counters of rc_reduce_kernel: the counters start at zero and the kernel leaves them there
            const size_t needRed = rc_reduce_scratch_bytes(B);
            if (needRed > ctx->dRed.cap) {
                CK(cudaStreamSynchronize(st));
                CK(ctx->dRed.ensure(needRed));
                CK(cudaMemsetAsync(ctx->dRed.p, 0, ctx->dRed.cap, st));
            }
        }
        CK(ctx->dAShort.ensure(needS));
    }

    unsigned char* db = (unsigned char*)ctx->dBlob.p;
    const RcModelDev* dM = (const RcModelDev*)(db + oModels);
    const RcSEvDev* dSev = (const RcSEvDev*)(db + oSev);
    const int* dSizeOff = (const int*)(db + oSizeOff);
    const int* dMaskOff = (const int*)(db + oMaskOff);
    const RcSizeStepDev* dSizes = (const RcSizeStepDev*)(db + oSizes);
    const RcMaskStepDev* dMasks = (const RcMaskStepDev*)(db + oMasks);
    const double* dW = (const double*)(db + oW);
    const RcModelEpochDev* dMep = (const RcModelEpochDev*)(db + oMep);

    if (ctx->firstChunk) CK(cudaEventRecord(ctx->evt[0], st));
    CK(cudaMemcpyAsync(db, hb, blobBytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->evt[5], st));
    ctx->blobInFlight = true;
    if (dProbs && ctx->world > 1) CK(cudaMemsetAsync(dProbs, 0, (size_t)B * ctx->nPatGlobal * sizeof(double), st));
    // ---- the launch chain.  Enqueued on the stream directly, or — from the second call with the same launch signature on —
    //      replayed as a CUDA graph captured from the same code: a step is 25 to 60 short launches over up to 8 streams
    //      with event joins, and for a small batch (a single `logl`) the host-side launch cost is most of the call.
    const int rowsGrid = rowsMax > 0 ? ((rowsMax + 63) / 64) * 64 : 0;      // the table kernel skips rows >= rowsAvail
    int launches = 0;
    bool capturing = false;
    auto rec_timing = [&](cudaEvent_t e, cudaStream_t s2) {
        // under capture a timing event must be an event-record NODE (it is recorded when the graph runs)
        return capturing ? cudaEventRecordWithFlags(e, s2, cudaEventRecordExternal) : cudaEventRecord(e, s2);
    };
    // fuse[c]: model range c of the plan group [pos, end) finishes its patterns in the launches of the last but one epoch
    auto fuse_flags = [&](const PlanEntry* pe, int pos, int end, int nChains, bool* fuse) {
        const int nE = pe->h.nEpochs, nModels = end - pos;
        bool fuseOK = ctx->optFuseLast && nE >= 2 && pe->allShortcut && pe->h.lastSingle && !pe->h.grpDirIn.empty();
        for (int v = RC_NGROUP * (nE - 1); fuseOK && v < RC_NGROUP * nE; ++v)
            if (pe->h.nKBins[(size_t)v] > 0 && (pe->h.epochMode[(size_t)v] < RC_MODE_PAT || !pe->h.grpDirIn[(size_t)v])) fuseOK = false;
        for (int c = 0; c < nChains; ++c) {
            const int p0 = pos + (int)((long long)nModels * c / nChains), p1 = pos + (int)((long long)nModels * (c + 1) / nChains);
            fuse[c] = fuseOK;
            for (int q = p0; q < p1 && fuse[c]; ++q) {
                if (hM[q].nSteps < 0 || hM[q].mode != 1) continue;          // skipped by the kernels
                const RcModelEpochDev& me = hMep[hM[q].mepOff + nE - 1];
                // (the step at which the last event fires is executed before the shortcut can apply: normally one grid step)
                if (hM[q].sShort < 0 || hM[q].sShort != me.end || me.end - me.beg > 4 || hM[q].sShort >= hM[q].nSteps) fuse[c] = false;
            }
        }
    };
    auto enqueue = [&](cudaStream_t st) -> int {
        launches = 0;
        CK(rec_timing(ctx->evt[1], st));
        if (rowsGrid > 0) {
            CK(rc_launch_tables(dM, dSizes, dSizeOff, dMasks, dMaskOff, (const double*)ctx->dDts.p, (double*)ctx->dTab.p, P, A1, rowsGrid, B, st));
            ++launches;
        }
        CK(rec_timing(ctx->evt[2], st));
        ctx->trajTimed = false;
        for (int pos = 0; pos < B;) {
            const PlanEntry* pe = mhs[(size_t)order[(size_t)pos]].plan;
            int end = pos + 1;
            while (end < B && mhs[(size_t)order[(size_t)end]].plan == pe) ++end;
            if (pe && ctx->nPat > 0) {
                bool anyKeyed = false, anySelf = false;
                for (int q = pos; q < end; ++q) (hM[q].mode == 1 ? anyKeyed : anySelf) = true;
                if (pe->d.nFastBins > 0 && anyKeyed) {
                    const bool timeIt = !ctx->trajTimed;      // trajectory kernels of the first plan group are timed
                    cudaStream_t ts = ctx->trajStream;
                    while ((int)ctx->evTraj.size() < pe->h.nEpochs) {
                        cudaEvent_t ev = nullptr;
                        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                        ctx->evTraj.push_back(ev);
                    }
                    CK(cudaEventRecord(ctx->evFork, st));     // tables written, previous work on st finished
                    CK(cudaStreamWaitEvent(ts, ctx->evFork, 0));
                    if (timeIt) CK(rec_timing(ctx->evt[6], ts));
                    CK(rc_launch_cum(pe->h.nEpochs, P, A1, dM, dMep, (const double*)ctx->dTab.p, (double*)ctx->dCum.p, pos, end - pos, ts));
                    ++launches;
                    for (int e = 0; e < pe->h.nEpochs; ++e) {
                        CK(rc_launch_traj(pe->d, e, pe->h.ep[(size_t)e].nThr, dM, dMep, dSev, (const double*)ctx->dTab.p, (const double*)ctx->dCum.p,
                                          (const double*)ctx->dDts.p, (double*)ctx->dKTab.p, (double*)ctx->dAEnd.p, (double*)ctx->dAShort.p,
                                          pos, end - pos, ts));
                        // (the timing record precedes the epoch's join event: a captured side stream must end in a join)
                        if (timeIt && e + 1 == pe->h.nEpochs) { CK(rec_timing(ctx->evt[7], ts)); ctx->trajTimed = true; }
                        CK(cudaEventRecord(ctx->evTraj[(size_t)e], ts));
                        ++launches;
                    }
                    // one launch per structural epoch and bin group; b ping-pongs between the two carry buffers.  A batch large
                    // enough is cut into RC_NCHAIN model ranges, each with its own chain of streams.
                    const int nModels = end - pos;
                    const int nChains = nModels >= 16 ? RC_NCHAIN : 1;
                    if (nChains > 1) {
                        CK(cudaEventRecord(ctx->evChain[0], st));         // everything queued so far (blob, tables) precedes the chains
                        for (int c = 1; c < nChains; ++c) CK(cudaStreamWaitEvent(ctx->chainStream[c], ctx->evChain[0], 0));
                    }
                    // Last epoch without a grid step (the shortcut applies at its first step, the usual case of a tree model) and
                    // reached by direct hand-over: the launches of the last but one epoch finish the patterns (RC_E_FUSE) and the
                    // last epoch is not launched.  Decided per model range.
                    const int nE = pe->h.nEpochs;
                    bool fuse[RC_NCHAIN];
                    fuse_flags(pe, pos, end, nChains, fuse);
                    for (int e = 0; e < pe->h.nEpochs; ++e) {
                        double* c0 = (double*)ctx->dBCarry.p;
                        double* c1 = c0 + (size_t)carrySlots * B;
                        for (int c = 0; c < nChains; ++c) {
                            if (fuse[c] && e == nE - 1) continue;
                            const int eArg = (fuse[c] && e == nE - 2) ? (e | RC_E_FUSE) : e;
                            cudaStream_t cs = c == 0 ? st : ctx->chainStream[c];
                            const int p0 = pos + (int)((long long)nModels * c / nChains), p1 = pos + (int)((long long)nModels * (c + 1) / nChains);
                            CK(cudaStreamWaitEvent(cs, ctx->evTraj[(size_t)e], 0));
                            if (eArg != e) CK(cudaStreamWaitEvent(cs, ctx->evTraj[(size_t)nE - 1], 0));      // the a-values at the shortcut step
                            // bin groups of the epoch: tiny / small / large patterns (one thread each), and the rest (rows or states).
                            // The group with the fewest, longest blocks goes first; every further group runs on its own stream.
                            struct Launch { int binBase, nBins, mode, maxFetch, maxB, z; };
                            Launch L[RC_NGROUP];
                            int nL = 0;
                            for (int v = RC_NGROUP * e; v < RC_NGROUP * (e + 1); ++v) {
                                if (pe->h.nKBins[(size_t)v] == 0) continue;
                                // A batch too small to fill the GPU cares about launches, not occupancy: the tiny-pattern bins then
                                // run in the same launch as the small-pattern bins (their headers are contiguous but for one unused
                                // entry).
                                const int nb0 = pe->h.nKBins[(size_t)v], nb1 = v + 1 < RC_NGROUP * (e + 1) ? pe->h.nKBins[(size_t)v + 1] : 0;
                                if (pe->h.epochMode[(size_t)v] == RC_MODE_PAT && nb1 > 0 && pe->h.epochMode[(size_t)v + 1] == RC_MODE_PAT + 1 &&
                                    pe->h.kBinBase[(size_t)v + 1] == pe->h.kBinBase[(size_t)v] + nb0 + 1 &&
                                    (long long)(nb0 + nb1) * nModels < 3LL * 7 * ctx->smCount) {
                                    L[nL++] = {pe->h.kBinBase[(size_t)v], nb0 + 1 + nb1, RC_MODE_PAT + 1,
                                               std::max(pe->h.grpMaxFetch[(size_t)v], pe->h.grpMaxFetch[(size_t)v + 1]),
                                               std::max(pe->h.grpMaxB[(size_t)v], pe->h.grpMaxB[(size_t)v + 1]), RC_PAT12_NP};
                                    ++v;
                                    continue;
                                }
                                L[nL++] = {pe->h.kBinBase[(size_t)v], nb0, pe->h.epochMode[(size_t)v], pe->h.grpMaxFetch[(size_t)v], pe->h.grpMaxB[(size_t)v], 1};
                            }
                            if (nL > 1) CK(cudaEventRecord(ctx->evEpoch[c], cs));
                            for (int q = nL - 1; q >= 0; --q) {
                                cudaStream_t gs = q == 0 ? cs : ctx->grpStream[c][q - 1];
                                if (q > 0) CK(cudaStreamWaitEvent(gs, ctx->evEpoch[c], 0));
                                CK(rc_launch_propagate_keyed(pe->d, eArg, L[q].binBase, L[q].nBins, L[q].mode, L[q].maxFetch, L[q].maxB, L[q].z, dM, dMep, dSev,
                                                             (const double*)ctx->dKTab.p, (const double*)ctx->dAShort.p, dW, (e & 1) ? c0 : c1,
                                                             (e & 1) ? c1 : c0, carrySlots, (double*)ctx->dDAcc.p, (double*)ctx->dPOut.p, p0, p1 - p0, gs));
                                ++launches;
                                if (q > 0) CK(cudaEventRecord(ctx->evGrp[c][q - 1], gs));
                            }
                            for (int q = 1; q < nL; ++q) CK(cudaStreamWaitEvent(cs, ctx->evGrp[c][q - 1], 0));
                        }
                    }
                    for (int c = 1; c < nChains; ++c) {
                        CK(cudaEventRecord(ctx->evChain[c], ctx->chainStream[c]));
                        CK(cudaStreamWaitEvent(st, ctx->evChain[c], 0));
                    }
                }
                if (pe->d.nFastBins > 0 && anySelf) {
                    CK(rc_launch_propagate_fast(pe->d, dM, dSev, (const double*)ctx->dTab.p, dW, (double*)ctx->dPOut.p, pos, end - pos, st));
                    ++launches;
                }
                if (pe->d.nBins > 0 && (anySelf || pe->d.nBins > pe->d.nBinsKB)) {
                    CK(rc_launch_propagate(pe->d, dM, dSev, (const double*)ctx->dTab.p, dW, (double*)ctx->dPOut.p, pos, end - pos,
                                           (unsigned char*)ctx->dBigScratch.p, ctx->dBigScratch.cap, st));
                    ++launches;
                }
            }
            pos = end;
        }
        CK(rec_timing(ctx->evt[3], st));
        CK(rc_launch_reduce((const double*)ctx->dPOut.p, (const long long*)ctx->dCounts.p, (const int*)ctx->dPatGlobal.p, dM,
                            ctx->nPat, ctx->nPatGlobal, B, dPartial, dProbs, ctx->dRed.p, st));
        ++launches;
        CK(rec_timing(ctx->evt[4], st));
        return RC_OK;
    };
    {
        // launch signature: everything baked into the kernel parameters and grid sizes of the chain
        std::string key;
        auto put = [&](const void* p, size_t n) { key.append(reinterpret_cast<const char*>(p), n); };
        auto putv = [&](long long v) { put(&v, sizeof v); };
        putv(B); putv(rowsGrid); putv(carrySlots); putv((long long)(dProbs != nullptr)); putv(ctx->nPat); putv((long long)ctx->dBigScratch.cap);
        const void* ptrs[] = {db + oModels, db + oSev, db + oMep, db + oW, db + oSizeOff, db + oMaskOff, db + oMasks, db + oSizes,
                              ctx->dTab.p, ctx->dPOut.p, ctx->dKTab.p, ctx->dAEnd.p, ctx->dAShort.p, ctx->dBCarry.p, ctx->dDAcc.p,
                              ctx->dDts.p, dPartial, dProbs, ctx->dCounts.p, ctx->dPatGlobal.p, ctx->dBigScratch.p, ctx->dCum.p, ctx->dRed.p};
        put(ptrs, sizeof ptrs);
        for (int pos = 0; pos < B;) {
            const PlanEntry* pe = mhs[(size_t)order[(size_t)pos]].plan;
            int end = pos + 1;
            while (end < B && mhs[(size_t)order[(size_t)end]].plan == pe) ++end;
            bool anyKeyed = false, anySelf = false;
            for (int q = pos; q < end; ++q) (hM[q].mode == 1 ? anyKeyed : anySelf) = true;
            putv((long long)(uintptr_t)pe); putv(end); putv(anyKeyed); putv(anySelf);
            if (pe && anyKeyed) {
                const int nCh = end - pos >= 16 ? RC_NCHAIN : 1;
                bool fuse[RC_NCHAIN];
                fuse_flags(pe, pos, end, nCh, fuse);
                for (int c = 0; c < nCh; ++c) putv(fuse[c]);
            }
            pos = end;
        }
        cudaGraphExec_t exec = nullptr;
        if (ctx->optGraphs) {
            auto it = ctx->graphs.find(key);
            if (it != ctx->graphs.end()) exec = it->second;
            else if (ctx->graphSeen.count(key)) {
                // second sighting: capture the chain on the context's own stream (the caller's may be the legacy default
                // stream, which cannot capture) and keep the executable graph
                if (ctx->graphs.size() >= 48) ctx->clear_graphs();
                cudaGraph_t g = nullptr;
                capturing = true;
                cudaError_t ce = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
                int rcq = ce == cudaSuccess ? enqueue(ctx->stream) : RC_ERR_CUDA;
                cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g);
                capturing = false;
                if (rcq == RC_OK && ee == cudaSuccess && g && cudaGraphInstantiate(&exec, g, 0) == cudaSuccess) {
                    ctx->graphs[key] = exec;
                    ctx->graphLaunches[key] = launches | (ctx->trajTimed ? 1 << 30 : 0);
                } else {
                    exec = nullptr;
                    cudaGetLastError();
                    ctx->optGraphs = 0;                       // capture is not usable here: stay on plain launches
                }
                if (g) cudaGraphDestroy(g);
            } else {
                if (ctx->graphSeen.size() >= 4096) ctx->graphSeen.clear();
                ctx->graphSeen.insert(key);
            }
        }
        if (exec) {
            CK(cudaGraphLaunch(exec, st));
            launches = ctx->graphLaunches[key] & ((1 << 30) - 1);
            ctx->trajTimed = (ctx->graphLaunches[key] >> 30) & 1;
            ctx->stats.graph_replay = 1;
        } else {
            int rcq = enqueue(st);
            if (rcq != RC_OK) return rcq;
        }
    }
    ctx->stats.launches = launches;
    ctx->stats.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tHost0).count();
    ctx->statsPending = true;
    return RC_OK;
}

int collect_stats(rc_ctx* ctx);

// Evaluates B models in chunks sized so that the trajectory tables of one chunk stay inside the memory budget
// ("ktab_total_mb"); chunks run back to back on the same stream.
int eval_impl(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
              const int32_t* noShortcut /*[B]*/, double* dPartial, double* dProbs, cudaStream_t st, int32_t* outStatus) {
    if (ctx->P == 0) { ctx->err = "rc_set_problem has not been called"; return RC_ERR_ARG; }
    if (ctx->times.empty()) { ctx->err = "rc_set_times has not been called"; return RC_ERR_ARG; }
    if (B <= 0) return RC_OK;
    const long long totalMax = (long long)(ctx->optKtabTotalMB * 1048576.0 / 16.0);
    // A batch is cut into chunks when its trajectory tables would not fit ("ktab_total_mb").  Chunks are queued back to back on
    // the stream; nothing waits for the device in between.  (Cutting host-heavy batches into chunks to overlap the host side
    // of one with the kernels of the previous was measured and dropped: the small workloads it targets are latency-bound on
    // the GPU, a chunk of 32 models takes about as long as one of 128 — C5 fell from 110k to 62k evaluations/s.)
    rc_stats acc;
    std::memset(&acc, 0, sizeof acc);
    acc.plan_cache_hit = 1;
    int done = 0, nChunks = 0;
    while (done < B) {
        int n = std::min(B - done, std::max(1, std::min(ctx->chunkHint, 32768)));
        long long maxPairs = 0;
        ctx->firstChunk = nChunks == 0;
        int rc = eval_chunk(ctx, n, ev, evOff + done, theta + done, disc + (size_t)done * ctx->P, noShortcut + done, dPartial + 2 * (size_t)done,
                            dProbs ? dProbs + (size_t)done * ctx->nPatGlobal : nullptr, st, outStatus ? outStatus + done : nullptr,
                            &maxPairs);
        if (rc != RC_OK) return rc;
        // a chunk without a keyed model says nothing about the table: let the hint grow back
        if (maxPairs > 0) ctx->chunkHint = (int)std::max<long long>(1, std::min<long long>(32768, totalMax / maxPairs));
        else ctx->chunkHint = std::min(32768, std::max(ctx->chunkHint, 256));
        done += n;
        ++nChunks;
        // counters of the chunks add up; the CUDA-event spans are read once, at the end of the call: tables / propagation /
        // trajectory spans are those of the LAST chunk, total_ms runs from the first chunk's upload to the last chunk's reduction
        acc.state_steps += ctx->stats.state_steps; acc.slots += ctx->stats.slots; acc.launches += ctx->stats.launches;
        acc.host_ms += ctx->stats.host_ms; acc.ktab_mb += ctx->stats.ktab_mb; acc.chunks += 1;
        acc.models_selfcontained += ctx->stats.models_selfcontained;
        acc.fp64_instr += ctx->stats.fp64_instr; acc.plan_misses += ctx->stats.plan_misses;
        acc.plan_evictions += ctx->stats.plan_evictions; acc.plan_build_ms += ctx->stats.plan_build_ms;
        acc.plan_cache_hit = acc.plan_cache_hit && ctx->stats.plan_cache_hit;
        acc.n_epochs = ctx->stats.n_epochs; acc.n_steps = ctx->stats.n_steps; acc.graph_replay = ctx->stats.graph_replay;
    }
    ctx->stats = acc;
    ctx->statsPending = true;
    return RC_OK;
}

int collect_stats(rc_ctx* ctx) {
    if (!ctx->statsPending) return RC_OK;
    CK(cudaEventSynchronize(ctx->evt[4]));
    // a span that cannot be read (it never fails the evaluation) is reported as 0
    auto span = [](cudaEvent_t a, cudaEvent_t b) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); ms = 0; }
        return (double)ms;
    };
    ctx->stats.tables_ms = span(ctx->evt[1], ctx->evt[2]);
    ctx->stats.kernel_ms = span(ctx->evt[2], ctx->evt[3]);
    ctx->stats.traj_ms = ctx->trajTimed ? span(ctx->evt[6], ctx->evt[7]) : 0.0;
    ctx->stats.total_ms = span(ctx->evt[0], ctx->evt[4]);
    ctx->statsPending = false;
    return RC_OK;
}

// Deal of patterns to ranks (host only).  Patterns are ordered lexicographically and cut into `world` contiguous
// chunks of (nearly) equal work, so that the patterns of a rank share their trajectory keys (each rank then needs about
// 1/world of the keys instead of all of them).  Work proxy: live states at t=0 = prod max(1, m_k).
std::vector<int> shard_patterns(int P, int n, const int32_t* pats, int rank, int world) {
    std::vector<int> order((size_t)n);
    std::vector<long long> work((size_t)n);
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        long long w = 1;
        for (int k = 0; k < P; ++k) w *= std::max(1, pats[(size_t)i * P + k]);
        work[(size_t)i] = w;
        total += w;
        order[(size_t)i] = i;
    }
    if (world > 1)
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const int32_t* x = pats + (size_t)a * P;
            const int32_t* y = pats + (size_t)b * P;
            for (int k = 0; k < P; ++k) if (x[k] != y[k]) return x[k] < y[k];
            return false;
        });
    std::vector<int> mine;
    long long run = 0;
    for (int j = 0; j < n; ++j) {
        // pattern j belongs to the rank whose work interval contains the midpoint of its own work
        const long long mid = run + work[(size_t)order[(size_t)j]] / 2;
        int r = total > 0 ? (int)((__int128)mid * world / total) : 0;
        if (r >= world) r = world - 1;
        if (r == rank) mine.push_back(order[(size_t)j]);
        run += work[(size_t)order[(size_t)j]];
    }
    std::sort(mine.begin(), mine.end());
    return mine;
}

int build_shard(rc_ctx* ctx) {
    const int P = ctx->P;
    ctx->clear_plans();
    ctx->patGlobalIdx = shard_patterns(P, ctx->nPatGlobal, ctx->patG.data(), ctx->rank, ctx->world);
    ctx->nPat = (int)ctx->patGlobalIdx.size();
    const int m = ctx->nPat;
    ctx->patS.resize((size_t)m * P);
    std::vector<int32_t> aInit((size_t)m * P);
    std::vector<uint8_t> support((size_t)m * P);
    std::vector<double> comb((size_t)m);
    std::vector<long long> cnt((size_t)m);
    for (int i = 0; i < m; ++i) {
        int g = ctx->patGlobalIdx[(size_t)i];
        double cf = 1.0;
        for (int k = 0; k < P; ++k) {
            int mk = ctx->patG[(size_t)g * P + k], nk = ctx->nVec[(size_t)k];
            ctx->patS[(size_t)i * P + k] = mk;
            aInit[(size_t)i * P + k] = nk - mk;
            support[(size_t)i * P + k] = mk > 0;
            cf = cf * rc::choose(nk, mk);        // Core.hs:49
        }
        comb[(size_t)i] = cf;
        cnt[(size_t)i] = ctx->cntG[(size_t)g];
    }
    size_t one = 1;
    CK(ctx->dAInit.ensure(std::max(one, aInit.size()) * sizeof(int32_t)));
    CK(ctx->dSupport.ensure(std::max(one, support.size())));
    CK(ctx->dCombFac.ensure(std::max(one, comb.size()) * sizeof(double)));
    CK(ctx->dCounts.ensure(std::max(one, cnt.size()) * sizeof(long long)));
    CK(ctx->dPatGlobal.ensure(std::max(one, (size_t)m) * sizeof(int)));
    if (m > 0) {
        CK(cudaMemcpy(ctx->dAInit.p, aInit.data(), aInit.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dSupport.p, support.data(), support.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dCombFac.p, comb.data(), comb.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dCounts.p, cnt.data(), cnt.size() * sizeof(long long), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->dPatGlobal.p, ctx->patGlobalIdx.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice));
    }
    return RC_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

const char* rc_version(void) { return "rarecoal_b200 0.1 (sm_100a)"; }

int rc_create(int device, rc_ctx** out) {
    if (!out) return RC_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RC_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return RC_ERR_CUDA;
    rc_ctx* ctx = new rc_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return RC_ERR_CUDA; }
    // the trajectory chain feeds every propagation launch: its blocks go first when both are runnable
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    if (cudaStreamCreateWithPriority(&ctx->trajStream, cudaStreamNonBlocking, std::getenv("RC_TRAJ_LOWPRIO") ? prLo : prHi) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return RC_ERR_CUDA; }
    for (int c = 0; c < RC_NCHAIN; ++c) {
        if ((c > 0 && cudaStreamCreateWithFlags(&ctx->chainStream[c], cudaStreamNonBlocking) != cudaSuccess) ||
            cudaEventCreateWithFlags(&ctx->evEpoch[c], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->evChain[c], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return RC_ERR_CUDA; }
        for (int g = 0; g < RC_NGROUP - 1; ++g)
            if (cudaStreamCreateWithFlags(&ctx->grpStream[c][g], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->evGrp[c][g], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return RC_ERR_CUDA; }
    }
    for (auto& e : ctx->evt)
        if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return RC_ERR_CUDA; }
    std::memset(&ctx->stats, 0, sizeof ctx->stats);
    *out = ctx;
    return RC_OK;
}

void rc_destroy(rc_ctx* ctx) {
    if (!ctx) return;
    if (ctx->multi) {
        rc_multi_destroy(ctx->multi);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();          // launches queued by rc_eval_partial on a caller's stream may still read the buffers
    if (ctx->child) rc_destroy(ctx->child);
    ctx->clear_graphs();
    if (ctx->comm) {
        std::string e;
        if (const rc::NcclApi* api = rc::nccl_api(e)) api->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->clear_plans();
    DevBuf* bufs[] = {&ctx->dAInit, &ctx->dCombFac, &ctx->dSupport, &ctx->dPatGlobal, &ctx->dCounts, &ctx->dDts,
                      &ctx->dBlob, &ctx->dTab, &ctx->dPOut, &ctx->dPartial, &ctx->dLogl, &ctx->dProbs, &ctx->dKTab, &ctx->dAEnd,
                      &ctx->dAShort, &ctx->dBCarry, &ctx->dDAcc, &ctx->dBigScratch, &ctx->dCum, &ctx->dRed};
    for (DevBuf* b : bufs) b->release();
    ctx->hBlob.release();
    ctx->hOut.release();
    ctx->hPlanStage.release();
    for (auto& e : ctx->evt) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->trajStream) cudaStreamDestroy(ctx->trajStream);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    for (int c = 0; c < RC_NCHAIN; ++c) {
        if (ctx->chainStream[c]) cudaStreamDestroy(ctx->chainStream[c]);
        if (ctx->evEpoch[c]) cudaEventDestroy(ctx->evEpoch[c]);
        if (ctx->evChain[c]) cudaEventDestroy(ctx->evChain[c]);
        for (int g = 0; g < RC_NGROUP - 1; ++g) {
            if (ctx->grpStream[c][g]) cudaStreamDestroy(ctx->grpStream[c][g]);
            if (ctx->evGrp[c][g]) cudaEventDestroy(ctx->evGrp[c][g]);
        }
    }
    for (cudaEvent_t ev : ctx->evTraj) cudaEventDestroy(ev);
    delete ctx;
}

const char* rc_last_error(rc_ctx* ctx) {
    if (!ctx) return "null context";
    if (ctx->multi) return rc_multi_last_error(ctx->multi);
    return ctx->err.c_str();
}

int rc_set_problem(rc_ctx* ctx, int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec,
                   const int64_t* counts, int64_t totalSites) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_set_problem(ctx->multi, P, maxAf, nPat, patterns, nVec, counts, totalSites);
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    if (P < 1 || P > RC_MAX_P || maxAf < 1 || maxAf > RC_MAX_AF || nPat < 0 || !nVec || (nPat > 0 && !patterns)) {
        ctx->err = "rc_set_problem: need 1 <= P <= 32, 1 <= maxAf <= 62";
        return RC_ERR_ARG;
    }
    if (!rc::key_fits(P, maxAf)) { ctx->err = "rc_set_problem: (maxAf+1)^P exceeds 63 bits"; return RC_ERR_UNSUPPORTED; }
    ctx->compErr = false;
    ctx->compErrMsg.clear();
    for (int i = 0; i < nPat; ++i) {
        int sum = 0;
        for (int k = 0; k < P; ++k) {
            int m = patterns[(size_t)i * P + k];
            if (m < 0) { ctx->err = "rc_set_problem: negative pattern entry"; return RC_ERR_ARG; }
            sum += m;
            // choose n k is 0 for k > n: "Overflow Error in getProb" (Core.hs:53-56)
            if (m > nVec[k] && !ctx->compErr) {
                ctx->compErr = true;
                ctx->compErrMsg = "Overflow Error in getProb: pattern " + std::to_string(i) + " exceeds nVec";
            }
        }
        if (sum < 1 || sum > maxAf) {
            ctx->err = "rc_set_problem: pattern " + std::to_string(i) + " is outside the state space (need 1 <= sum <= maxAf)";
            return RC_ERR_ARG;
        }
    }
    ctx->P = P;
    ctx->maxAf = maxAf;
    ctx->nPatGlobal = nPat;
    ctx->patG.assign(patterns, patterns + (size_t)nPat * P);
    ctx->nVec.assign(nVec, nVec + P);
    ctx->cntG.assign((size_t)nPat, 0);
    ctx->sumCounts = 0;
    if (counts)
        for (int i = 0; i < nPat; ++i) {
            ctx->cntG[(size_t)i] = counts[i];
            ctx->sumCounts += counts[i];
        }
    ctx->totalSites = totalSites;
    ctx->rank = ctx->comm ? ctx->commRank : 0;          // a context that joined a communicator keeps its shard
    ctx->world = ctx->comm ? ctx->commRanks : 1;
    return build_shard(ctx);
}

int rc_set_shard(rc_ctx* ctx, int rank, int world) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return RC_ERR_ARG;                   // a multi-device context shards by itself
    ctx->err.clear();
    if (ctx->comm && (rank != ctx->commRank || world != ctx->commRanks)) {
        ctx->err = "rc_set_shard: the context's shard is fixed by rc_comm_init";
        return RC_ERR_ARG;
    }
    if (ctx->P == 0) { ctx->err = "rc_set_shard: call rc_set_problem first"; return RC_ERR_ARG; }
    if (world < 1 || rank < 0 || rank >= world) { ctx->err = "rc_set_shard: bad rank/world"; return RC_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    ctx->rank = rank;
    ctx->world = world;
    return build_shard(ctx);
}

int rc_set_times(rc_ctx* ctx, int T, const double* t) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_set_times(ctx->multi, T, t);
    ctx->err.clear();
    if (T < 1 || !t) { ctx->err = "rc_set_times: empty grid"; return RC_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    // validate the whole grid before anything in the context changes
    std::vector<double> dts((size_t)T);
    // deltaT = nextTime - t; t += deltaT (Core.hs:129,135) — the running clock is kept as the reference keeps it
    double clock = 0.0;
    for (int s = 0; s < T; ++s) {
        if (!std::isfinite(t[s])) { ctx->err = "rc_set_times: grid must be finite"; return RC_ERR_ARG; }
        if (s > 0 && !(t[s] >= t[s - 1])) { ctx->err = "rc_set_times: grid must be non-decreasing"; return RC_ERR_ARG; }
        double dt = t[s] - clock;
        dts[(size_t)s] = dt;
        if (dt > 0) clock = clock + dt;
    }
    CK(cudaDeviceSynchronize());        // launches of an earlier call may still read the old grid
    CK(ctx->dDts.ensure((size_t)T * sizeof(double)));
    CK(cudaMemcpy(ctx->dDts.p, dts.data(), (size_t)T * sizeof(double), cudaMemcpyHostToDevice));
    ctx->times.assign(t, t + T);
    ctx->dts.swap(dts);
    ctx->clear_plans();
    return RC_OK;
}

int rc_time_steps(int n0, int lingen, double tMax, double* out, int cap) {
    // getTimeSteps (Utils.hs:64-73)
    double tMin = 1.0 / (2.0 * (double)n0);
    double alpha = (double)lingen / (2.0 * (double)n0);
    double lg = std::log(1.0 + tMax / alpha);
    int nr = (int)std::floor(lg / std::log(1.0 + tMin / alpha));
    int n = nr - 1 > 0 ? nr - 1 : 0;
    for (int i = 1; i <= n && i <= cap; ++i)
        if (out) out[i - 1] = alpha * std::exp((double)i / (double)nr * lg) - alpha;
    return n;
}

int rc_eval_partial(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta,
                    const double* disc, int noShortcut, double* dPartial, double* dProbs, void* stream,
                    int32_t* outStatus) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return RC_ERR_ARG;                   // device pointers belong to one device: use rc_eval on a multi-device context
    ctx->err.clear();
    if (B < 0 || !evOff || !theta || !disc || !dPartial) { ctx->err = "rc_eval_partial: null argument"; return RC_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    std::vector<int32_t> ns((size_t)std::max(B, 1), noShortcut ? 1 : 0);
    return eval_impl(ctx, B, ev, evOff, theta, disc, ns.data(), dPartial, dProbs, (cudaStream_t)stream, outStatus);
}

int rc_finalize(rc_ctx* ctx, int B, const double* dPartialReduced, double siteRed, void* stream, double* outLogl) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return RC_ERR_ARG;
    ctx->err.clear();
    if (B <= 0) return RC_OK;
    if (!dPartialReduced || !outLogl) { ctx->err = "rc_finalize: null argument"; return RC_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(ctx->dLogl.ensure((size_t)B * sizeof(double)));
    CK(ctx->hOut.ensure((size_t)B * sizeof(double)));
    double other = (double)(ctx->totalSites - ctx->sumCounts);     // MaxUtils.hs:90
    CK(rc_launch_finalize(dPartialReduced, B, other, siteRed, (double*)ctx->dLogl.p, st));
    CK(cudaMemcpyAsync(ctx->hOut.p, ctx->dLogl.p, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::memcpy(outLogl, ctx->hOut.p, (size_t)B * sizeof(double));
    ctx->stats.launches += 1;
    return collect_stats(ctx);
}

int rc_eval(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
            int noShortcut, double siteRed, double* outProbs, double* outLogl, int32_t* outStatus) {
    std::vector<int32_t> ns((size_t)std::max(B, 1), noShortcut ? 1 : 0);
    return rc_eval_models(ctx, B, ev, evOff, theta, disc, ns.data(), siteRed, outProbs, outLogl, outStatus);
}

int rc_eval_models(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
                   const int32_t* noShortcut, double siteRed, double* outProbs, double* outLogl, int32_t* outStatus) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_eval_models(ctx->multi, B, ev, evOff, theta, disc, noShortcut, siteRed, outProbs, outLogl, outStatus);
    ctx->err.clear();
    if (B < 0 || !evOff || !theta || !disc || !outLogl || !noShortcut) { ctx->err = "rc_eval: null argument"; return RC_ERR_ARG; }
    if (B == 0) return RC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dPartial.ensure((size_t)2 * B * sizeof(double)));
    double* dProbs = nullptr;
    size_t probBytes = (size_t)B * std::max(ctx->nPatGlobal, 1) * sizeof(double);
    if (outProbs) {
        CK(ctx->dProbs.ensure(probBytes));
        dProbs = (double*)ctx->dProbs.p;
    }
    int rc = eval_impl(ctx, B, ev, evOff, theta, disc, noShortcut, (double*)ctx->dPartial.p, dProbs, ctx->stream, outStatus);
    if (ctx->comm) {
        // one rank of a communicator: the partial sums (and probabilities) of all pattern shards are summed by NCCL on the
        // evaluation stream.  A rank whose evaluation failed still takes part — with NaNs — so that its peers do not wait
        // for it forever; it then reports its own error.
        std::string nerr;
        const rc::NcclApi* api = rc::nccl_api(nerr);
        if (!api) { ctx->err = nerr; return RC_ERR_CUDA; }
        if (rc != RC_OK) {
            std::vector<double> nanv((size_t)2 * B, std::nan(""));
            cudaMemcpyAsync(ctx->dPartial.p, nanv.data(), nanv.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
        int nr = api->AllReduce(ctx->dPartial.p, ctx->dPartial.p, (size_t)2 * B, rc::kNcclFloat64, rc::kNcclSum, ctx->comm, ctx->stream);
        if (nr == 0 && dProbs && ctx->nPatGlobal > 0)
            nr = api->AllReduce(dProbs, dProbs, (size_t)B * ctx->nPatGlobal, rc::kNcclFloat64, rc::kNcclSum, ctx->comm, ctx->stream);
        if (nr != 0 && rc == RC_OK) { ctx->err = std::string("ncclAllReduce: ") + api->GetErrorString(nr); rc = RC_ERR_CUDA; }
        ctx->stats.launches += dProbs ? 2 : 1;
    }
    if (rc != RC_OK) return rc;
    if (outProbs && ctx->nPatGlobal > 0) {
        CK(ctx->hOut.ensure(std::max(probBytes, (size_t)B * sizeof(double))));
        CK(cudaMemcpyAsync(ctx->hOut.p, dProbs, probBytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::memcpy(outProbs, ctx->hOut.p, probBytes);
    }
    return rc_finalize(ctx, B, (const double*)ctx->dPartial.p, siteRed, ctx->stream, outLogl);
}

int rc_get_prob(rc_ctx* ctx, int P, int maxAf, const int32_t* nVec, const int32_t* config, const rc_event* ev, int nEv,
                double theta, const double* disc, int noShortcut, double* outProb) {
    if (!ctx) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_get_prob(ctx->multi, P, maxAf, nVec, config, ev, nEv, theta, disc, noShortcut, outProb);
    ctx->err.clear();
    if (ctx->times.empty()) { ctx->err = "rc_get_prob: rc_set_times has not been called"; return RC_ERR_ARG; }
    if (!ctx->child) {
        int rc = rc_create(ctx->device, &ctx->child);
        if (rc != RC_OK) { ctx->err = "rc_get_prob: cannot create scratch context"; return rc; }
    }
    rc_ctx* c = ctx->child;
    int64_t zero = 0;
    int rc = rc_set_problem(c, P, maxAf, 1, config, nVec, &zero, 0);
    if (rc == RC_OK) rc = rc_set_times(c, (int)ctx->times.size(), ctx->times.data());
    int32_t off[2] = {0, nEv};
    int32_t status = 0;
    double logl = 0, prob = 0;
    if (rc == RC_OK) rc = rc_eval(c, 1, ev, off, &theta, disc, noShortcut, 1.0, &prob, &logl, &status);
    if (rc != RC_OK) { ctx->err = c->err; return rc; }
    if (status != RC_OK) { ctx->err = c->err; return status; }
    *outProb = prob;
    return RC_OK;
}

int rc_set_option(rc_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_set_option(ctx->multi, name, value);
    ctx->err.clear();
    std::string n(name);
    if (n == "keyed") ctx->optKeyed = value != 0.0;
    else if (n == "pattern_mode") {
        const int v = value != 0.0;
        if (v != ctx->optPattern) {
            // plans depend on it: finish what is queued, then drop them
            cudaSetDevice(ctx->device);
            cudaDeviceSynchronize();
            ctx->clear_plans();
            ctx->optPattern = v;
        }
    }
    else if (n == "fuse_last") ctx->optFuseLast = value != 0.0;
    else if (n == "pattern_embed") {
        const int v = value != 0.0;
        if (v != ctx->optPatternEmbed) {
            cudaSetDevice(ctx->device);
            cudaDeviceSynchronize();
            ctx->clear_plans();
            ctx->optPatternEmbed = v;
        }
    }
    else if (n == "graphs") { ctx->optGraphs = value != 0.0; if (!ctx->optGraphs) { cudaSetDevice(ctx->device); cudaDeviceSynchronize(); ctx->clear_graphs(); } }
    else if (n == "plan_cache_max") ctx->optPlanCacheMax = std::max(1, (int)value);
    else if (n == "plan_cache_mb") ctx->optPlanCacheMB = std::max(1.0, value);
    else if (n == "ktab_model_mb") { ctx->optKtabModelMB = value; ctx->chunkHint = 256; }
    else if (n == "big_smem_kb") {
        if ((int)value != ctx->optBigSmemKB) { cudaSetDevice(ctx->device); cudaDeviceSynchronize(); ctx->clear_plans(); ctx->optBigSmemKB = (int)value; }
    }
    else if (n == "ktab_total_mb") { ctx->optKtabTotalMB = value; ctx->chunkHint = 256; }
    else { ctx->err = "rc_set_option: unknown option " + n; return RC_ERR_ARG; }
    return RC_OK;
}

int rc_last_stats(rc_ctx* ctx, rc_stats* out) {
    if (!ctx || !out) return RC_ERR_ARG;
    if (ctx->multi) return rc_multi_last_stats(ctx->multi, out);
    int rc = collect_stats(ctx);
    *out = ctx->stats;
    return rc;
}

int rc_fp64_peak(rc_ctx* ctx, double* outTflops) {
    if (!ctx || !outTflops) return RC_ERR_ARG;
    if (ctx->multi) return rc_fp64_peak(rc_multi_first(ctx->multi), outTflops);
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->smCount * 16, iters = 4096;
    double* d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * 256 * sizeof(double)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, ctx->stream));
        CK(rc_launch_fp64_peak(d, blocks, iters, ctx->stream));
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *outTflops = best;
    return RC_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
int rc_comm_unique_id(void* out128) {
    if (!out128) return RC_ERR_ARG;
    std::string e;
    const rc::NcclApi* api = rc::nccl_api(e);
    if (!api) return RC_ERR_CUDA;
    rc::NcclUniqueId id;
    if (api->GetUniqueId(&id) != 0) return RC_ERR_CUDA;
    std::memcpy(out128, id.internal, sizeof id.internal);
    return RC_OK;
}

int rc_comm_init(rc_ctx* ctx, int nRanks, int rank, const void* uniqueId128) {
    if (!ctx || ctx->multi) return RC_ERR_ARG;
    ctx->err.clear();
    if (nRanks < 1 || rank < 0 || rank >= nRanks || !uniqueId128) { ctx->err = "rc_comm_init: bad rank / size / id"; return RC_ERR_ARG; }
    if (ctx->comm) { ctx->err = "rc_comm_init: the context already belongs to a communicator"; return RC_ERR_ARG; }
    const rc::NcclApi* api = rc::nccl_api(ctx->err);
    if (!api) return RC_ERR_CUDA;
    CK(cudaSetDevice(ctx->device));
    rc::NcclUniqueId id;
    std::memcpy(id.internal, uniqueId128, sizeof id.internal);
    rc::NcclComm comm = nullptr;
    int nr = api->CommInitRank(&comm, nRanks, id, rank);
    if (nr != 0) { ctx->err = std::string("ncclCommInitRank: ") + api->GetErrorString(nr); return RC_ERR_CUDA; }
    ctx->comm = comm;
    ctx->commRanks = nRanks;
    ctx->commRank = rank;
    ctx->rank = rank;
    ctx->world = nRanks;
    return ctx->P > 0 ? build_shard(ctx) : RC_OK;
}

int rc_comm_allreduce(rc_ctx* ctx, double* dBuf, int64_t n, void* stream) {
    if (!ctx || ctx->multi) return RC_ERR_ARG;
    ctx->err.clear();
    if (!ctx->comm) { ctx->err = "rc_comm_allreduce: rc_comm_init has not been called"; return RC_ERR_ARG; }
    if (n <= 0) return RC_OK;
    if (!dBuf) { ctx->err = "rc_comm_allreduce: null buffer"; return RC_ERR_ARG; }
    const rc::NcclApi* api = rc::nccl_api(ctx->err);
    if (!api) return RC_ERR_CUDA;
    CK(cudaSetDevice(ctx->device));
    int nr = api->AllReduce(dBuf, dBuf, (size_t)n, rc::kNcclFloat64, rc::kNcclSum, ctx->comm, (cudaStream_t)stream);
    if (nr != 0) { ctx->err = std::string("ncclAllReduce: ") + api->GetErrorString(nr); return RC_ERR_CUDA; }
    return RC_OK;
}

int rc_create_multi(int nDevices, const int32_t* devices, rc_ctx** out) {
    if (!out) return RC_ERR_ARG;
    *out = nullptr;
    RcMulti* m = nullptr;
    int rc = rc_multi_create(nDevices, devices, &m);
    if (rc != RC_OK) return rc;
    rc_ctx* ctx = new rc_ctx();
    ctx->multi = m;
    *out = ctx;
    return RC_OK;
}

int rc_device_count(rc_ctx* ctx) {
    if (!ctx) return RC_ERR_ARG;
    return ctx->multi ? rc_multi_count(ctx->multi) : 1;
}

// ---- host-side mirrors -------------------------------------------------------------------------
int rc_all_configs(int maxAf, int P, const int32_t* nVec, int32_t* out, int cap) {
    if (P < 1 || maxAf < 1 || !nVec) return RC_ERR_ARG;
    auto v = rc::all_configs(maxAf, std::vector<int>(nVec, nVec + P));
    if (out)
        for (size_t i = 0; i < v.size() && (int)i < cap; ++i)
            for (int k = 0; k < P; ++k) out[i * P + k] = v[i][k];
    return (int)v.size();
}

int rc_validate_model(int P, const double* disc, const rc_event* ev, int nEv, char* err, int errCap) {
    if (P < 1 || P > RC_MAX_P || !disc || nEv < 0) return RC_ERR_ARG;
    std::string msg;
    int rc = validate_model(P, disc, sort_events(ev, nEv), msg);
    if (err && errCap > 0) std::snprintf(err, (size_t)errCap, "%s", msg.c_str());
    return rc;
}

double rc_reg_penalty(int P, double reg, const rc_event* ev, int nEv) {
    // getRegularizationPenalty (StateSpace.hs:199-221): sizes tracked through SetPopSize events only
    if (P < 1 || nEv < 0 || (nEv > 0 && !ev)) return std::nan("");
    std::vector<rc_event> s = sort_events(ev, nEv);
    std::vector<double> size((size_t)std::max(P, 1), 1.0);
    auto term = [reg](double from, double to) {
        double ratio = to > from ? to / from : from / to;
        return reg * ((ratio - 1.0) * (ratio - 1.0));
    };
    double total = 0.0;
    // the reference indexes its size vector with V.! / V.// and fails on a bad branch (StateSpace.hs:205-215); here a bad index
    // yields NaN, which minFunc turns into the 1e20 penalty (MaxUtils.hs:72-75) and the Python layer into RarecoalModelException
    auto bad = [P](int i) { return i < 0 || i >= P; };
    for (const rc_event& e : s) {
        if ((e.type == RC_EV_POPSIZE && bad(e.k)) || (e.type == RC_EV_JOIN && (bad(e.k) || bad(e.l))))
            return std::nan("");
        if (e.type == RC_EV_POPSIZE) {
            if (e.t != 0.0) total = total + term(size[(size_t)e.k], e.v);
            size[(size_t)e.k] = e.v;
        } else if (e.type == RC_EV_JOIN) {
            total = total + term(size[(size_t)e.l], size[(size_t)e.k]);
        }
    }
    return total;
}

double rc_choose(int n, int k) { return rc::choose(n, k); }

int rc_desugar_growth(int T, const double* times, const rc_event* ev, int nEv, rc_event* out, int cap) {
    if (T < 0 || (T > 0 && !times) || nEv < 0 || (nEv > 0 && !ev)) return RC_ERR_ARG;
    std::vector<rc_event> r = desugar_growth(std::vector<double>(times, times + T), sort_events(ev, nEv));
    if (out) for (size_t i = 0; i < r.size() && (int)i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

int rc_shard_patterns(int P, int nPat, const int32_t* patterns, int rank, int world, int32_t* out, int cap) {
    if (P < 1 || nPat < 0 || (nPat > 0 && !patterns) || world < 1 || rank < 0 || rank >= world) return RC_ERR_ARG;
    std::vector<int> mine = shard_patterns(P, nPat, patterns, rank, world);
    if (out)
        for (size_t i = 0; i < mine.size() && (int)i < cap; ++i) out[i] = mine[i];
    return (int)mine.size();
}

static thread_local uint64_t tlsPlanChecksum = 0;

uint64_t rc_plan_last_checksum(void) { return tlsPlanChecksum; }

int rc_plan_stats(int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec, int T, const double* times,
                  const rc_event* ev, int nEv, int noShortcut, int64_t* stateSteps, int64_t* slots0, int32_t* maxLive,
                  int32_t* nEpochs, int32_t* nBins) {
    if (P < 1 || P > RC_MAX_P || maxAf < 1 || maxAf > RC_MAX_AF || nPat < 1 || !patterns || !nVec || T < 1 || !times)
        return RC_ERR_ARG;
    rc_ctx tmp;   // host-only use: no CUDA call is made on this object
    tmp.P = P;
    tmp.maxAf = maxAf;
    tmp.times.assign(times, times + T);
    std::vector<double> disc((size_t)P, 1.0);
    ModelHost mh;
    resolve_model(&tmp, ev, nEv, disc.data(), noShortcut, mh);
    if (mh.status != RC_OK) return mh.status;
    rc::PlanHost h;
    const auto tp0 = std::chrono::steady_clock::now();
    std::string e = rc::build_plan(P, maxAf, nPat, patterns, nVec, mh.sev, mh.gaps, mh.hasSplit, 227 * 1024, h);
    if (!e.empty()) { std::fprintf(stderr, "%s\n", e.c_str()); return RC_ERR_UNSUPPORTED; }
    const auto tp1 = std::chrono::steady_clock::now();
    e = rc::check_plan(h);
    if (std::getenv("RC_PLAN_TIMING"))
        std::fprintf(stderr, "[rc plan] build_plan %.1f ms, check_plan %.1f ms\n", std::chrono::duration<double, std::milli>(tp1 - tp0).count(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp1).count());
    if (!e.empty()) { std::fprintf(stderr, "%s\n", e.c_str()); return RC_ERR_UNSUPPORTED; }
    tlsPlanChecksum = rc::plan_checksum(h);
    if (stateSteps) *stateSteps = count_state_steps(mh, h, T);
    if (slots0) *slots0 = h.epochSlots[0];
    if (maxLive) *maxLive = h.maxLive;
    if (nEpochs) *nEpochs = h.nEpochs;
    if (nBins) *nBins = h.nBins + h.nFastBins;
    return RC_OK;
}

rc_statespace* rc_ss_create(int P, int maxAf) {
    rc::StateSpace* s = rc::make_state_space(P, maxAf);
    if (!s) return nullptr;
    rc_statespace* h = new rc_statespace();
    h->ss = s;
    return h;
}
void rc_ss_destroy(rc_statespace* ss) {
    if (!ss) return;
    delete ss->ss;
    delete ss;
}
int rc_ss_nr_states(const rc_statespace* ss) { return ss ? ss->ss->nrStates : RC_ERR_ARG; }
int rc_ss_state_to_id(const rc_statespace* ss, const int32_t* x) { return ss && x ? ss->ss->state_to_id(x) : RC_ERR_ARG; }
int rc_ss_id_to_state(const rc_statespace* ss, int id, int32_t* x) {
    if (!ss || !x || id < 0 || id >= ss->ss->nrStates) return RC_ERR_ARG;
    for (int k = 0; k < ss->ss->P; ++k) x[k] = ss->ss->states[(size_t)id * ss->ss->P + k];
    return RC_OK;
}
int rc_ss_x1up(const rc_statespace* ss, int id, int32_t* out) {
    if (!ss || !out || id < 0 || id >= ss->ss->nrStates) return RC_ERR_ARG;
    for (int k = 0; k < ss->ss->P; ++k) out[k] = ss->ss->x1up[(size_t)id * ss->ss->P + k];
    return RC_OK;
}
int rc_ss_x1(const rc_statespace* ss, int k) { return ss && k >= 0 && k < ss->ss->P ? ss->ss->x1[(size_t)k] : RC_ERR_ARG; }
int rc_ss_fill_up(const rc_statespace* ss, const int32_t* ids, int n, int32_t* out, int cap) {
    if (!ss || (n > 0 && !ids)) return RC_ERR_ARG;
    std::vector<int> r = ss->ss->fill_up(std::vector<int>(ids, ids + n));
    if (out) for (size_t i = 0; i < r.size() && (int)i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

}  // extern "C"
