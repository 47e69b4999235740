// rc_multi.cpp — rc_create_multi: one context that drives N GPUs of this process.
//
// The reference fans one likelihood out over the cores of one machine inside the library call (parMap over patterns,
// MaxUtils.hs:108-110); this is the same seam over GPUs, behind the same rc_eval call:
//   * models sharded  (B >= N): device d evaluates models [B d / N, B (d+1) / N) on all patterns and writes its slice of the
//     caller's output arrays — independent objects, no collective;
//   * patterns sharded (B < N, e.g. a single `logl`): device d evaluates its pattern shard of every model; the per-model
//     partial sums (sum cnt log p, sum p) — and the probabilities when asked for — are summed across the devices with
//     ncclAllReduce over NVLink, each device on its own evaluation stream, and device 0 finalises into the caller's arrays.
// One host thread per device issues that device's work (the planner and the launch chain are host code), so the devices
// run concurrently.  Built on the public C ABI of the per-device contexts only.
#include "rc_multi.h"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has = false, quit = false, done = true;
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has || quit; });
            if (quit) return;
            has = false;
            lk.unlock();
            job();
            lk.lock();
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<void()> j) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has = true;
        done = false;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
    }
};

}  // namespace

struct RcMulti {
    int n = 0;
    std::vector<int> devices;
    std::vector<rc_ctx*> full, shard;            // shard: members of one NCCL communicator, created on first use
    std::vector<Worker*> workers;
    int mode = 0;                                // "multi_shard": 0 auto, 1 models, 2 patterns
    int lastMode = 0;
    std::string err;
    // what the children were given, for the lazily created pattern-shard contexts
    bool haveProblem = false, haveTimes = false;
    int P = 0, maxAf = 0, nPat = 0;
    std::vector<int32_t> pats, nVec;
    std::vector<int64_t> counts;
    int64_t total = 0;
    std::vector<double> times;
    std::vector<std::pair<std::string, double>> options;
    std::vector<std::vector<double>> scratchProbs, scratchLogl;
    std::vector<std::vector<int32_t>> scratchStatus;
    rc_stats stats;

    // runs fn(d) for every device on that device's thread; returns the first non-zero code (and keeps its message)
    int run_all(const std::function<int(int)>& fn, const std::vector<rc_ctx*>& ctxs) {
        std::vector<int> rcs((size_t)n, 0);
        for (int d = 0; d < n; ++d) workers[(size_t)d]->submit([&, d] { rcs[(size_t)d] = fn(d); });
        for (int d = 0; d < n; ++d) workers[(size_t)d]->wait();
        for (int d = 0; d < n; ++d)
            if (rcs[(size_t)d] != RC_OK) {
                err = "device " + std::to_string(devices[(size_t)d]) + ": " + (ctxs[(size_t)d] ? rc_last_error(ctxs[(size_t)d]) : "");
                return rcs[(size_t)d];
            }
        return RC_OK;
    }
};

static int ensure_shards(RcMulti* m) {
    if (!m->shard.empty()) return RC_OK;
    unsigned char uid[128];
    int rc = rc_comm_unique_id(uid);
    if (rc != RC_OK) { m->err = "NCCL is not available: the pattern-sharded multi-GPU path needs libnccl.so.2"; return rc; }
    std::vector<rc_ctx*> ctxs((size_t)m->n, nullptr);
    for (int d = 0; d < m->n; ++d) {
        rc = rc_create(m->devices[(size_t)d], &ctxs[(size_t)d]);
        if (rc != RC_OK) {
            for (rc_ctx* c : ctxs) if (c) rc_destroy(c);
            m->err = "rc_create failed on device " + std::to_string(m->devices[(size_t)d]);
            return rc;
        }
    }
    // ncclCommInitRank blocks until every rank has joined: all devices at once, each on its own thread
    rc = m->run_all([&](int d) {
        rc_ctx* c = ctxs[(size_t)d];
        for (auto& o : m->options) rc_set_option(c, o.first.c_str(), o.second);
        int r = rc_comm_init(c, m->n, d, uid);
        if (r == RC_OK && m->haveProblem)
            r = rc_set_problem(c, m->P, m->maxAf, m->nPat, m->pats.data(), m->nVec.data(), m->counts.data(), m->total);
        if (r == RC_OK && m->haveTimes) r = rc_set_times(c, (int)m->times.size(), m->times.data());
        return r;
    }, ctxs);
    if (rc != RC_OK) {
        for (rc_ctx* c : ctxs) if (c) rc_destroy(c);
        return rc;
    }
    m->shard = ctxs;
    return RC_OK;
}

int rc_multi_create(int nDevices, const int32_t* devices, RcMulti** out) {
    if (nDevices < 1 || nDevices > 64 || !out) return RC_ERR_ARG;
    RcMulti* m = new RcMulti();
    m->n = nDevices;
    std::memset(&m->stats, 0, sizeof m->stats);
    for (int d = 0; d < nDevices; ++d) m->devices.push_back(devices ? devices[d] : d);
    for (int d = 0; d < nDevices; ++d) {
        rc_ctx* c = nullptr;
        int rc = rc_create(m->devices[(size_t)d], &c);
        if (rc != RC_OK) {
            for (rc_ctx* x : m->full) rc_destroy(x);
            delete m;
            return rc;
        }
        m->full.push_back(c);
    }
    for (int d = 0; d < nDevices; ++d) {
        Worker* w = new Worker();
        w->th = std::thread([w] { w->loop(); });
        m->workers.push_back(w);
    }
    m->scratchProbs.resize((size_t)nDevices);
    m->scratchLogl.resize((size_t)nDevices);
    m->scratchStatus.resize((size_t)nDevices);
    *out = m;
    return RC_OK;
}

void rc_multi_destroy(RcMulti* m) {
    if (!m) return;
    // communicator members are torn down together (ncclCommDestroy), each on its device's thread
    if (!m->shard.empty()) m->run_all([&](int d) { rc_destroy(m->shard[(size_t)d]); return 0; }, m->shard);
    for (Worker* w : m->workers) {
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        w->th.join();
        delete w;
    }
    for (rc_ctx* c : m->full) rc_destroy(c);
    delete m;
}

const char* rc_multi_last_error(RcMulti* m) { return m->err.c_str(); }
int rc_multi_count(RcMulti* m) { return m->n; }
rc_ctx* rc_multi_first(RcMulti* m) { return m->full[0]; }

int rc_multi_set_problem(RcMulti* m, int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec,
                         const int64_t* counts, int64_t totalSites) {
    m->err.clear();
    if (P < 1 || nPat < 0 || !nVec || (nPat > 0 && !patterns)) { m->err = "rc_set_problem: bad arguments"; return RC_ERR_ARG; }
    m->P = P; m->maxAf = maxAf; m->nPat = nPat; m->total = totalSites;
    m->pats.assign(patterns, patterns + (size_t)nPat * P);
    m->nVec.assign(nVec, nVec + P);
    m->counts.assign((size_t)nPat, 0);
    if (counts) m->counts.assign(counts, counts + nPat);
    m->haveProblem = false;
    int rc = m->run_all([&](int d) { return rc_set_problem(m->full[(size_t)d], P, maxAf, nPat, patterns, nVec, counts, totalSites); }, m->full);
    if (rc == RC_OK && !m->shard.empty())
        rc = m->run_all([&](int d) { return rc_set_problem(m->shard[(size_t)d], P, maxAf, nPat, patterns, nVec, counts, totalSites); }, m->shard);
    m->haveProblem = rc == RC_OK;
    return rc;
}

int rc_multi_set_times(RcMulti* m, int T, const double* t) {
    m->err.clear();
    if (T < 1 || !t) { m->err = "rc_set_times: empty grid"; return RC_ERR_ARG; }
    m->haveTimes = false;
    int rc = m->run_all([&](int d) { return rc_set_times(m->full[(size_t)d], T, t); }, m->full);
    if (rc == RC_OK && !m->shard.empty()) rc = m->run_all([&](int d) { return rc_set_times(m->shard[(size_t)d], T, t); }, m->shard);
    if (rc == RC_OK) { m->times.assign(t, t + T); m->haveTimes = true; }
    return rc;
}

int rc_multi_set_option(RcMulti* m, const char* name, double value) {
    m->err.clear();
    if (std::string(name) == "multi_shard") {
        if (value != 0.0 && value != 1.0 && value != 2.0) { m->err = "multi_shard: 0 (auto), 1 (models) or 2 (patterns)"; return RC_ERR_ARG; }
        m->mode = (int)value;
        return RC_OK;
    }
    int rc = m->run_all([&](int d) { return rc_set_option(m->full[(size_t)d], name, value); }, m->full);
    if (rc == RC_OK && !m->shard.empty()) rc = m->run_all([&](int d) { return rc_set_option(m->shard[(size_t)d], name, value); }, m->shard);
    if (rc == RC_OK) m->options.emplace_back(name, value);
    return rc;
}

static void fold_stats(RcMulti* m, const std::vector<rc_ctx*>& ctxs, const std::vector<char>& used) {
    rc_stats acc;
    std::memset(&acc, 0, sizeof acc);
    acc.plan_cache_hit = 1;
    for (int d = 0; d < m->n; ++d) {
        if (!used[(size_t)d]) continue;
        rc_stats s;
        if (rc_last_stats(ctxs[(size_t)d], &s) != RC_OK) continue;
        acc.state_steps += s.state_steps; acc.slots += s.slots; acc.fp64_instr += s.fp64_instr;
        acc.kernel_ms = std::max(acc.kernel_ms, s.kernel_ms); acc.traj_ms = std::max(acc.traj_ms, s.traj_ms);
        acc.tables_ms = std::max(acc.tables_ms, s.tables_ms); acc.total_ms = std::max(acc.total_ms, s.total_ms);
        acc.plan_build_ms = std::max(acc.plan_build_ms, s.plan_build_ms);
        acc.host_ms = std::max(acc.host_ms, s.host_ms); acc.ktab_mb += s.ktab_mb; acc.chunks = std::max(acc.chunks, s.chunks); acc.models_selfcontained += s.models_selfcontained;
        acc.launches += s.launches; acc.plan_misses += s.plan_misses; acc.plan_evictions += s.plan_evictions;
        acc.plan_cache_hit = acc.plan_cache_hit && s.plan_cache_hit;
        acc.n_epochs = s.n_epochs; acc.n_steps = s.n_steps;
    }
    m->stats = acc;
}

int rc_multi_eval_models(RcMulti* m, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
                         const int32_t* noShortcut, double siteRed, double* outProbs, double* outLogl, int32_t* outStatus) {
    m->err.clear();
    if (B < 0 || !evOff || !theta || !disc || !outLogl || !noShortcut) { m->err = "rc_eval: null argument"; return RC_ERR_ARG; }
    if (!m->haveProblem) { m->err = "rc_set_problem has not been called"; return RC_ERR_ARG; }
    if (!m->haveTimes) { m->err = "rc_set_times has not been called"; return RC_ERR_ARG; }
    if (B == 0) return RC_OK;
    const int n = m->n, P = m->P;
    const bool byPattern = n > 1 && (m->mode == 2 || (m->mode == 0 && B < n));
    m->lastMode = byPattern ? 2 : 1;
    std::vector<char> used((size_t)n, 0);
    if (!byPattern) {
        int rc = m->run_all([&](int d) {
            const int b0 = (int)((long long)B * d / n), b1 = (int)((long long)B * (d + 1) / n);
            if (b1 == b0) return (int)RC_OK;
            used[(size_t)d] = 1;
            return rc_eval_models(m->full[(size_t)d], b1 - b0, ev, evOff + b0, theta + b0, disc + (size_t)b0 * P, noShortcut + b0, siteRed,
                                  outProbs ? outProbs + (size_t)b0 * m->nPat : nullptr, outLogl + b0, outStatus ? outStatus + b0 : nullptr);
        }, m->full);
        fold_stats(m, m->full, used);
        return rc;
    }
    int rc = ensure_shards(m);
    if (rc != RC_OK) return rc;
    for (int d = 1; d < n; ++d) {          // every rank takes part in the same collectives; ranks > 0 write to scratch
        if (outProbs) m->scratchProbs[(size_t)d].resize((size_t)B * std::max(m->nPat, 1));
        m->scratchLogl[(size_t)d].resize((size_t)B);
        m->scratchStatus[(size_t)d].resize((size_t)B);
    }
    rc = m->run_all([&](int d) {
        used[(size_t)d] = 1;
        double* pr = !outProbs ? nullptr : d == 0 ? outProbs : m->scratchProbs[(size_t)d].data();
        return rc_eval_models(m->shard[(size_t)d], B, ev, evOff, theta, disc, noShortcut, siteRed, pr,
                              d == 0 ? outLogl : m->scratchLogl[(size_t)d].data(),
                              d == 0 ? outStatus : m->scratchStatus[(size_t)d].data());
    }, m->shard);
    fold_stats(m, m->shard, used);
    return rc;
}

int rc_multi_get_prob(RcMulti* m, int P, int maxAf, const int32_t* nVec, const int32_t* config, const rc_event* ev, int nEv,
                      double theta, const double* disc, int noShortcut, double* outProb) {
    int rc = rc_get_prob(m->full[0], P, maxAf, nVec, config, ev, nEv, theta, disc, noShortcut, outProb);
    if (rc != RC_OK) m->err = rc_last_error(m->full[0]);
    return rc;
}

int rc_multi_last_stats(RcMulti* m, rc_stats* out) {
    *out = m->stats;
    return RC_OK;
}
