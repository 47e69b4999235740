// rarecoal_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A literal CPU restatement of the reference's likelihood hot path, used only as the
// checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs.  Nothing under rarecoal_b200/ may include, link or call this file.
//
// Parity status: PINNED.  The restatement reproduces every known answer the reference
// ships for this path (tests/test_oracle_golden.py):
//   * src/RarecoalLib/Core/Test.hs:177,181 (+ the three commented-out values :178-180)
//   * exampleData/fourPop.m4.mcmc.paramEstimates.txt:4  (whole-path Score = -logl)
//   * src/RarecoalLib/StateSpace/Test.hs:82-95 (fillUpStateSpace), :66-80 (enumeration)
// The Haskell reference itself cannot be built in this image (no ghc/cabal/stack).
//
// Every function cites the reference file:line it follows (paths relative to the
// reference checkout).  Floating-point operation ORDER follows the Haskell source
// (left folds, `1.0 - exp(-x)` rather than expm1, GHC's (^) by repeated squaring);
// build with -ffp-contract=off so the compiler does not fuse multiply-adds.
//
// Only liberty taken (same as SURVEY.md §8d): the per-step `VM.set bTemp 0` over all
// nrStates (Core.hs:247) is not executed literally — entries outside the live list are
// never read before being overwritten, so results are identical; the dense `b` vector
// and the live-list order (fillUpStateSpace + nub) are kept.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <mutex>
#include <thread>

namespace {

typedef std::vector<int> IVec;

// ---------------------------------------------------------------------------------------
// Utils.hs:121-127  allSubsets n k (list-monad order: subsets(n-1,k) ++ map (++[n]) subsets(n-1,k-1))
void all_subsets(int n, int k, std::vector<IVec>& out) {
    if (k == 0) { out.push_back(IVec()); return; }
    if (!(n > 0)) return;                       // `True <- return $ n > 0` fails the monad -> []
    all_subsets(n - 1, k, out);
    std::vector<IVec> rest;
    all_subsets(n - 1, k - 1, rest);
    for (auto& s : rest) { s.push_back(n); out.push_back(s); }
}

// Utils.hs:106-112  allPatternsForAF
void all_patterns_for_af(int nrPops, int af, std::vector<IVec>& out) {
    std::vector<IVec> subs;
    all_subsets(nrPops + af - 1, nrPops - 1, subs);
    for (auto& s : subs) {
        IVec ext; ext.push_back(0);
        ext.insert(ext.end(), s.begin(), s.end());
        ext.push_back(nrPops + af);
        IVec diffs;
        for (size_t i = 0; i + 1 < ext.size(); ++i) diffs.push_back(ext[i + 1] - ext[i] - 1);
        out.push_back(diffs);
    }
}

// Utils.hs:100-104  computeAllConfigs
std::vector<IVec> compute_all_configs(int maxAf, const IVec& nVec) {
    int P = (int)nVec.size();
    std::vector<IVec> all, out;
    for (int af = 1; af <= maxAf; ++af) all_patterns_for_af(P, af, all);
    for (auto& v : all) {
        bool ok = true;
        for (int k = 0; k < P; ++k) if (v[k] > nVec[k]) { ok = false; break; }
        if (ok) out.push_back(v);
    }
    return out;
}

// Utils.hs:208-215  choose / chooseCont : product [ (n+1-j)/j | j<-[1..k] ]  (left fold from 1)
double choose_int(int n, int k) {
    if (k == 0) return 1.0;
    double p = 1.0;
    for (int j = 1; j <= k; ++j) p = p * ((double)(n + 1 - j) / (double)j);
    return p;
}
double choose_cont(double n, int k) {
    if (k == 0) return 1.0;
    double p = 1.0;
    for (int j = 1; j <= k; ++j) p = p * ((n + 1.0 - (double)j) / (double)j);
    return p;
}

// GHC.Real (^) — exponentiation by repeated squaring, same multiplication order
// (used at Core.hs:216 `m ^ s * (1.0 - m) ^ (xl - s)`).
double hs_pow_acc(double x, int y, double z) {
    for (;;) {
        if ((y & 1) == 0) { x = x * x; y = y / 2; }
        else if (y == 1) return x * z;
        else { z = x * z; x = x * x; y = y / 2; }
    }
}
double hs_pow(double x, int y) {
    if (y == 0) return 1.0;
    for (;;) {
        if ((y & 1) == 0) { x = x * x; y = y / 2; }
        else if (y == 1) return x;
        else return hs_pow_acc(x * x, y / 2, x);
    }
}

// ---------------------------------------------------------------------------------------
// StateSpace.hs:20-28,73-124  JointStateSpace (ids = position in computeAllConfigs maxAf [maxAf]*P)
struct StateSpace {
    int P, maxAf, nrStates;
    std::vector<IVec> idToState;
    std::unordered_map<std::string, int> stateToId;
    std::vector<int> x1up;   // [nrStates][P], -1 if sum+1 > maxAf   (StateSpace.hs:77-79)
    std::vector<int> x1;     // [P]                                   (StateSpace.hs:80,123-124)
    static std::string key(const int* x, int P) {
        std::string s((size_t)P, '\0');
        for (int k = 0; k < P; ++k) s[k] = (char)x[k];
        return s;
    }
    int lookup(const int* x) const {
        auto it = stateToId.find(key(x, P));
        return it == stateToId.end() ? -1 : it->second;
    }
};

StateSpace* make_state_space(int P, int maxAf) {
    StateSpace* ss = new StateSpace();
    ss->P = P; ss->maxAf = maxAf;
    ss->idToState = compute_all_configs(maxAf, IVec(P, maxAf));
    ss->nrStates = (int)ss->idToState.size();
    for (int i = 0; i < ss->nrStates; ++i) ss->stateToId[StateSpace::key(ss->idToState[i].data(), P)] = i;
    ss->x1up.assign((size_t)ss->nrStates * P, -1);
    for (int i = 0; i < ss->nrStates; ++i) {
        IVec x = ss->idToState[i];
        int sum = 0; for (int k = 0; k < P; ++k) sum += x[k];
        for (int k = 0; k < P; ++k) {
            if (sum + 1 <= maxAf) { x[k] += 1; ss->x1up[(size_t)i * P + k] = ss->lookup(x.data()); x[k] -= 1; }
        }
    }
    ss->x1.resize(P);
    for (int k = 0; k < P; ++k) { IVec e(P, 0); e[k] = 1; ss->x1[k] = ss->lookup(e.data()); }
    return ss;
}

// StateSpace.hs:145-160  fillUpStateSpace: nub (concatMap expandPattern states); expandPattern is a
// list-monad foldM over k=0..P-1, so component 0 varies slowest.
void fill_up(const StateSpace& ss, const std::vector<int>& ids, std::vector<int>& out) {
    out.clear();
    std::vector<char> seen((size_t)ss.nrStates, 0);
    int P = ss.P;
    IVec cur(P), hi(P);
    for (int id : ids) {
        const IVec& x = ss.idToState[id];
        for (int k = 0; k < P; ++k) { hi[k] = x[k]; cur[k] = x[k] == 0 ? 0 : 1; }
        for (;;) {
            int y = ss.lookup(cur.data());
            if (!seen[y]) { seen[y] = 1; out.push_back(y); }
            int k = P - 1;
            for (; k >= 0; --k) {
                if (hi[k] == 0) continue;
                if (cur[k] < hi[k]) { cur[k]++; break; }
                cur[k] = 1;
            }
            if (k < 0) break;
        }
    }
}

// ---------------------------------------------------------------------------------------
// StateSpace.hs:30-51 ModelEvent / EventType
struct Event { double t; int32_t type; int32_t k; int32_t l; double v; };
enum { EV_JOIN = 0, EV_SPLIT = 1, EV_POPSIZE = 2, EV_FREEZE = 3 };

void set_err(char* buf, int cap, const std::string& s) {
    if (buf && cap > 0) { std::snprintf(buf, (size_t)cap, "%s", s.c_str()); }
}

// Core.hs:292-294 / StateSpace.hs:167-168: stable sortBy time
std::vector<Event> sorted_events(const Event* ev, int n) {
    std::vector<Event> v(ev, ev + n);
    std::stable_sort(v.begin(), v.end(), [](const Event& a, const Event& b) { return a.t < b.t; });
    return v;
}

// StateSpace.hs:163-197 validateModel.  Returns 0 ok, 1 = RarecoalModelException.
// Note the Split pattern match at :190 binds `Split l k m` — only index ranges are checked, so the
// naming swap has no effect.
int validate_model(int nDisc, const double* disc, const Event* ev, int nEv, char* err, int cap) {
    for (int i = 0; i < nEv; ++i) if (ev[i].t < 0) { set_err(err, cap, "Negative event times"); return 1; }
    for (int i = 0; i < nDisc; ++i) if (disc[i] <= 0 || disc[i] > 1) { set_err(err, cap, "illegal discovery Rate"); return 1; }
    std::vector<Event> s = sorted_events(ev, nEv);
    int n = nDisc;
    for (int i = 0; i < nEv; ++i) {
        const Event& e = s[i];
        switch (e.type) {
        case EV_JOIN: {
            int k = e.k, l = e.l;
            if (k >= n || l >= n || k < 0 || l < 0) { set_err(err, cap, "illegal branch indices in event"); return 1; }
            bool illegal = false;
            for (int j = i + 1; j < nEv; ++j) {
                const Event& r = s[j];
                if (r.type == EV_JOIN || r.type == EV_SPLIT) illegal = illegal || r.k == l || r.l == l;
                else illegal = illegal || r.k == l;
            }
            if (k == l || illegal) { set_err(err, cap, "Illegal join"); return 1; }
            break; }
        case EV_POPSIZE:
            if (e.k >= n || e.k < 0) { set_err(err, cap, "illegal branch indices in event"); return 1; }
            if (e.v <= 0) { set_err(err, cap, "Illegal population size"); return 1; }
            break;
        case EV_SPLIT:
            if (e.k >= n || e.l >= n || e.k < 0 || e.l < 0) { set_err(err, cap, "illegal branch indices in event"); return 1; }
            if (e.v < 0.0 || e.v > 1.0) { set_err(err, cap, "Illegal split rate"); return 1; }
            break;
        case EV_FREEZE:
            if (e.k >= n || e.k < 0) { set_err(err, cap, "illegal branch indices in event"); return 1; }
            break;
        default:
            set_err(err, cap, "unknown event type"); return 1;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// Core.hs:22-36  ModelState / CoalState
struct Coal {
    const StateSpace* ss;
    std::vector<double> a, b, bTemp, popSize;
    std::vector<char> freeze;
    std::vector<int> live;
    double d, t;
    long long stateSteps;    // instrumentation only: executed b(x) updates (SURVEY.md §8d unit)
};

// Core.hs:165-170
void pop_join_a(double* a, int k, int l) { double al = a[l], ak = a[k]; a[l] = 0.0; a[k] = al + ak; }
// Core.hs:195-200
void pop_split_a(double* a, int k, int l, double m) { double al = a[l], ak = a[k]; a[l] = (1.0 - m) * al; a[k] = m * al + ak; }

// Core.hs:172-188
void pop_join_b(Coal& c, int k, int l) {
    const StateSpace& ss = *c.ss; int P = ss.P;
    std::fill(c.bTemp.begin(), c.bTemp.end(), 0.0);
    std::vector<int> newIds; newIds.reserve(c.live.size());
    IVec x(P);
    for (int xId : c.live) {
        double oldProb = c.b[xId];
        x = ss.idToState[xId];
        int xl = x[l], xk = x[k];
        x[k] = xk + xl; x[l] = 0;
        int id2 = ss.lookup(x.data());
        c.bTemp[id2] = c.bTemp[id2] + oldProb;
        newIds.push_back(id2);
    }
    c.b = c.bTemp;
    // nub, then fillUp
    std::vector<int> uniq; std::vector<char> seen((size_t)ss.nrStates, 0);
    for (int id : newIds) if (!seen[id]) { seen[id] = 1; uniq.push_back(id); }
    fill_up(ss, uniq, c.live);
}

// Core.hs:202-221
void pop_split_b(Coal& c, int k, int l, double m) {
    const StateSpace& ss = *c.ss; int P = ss.P;
    std::fill(c.bTemp.begin(), c.bTemp.end(), 0.0);
    std::vector<int> newIds;
    IVec x(P);
    for (int xId : c.live) {
        double oldProb = c.b[xId];
        const IVec& x0 = ss.idToState[xId];
        int xl = x0[l], xk = x0[k];
        for (int s = 0; s <= xl; ++s) {
            x = x0;
            // V.// applies updates left to right: [(k, xk+s), (l, xl-s)] — if k == l the second wins
            x[k] = xk + s; x[l] = xl - s;
            int id2 = ss.lookup(x.data());
            double binomialFactor = hs_pow(m, s) * hs_pow(1.0 - m, xl - s) * choose_int(xl, s);
            c.bTemp[id2] = c.bTemp[id2] + oldProb * binomialFactor;
            newIds.push_back(id2);
        }
    }
    c.b = c.bTemp;
    std::vector<int> uniq; std::vector<char> seen((size_t)ss.nrStates, 0);
    for (int id : newIds) if (c.b[id] > 0.0 && !seen[id]) { seen[id] = 1; uniq.push_back(id); }
    fill_up(ss, uniq, c.live);
}

// Core.hs:229-240
void update_a(Coal& c, double dt) {
    int P = c.ss->P;
    for (int k = 0; k < P; ++k) {
        double aK = c.a[k];
        if (c.freeze[k] || aK < 1.0) continue;
        c.a[k] = 1.0 / (1.0 + ((1.0 / aK) - 1.0) * std::exp(-(0.5 * dt / c.popSize[k])));
    }
}

// Core.hs:242-280
void update_b(Coal& c, double dt) {
    const StateSpace& ss = *c.ss; int P = ss.P;
    for (int xId : c.live) {
        const int* up = &ss.x1up[(size_t)xId * P];
        const IVec& x = ss.idToState[xId];
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < P; ++i) {
            double xx = (double)x[i], pp = c.popSize[i], aa = c.a[i];
            double add = c.freeze[i] ? 0.0 : xx * (xx - 1.0) / 2.0 * (1.0 / pp) + xx * aa * (1.0 / pp);
            t1 = t1 + add;
        }
        for (int i = 0; i < P; ++i) {
            double bb = up[i] == -1 ? 0.0 : c.b[up[i]];
            double xx = (double)x[i], pp = c.popSize[i];
            double add = c.freeze[i] ? 0.0 : bb * (1.0 - std::exp(-(xx * (xx + 1.0) / 2.0 * (1.0 / pp) * dt)));
            t2 = t2 + add;
        }
        c.bTemp[xId] = c.b[xId] * std::exp(-(t1 * dt)) + t2;
    }
    for (int xId : c.live) c.b[xId] = c.bTemp[xId];
    c.stateSteps += (long long)c.live.size();
}

// Core.hs:282-288
void update_d(Coal& c, double dt) {
    int P = c.ss->P;
    double s = 0.0;
    for (int k = 0; k < P; ++k) s = s + c.b[c.ss->x1[k]];
    c.d = c.d + dt * s;
}

// Core.hs:80-89
bool can_use_shortcut(const Coal& c, bool queueEmpty) {
    int P = c.ss->P;
    bool allDerived = true;
    for (int xId : c.live) {
        int nz = 0; const IVec& x = c.ss->idToState[xId];
        for (int k = 0; k < P; ++k) nz += x[k] > 0;
        if (nz != 1) { allDerived = false; break; }
    }
    int na = 0; for (int k = 0; k < P; ++k) na += c.a[k] > 0.0;
    return allDerived && na == 1 && queueEmpty;
}

// Core.hs:91-118
void propagate_shortcut(Coal& c) {
    int P = c.ss->P;
    double nrA = 0.0; for (int k = 0; k < P; ++k) nrA = nrA + c.a[k];
    int popIndex = 0; while (!(c.a[popIndex] > 0.0)) ++popIndex;
    double popSize = c.popSize[popIndex];
    double add = 0.0;
    for (int xId : c.live) {
        const IVec& x = c.ss->idToState[xId];
        int nrDerived = 0; for (int k = 0; k < P; ++k) nrDerived += x[k];
        double withComb = 2.0 * popSize / (double)nrDerived;
        double combFactor = choose_cont(nrA + (double)nrDerived, nrDerived);
        add = add + c.b[xId] * (withComb / combFactor);
    }
    std::fill(c.b.begin(), c.b.end(), 0.0);
    c.d = c.d + add;
    c.a[popIndex] = 1.0;
}

// Core.hs:38-56 getProb (with makeInitModelState :290-298, makeInitCoalState :58-70,
// propagateStates :72-78, singleStep :120-135, performEvent :147-158).
// Returns 0 ok, 1 RarecoalModelException, 2 RarecoalCompException.
int get_prob(const StateSpace& ss, int T, const double* times, double theta, const double* disc,
             int noShortcut, const Event* ev, int nEv, const int* nVec, const int* config,
             double* out, long long* stateStepsOut, char* err, int cap) {
    int P = ss.P;
    int rc = validate_model(P, disc, ev, nEv, err, cap);
    if (rc) return rc;
    int sum = 0; for (int k = 0; k < P; ++k) sum += config[k];
    for (int k = 0; k < P; ++k) if (config[k] < 0) { set_err(err, cap, "illegal sample configuration given"); return 2; }
    if (sum < 1 || sum > ss.maxAf) { set_err(err, cap, "cannot find state (pattern outside the state space)"); return 2; }
    std::vector<Event> q = sorted_events(ev, nEv);
    size_t qpos = 0;
    Coal c; c.ss = &ss;
    c.popSize.assign(P, 1.0); c.freeze.assign(P, 0); c.t = 0.0; c.d = 0.0; c.stateSteps = 0;
    c.a.resize(P); for (int k = 0; k < P; ++k) c.a[k] = (double)(nVec[k] - config[k]);
    c.b.assign((size_t)ss.nrStates, 0.0); c.bTemp.assign((size_t)ss.nrStates, 0.0);
    int initialId = ss.lookup(config);
    c.b[initialId] = 1.0;
    fill_up(ss, std::vector<int>(1, initialId), c.live);
    for (int s = 0; s < T; ++s) {
        double nextTime = times[s];
        if (can_use_shortcut(c, qpos == q.size()) && !noShortcut) { propagate_shortcut(c); break; }
        while (qpos < q.size() && q[qpos].t < nextTime) {
            const Event& e = q[qpos];
            switch (e.type) {
            case EV_JOIN: pop_join_a(c.a.data(), e.k, e.l); pop_join_b(c, e.k, e.l); break;
            case EV_SPLIT: pop_split_a(c.a.data(), e.k, e.l, e.v); pop_split_b(c, e.k, e.l, e.v); break;
            case EV_POPSIZE: c.popSize[e.k] = e.v; break;
            case EV_FREEZE: c.freeze[e.k] = e.v != 0.0; break;
            }
            ++qpos;
        }
        double deltaT = nextTime - c.t;
        if (deltaT > 0) {
            update_a(c, deltaT); update_b(c, deltaT); update_d(c, deltaT);
            c.t = c.t + deltaT;
        }
    }
    double combFac = 1.0;
    for (int k = 0; k < P; ++k) combFac = combFac * choose_int(nVec[k], config[k]);
    double discFac = 1.0;
    for (int k = 0; k < P; ++k) discFac = discFac * (config[k] > 0 ? disc[k] : 1.0);
    if (stateStepsOut) *stateStepsOut = c.stateSteps;
    if (!(combFac > 0)) { set_err(err, cap, "Overflow Error in getProb"); return 2; }
    *out = c.d * theta * combFac * discFac;
    return 0;
}

}  // namespace

// =======================================================================================
extern "C" {

typedef struct { double t; int32_t type; int32_t k; int32_t l; double v; } orc_event;

// Utils.hs:64-73 getTimeSteps.  Returns the number of grid points (nr_steps - 1); fills up to cap.
int orc_time_steps(int n0, int lingen, double tMax, double* out, int cap) {
    double tMin = 1.0 / (2.0 * (double)n0);
    double alpha = (double)lingen / (2.0 * (double)n0);
    int nr = (int)std::floor(std::log(1.0 + tMax / alpha) / std::log(1.0 + tMin / alpha));
    int n = 0;
    for (int i = 1; i <= nr - 1; ++i, ++n)
        if (out && n < cap) out[n] = alpha * std::exp((double)i / (double)nr * std::log(1.0 + tMax / alpha)) - alpha;
    return n;
}

// Utils.hs:100-104.  out = [n][P] row-major; returns n.
int orc_all_configs(int maxAf, int P, const int* nVec, int* out, int cap) {
    std::vector<IVec> v = compute_all_configs(maxAf, IVec(nVec, nVec + P));
    if (out) for (size_t i = 0; i < v.size() && (int)i < cap; ++i) for (int k = 0; k < P; ++k) out[i * P + k] = v[i][k];
    return (int)v.size();
}

double orc_choose(int n, int k) { return choose_int(n, k); }
double orc_choose_cont(double n, int k) { return choose_cont(n, k); }

void* orc_ss_create(int P, int maxAf) { return make_state_space(P, maxAf); }
void orc_ss_destroy(void* h) { delete (StateSpace*)h; }
int orc_ss_nr_states(void* h) { return ((StateSpace*)h)->nrStates; }
int orc_ss_state_to_id(void* h, const int* x) { return ((StateSpace*)h)->lookup(x); }
void orc_ss_id_to_state(void* h, int id, int* x) { auto* s = (StateSpace*)h; for (int k = 0; k < s->P; ++k) x[k] = s->idToState[id][k]; }
void orc_ss_x1up(void* h, int id, int* out) { auto* s = (StateSpace*)h; for (int k = 0; k < s->P; ++k) out[k] = s->x1up[(size_t)id * s->P + k]; }
int orc_ss_x1(void* h, int k) { return ((StateSpace*)h)->x1[k]; }
int orc_ss_fill_up(void* h, const int* ids, int n, int* out, int cap) {
    std::vector<int> o; fill_up(*(StateSpace*)h, std::vector<int>(ids, ids + n), o);
    for (size_t i = 0; i < o.size() && (int)i < cap; ++i) out[i] = o[i];
    return (int)o.size();
}

int orc_validate_model(int P, const double* disc, const orc_event* ev, int nEv, char* err, int cap) {
    return validate_model(P, disc, (const Event*)ev, nEv, err, cap);
}

// Core.hs:165-170,195-200 (exported for the conservation property tests, Core/Test.hs:35-127)
void orc_pop_join_a(double* a, int k, int l) { pop_join_a(a, k, l); }
void orc_pop_split_a(double* a, int k, int l, double m) { pop_split_a(a, k, l, m); }
// popJoinB / popSplitB on a dense b vector with an explicit live list; live list is rewritten.
int orc_pop_join_b(void* h, double* b, int* live, int nLive, int cap, int k, int l) {
    Coal c; c.ss = (StateSpace*)h; int n = c.ss->nrStates;
    c.b.assign(b, b + n); c.bTemp.assign(n, 0.0); c.live.assign(live, live + nLive);
    pop_join_b(c, k, l);
    std::copy(c.b.begin(), c.b.end(), b);
    for (size_t i = 0; i < c.live.size() && (int)i < cap; ++i) live[i] = c.live[i];
    return (int)c.live.size();
}
int orc_pop_split_b(void* h, double* b, int* live, int nLive, int cap, int k, int l, double m) {
    Coal c; c.ss = (StateSpace*)h; int n = c.ss->nrStates;
    c.b.assign(b, b + n); c.bTemp.assign(n, 0.0); c.live.assign(live, live + nLive);
    pop_split_b(c, k, l, m);
    std::copy(c.b.begin(), c.b.end(), b);
    for (size_t i = 0; i < c.live.size() && (int)i < cap; ++i) live[i] = c.live[i];
    return (int)c.live.size();
}

// Core.hs:38-56
int orc_get_prob(void* h, int T, const double* times, double theta, const double* disc, int noShortcut,
                 const orc_event* ev, int nEv, const int* nVec, const int* config, double* out,
                 long long* stateSteps, char* err, int cap) {
    return get_prob(*(StateSpace*)h, T, times, theta, disc, noShortcut, (const Event*)ev, nEv, nVec, config,
                    out, stateSteps, err, cap);
}

// StateSpace.hs:199-221 getRegularizationPenalty
double orc_reg_penalty(int P, double reg, const orc_event* ev, int nEv) {
    std::vector<Event> s = sorted_events((const Event*)ev, nEv);
    std::vector<double> ps((size_t)P, 1.0);
    double res = 0.0;
    auto regFunc = [&](double oldP, double newP) {
        double r = newP > oldP ? newP / oldP - 1.0 : oldP / newP - 1.0;
        return reg * (r * r);
    };
    for (auto& e : s) {
        if (e.type == EV_POPSIZE) {
            double oldP = ps[e.k];
            if (e.t != 0.0) res = res + regFunc(oldP, e.v);
            ps[e.k] = e.v;
        } else if (e.type == EV_JOIN) {
            // `Join l k` at :210 — first field (to) named l, second (from) named k
            double fromP = ps[e.l], toP = ps[e.k];
            res = res + regFunc(fromP, toP);
        }
    }
    return res;
}

// MaxUtils.hs:93-119 computeFrequencySpectrum (patterns already in model-branch / standard order) +
// MaxUtils.hs:85-91 computeLogLikelihoodFromSpec.  The per-pattern fan-out (`parMap rdeepseq`,
// MaxUtils.hs:109-110) is a dynamic work queue over patterns served by `threads` std::threads (0 = all cores).
// probs[nPat] out; returns status of the first failing pattern (0 ok).
int orc_logl(int P, int maxAf, int T, const double* times, double theta, const double* disc, int noShortcut,
             const orc_event* ev, int nEv, const int* nVec, int nPat, const int* patterns,
             const long long* counts, long long totalSites, double siteRed, int threads,
             double* probs, double* logl, long long* stateSteps, char* err, int cap) {
    // MaxUtils.hs:107: the state space is rebuilt for every likelihood call
    StateSpace* ss = make_state_space(P, maxAf);
    int status = 0;
    long long W = 0;
    int nThreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nThreads < 1) nThreads = 1;
    if (nThreads > nPat) nThreads = nPat > 0 ? nPat : 1;
    std::atomic<int> next(0);
    std::atomic<long long> Wacc(0);
    std::mutex mu;
    auto worker = [&]() {
        long long wLocal = 0;
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= nPat) break;
            char e[256]; e[0] = 0; long long w = 0; double p = 0.0;
            int rc = get_prob(*ss, T, times, theta, disc, noShortcut, (const Event*)ev, nEv, nVec,
                              patterns + (size_t)i * P, &p, &w, e, 256);
            probs[i] = p; wLocal += w;
            if (rc) { std::lock_guard<std::mutex> g(mu); if (!status) { status = rc; set_err(err, cap, e); } }
        }
        Wacc += wLocal;
    };
    if (nThreads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nThreads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    W = Wacc.load();
    delete ss;
    if (stateSteps) *stateSteps = W;
    if (status) return status;
    // MaxUtils.hs:85-91 (left folds)
    double ll = 0.0, sp = 0.0; long long sc = 0;
    for (int i = 0; i < nPat; ++i) { ll = ll + std::log(probs[i]) * (double)counts[i]; sp = sp + probs[i]; sc += counts[i]; }
    long long other = totalSites - sc;
    *logl = siteRed * (ll + (double)other * std::log(1.0 - sp));
    return 0;
}

int orc_max_threads() { int n = (int)std::thread::hardware_concurrency(); return n < 1 ? 1 : n; }

}  // extern "C"
