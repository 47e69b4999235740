// rc_plan.cpp — host planner (see rc_plan.h).
//
// The planner resolves, once per (problem, event structure), everything about the propagation that
// does not depend on continuous parameters:
//   * the live-state list of every pattern in every structural epoch, in the reference's own order
//     (fillUpStateSpace . nub, StateSpace.hs:145-160; Core.hs:67,188,221);
//   * for each live state the neighbour slots x + e_k read by updateB (Core.hs:272-280; the
//     reference's dense `x1up` table, StateSpace.hs:77-79, restricted to the live set);
//   * the Join / Split transition tables (Core.hs:172-221) as gather lists: new_b[y] = sum over
//     sources of old_b[x] * w, with w = 1 (Join) or an index into the per-model binomial table
//     m^s (1-m)^(x_l-s) choose(x_l, s) (Split);
//   * whether the pattern can take the closed-form shortcut (Core.hs:80-89) once the event queue is
//     empty.
// The kernels then do arithmetic only.
#include "rc_plan.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <thread>

namespace rc {

// host threads of the planner: the machine's, at most 16; RC_PLAN_THREADS overrides (tests: a plan must not depend on it)
static int plan_threads() {
    static const int n = [] {
        const char* env = std::getenv("RC_PLAN_THREADS");
        int v = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(v, 16));
    }();
    return n;
}

// ---- enumeration -------------------------------------------------------------------------------
// computeAllConfigs (Utils.hs:100-112,121-127).  The reference enumerates (P-1)-subsets of
// {1..P+af-1} in `allSubsets` order (those without the largest element first) and takes gaps; that
// is the same as: last component descending from af to 0, and for each value the (P-1)-prefixes of
// the remainder in the same order.
static void compositions(int P, int af, std::vector<int>& cur, std::vector<std::vector<int>>& out) {
    if (P == 1) {
        cur[0] = af;
        out.push_back(cur);
        return;
    }
    for (int v = af; v >= 0; --v) {
        cur[P - 1] = v;
        compositions(P - 1, af - v, cur, out);
    }
}

std::vector<std::vector<int>> all_configs(int maxAf, const std::vector<int>& nVec) {
    int P = (int)nVec.size();
    std::vector<std::vector<int>> all, out;
    std::vector<int> cur((size_t)P, 0);
    for (int af = 1; af <= maxAf; ++af) compositions(P, af, cur, all);
    for (auto& v : all) {
        bool ok = true;
        for (int k = 0; k < P && ok; ++k) ok = v[k] <= nVec[k];
        if (ok) out.push_back(v);
    }
    return out;
}

// choose n k (Utils.hs:208-210): left product of (n+1-j)/j.
double choose(int n, int k) {
    double p = 1.0;
    for (int j = 1; j <= k; ++j) p = p * ((double)(n + 1 - j) / (double)j);
    return p;
}

// ---- state space -------------------------------------------------------------------------------
bool key_fits(int P, int maxAf) {
    long double v = 1.0L;
    for (int k = 0; k < P; ++k) v *= (long double)(maxAf + 1);
    return v < 9.0e18L;
}

uint64_t StateSpace::key(const int32_t* x) const {
    uint64_t h = 0;
    for (int k = 0; k < P; ++k) h = h * (uint64_t)(maxAf + 1) + (uint64_t)x[k];
    return h;
}

int StateSpace::state_to_id(const int32_t* x) const {
    for (int k = 0; k < P; ++k)
        if (x[k] < 0 || x[k] > maxAf) return -1;
    auto it = toId.find(key(x));
    return it == toId.end() ? -1 : it->second;
}

StateSpace* make_state_space(int P, int maxAf) {
    if (P < 1 || P > RC_MAX_P || maxAf < 1 || maxAf > RC_MAX_AF || !key_fits(P, maxAf)) return nullptr;
    StateSpace* ss = new StateSpace();
    ss->P = P;
    ss->maxAf = maxAf;
    auto cfg = all_configs(maxAf, std::vector<int>((size_t)P, maxAf));
    ss->nrStates = (int)cfg.size();
    ss->states.resize((size_t)ss->nrStates * P);
    std::vector<int32_t> x((size_t)P);
    for (int i = 0; i < ss->nrStates; ++i) {
        for (int k = 0; k < P; ++k) {
            ss->states[(size_t)i * P + k] = (uint8_t)cfg[i][k];
            x[k] = cfg[i][k];
        }
        ss->toId[ss->key(x.data())] = i;
    }
    ss->x1up.assign((size_t)ss->nrStates * P, -1);
    for (int i = 0; i < ss->nrStates; ++i) {
        int sum = 0;
        for (int k = 0; k < P; ++k) {
            x[k] = ss->states[(size_t)i * P + k];
            sum += x[k];
        }
        if (sum + 1 > maxAf) continue;   // StateSpace.hs:79
        for (int k = 0; k < P; ++k) {
            x[k] += 1;
            ss->x1up[(size_t)i * P + k] = ss->state_to_id(x.data());
            x[k] -= 1;
        }
    }
    ss->x1.resize((size_t)P);
    for (int k = 0; k < P; ++k) {
        std::fill(x.begin(), x.end(), 0);
        x[k] = 1;
        ss->x1[k] = ss->state_to_id(x.data());
    }
    return ss;
}

std::vector<int> StateSpace::fill_up(const std::vector<int>& ids) const {
    std::vector<int> out;
    std::vector<char> seen((size_t)nrStates, 0);
    std::vector<int32_t> lo((size_t)P), hi((size_t)P);
    for (int id : ids) {
        if (id < 0 || id >= nrStates) continue;
        for (int k = 0; k < P; ++k) {
            hi[k] = states[(size_t)id * P + k];
            lo[k] = hi[k] ? 1 : 0;
        }
        for (;;) {
            int y = state_to_id(lo.data());
            if (y >= 0 && !seen[y]) {
                seen[y] = 1;
                out.push_back(y);
            }
            int k = P - 1;
            while (k >= 0 && (hi[k] == 0 || lo[k] == hi[k])) {
                if (hi[k]) lo[k] = 1;
                --k;
            }
            if (k < 0) break;
            lo[k] += 1;
        }
    }
    return out;
}

// ---- per-pattern plan ----------------------------------------------------------------------------
namespace {

struct St {
    uint8_t x[RC_MAX_P];
};

struct PatPlan {
    std::vector<int> live;          // [nEpochs]
    std::vector<uint32_t> words;    // per epoch: live * stride
    std::vector<uint32_t> meta;     // per epoch: live
    std::vector<int> trCnt;         // per transition: live_{e+1} counts
    std::vector<uint32_t> trEnt;
    std::vector<uint8_t> mx;        // per epoch: P bytes, largest x_k over the live states
    std::vector<uint8_t> mn;        // per epoch: P bytes, 1 if every live state has x_k >= 1 (and some has), else 0
    // row view of a box-shaped live set (one thread per row along the widest branch), per epoch
    std::vector<uint8_t> rowOK, rowM, rowK, rowNO;   // usable / row length / row branch / other non-zero branches
    std::vector<uint8_t> fullBox;                    // per epoch: the live set is exactly prod_k {1..mx_k} (every state has every live branch)
    std::vector<int> rowStride, rowCnt;              // reference-order stride of the row branch / number of rows
    std::vector<uint32_t> rowWords;                  // per row RC_ROW_NO words (k | x<<8 | neighbour row<<16)
    std::vector<uint16_t> rowRef;                    // per row: reference-order index of its x_r = 1 state
    std::vector<uint16_t> rowRefs;                   // per row RC_ROW_MAXM reference-order indices (x_r = 1 .. M; 0xFFFF beyond the row)
    std::vector<uint8_t> rowBr;                      // per row: its row branch (the widest branch of the row's part)
    std::vector<uint8_t> rowPm;                      // per row: longest row of its part (rows of a part are stored with one stride)
    std::vector<int> rowPad;                         // per epoch: sum over the rows of (rowPm | 1) — b doubles the pattern needs in a row bin
    bool shortcutOK = false;
    int maxLive = 0;
    int maxNnz = 0;
    int maxR = 0;                   // largest sum_k mx[k] over epochs (rows of the per-step rate table)
    std::string err;
};

struct Ctx {
    int P, maxAf, stride;
    uint64_t base;
    uint64_t pw[RC_MAX_P];          // weight of x[k] in key(): base^(P-1-k)
    uint64_t key(const St& s) const {
        uint64_t h = 0;
        for (int k = 0; k < P; ++k) h = h * base + s.x[k];
        return h;
    }
};

// state key -> index.  Open addressing over a flat array that keeps its capacity across clear(): the planner builds a few of
// these per pattern and epoch (live sets of 1 - 700 states), and node allocation made std::unordered_map the largest item of
// the per-pattern phase.  The subset of the map interface the planner uses; keys are < 2^63, so ~0 marks an empty slot.
class Index {
public:
    typedef std::pair<uint64_t, int> Slot;
    typedef Slot* iterator;
    typedef const Slot* const_iterator;
    Index() { rehash(32); }
    void clear() {
        for (Slot& s : t_) s.first = EMPTY;
        n_ = 0;
    }
    iterator end() { return nullptr; }
    const_iterator end() const { return nullptr; }
    iterator find(uint64_t k) { return const_cast<iterator>(static_cast<const Index*>(this)->find(k)); }
    const_iterator find(uint64_t k) const {
        for (size_t i = hash(k);; i = (i + 1) & mask_) {
            if (t_[i].first == k) return &t_[i];
            if (t_[i].first == EMPTY) return nullptr;
        }
    }
    std::pair<iterator, bool> emplace(uint64_t k, int v) {
        if ((n_ + 1) * 2 > t_.size()) rehash(t_.size() * 2);
        for (size_t i = hash(k);; i = (i + 1) & mask_) {
            if (t_[i].first == k) return {&t_[i], false};
            if (t_[i].first == EMPTY) { t_[i] = Slot(k, v); ++n_; return {&t_[i], true}; }
        }
    }
    int& operator[](uint64_t k) { return emplace(k, 0).first->second; }
    size_t size() const { return n_; }
    void swap(Index& o) { t_.swap(o.t_); std::swap(n_, o.n_); std::swap(mask_, o.mask_); }
private:
    static constexpr uint64_t EMPTY = ~0ull;
    std::vector<Slot> t_;
    size_t n_ = 0, mask_ = 0;
    size_t hash(uint64_t k) const { return (size_t)((k * 0x9E3779B97F4A7C15ull) >> 29) & mask_; }
    void rehash(size_t cap) {
        std::vector<Slot> old;
        old.swap(t_);
        t_.assign(cap, Slot(EMPTY, 0));
        mask_ = cap - 1;
        n_ = 0;
        for (const Slot& s : old) if (s.first != EMPTY) emplace(s.first, s.second);
    }
};

// fillUpStateSpace over explicit states: for every seed, every y with 1 <= y_k <= x_k where x_k > 0
// (component 0 slowest), first occurrence kept (nub).
void fill_up_states(const Ctx& c, const std::vector<St>& seeds, std::vector<St>& out, Index& idx) {
    out.clear();
    idx.clear();
    St cur;
    for (const St& hi : seeds) {
        for (int k = 0; k < c.P; ++k) cur.x[k] = hi.x[k] ? 1 : 0;
        for (;;) {
            uint64_t h = c.key(cur);
            if (idx.find(h) == idx.end()) {
                idx.emplace(h, (int)out.size());
                out.push_back(cur);
            }
            int k = c.P - 1;
            while (k >= 0 && (hi.x[k] == 0 || cur.x[k] == hi.x[k])) {
                if (hi.x[k]) cur.x[k] = 1;
                --k;
            }
            if (k < 0) break;
            cur.x[k] += 1;
        }
    }
}

void emit_epoch(const Ctx& c, const std::vector<St>& L, const Index& idx, PatPlan& pp) {
    pp.live.push_back((int)L.size());
    pp.maxLive = std::max(pp.maxLive, (int)L.size());
    size_t w0 = pp.words.size();
    pp.words.resize(w0 + L.size() * (size_t)c.stride, 0u);
    size_t m0 = pp.mx.size();
    pp.mx.resize(m0 + (size_t)c.P, 0);
    pp.mn.resize(m0 + (size_t)c.P, L.empty() ? 0 : 1);
    for (size_t i = 0; i < L.size(); ++i) {
        const St& s = L[i];
        const uint64_t hs = c.key(s);
        int nnz = 0, sum = 0;
        for (int k = 0; k < c.P; ++k) {
            int xk = s.x[k];
            if (xk > pp.mx[m0 + (size_t)k]) pp.mx[m0 + (size_t)k] = (uint8_t)xk;
            if (!xk) pp.mn[m0 + (size_t)k] = 0;
            if (!xk) continue;
            sum += xk;
            uint32_t up = RC_NO_UP;
            if (xk + 1 <= c.maxAf) {
                auto it = idx.find(hs + c.pw[k]);            // key of x + e_k
                if (it != idx.end()) up = (uint32_t)it->second;
            }
            pp.words[w0 + i * c.stride + nnz] = (uint32_t)k | ((uint32_t)xk << 8) | (up << 16);
            ++nnz;
        }
        pp.maxNnz = std::max(pp.maxNnz, nnz);
        uint32_t m = (uint32_t)nnz | ((uint32_t)sum << 16);
        if (nnz == 1 && sum == 1) m |= RC_META_UNIT;
        pp.meta.push_back(m);
    }
    int r = 0;
    for (int k = 0; k < c.P; ++k) r += pp.mx[m0 + (size_t)k];
    pp.maxR = std::max(pp.maxR, r);
    // ---- row view ----
    // updateB couples a state only to x + e_k for its NON-ZERO components (Core.hs:272-280: a zero component contributes
    // nothing to t1 or t2), so the live set falls apart into independent parts by SUPPORT (the set of non-zero branches).
    // fillUp makes every part down-closed (with x every y <= x of the same support): a box for a tree model, a union of
    // boxes — a staircase — after a Split (Core.hs:220-221).  Every part is cut into rows along its own widest branch, x_r =
    // 1 .. the row's own length; a neighbouring row may be shorter, the states it lacks are not live, i.e. b = 0.
    // One thread of the row kernel owns a row.
    std::map<uint32_t, std::vector<int>> parts;          // support mask -> states (reference order)
    for (size_t i = 0; i < L.size(); ++i) {
        uint32_t m = 0;
        for (int k = 0; k < c.P; ++k) if (L[i].x[k]) m |= 1u << k;
        parts[m].push_back((int)i);
    }
    bool ok = !L.empty();
    int maxM = 0, maxNO = 0, firstRk = 0, nz = 0;
    long long box = 1;
    for (int k = 0; k < c.P; ++k) { int mk = pp.mx[m0 + (size_t)k]; if (mk) { ++nz; box *= mk; } }
    const size_t rowStart = pp.rowRef.size();
    std::vector<uint16_t> refsAll;
    std::vector<uint8_t> brAll, pmAll;
    int padSum = 0;
    {
        std::unordered_map<int, int> rowOfRef;           // reference index of a row's first state -> row (inside the pattern)
        struct Part { uint32_t mask; int rk, rm; };
        std::vector<Part> plist;
        // rows of all parts first (neighbour rows are looked up across the whole pattern's row list)
        for (auto& kv : parts) {
            if (!ok) break;
            int ext[RC_MAX_P] = {0}, rk = -1, rm = 0, no = -1;
            long long cnt = 1;
            for (int st : kv.second) for (int k = 0; k < c.P; ++k) ext[k] = std::max(ext[k], (int)L[(size_t)st].x[k]);
            for (int k = 0; k < c.P; ++k) if (ext[k]) { ++no; cnt *= ext[k]; if (ext[k] > rm) { rm = ext[k]; rk = k; } }
            (void)cnt;
            ok = no <= RC_ROW_NO && rm <= RC_ROW_MAXM && (no <= 3 || no + rm <= RC_ROW_SUM);
            if (!ok) break;
            plist.push_back({kv.first, rk, rm});
            if (plist.size() == 1) firstRk = rk;
            maxM = std::max(maxM, rm);
            maxNO = std::max(maxNO, no);
            for (int st : kv.second)
                if (L[(size_t)st].x[rk] == 1) {
                    rowOfRef.emplace(st, (int)(pp.rowRef.size() - rowStart));
                    pp.rowRef.push_back((uint16_t)st);
                    brAll.push_back((uint8_t)rk);
                    pmAll.push_back((uint8_t)rm);
                    padSum += rm | 1;
                }
        }
        for (size_t q = rowStart; q < pp.rowRef.size() && ok; ++q) {
            St s = L[pp.rowRef[q]];
            const int rk = brAll[q - rowStart];
            int no = 0;
            uint32_t w[RC_ROW_NO];
            for (int d = 0; d < RC_ROW_NO; ++d) w[d] = 0xFFFF0000u;     // k = 0, x = 0: unused entry
            for (int k = 0; k < c.P; ++k) {
                int xk = s.x[k];
                if (!xk || k == rk) continue;
                uint32_t up = RC_NO_UP;
                s.x[k] = (uint8_t)(xk + 1);
                auto it = idx.find(c.key(s));
                s.x[k] = (uint8_t)xk;
                if (it != idx.end()) {
                    auto rt = rowOfRef.find(it->second);
                    if (rt == rowOfRef.end()) { ok = false; break; }
                    up = (uint32_t)rt->second;
                }
                w[no++] = (uint32_t)k | ((uint32_t)xk << 8) | (up << 16);
            }
            for (int d = 0; d < RC_ROW_NO; ++d) pp.rowWords.push_back(w[d]);
        }
        // every row must be complete (x_r = 1 .. the extent of its part); the reference-order index of each of its states is
        // kept, since after a Join the live list is no longer in mixed-radix order (fillUp over several seeds, Core.hs:188)
        const size_t nRows = pp.rowRef.size() - rowStart;
        refsAll.assign(nRows * RC_ROW_MAXM, 0xFFFF);
        size_t covered = 0;
        for (size_t q = 0; q < nRows && ok; ++q) {
            St s = L[pp.rowRef[rowStart + q]];
            const int rk = brAll[q];
            uint32_t m = 0;
            for (int k = 0; k < c.P; ++k) if (s.x[k]) m |= 1u << k;
            int rm = 0;
            for (const Part& pt : plist) if (pt.mask == m) rm = pt.rm;
            for (int j = 0; j < rm; ++j) {                     // the row ends at its first state that is not live
                s.x[rk] = (uint8_t)(j + 1);
                auto it = idx.find(c.key(s));
                if (it == idx.end()) break;
                refsAll[q * RC_ROW_MAXM + (size_t)j] = (uint16_t)it->second;
                ++covered;
            }
        }
        ok = ok && covered == L.size();
    }
    if (ok) {
        pp.rowRefs.insert(pp.rowRefs.end(), refsAll.begin(), refsAll.end());
        pp.rowBr.insert(pp.rowBr.end(), brAll.begin(), brAll.end());
        pp.rowPm.insert(pp.rowPm.end(), pmAll.begin(), pmAll.end());
    } else {
        pp.rowRef.resize(rowStart);
        pp.rowWords.resize(rowStart * RC_ROW_NO);
    }
    // a single part whose support is every live branch: the box of pattern mode
    bool full = ok && parts.size() == 1 && box == (long long)L.size();
    int stride = 1;
    for (int k = firstRk + 1; k < c.P && full; ++k) if (pp.mx[m0 + (size_t)k]) stride *= pp.mx[m0 + (size_t)k];
    pp.fullBox.push_back(full ? 1 : 0);
    pp.rowOK.push_back(ok ? 1 : 0);
    pp.rowM.push_back((uint8_t)(ok ? maxM : 0));
    pp.rowK.push_back((uint8_t)(ok ? firstRk : 0));
    pp.rowNO.push_back((uint8_t)(ok ? maxNO : 0));
    pp.rowStride.push_back(full ? stride : 0);
    pp.rowCnt.push_back(ok ? (int)(pp.rowRef.size() - rowStart) : 0);
    pp.rowPad.push_back(ok ? padSum : 0);
    (void)nz;
}

void plan_pattern(const Ctx& c, const int32_t* m, const int32_t* nVec, const std::vector<StructEv>& sev,
                  const std::vector<GapInfo>& gaps, bool hasSplit, PatPlan& pp) {
    const int P = c.P;
    std::vector<St> L, L2, seeds;
    Index idx, idx2;
    St m0;
    std::memset(&m0, 0, sizeof(m0));
    for (int k = 0; k < P; ++k) m0.x[k] = (uint8_t)m[k];
    seeds.assign(1, m0);
    fill_up_states(c, seeds, L, idx);                    // Core.hs:67
    std::vector<char> Z(L.size(), 0), Z2;                // Z[i] <=> b(L[i]) > 0 (symbolically)
    Z[idx[c.key(m0)]] = 1;                               // Core.hs:66
    bool apos[RC_MAX_P];
    for (int k = 0; k < P; ++k) apos[k] = (nVec[k] - m[k]) > 0;
    emit_epoch(c, L, idx, pp);

    for (size_t g = 0; g < sev.size(); ++g) {
        // updateB spreads positivity one level down per executed step (Core.hs:272-280)
        if (hasSplit) {
            for (unsigned mask : gaps[g].masks) {
                bool all = true, changed = false;
                Z2 = Z;
                for (size_t i = 0; i < L.size(); ++i) {
                    if (Z[i]) continue;
                    St s = L[i];
                    for (int k = 0; k < P && !Z2[i]; ++k) {
                        if (!s.x[k] || ((mask >> k) & 1u)) continue;
                        s.x[k] += 1;
                        auto it = idx.find(c.key(s));
                        if (it != idx.end() && Z[it->second]) Z2[i] = 1;
                        s.x[k] -= 1;
                    }
                    if (Z2[i]) changed = true; else all = false;
                }
                Z.swap(Z2);
                if (all || !changed) break;
            }
        }
        const StructEv& e = sev[g];
        const int k = e.k, l = e.l;
        // (target, src, widx) records in the reference's accumulation order
        struct Rec { uint64_t key; int src; uint32_t widx; };
        std::vector<Rec> recs;
        seeds.clear();
        Index pos;                                       // target -> b' > 0 ?
        if (e.type == RC_EV_JOIN_D) {                      // Core.hs:172-188
            Index seen;
            for (size_t i = 0; i < L.size(); ++i) {
                St y = L[i];
                int v = y.x[k] + y.x[l];
                if (v > c.maxAf) { pp.err = "state outside the state space after Join"; return; }
                y.x[k] = (uint8_t)v;
                y.x[l] = 0;
                uint64_t h = c.key(y);
                recs.push_back({h, (int)i, RC_W_ONE});
                if (seen.emplace(h, 0).second) seeds.push_back(y);   // nub newNonZeroStateIds (no b>0 filter)
                int& pz = pos[h];
                pz = pz || Z[i];
            }
            apos[k] = apos[k] || apos[l];                // Core.hs:165-170
            apos[l] = false;
        } else {                                         // Core.hs:202-221
            std::vector<std::pair<uint64_t, St>> targets;
            for (size_t i = 0; i < L.size(); ++i) {
                int xl = L[i].x[l], xk = L[i].x[k];
                for (int s = 0; s <= xl; ++s) {
                    St y = L[i];
                    y.x[k] = (uint8_t)(xk + s);
                    y.x[l] = (uint8_t)(xl - s);
                    uint64_t h = c.key(y);
                    recs.push_back({h, (int)i, (uint32_t)(xl * (c.maxAf + 1) + s)});
                    bool wpos = e.mclass == 2 || (e.mclass == 0 && s == 0) || (e.mclass == 1 && s == xl);
                    int& pz = pos[h];
                    pz = pz || (Z[i] && wpos);
                    targets.emplace_back(h, y);
                }
            }
            Index seen;
            for (auto& t : targets)                      // filterM (b > 0) then nub, Core.hs:220-221
                if (pos[t.first] && seen.emplace(t.first, 0).second) seeds.push_back(t.second);
            bool al = apos[l];                           // Core.hs:195-200
            apos[k] = apos[k] || (al && e.mclass != 0);
            apos[l] = al && e.mclass != 1;
        }
        fill_up_states(c, seeds, L2, idx2);
        if (L2.size() > 0xFFFEu) { pp.err = "live set of one pattern exceeds 65534 states"; return; }
        // gather lists per new slot, preserving source order
        std::vector<int> cnt(L2.size(), 0);
        std::vector<int> tgt(recs.size(), -1);
        for (size_t r = 0; r < recs.size(); ++r) {
            auto it = idx2.find(recs[r].key);
            if (it == idx2.end()) continue;              // target filtered out (its b' is exactly 0)
            tgt[r] = it->second;
            cnt[it->second]++;
        }
        std::vector<int> off(L2.size() + 1, 0);
        for (size_t i = 0; i < L2.size(); ++i) off[i + 1] = off[i] + cnt[i];
        size_t e0 = pp.trEnt.size();
        pp.trEnt.resize(e0 + (size_t)off[L2.size()]);
        std::vector<int> fillp(off.begin(), off.end() - 1);
        for (size_t r = 0; r < recs.size(); ++r) {
            if (tgt[r] < 0) continue;
            pp.trEnt[e0 + (size_t)fillp[tgt[r]]++] = (uint32_t)recs[r].src | (recs[r].widx << 16);
        }
        pp.trCnt.insert(pp.trCnt.end(), cnt.begin(), cnt.end());
        Z2.assign(L2.size(), 0);
        for (size_t i = 0; i < L2.size(); ++i) {
            auto it = pos.find(c.key(L2[i]));
            Z2[i] = it != pos.end() && it->second;
        }
        L.swap(L2);
        idx.swap(idx2);
        Z.swap(Z2);
        emit_epoch(c, L, idx, pp);
    }
    // canUseShortcut (Core.hs:80-89), evaluated when the event queue is empty
    bool single = true;
    size_t last = pp.meta.size() - L.size();
    for (size_t i = 0; i < L.size(); ++i) single = single && ((pp.meta[last + i] & 0xFFu) == 1u);
    int na = 0;
    for (int k = 0; k < P; ++k) na += apos[k] ? 1 : 0;
    pp.shortcutOK = single && na == 1;
}

}  // namespace

// Pattern mode: shape of a box-shaped live set whose states all fit the registers of one thread.  A = the row branch (the
// widest), then up to four more branches with m >= 2 in descending order (B, C, D, E); every other non-zero branch has
// m = 1 and only contributes a decay factor.  Returns MA | MB << 4 | MC << 8 | MD << 12 | ME << 16, or 0 if the pattern has
// more than maxStates states or does not qualify; dims (optional) receives the branch ids in the kernel's order: A, B..,
// then the m = 1 branches.
static uint32_t pat_shape(const uint8_t* mx, int P, int rk, int maxStates, int* dims = nullptr, int* nShaped = nullptr, int* nUnit = nullptr) {
    const int MA = mx[rk];
    if (MA < 1 || MA > RC_ROW_MAXM) return 0;
    int sh[4] = {0, 0, 0, 0}, shK[4] = {0, 0, 0, 0}, nS = 0, nU = 0, un[RC_MAX_P];
    for (int k = 0; k < P; ++k) {
        if (k == rk || !mx[k]) continue;
        if (mx[k] == 1) { un[nU++] = k; continue; }
        if (nS == 4) return 0;
        int q = nS++;
        while (q > 0 && sh[q - 1] < mx[k]) { sh[q] = sh[q - 1]; shK[q] = shK[q - 1]; --q; }     // descending, stable in k
        sh[q] = mx[k];
        shK[q] = k;
    }
    long long prod = MA;
    for (int q = 0; q < nS; ++q) prod *= sh[q];
    if (prod > maxStates || 1 + nS + nU > 8 || (nS && sh[0] > MA)) return 0;
    if (dims) {
        dims[0] = rk;
        for (int q = 0; q < nS; ++q) dims[1 + q] = shK[q];
        for (int q = 0; q < nU; ++q) dims[1 + nS + q] = un[q];
    }
    if (nShaped) *nShaped = nS;
    if (nUnit) *nUnit = nU;
    return (uint32_t)MA | ((uint32_t)sh[0] << 4) | ((uint32_t)sh[1] << 8) | ((uint32_t)sh[2] << 12) | ((uint32_t)sh[3] << 16);
}

// Pattern mode for a live set that is not a box of all live branches (what a Split leaves: a union of boxes by support, down-
// closed): the pattern runs on its BOUNDING box, a branch whose coordinate reaches 0 taking the values 0 .. mx (pair row of
// x = 0: {1, 0}, the table's leading pair of every key when the plan has a Split).  updateB couples x only to x + e_k of its
// non-zero components (Core.hs:272-280) and the live set is down-closed within a support, so the states of the box that are not
// live receive nothing and stay exactly 0; the lanes carry them along.  The row branch A must never reach 0 (then e_A is the
// only unit state the box can hold, at local index 0).  Returns the shape (sizes including the value 0) or 0.
struct PatBox {
    uint32_t shape = 0;
    int rk = 0, dims[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nS = 0, nU = 0, n = 0;
    uint32_t zeroMask = 0;      // bit q: dims[q] (q >= 1, a shaped branch) includes the value 0
    bool unit = false;          // e_A is a state of the box: no unit branch and every shaped branch includes 0
};
static bool pat_box(const uint8_t* mx, const uint8_t* mn, bool fullBox, int rowK, int P, int maxStates, PatBox& B) {
    B = PatBox();
    if (fullBox) {
        B.rk = rowK;
        B.shape = pat_shape(mx, P, rowK, maxStates, B.dims, &B.nS, &B.nU);
        if (!B.shape) return false;
    } else {
        int rk = -1;
        for (int k = 0; k < P; ++k) if (mx[k] && mn[k] && (rk < 0 || mx[k] > mx[rk])) rk = k;
        if (rk < 0) return false;
        const int MA = mx[rk];
        if (MA > RC_ROW_MAXM) return false;
        int sh[4] = {0, 0, 0, 0}, shK[4] = {0, 0, 0, 0}, shZ[4] = {0, 0, 0, 0}, nS = 0, nU = 0, un[RC_MAX_P];
        for (int k = 0; k < P; ++k) {
            if (k == rk || !mx[k]) continue;
            const int sz = mx[k] + (mn[k] ? 0 : 1);
            if (sz == 1) { un[nU++] = k; continue; }
            if (nS == 4) return false;
            int q = nS++;
            while (q > 0 && sh[q - 1] < sz) { sh[q] = sh[q - 1]; shK[q] = shK[q - 1]; shZ[q] = shZ[q - 1]; --q; }
            sh[q] = sz;
            shK[q] = k;
            shZ[q] = mn[k] ? 0 : 1;
        }
        long long prod = MA;
        for (int q = 0; q < nS; ++q) prod *= sh[q];
        if (prod > maxStates || 1 + nS + nU > 8 || (nS && sh[0] > MA)) return false;
        B.rk = rk;
        B.dims[0] = rk;
        for (int q = 0; q < nS; ++q) { B.dims[1 + q] = shK[q]; if (shZ[q]) B.zeroMask |= 1u << (1 + q); }
        for (int q = 0; q < nU; ++q) B.dims[1 + nS + q] = un[q];
        B.nS = nS;
        B.nU = nU;
        B.shape = (uint32_t)MA | ((uint32_t)sh[0] << 4) | ((uint32_t)sh[1] << 8) | ((uint32_t)sh[2] << 12) | ((uint32_t)sh[3] << 16);
    }
    B.n = (int)(B.shape & 0xFu);
    for (int q = 1; q < 5; ++q) B.n *= std::max(1, (int)((B.shape >> (4 * q)) & 0xFu));
    B.unit = B.nU == 0;
    for (int q = 1; q <= B.nS; ++q) B.unit = B.unit && ((B.zeroMask >> q) & 1u);
    return true;
}

// Bin images of the keyed tier (record layouts in rc_types.h).  Everything the set-up of a block used to look up through
// pattern ids, slot offsets, packed words and transition offsets is resolved here once, in the order the lanes of the
// block consume it: the block reads its header, every lane one 32-byte record, and the gather goes straight to the
// transition entries.
static std::string build_bin_images(PlanHost& out, const std::vector<StructEv>& sev) {
    const int P = out.P, nPat = out.nPat, nE = out.nEpochs, nnzw = out.nnzw;
    out.kBinHdr.assign(out.kBinPatOff.size() * (size_t)RC_HDR_N, 0);
    out.kPatRec.assign(out.kBinPats.size() * 2, 0);
    out.kLane.clear();
    out.kElem.clear();
    out.grpMaxFetch.assign((size_t)out.nEpochs * RC_NGROUP, 0);
    out.grpMaxB.assign((size_t)out.nEpochs * RC_NGROUP, 0);
    if (out.kBinBase.empty()) return "";
    if (out.maxEpochSlots >= (1LL << RC_IMG_CNT_SHIFT)) return "too many live states in one epoch for the keyed tier";
    if ((long long)out.trEnt.size() >= 0xFFFFFFFFLL) return "too many transition entries for the keyed tier";
    // pattern-mode lanes: where the record of (epoch, pattern) went, its group and its first element record
    std::vector<long long> recPos((size_t)nE * nPat, -1);
    std::vector<int> elPos((size_t)nE * nPat, 0), grpOf((size_t)nE * nPat, 0), boxN((size_t)nE * nPat, 0), laneOf((size_t)nE * nPat, 0),
        binOf((size_t)nE * nPat, 0);
    // gather record of the state with reference index `ref` of pattern i in epoch e
    auto gather_rec = [&](int e, int i, int ref, int live, uint32_t& r0, uint32_t& r1) -> bool {
        if (e == 0) {
            // the seed state m is the last one fillUp emits (Core.hs:66)
            r0 = ref == live - 1 ? RC_IMG_SEED : 0u;
            r1 = 0u;
            return true;
        }
        const int* trOff = &out.trOff[(size_t)out.trBase[(size_t)e - 1]];
        const int sl = out.slotOff[(size_t)e * (nPat + 1) + i] + ref;
        const int a = trOff[sl], b = trOff[sl + 1];
        if (b - a >= (1 << (32 - RC_IMG_CNT_SHIFT))) return false;
        r0 = (uint32_t)(out.trEntBase[(size_t)e - 1] + a);
        r1 = (uint32_t)out.slotOff[(size_t)(e - 1) * (nPat + 1) + i] | ((uint32_t)(b - a) << RC_IMG_CNT_SHIFT);
        return true;
    };
    // every bin group builds its records into vectors of its own (the groups are independent: a pattern is in one group per
    // epoch), side by side on the planner's threads; positions are local until the groups are put one behind the other below
    const int nGrp = nE * RC_NGROUP;
    std::vector<std::vector<uint32_t>> grpLane((size_t)nGrp), grpElem((size_t)nGrp);
    std::vector<std::string> grpErr((size_t)nGrp);
    auto do_group = [&](int v) {
        std::vector<uint32_t>& GL = grpLane[(size_t)v];
        std::vector<uint32_t>& GE = grpElem[(size_t)v];
        std::string& gerr = grpErr[(size_t)v];
        const int e = v / RC_NGROUP;
        const bool rows = out.epochMode[(size_t)v] != 0, patM = out.epochMode[(size_t)v] >= 3;
        const int patMax = rc_pat_cap(out.epochMode[(size_t)v]);
        const int* so = &out.slotOff[(size_t)e * (nPat + 1)];
        const int* ro = &out.rowOff[(size_t)e * (nPat + 1)];
        const int b0 = out.kBinBase[(size_t)v];
        for (int bI = b0; bI < b0 + out.nKBins[(size_t)v]; ++bI) {
            const int p0 = out.kBinPatOff[(size_t)bI], p1 = out.kBinPatOff[(size_t)bI + 1];
            int* hdr = &out.kBinHdr[(size_t)bI * RC_HDR_N];
            hdr[0] = p0;
            hdr[1] = p1 - p0;
            hdr[4] = out.kFetchOff[(size_t)bI];
            hdr[5] = out.kFetchOff[(size_t)bI + 1] - out.kFetchOff[(size_t)bI];
            hdr[6] = (int)(GL.size() / 4);
            hdr[7] = (int)(GE.size() / 2);
            hdr[8] = out.binSegOff[(size_t)bI];
            hdr[9] = out.binSegCnt[(size_t)bI];
            hdr[10] = out.binSegBytes[(size_t)bI];
            int lanes = 0, bOff = 0;
            uint32_t binShape = 0;
            for (int q = p0; q < p1; ++q) {
                const int i = out.kBinPats[(size_t)q];
                const int live = so[i + 1] - so[i];
                const uint16_t* rb = &out.kRowBase[(size_t)q * P];
                out.kPatRec[(size_t)q * 2] = i;
                out.kPatRec[(size_t)q * 2 + 1] = lanes | (bOff << 16);
                if (patM) {
                    // pattern mode: one lane per pattern; its states in mixed-radix order over (A, B, C, D, E), see rc_types.h
                    const uint32_t pr = out.patRow[(size_t)e * nPat + i];
                    const bool full = out.patFull[(size_t)e * nPat + i] != 0;
                    PatBox pb;
                    if (!pat_box(&out.maxX[((size_t)e * nPat + i) * P], &out.minX[((size_t)e * nPat + i) * P], full, (int)((pr >> 8) & 0xFFu), P, patMax, pb))
                        { gerr = "pattern-mode bin with an unsuitable pattern"; return; }
                    const uint32_t shape = pb.shape;
                    const int nS = pb.nS, nU = pb.nU, nBox = pb.n;
                    const int* dims = pb.dims;
                    const int MA = (int)(shape & 0xFu);
                    int Mp[4];
                    for (int q = 0; q < 4; ++q) Mp[q] = std::max(1, (int)((shape >> (4 * (q + 1))) & 0xFu));
                    const int SA = Mp[0] * Mp[1] * Mp[2] * Mp[3];
                    if (live > nBox || (full && live != nBox)) { gerr = "pattern-mode image: shape does not match the live set"; return; }
                    uint32_t rec[RC_PLANE_N];
                    for (int q = 0; q < RC_PLANE_N; ++q) rec[q] = 0u;
                    const bool direct = out.grpDirect[(size_t)v] != 0;
                    for (int q = 0; q < 8; ++q) {
                        // an unused slot points at pair row 0, {1, dt}: a neutral decay factor (direct bins have no ring: marker);
                        // a branch that includes the value 0 starts one row earlier, at the key's leading {1, 0} pair
                        const uint32_t row = direct ? (uint32_t)(RC_KROWS - 1) : q < 1 + nS + nU ? (uint32_t)rb[dims[q]] - ((pb.zeroMask >> q) & 1u) : 0u;
                        if (row > 0xFFu) { gerr = "pattern-mode image out of range (pair row)"; return; }
                        rec[q >> 2] |= row << (8 * (q & 3));
                    }
                    if (direct) {
                        // first pair of the pattern's only key inside a step row of the trajectory table
                        if (nS + nU != 0) { gerr = "direct group with a multi-branch pattern"; return; }
                        rec[15] = (uint32_t)out.keyPairBase[(size_t)out.ep[(size_t)e].keyOff + out.keyId[((size_t)e * nPat + i) * P + pb.rk]];
                    }
                    // every live state (reference order) at its place in the box; the places without one stay empty (0xFF: b = 0)
                    for (int w = 6; w < 15; ++w) rec[w] = 0xFFFFFFFFu;
                    std::vector<uint32_t> el((size_t)nBox * 2, 0u);
                    for (int ref = 0; ref < live; ++ref) {
                        const long long g = out.epochBase[(size_t)e] + so[i] + ref;
                        const int nnz = (int)(out.meta[(size_t)g] & 0xFFu);
                        int x[RC_MAX_P];
                        for (int k = 0; k < P; ++k) x[k] = 0;
                        for (int d = 0; d < nnz; ++d) {
                            const uint32_t w = out.words[(size_t)g * nnzw + d];
                            x[w & 0xFFu] = (int)((w >> 8) & 0xFFu);
                        }
                        int xs[4] = {0, 0, 0, 0};
                        for (int q = 0; q < nS; ++q) xs[q] = x[dims[1 + q]] - (((pb.zeroMask >> (1 + q)) & 1u) ? 0 : 1);
                        const int a = x[pb.rk] - 1;
                        const int other = ((xs[0] * Mp[1] + xs[1]) * Mp[2] + xs[2]) * Mp[3] + xs[3];
                        bool in = a >= 0 && a < MA && other >= 0 && other < SA;
                        for (int q = 0; q < nS && in; ++q) in = xs[q] >= 0 && xs[q] < Mp[q];
                        for (int q = 0; q < nU && in; ++q) in = x[dims[1 + nS + q]] == 1;
                        if (!in) { gerr = "pattern-mode image out of range (state outside the box)"; return; }
                        const int idx = a * SA + other;
                        if (ref > 0xFE || ((rec[6 + (idx >> 2)] >> (8 * (idx & 3))) & 0xFFu) != 0xFFu) { gerr = "pattern-mode image out of range (state index)"; return; }
                        rec[6 + (idx >> 2)] = (rec[6 + (idx >> 2)] & ~(0xFFu << (8 * (idx & 3)))) | ((uint32_t)ref << (8 * (idx & 3)));
                        if (!gather_rec(e, i, ref, live, el[(size_t)idx * 2], el[(size_t)idx * 2 + 1])) { gerr = "too many transition entries per state"; return; }
                    }
                    static_assert(6 * 4 + RC_PAT_MAXL <= RC_PLANE_N * 4, "state indices fit the lane record");
                    rec[2] = shape | ((uint32_t)nU << 20) | (pb.unit ? RC_PLANE_UNIT : 0u);
                    rec[3] = (uint32_t)so[i];
                    rec[4] = (uint32_t)i;
                    rec[5] = (uint32_t)bOff;
                    recPos[(size_t)e * nPat + i] = (long long)GL.size();
                    elPos[(size_t)e * nPat + i] = hdr[7] + bOff;
                    grpOf[(size_t)e * nPat + i] = v;
                    boxN[(size_t)e * nPat + i] = nBox;
                    laneOf[(size_t)e * nPat + i] = lanes;
                    binOf[(size_t)e * nPat + i] = bI;
                    if (lanes == 0) binShape = shape;
                    else if (binShape != shape) { gerr = "pattern-mode bin with more than one shape"; return; }
                    GL.insert(GL.end(), rec, rec + RC_PLANE_N);
                    GE.insert(GE.end(), el.begin(), el.end());
                    lanes += 1;
                    bOff += nBox;
                } else if (!rows) {
                    for (int s = 0; s < live; ++s) {
                        const long long g = out.epochBase[(size_t)e] + so[i] + s;
                        const uint32_t mt = out.meta[(size_t)g];
                        const int n = (int)(mt & 0xFFu);
                        if (n > 8) { gerr = "keyed bin with more than 8 components"; return; }
                        uint32_t rec[RC_LANE_N] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int d = 0; d < 8; ++d) {
                            uint32_t row = RC_KROWS - 1, up = (uint32_t)(lanes + s);
                            if (d < n) {
                                const uint32_t w = out.words[(size_t)g * nnzw + d];
                                const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                                row = (uint32_t)((int)rb[k] + x - 1);
                                if ((w >> 16) != RC_NO_UP) up = (uint32_t)(lanes + (int)(w >> 16));
                            }
                            if (row > 0xFFu || up > 0xFFu) { gerr = "keyed bin image out of range"; return; }
                            rec[d >> 2] |= row << (8 * (d & 3));
                            rec[2 + (d >> 2)] |= up << (8 * (d & 3));
                        }
                        rec[4] = mt;
                        rec[5] = (uint32_t)(so[i] + s);
                        if (!gather_rec(e, i, s, live, rec[6], rec[7])) { gerr = "too many transition entries per state"; return; }
                        GL.insert(GL.end(), rec, rec + RC_LANE_N);
                    }
                    lanes += live;
                } else {
                    // rows of one pattern share the stride of its longest row; length, row branch and other branches are per row
                    // (the parts of a live set left by a Split differ in all three)
                    const int nr = ro[i + 1] - ro[i];
                    // rows of one part share a stride (its longest row, made odd): a row reads its neighbour rows — always
                    // of its own part — at its own element index, and finds zeros where the neighbour is shorter
                    std::vector<int> rOff((size_t)nr + 1, 0);
                    for (int r = 0; r < nr; ++r)
                        rOff[(size_t)r + 1] = rOff[(size_t)r] + (out.rowPmAll[(size_t)(out.rowEpochBase[(size_t)e] + ro[i] + r)] | 1);
                    for (int r = 0; r < nr; ++r) {
                        const long long grow = out.rowEpochBase[(size_t)e] + ro[i] + r;
                        const int rk = out.rowBrAll[(size_t)grow];
                        const int Mp = out.rowPmAll[(size_t)grow] | 1;
                        int M = 0, NO = 0;
                        while (M < RC_ROW_MAXM && out.rowRefAll[(size_t)grow * RC_ROW_MAXM + M] != 0xFFFF) ++M;
                        while (NO < RC_ROW_NO && ((out.rowWordsAll[(size_t)grow * RC_ROW_NO + NO] >> 8) & 0xFFu) != 0) ++NO;
                        uint32_t rec[RC_RLANE_N] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                        uint32_t off[RC_ROW_NO + 1];
                        off[0] = (uint32_t)(bOff + rOff[(size_t)r]);
                        static_assert(RC_ROW_NO == 7, "row lane record packs seven other branches");
                        for (int d = 0; d < RC_ROW_NO; ++d) {
                            uint32_t row = RC_KROWS - 1;
                            off[d + 1] = RC_ROW_BCAP;
                            if (d < NO) {
                                const uint32_t w = out.rowWordsAll[(size_t)grow * RC_ROW_NO + d];
                                const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                                row = (uint32_t)((int)rb[k] + x - 1);
                                if ((w >> 16) != RC_NO_UP) off[d + 1] = (uint32_t)(bOff + rOff[(size_t)(w >> 16)]);
                            }
                            if (row > 0xFFu || off[d + 1] > RC_ROW_BCAP) { gerr = "row bin image out of range"; return; }
                            rec[d >> 2] |= row << (8 * (d & 3));
                        }
                        if (rb[rk] > 0xFFu || off[0] > RC_ROW_BCAP) { gerr = "row bin image out of range"; return; }
                        rec[1] |= (uint32_t)rb[rk] << 24;
                        rec[2] = (uint32_t)M | ((uint32_t)NO << 8);
                        for (int q = 0; q < RC_ROW_NO + 1; ++q) rec[3 + (q >> 1)] |= off[q] << (16 * (q & 1));
                        rec[7] = (uint32_t)so[i];
                        if (grow > 0xFFFFFFFFLL) { gerr = "bin images too large"; return; }
                        rec[8] = (uint32_t)grow;
                        rec[9] = (uint32_t)i;
                        GL.insert(GL.end(), rec, rec + RC_RLANE_N);
                        for (int j = 0; j < Mp; ++j) {
                            uint32_t g0 = 0u, g1 = 0u;
                            if (j < M) {
                                const int ref = (int)out.rowRefAll[(size_t)grow * RC_ROW_MAXM + j];
                                if (!gather_rec(e, i, ref, live, g0, g1)) { gerr = "too many transition entries per state"; return; }
                            }
                            GE.push_back(g0);
                            GE.push_back(g1);
                        }
                    }
                    lanes += nr;
                    bOff += rOff[(size_t)nr];
                }
            }
            hdr[2] = lanes;
            hdr[3] = bOff;
            if (patM) hdr[11] = (int)binShape;        // one shape per pattern bin: the block dispatches on it without waiting for a lane record
            if (patM && e == 0) {
                // the seed state m is the last one in the lane's own order too (all coordinates at their maximum)
                bool seedLast = true;
                size_t el0 = (size_t)hdr[7] * 2;
                for (int q = p0; q < p1 && seedLast; ++q) {
                    const int nb = boxN[(size_t)e * nPat + out.kBinPats[(size_t)q]];
                    for (int x = 0; x < nb; ++x)
                        if ((GE[el0 + (size_t)x * 2] == RC_IMG_SEED) != (x == nb - 1)) seedLast = false;
                    el0 += (size_t)nb * 2;
                }
                if (seedLast) hdr[3] |= RC_HDR_SEED;
            }
            out.grpMaxFetch[(size_t)v] = std::max(out.grpMaxFetch[(size_t)v], hdr[5]);
            out.grpMaxB[(size_t)v] = std::max(out.grpMaxB[(size_t)v], bOff);
        }
    };
    {
        int nT = plan_threads();
        if (nPat < 256) nT = 1;
        // largest groups first: the longest task bounds the phase
        std::vector<int> bysize((size_t)nGrp);
        std::vector<long long> gsz((size_t)nGrp, 0);
        for (int v = 0; v < nGrp; ++v) {
            bysize[(size_t)v] = v;
            const int b0 = out.kBinBase[(size_t)v], nb = out.nKBins[(size_t)v];
            if (nb > 0) gsz[(size_t)v] = (long long)out.kBinPatOff[(size_t)(b0 + nb)] - out.kBinPatOff[(size_t)b0];
        }
        std::stable_sort(bysize.begin(), bysize.end(), [&](int a, int b) { return gsz[(size_t)a] > gsz[(size_t)b]; });
        std::vector<double> grpMs((size_t)nGrp, 0.0);
        const auto tAll0 = std::chrono::steady_clock::now();
        std::atomic<int> nxt(0);
        auto w = [&]() {
            for (;;) {
                const int q = nxt.fetch_add(1);
                if (q >= nGrp) break;
                const auto tg0 = std::chrono::steady_clock::now();
                do_group(bysize[(size_t)q]);
                grpMs[(size_t)bysize[(size_t)q]] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg0).count();
            }
        };
        if (nT == 1) w();
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nT; ++t) pool.emplace_back(w);
            for (auto& th : pool) th.join();
        }
        if (std::getenv("RC_PLAN_TIMING")) {
            double mx = 0.0, sum = 0.0;
            for (double m : grpMs) { mx = std::max(mx, m); sum += m; }
            std::fprintf(stderr, "[rc plan]   bin images: groups %.1f ms wall, longest %.1f ms, sum %.1f ms\n",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tAll0).count(), mx, sum);
        }
        for (int v = 0; v < nGrp; ++v) if (!grpErr[(size_t)v].empty()) return grpErr[(size_t)v];
        // one behind the other, in group order: local positions become global ones
        size_t nLane = 0, nElem = 0;
        for (int v = 0; v < nGrp; ++v) { nLane += grpLane[(size_t)v].size(); nElem += grpElem[(size_t)v].size(); }
        if (nLane / 4 > 0x7FFFFFFFull || nElem / 2 > 0x7FFFFFFFull) return "bin images too large";
        out.kLane.resize(nLane);
        out.kElem.resize(nElem);
        size_t lp = 0, ep = 0;
        for (int v = 0; v < nGrp; ++v) {
            const int e = v / RC_NGROUP;
            if (v % RC_NGROUP == 0) out.ep[(size_t)e].elBase = (int)(ep / 2);
            const int b0 = out.kBinBase[(size_t)v];
            for (int bI = b0; bI < b0 + out.nKBins[(size_t)v]; ++bI) {
                int* hdr = &out.kBinHdr[(size_t)bI * RC_HDR_N];
                hdr[6] += (int)(lp / 4);
                hdr[7] += (int)(ep / 2);
                if (out.epochMode[(size_t)v] >= 3)
                    for (int q = out.kBinPatOff[(size_t)bI]; q < out.kBinPatOff[(size_t)bI + 1]; ++q) {
                        const int i = out.kBinPats[(size_t)q];
                        recPos[(size_t)e * nPat + i] += (long long)lp;
                        elPos[(size_t)e * nPat + i] += (int)(ep / 2);
                    }
            }
            std::copy(grpLane[(size_t)v].begin(), grpLane[(size_t)v].end(), out.kLane.begin() + (long long)lp);
            std::copy(grpElem[(size_t)v].begin(), grpElem[(size_t)v].end(), out.kElem.begin() + (long long)ep);
            lp += grpLane[(size_t)v].size();
            ep += grpElem[(size_t)v].size();
            std::vector<uint32_t>().swap(grpLane[(size_t)v]);
            std::vector<uint32_t>().swap(grpElem[(size_t)v]);
        }
    }
    // ---- direct hand-over across a Join between pattern-mode epochs.  The lane that owns a pattern holds all its b(x): it
    //      applies the Join itself (new_b[J(x)] += b(x), Core.hs:172-188) and writes the result where the pattern's lane of the
    //      next epoch reads it — contiguous, in that lane's own state order — instead of writing slots that the next block
    //      gathers through element records and transition entries (three dependent loads per state at the start of every block).
    //      A bin group reads this way if every pattern of it was in pattern mode in the previous epoch. ----
    const auto tDir0 = std::chrono::steady_clock::now();
    out.grpDirIn.assign((size_t)nE * RC_NGROUP, 0);
    std::vector<long long> dirRun((size_t)nE, 0);
    out.dirOff = out.maxEpochSlots;
    out.dirSize = 0;
    static const bool dirEnv = std::getenv("RC_PAT_DIRHAND") ? std::atoi(std::getenv("RC_PAT_DIRHAND")) != 0 : true;
    for (int e = 1; e < nE && dirEnv; ++e) {
        if (sev[(size_t)e - 1].type != RC_EV_JOIN_D) continue;
        const int* trOff = &out.trOff[(size_t)out.trBase[(size_t)e - 1]];
        for (int g = 0; g < RC_NPATG; ++g) {
            const int v = e * RC_NGROUP + g;
            if (out.epochMode[(size_t)v] < RC_MODE_PAT || out.nKBins[(size_t)v] == 0) continue;
            const int b0 = out.kBinBase[(size_t)v];
            const int q0 = out.kBinPatOff[(size_t)b0], q1 = out.kBinPatOff[(size_t)(b0 + out.nKBins[(size_t)v])];
            std::vector<uint8_t> maps((size_t)(q1 - q0) * RC_PAT_MAXL, 0);
            std::atomic<bool> okAll(true);
            // (the maps of the patterns are independent of each other: built side by side)
            auto build_maps = [&](int qa, int qb) {
            bool ok = true;
            for (int q = qa; q < qb && ok && okAll.load(std::memory_order_relaxed); ++q) {
                const int i = out.kBinPats[(size_t)q];
                const long long rpN = recPos[(size_t)e * nPat + i], rpO = recPos[(size_t)(e - 1) * nPat + i];
                if (rpN < 0 || rpO < 0) { ok = false; break; }
                const uint32_t* rn = &out.kLane[(size_t)rpN];
                const uint32_t* ro = &out.kLane[(size_t)rpO];
                const int liveO = out.slotOff[(size_t)(e - 1) * (nPat + 1) + i + 1] - out.slotOff[(size_t)(e - 1) * (nPat + 1) + i];
                const int nN = boxN[(size_t)e * nPat + i], nO = boxN[(size_t)(e - 1) * nPat + i];
                if (nN > RC_PAT_MAXL || nO > RC_PAT_MAXL || liveO > nO) { ok = false; break; }
                // (places of a bounding box without a live state: 0xFF on both sides)
                uint8_t localOfRef[RC_PAT_MAXL];
                for (int x = 0; x < nO; ++x) {
                    const uint32_t ref = (ro[6 + (x >> 2)] >> (8 * (x & 3))) & 0xFFu;
                    if (ref != 0xFFu) localOfRef[ref] = (uint8_t)x;
                }
                uint8_t* mp = &maps[(size_t)(q - q0) * RC_PAT_MAXL];
                for (int x = 0; x < RC_PAT_MAXL; ++x) mp[x] = 0xFF;
                int mapped = 0;
                for (int y = 0; y < nN && ok; ++y) {
                    const int ref = (int)((rn[6 + (y >> 2)] >> (8 * (y & 3))) & 0xFFu);
                    if (ref == 0xFF) continue;
                    const int sl = out.slotOff[(size_t)e * (nPat + 1) + i] + ref;
                    for (int t = trOff[sl]; t < trOff[sl + 1]; ++t) {
                        const uint32_t ent = out.trEnt[(size_t)(out.trEntBase[(size_t)e - 1] + t)];
                        const int src = (int)(ent & 0xFFFFu);
                        if ((ent >> 16) != RC_W_ONE || src >= liveO || mp[localOfRef[src]] != 0xFF) { ok = false; break; }
                        mp[localOfRef[src]] = (uint8_t)y;
                        ++mapped;
                    }
                }
                ok = ok && mapped == liveO;                            // a Join maps every live state somewhere
            }
            if (!ok) okAll.store(false, std::memory_order_relaxed);
            };
            {
                int nT = plan_threads();
                const int chunk = 512;
                if (q1 - q0 < 4 * chunk) nT = 1;
                std::atomic<int> nxt(q0);
                auto w = [&]() {
                    for (;;) {
                        const int qa = nxt.fetch_add(chunk);
                        if (qa >= q1) break;
                        build_maps(qa, std::min(q1, qa + chunk));
                    }
                };
                std::vector<std::thread> pool;
                for (int t = 1; t < nT; ++t) pool.emplace_back(w);
                w();
                for (auto& th : pool) th.join();
            }
            if (!okAll.load()) continue;
            out.grpDirIn[(size_t)v] = 1;
            // the bin's region of the lane-ordered part: [state of the box][lane], RC_DIR_LANES lanes — a lane reads its states at
            // that stride, a warp 256 contiguous bytes per state, straight into registers
            for (int bI = b0; bI < b0 + out.nKBins[(size_t)v]; ++bI) {
                int* hdr = &out.kBinHdr[(size_t)bI * RC_HDR_N];
                hdr[3] |= RC_HDR_DIRIN;
                const uint32_t sh = (uint32_t)hdr[11];
                int nb = (int)(sh & 0xFu);
                for (int w = 1; w < 5; ++w) nb *= std::max(1, (int)((sh >> (4 * w)) & 0xFu));
                hdr[12] = (int)dirRun[(size_t)e];
                dirRun[(size_t)e] += (long long)RC_DIR_LANES * nb;
            }
            for (int q = q0; q < q1; ++q) {
                const int i = out.kBinPats[(size_t)q];
                uint32_t* ro = &out.kLane[(size_t)recPos[(size_t)(e - 1) * nPat + i]];
                const int nN = boxN[(size_t)e * nPat + i], nO = boxN[(size_t)(e - 1) * nPat + i];
                const uint8_t* mp = &maps[(size_t)(q - q0) * RC_PAT_MAXL];
                for (int w = 6; w < 15; ++w) ro[w] = 0xFFFFFFFFu;
                for (int x = 0; x < nO; ++x) ro[6 + (x >> 2)] = (ro[6 + (x >> 2)] & ~(0xFFu << (8 * (x & 3)))) | ((uint32_t)mp[x] << (8 * (x & 3)));
                ro[2] |= RC_PLANE_DIROUT | ((uint32_t)nN << 24);
                ro[3] = (uint32_t)(out.kBinHdr[(size_t)binOf[(size_t)e * nPat + i] * RC_HDR_N + 12] + laneOf[(size_t)e * nPat + i]);
                // the lane sums into [state][lane] doubles of its block's shared memory
                int& mb = out.grpMaxB[(size_t)grpOf[(size_t)(e - 1) * nPat + i]];
                mb = std::max(mb, rc_pat_lanes(out.epochMode[(size_t)grpOf[(size_t)(e - 1) * nPat + i]]) * nN);
            }
            out.dirSize = std::max(out.dirSize, dirRun[(size_t)e]);
        }
    }
    if (out.dirSize > 0) {
        out.dirOff = (out.maxEpochSlots + 1) & ~1LL;
        out.maxEpochSlots = out.dirOff + out.dirSize;
    }
    if (std::getenv("RC_PLAN_TIMING"))
        std::fprintf(stderr, "[rc plan]   bin images: hand-over maps %.1f ms\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tDir0).count());
    return "";
}

std::string build_plan(int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec,
                       const std::vector<StructEv>& sev, const std::vector<GapInfo>& gaps, bool hasSplit,
                       int smemBudgetBytes, PlanHost& out, int patternMode) {
    if (!key_fits(P, maxAf)) return "state key (maxAf+1)^P does not fit 63 bits";
    Ctx c;
    c.P = P;
    c.maxAf = maxAf;
    c.stride = std::min(P, maxAf);
    c.base = (uint64_t)(maxAf + 1);
    for (int k = P - 1; k >= 0; --k) c.pw[k] = k == P - 1 ? 1 : c.pw[k + 1] * c.base;
    const int nEpochs = (int)sev.size() + 1;
    std::vector<PatPlan> pps((size_t)nPat);
    // RC_PLAN_TIMING=1: wall time of the planner's phases on stderr
    static const bool timing = std::getenv("RC_PLAN_TIMING") != nullptr;
    auto tPhase = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rc plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tPhase).count());
        tPhase = now;
    };

    int nThreads = plan_threads();
    if (nPat < 256) nThreads = 1;
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            int i0 = next.fetch_add(64);
            if (i0 >= nPat) break;
            int i1 = std::min(nPat, i0 + 64);
            for (int i = i0; i < i1; ++i)
                plan_pattern(c, patterns + (size_t)i * P, nVec, sev, gaps, hasSplit, pps[(size_t)i]);
        }
    };
    if (nThreads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nThreads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    lap("live sets per pattern");
    // the same pool for the later per-pattern copy loops: fn(i0, i1) on chunks of patterns with disjoint outputs
    auto par_for = [&](int n, int chunk, const std::function<void(int, int)>& fn) {
        if (nThreads == 1 || n <= chunk) { fn(0, n); return; }
        std::atomic<int> nxt(0);
        auto w = [&]() {
            for (;;) {
                const int i0 = nxt.fetch_add(chunk);
                if (i0 >= n) break;
                fn(i0, std::min(n, i0 + chunk));
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < nThreads; ++t) pool.emplace_back(w);
        for (auto& th : pool) th.join();
    };
    int nnzw = 1, maxLive = 1;
    for (auto& pp : pps) {
        if (!pp.err.empty()) return pp.err;
        nnzw = std::max(nnzw, pp.maxNnz);
        maxLive = std::max(maxLive, pp.maxLive);
    }

    out = PlanHost();
    out.P = P; out.maxAf = maxAf; out.nPat = nPat; out.nEpochs = nEpochs; out.nnzw = nnzw; out.maxLive = maxLive;
    // patternMode & 2 ("pattern_embed" option, or RC_PAT_EMBED=1): live sets that are not boxes run in pattern mode on their
    // bounding boxes (pat_box) instead of row / state mode.  Off by default: measured slower on the reference's fourPop data
    // (C2, 64 models: 5.76 against 5.12 ms) and on C5 (0.94 against 0.82 ms) — 1 000 patterns spread over so many shape classes
    // that a 64-lane bin holds 3 - 10 of them, and the boxes carry states that are not live.
    static const bool embEnv = std::getenv("RC_PAT_EMBED") ? std::atoi(std::getenv("RC_PAT_EMBED")) != 0 : false;
    out.zeroRow = (hasSplit && (embEnv || (patternMode & 2)) && patternMode != 0) ? 1 : 0;

    // Two tiers.  Fast tier (rc_propagate_fast_kernel): one live state per thread, <= RC_BLOCK states per
    // pattern, <= 8 non-zero components per state.  Big tier (rc_propagate_kernel): several states per thread,
    // block capacity = the largest pattern rounded up to the block size, bounded by shared memory.
    // A pattern too large for one state per thread can still run on the trajectory-key tier if in EVERY epoch its live set
    // cuts into rows that fit one row-mode bin (`keyedBig`; live sets of several hundred states after admixture pulses).
    std::vector<int> fastIds, bigIds, keyedBig;
    std::vector<char> isKeyedBig((size_t)nPat, 0);
    int maxLiveBig = 0, maxLiveFast = 1;
    const bool fastOK = nnzw <= 8 && rc_row_len(P, maxAf + 1) <= RC_ROWCAP;
    for (int i = 0; i < nPat; ++i) {
        const PatPlan& pp = pps[(size_t)i];
        if (fastOK && pp.maxLive <= RC_BLOCK && pp.maxR <= 2 * RC_BLOCK - 2) {
            fastIds.push_back(i);
            maxLiveFast = std::max(maxLiveFast, pp.maxLive);
        } else {
            bigIds.push_back(i);
            maxLiveBig = std::max(maxLiveBig, pp.maxLive);
            bool rowsFit = fastOK && patternMode >= 0;
            for (int e = 0; e < nEpochs && rowsFit; ++e) {
                int rr = 0;
                for (int k = 0; k < P; ++k) rr += pp.mx[(size_t)e * P + k];
                rowsFit = pp.rowOK[(size_t)e] && pp.rowCnt[(size_t)e] <= RC_BLOCK && pp.rowPad[(size_t)e] <= RC_ROW_BCAP && rr <= RC_KROWS - 2;
            }
            if (rowsFit) { keyedBig.push_back(i); isKeyedBig[(size_t)i] = 1; }
        }
    }
    // smallest block that holds the largest fast-tier pattern: fewer warps per barrier, finer packing
    (void)maxLiveFast;
    int fastBlock = 256;   // measured on C3: 256-thread blocks beat 64-thread blocks (141 vs 226 ms per 44 models)
    for (int i : fastIds)
        while (pps[(size_t)i].maxR > 2 * fastBlock - 2) fastBlock *= 2;
    out.fastBlock = fastBlock;
    const int maxPatFast = std::getenv("RC_MAXPAT") ? std::atoi(std::getenv("RC_MAXPAT")) : std::max(1, std::min(RC_NPB, fastBlock / P));
    int cap = ((std::max(maxLiveBig, 1) + RC_BLOCK - 1) / RC_BLOCK) * RC_BLOCK;
    // bytes per slot in the big-tier kernel: 2 b buffers + d accumulator (24) + words (4*nnzw) + meta (4) + owner (2)
    int perSlot = 24 + 4 * nnzw + 4 + 2;
    int fixed = RC_NPB * P * 16 + RC_NPB * 32 + rc_row_len(P, maxAf + 1) * 8 + 1024;
    // a live set beyond the shared-memory tier: the same kernel with its per-slot arrays in global memory (rc_kernels.cu)
    out.bigGlobal = (!bigIds.empty() && (long long)cap * perSlot + fixed > smemBudgetBytes) ? 1 : 0;
    if (maxLiveBig >= (int)RC_NO_UP) return "live set of one pattern (" + std::to_string(maxLiveBig) + " states) exceeds 16-bit slot indices";
    out.cap = cap;

    // best-fit decreasing on the per-pattern maximum live size, under the tier's pattern / rate-row limits
    auto pack = [&](const std::vector<int>& ids, int capSlots, int maxPats, int capR, std::vector<int>& patOff,
                    std::vector<int>& pats) {
        std::vector<int> order(ids);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pps[(size_t)a].maxLive > pps[(size_t)b].maxLive; });
        std::vector<std::vector<int>> bins;
        std::vector<int> freeSlots, freeR;
        std::multimap<int, int> open;   // free slots -> bin
        for (int i : order) {
            int sz = pps[(size_t)i].maxLive, rr = pps[(size_t)i].maxR;
            int b = -1, tries = 0;
            for (auto it = open.lower_bound(sz); it != open.end() && tries < 8; ++it, ++tries)
                if (freeR[(size_t)it->second] >= rr) { b = it->second; open.erase(it); break; }
            if (b < 0) {
                b = (int)bins.size();
                bins.emplace_back();
                freeSlots.push_back(capSlots);
                freeR.push_back(capR);
            }
            bins[(size_t)b].push_back(i);
            freeSlots[(size_t)b] -= sz;
            freeR[(size_t)b] -= rr;
            if ((int)bins[(size_t)b].size() < maxPats && freeSlots[(size_t)b] > 0 && freeR[(size_t)b] > 0)
                open.emplace(freeSlots[(size_t)b], b);
        }
        // heaviest bins first so the hardware block scheduler ends with the light ones
        std::vector<int> border(bins.size());
        for (size_t b = 0; b < bins.size(); ++b) border[b] = (int)b;
        std::stable_sort(border.begin(), border.end(), [&](int a, int b) { return freeSlots[(size_t)a] < freeSlots[(size_t)b]; });
        patOff.assign(1, 0);
        pats.clear();
        for (int b : border) {
            for (int i : bins[(size_t)b]) pats.push_back(i);
            patOff.push_back((int)pats.size());
        }
        return (int)bins.size();
    };
    // ---- trajectory keys (see rc_plan.h): a_k(t) of a pattern depends only on the pattern's derived-allele counts
    //      on the leaves merged into branch k so far, so it is computed once per distinct (branch, sub-tuple)
    {
        std::vector<uint32_t> leaf((size_t)P);
        std::vector<char> dead((size_t)P, 0);
        for (int k = 0; k < P; ++k) leaf[(size_t)k] = 1u << k;
        out.keyId.assign((size_t)nEpochs * nPat * P, 0);
        out.ep.assign((size_t)nEpochs, PlanHost::Epoch());
        std::vector<int> prevKeyOff(1, 0);
        for (int e = 0; e < nEpochs; ++e) {
            if (e > 0) {
                const StructEv& ev = sev[(size_t)e - 1];
                leaf[(size_t)ev.k] |= leaf[(size_t)ev.l];
                if (ev.type == RC_EV_JOIN_D) { leaf[(size_t)ev.l] = 0; dead[(size_t)ev.l] = 1; }
            }
            PlanHost::Epoch& E = out.ep[(size_t)e];
            E.keyOff = (int)out.keyBranch.size();
            // key of (branch k, counts on the leaves merged into k): the counts as digits base maxAf + 1 (fits 63 bits:
            // key_fits), one index per branch; ids in order of first appearance over the patterns
            std::vector<Index> ids((size_t)P);
            int nIds = 0;
            const uint64_t deadKey = ~(uint64_t)0 - 1;
            for (int i = 0; i < nPat; ++i) {
                const int32_t* m = patterns + (size_t)i * P;
                for (int k = 0; k < P; ++k) {
                    uint64_t ks = 0;
                    if (dead[(size_t)k]) ks = deadKey;
                    else for (int q = 0; q < P; ++q) if ((leaf[(size_t)k] >> q) & 1u) ks = ks * c.base + (uint64_t)m[q];
                    auto it = ids[(size_t)k].find(ks);
                    int id;
                    if (it == ids[(size_t)k].end()) {
                        id = nIds++;
                        ids[(size_t)k].emplace(ks, id);
                        out.keyBranch.push_back((uint8_t)k);
                        out.keyMaxJ.push_back(0);
                        int pa = -1, pb = -1, op = RC_KOP_COPY, init = 0;
                        if (e == 0) { op = RC_KOP_INIT; init = nVec[k] - m[k]; }
                        else {
                            const StructEv& ev = sev[(size_t)e - 1];
                            const int* prev = &out.keyId[((size_t)(e - 1) * nPat + i) * P];
                            pa = prev[k];
                            if (ev.type == RC_EV_JOIN_D) {
                                if (k == ev.k) { op = RC_KOP_JOIN; pb = prev[ev.l]; }
                                else if (k == ev.l) { op = RC_KOP_ZERO; }
                            } else {
                                if (k == ev.k) { op = RC_KOP_SPLIT_K; pb = prev[ev.l]; }
                                else if (k == ev.l) { op = RC_KOP_SPLIT_L; }
                            }
                        }
                        out.keyPa.push_back(pa);
                        out.keyPb.push_back(pb);
                        out.keyOp.push_back((uint8_t)op);
                        out.keyInit.push_back(init);
                    } else id = it->second;
                    out.keyId[((size_t)e * nPat + i) * P + k] = id;
                    uint8_t mxk = pps[(size_t)i].mx[(size_t)e * P + k];
                    uint8_t& mj = out.keyMaxJ[(size_t)E.keyOff + id];
                    if (mxk > mj) mj = mxk;
                }
            }
            E.nKeys = nIds;
            E.thrOff = (int)out.thrKey.size();
            int pairs = 1;   // pair 0 of every step row is {dt, 0}
            for (int q = 0; q < E.nKeys; ++q) {
                int mj = out.keyMaxJ[(size_t)E.keyOff + q];
                out.keyThrBase.push_back((int)out.thrKey.size() - E.thrOff);
                // (a plan with a Split: the pair of x = 0, {1, 0}, leads every key — pattern mode on bounding boxes, pat_box)
                out.keyPairBase.push_back(pairs + out.zeroRow);
                pairs += rc_pair_stride(mj + out.zeroRow);        // odd stride: see rc_types.h (RC_HDR_N)
                for (int t = 0; t < std::max(1, mj); ++t) out.thrKey.push_back(q);
            }
            E.nThr = (int)out.thrKey.size() - E.thrOff;
            E.rowPairs = pairs;
        }
        out.totalKeys = (int)out.keyBranch.size();
        out.lastKeyId.assign(out.keyId.end() - (size_t)nPat * P, out.keyId.end());
        out.lastPairBase.assign((size_t)nPat, -1);
        out.lastSingle = nPat > 0;
        for (int i = 0; i < nPat; ++i) {
            int rk = -1, nz = 0;
            for (int k = 0; k < P; ++k) if (pps[(size_t)i].mx[(size_t)(nEpochs - 1) * P + k]) { rk = k; ++nz; }
            if (nz == 1)
                out.lastPairBase[(size_t)i] = out.keyPairBase[(size_t)out.ep[(size_t)nEpochs - 1].keyOff + out.lastKeyId[(size_t)i * P + rk]];
            else out.lastSingle = false;
        }
    }

    lap("trajectory keys");
    // ---- fast-tier bins: patterns in lexicographic order (neighbours share trajectory keys), packed first-fit over a
    //      short window of open bins under the slot / pattern / pair-row limits of the epochs the bins are used for.
    //      Two sets: one spanning all epochs (self-contained kernel: events handled inside the block) and one per
    //      epoch (keyed kernel: one launch per epoch, b carried through global memory, bins sized for that epoch only).
    {
        std::vector<int> order(fastIds);
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            const int32_t* x = patterns + (size_t)a * P;
            const int32_t* y = patterns + (size_t)b * P;
            for (int k = 0; k < P; ++k) if (x[k] != y[k]) return x[k] < y[k];
            return a < b;
        });
        // per-epoch (keyed) bins also hold the big patterns that run in row mode
        std::vector<int> orderK(order);
        orderK.insert(orderK.end(), keyedBig.begin(), keyedBig.end());
        std::sort(orderK.begin(), orderK.end(), [&](int a, int b) {
            const int32_t* x = patterns + (size_t)a * P;
            const int32_t* y = patterns + (size_t)b * P;
            for (int k = 0; k < P; ++k) if (x[k] != y[k]) return x[k] < y[k];
            return a < b;
        });
        const int rowCap = RC_KROWS - 2;          // smem pair rows per parity minus the dt row and the padding row
        struct OpenBin {
            std::vector<int> ids, slotsE, rowsUsed, rowsSelf;
            int rowStates = 0, cls = 0;
            std::vector<std::unordered_map<int, int>> keyRows;   // per epoch: key -> pair rows needed
        };
        // measured on C3 (44 models): 36.7 ms with lexicographic bins, 33.4 ms with cost-class bins; RC_BINSORT=0 restores the former
        const bool binSort = std::getenv("RC_BINSORT") ? std::atoi(std::getenv("RC_BINSORT")) != 0 : true;
        // mode: 0 one state per lane, 1 one row per lane, 3 one pattern per lane (bins of one cost class only)
        // packDirect: a direct group needs no pair rows in shared memory (no row limit)
        auto pack_range = [&](int e0, int e1, int maxPats, int mode, bool keyedBins, const std::vector<int>& ids, std::vector<int>& patOff, std::vector<int>& pats,
                              bool packDirect = false) {
            const int nE = e1 - e0 + 1;
            const bool rowMode = mode != 0, patMode = mode >= 3;
            const int patMaxStates = rc_pat_cap(mode);
            auto lanesOf = [&](const PatPlan& pp, int e) { return patMode ? 1 : rowMode ? pp.rowCnt[(size_t)e] : pp.live[(size_t)e]; };
            auto clsCompute = [&](int i) {
                const PatPlan& pp = pps[(size_t)i];
                if (patMode) {
                    PatBox pb;
                    pat_box(&pp.mx[(size_t)e0 * P], &pp.mn[(size_t)e0 * P], pp.fullBox[(size_t)e0] != 0, pp.rowK[(size_t)e0], P, patMaxStates, pb);
                    const uint32_t sh = pb.shape;
                    const int nU = pb.nU;
                    static const bool nuClass = std::getenv("RC_PAT_NUCLASS") ? std::atoi(std::getenv("RC_PAT_NUCLASS")) != 0 : false;
                    return (int)(sh | (nuClass ? (uint32_t)nU << 20 : 0u));   // blocks are uniform in shape (the unrolled update)
                }
                if (rowMode) return pp.rowNO[(size_t)e0] * 64 + pp.rowM[(size_t)e0];
                int c = 0;
                for (int k = 0; k < P; ++k) c += pp.mx[(size_t)e0 * P + k] ? 1 : 0;
                return c;
            };
            // (the class of a pattern is asked for by every comparison of the sorts below and by every probe of an open bin)
            std::vector<int> clsV((size_t)nPat, 0);
            for (int i : ids) clsV[(size_t)i] = clsCompute(i);
            auto clsOf = [&](int i) { return clsV[(size_t)i]; };
            const size_t WINDOW = 12;
            std::vector<OpenBin> open;
            auto flush = [&](OpenBin& ob) {
                if (nE == 1) {
                    // a warp visits as many components as its widest state: keep patterns with equally many non-zero
                    // branches next to each other inside the bin
                    auto nnzOf = [&](int i) {
                        int c = 0;
                        for (int k = 0; k < P; ++k) c += pps[(size_t)i].mx[(size_t)e0 * P + k] ? 1 : 0;
                        return c;
                    };
                    if (patMode) {
                    } else if (rowMode)
                        std::stable_sort(ob.ids.begin(), ob.ids.end(), [&](int a, int b) {
                            const PatPlan &x = pps[(size_t)a], &y = pps[(size_t)b];
                            if (x.rowNO[(size_t)e0] != y.rowNO[(size_t)e0]) return x.rowNO[(size_t)e0] > y.rowNO[(size_t)e0];
                            return x.rowM[(size_t)e0] > y.rowM[(size_t)e0];
                        });
                    else
                        std::stable_sort(ob.ids.begin(), ob.ids.end(), [&](int a, int b) { return nnzOf(a) > nnzOf(b); });
                }
                for (int i : ob.ids) pats.push_back(i);
                patOff.push_back((int)pats.size());
            };
            auto try_add = [&](OpenBin& ob, int i) -> bool {
                const PatPlan& pp = pps[(size_t)i];
                if ((int)ob.ids.size() + 1 > maxPats) return false;
                for (int q = 0; q < nE; ++q)
                    if (ob.slotsE[(size_t)q] + lanesOf(pp, e0 + q) > fastBlock) return false;
                if (patMode && !ob.ids.empty() && ob.cls != clsOf(i)) return false;
                // padded states of the rows (odd stride) must fit the block's b buffer
                const int padAdd = (rowMode && !patMode) ? pp.rowPad[(size_t)e0] : 0;
                if (ob.rowStates + padAdd > RC_ROW_BCAP) return false;
                int addS[16] = {0}, selfS[16] = {0};
                std::vector<int> addV, selfV;
                if (nE > 16) { addV.assign((size_t)nE, 0); selfV.assign((size_t)nE, 0); }
                int* add = nE > 16 ? addV.data() : addS;
                int* addSelf = nE > 16 ? selfV.data() : selfS;
                for (int q = 0; q < nE; ++q) {
                    const int e = e0 + q;
                    for (int k = 0; k < P; ++k) {
                        int mxk = pp.mx[(size_t)e * P + k];
                        if (!mxk) continue;
                        addSelf[(size_t)q] += mxk;
                        int key = out.keyId[((size_t)e * nPat + i) * P + k];
                        auto it = ob.keyRows[(size_t)q].find(key);
                        int have = it == ob.keyRows[(size_t)q].end() ? 0 : it->second;
                        if (mxk > have) add[(size_t)q] += patMode ? rc_pair_stride(mxk + out.zeroRow) - (have ? rc_pair_stride(have + out.zeroRow) : 0) : mxk - have;
                    }
                    if (!packDirect && ob.rowsUsed[(size_t)q] + add[(size_t)q] > rowCap) return false;
                    // (pair rows without de-duplication: a limit of the self-contained kernel's rate table only)
                    if (!keyedBins && ob.rowsSelf[(size_t)q] + addSelf[(size_t)q] > 2 * fastBlock - 2) return false;
                }
                for (int q = 0; q < nE; ++q) {
                    const int e = e0 + q;
                    for (int k = 0; k < P; ++k) {
                        int mxk = pp.mx[(size_t)e * P + k];
                        if (!mxk) continue;
                        int key = out.keyId[((size_t)e * nPat + i) * P + k];
                        int& have = ob.keyRows[(size_t)q][key];
                        if (mxk > have) { ob.rowsUsed[(size_t)q] += patMode ? rc_pair_stride(mxk + out.zeroRow) - (have ? rc_pair_stride(have + out.zeroRow) : 0) : mxk - have; have = mxk; }
                    }
                    ob.rowsSelf[(size_t)q] += addSelf[(size_t)q];
                    ob.slotsE[(size_t)q] += lanesOf(pp, e);
                }
                ob.rowStates += padAdd;
                if (ob.ids.empty()) ob.cls = clsOf(i);
                ob.ids.push_back(i);
                return true;
            };
            // One-epoch bins: a block advances at the pace of its slowest warp (one barrier per grid step), and a warp at
            // the pace of its widest lane.  Patterns of equal cost class (row mode: other branches, row length; state
            // mode: non-zero branches) are therefore binned together; inside a class the lexicographic order keeps
            // patterns that share trajectory keys next to each other.
            std::vector<int> ord(ids);
            if (nE == 1 && (binSort || patMode)) {
                std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return clsOf(a) > clsOf(b); });
                // Patterns with exactly two live branches (the last epochs of a tree): inside a shape class every pattern is a
                // pair (key of A, key of B).  Lexicographic order keeps A fixed and walks through all B keys, so a bin of 64
                // patterns needs ~64 distinct B keys; 8 x 8 tiles of the (A key, B key) grid need 16 keys.
                static const bool tile = std::getenv("RC_PAT_TILE") ? std::atoi(std::getenv("RC_PAT_TILE")) != 0 : true;
                if (patMode && tile) {
                    auto two = [&](int i, int& ka, int& kb) {
                        const PatPlan& pp = pps[(size_t)i];
                        const int rk = pp.rowK[(size_t)e0];
                        int other = -1, n = 0;
                        for (int k = 0; k < P; ++k)
                            if (pp.mx[(size_t)e0 * P + k] && k != rk) { other = k; ++n; }
                        if (n != 1) return false;
                        ka = out.keyId[((size_t)e0 * nPat + i) * P + rk];
                        kb = out.keyId[((size_t)e0 * nPat + i) * P + other];
                        return true;
                    };
                    std::vector<int> twoA((size_t)nPat, -1), twoB((size_t)nPat, -1);
                    for (int i : ids) {
                        int ka = 0, kb = 0;
                        if (two(i, ka, kb)) { twoA[(size_t)i] = ka; twoB[(size_t)i] = kb; }
                    }
                    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
                        const int ca = clsOf(a), cb = clsOf(b);
                        if (ca != cb) return ca > cb;
                        const int a1 = twoA[(size_t)a], a2 = twoB[(size_t)a], b1 = twoA[(size_t)b], b2 = twoB[(size_t)b];
                        const bool ta = a1 >= 0, tb = b1 >= 0;
                        if (ta != tb) return ta;                        // a consistent order: two-branch patterns first,
                        if (!ta) return false;                          // the others keep their lexicographic order
                        if ((a1 >> 3) != (b1 >> 3)) return (a1 >> 3) < (b1 >> 3);
                        if ((a2 >> 3) != (b2 >> 3)) return (a2 >> 3) < (b2 >> 3);
                        if (a1 != b1) return a1 < b1;
                        return a2 < b2;
                    });
                }
            }
            for (int i : ord) {
                bool placed = false;
                for (size_t q = open.size(); q-- > 0 && !placed;) placed = try_add(open[q], i);
                if (placed) continue;
                if (open.size() >= WINDOW) {
                    flush(open.front());
                    open.erase(open.begin());
                }
                open.emplace_back();
                OpenBin& ob = open.back();
                ob.slotsE.assign((size_t)nE, 0);
                ob.rowsUsed.assign((size_t)nE, 0);
                ob.rowsSelf.assign((size_t)nE, 0);
                ob.keyRows.resize((size_t)nE);
                if (!try_add(ob, i)) ob.ids.push_back(i);     // a single pattern always fits its tier
            }
            for (OpenBin& ob : open) flush(ob);
        };
        // de-duplicated pair rows a bin fetches in epoch e, and the first row of every (pattern, branch)
        auto fetch_lists = [&](int e, const std::vector<int>& patOff, const std::vector<int>& pats, int b0, int b1,
                               std::vector<int>& fOff, std::vector<int>& fRows, std::vector<uint16_t>& rBase, size_t rBaseOff, bool pad) {
            const PlanHost::Epoch& E = out.ep[(size_t)e];
            Index rowsOf, startOf;                   // (flat, cleared per bin: a bin has a few dozen keys)
            std::vector<int> keyOrder;
            for (int bI = b0; bI < b1; ++bI) {
                fOff.push_back((int)fRows.size());
                rowsOf.clear();
                startOf.clear();
                keyOrder.clear();
                for (int q = patOff[(size_t)bI]; q < patOff[(size_t)bI + 1]; ++q) {
                    int i = pats[(size_t)q];
                    for (int k = 0; k < P; ++k) {
                        int mxk = pps[(size_t)i].mx[(size_t)e * P + k];
                        if (!mxk) continue;
                        int key = out.keyId[((size_t)e * nPat + i) * P + k];
                        auto it = rowsOf.find(key);
                        if (it == rowsOf.end()) { rowsOf.emplace(key, mxk); keyOrder.push_back(key); }
                        else if (mxk > it->second) it->second = mxk;
                    }
                }
                fRows.push_back(0);                 // smem pair row 0 <- {dt, 0}
                int row = 1;
                if (pad) {
                    // Pattern mode: ring rows in table order, every key with its (odd) table stride — the lanes of a quarter warp
                    // read the pairs of eight consecutive keys with one LDS.128 without bank conflicts, and what is contiguous in
                    // the table is contiguous in the ring: the bin's segments, one bulk copy each per grid step
                    std::sort(keyOrder.begin(), keyOrder.end(), [&](int a, int b) {
                        return out.keyPairBase[(size_t)E.keyOff + a] < out.keyPairBase[(size_t)E.keyOff + b];
                    });
                    out.binSegOff[(size_t)bI] = (int)(out.kSegs.size() / 2);
                    int segSrc = 0, segDst = 0, segLen = 1, nSeg = 0, bytes = 0;      // row 0 = pair 0 of the step row
                    auto close = [&]() {
                        out.kSegs.push_back((uint32_t)segSrc);
                        out.kSegs.push_back((uint32_t)segDst | ((uint32_t)segLen << 16));
                        ++nSeg;
                        bytes += segLen * 16;
                    };
                    for (int key : keyOrder) {
                        // (with a leading pair of x = 0 when the plan has one: the key's first row then precedes startOf)
                        startOf[key] = row + out.zeroRow;
                        const int n = rowsOf[key], st = rc_pair_stride(n + out.zeroRow), src = out.keyPairBase[(size_t)E.keyOff + key] - out.zeroRow;
                        for (int j = 0; j < st; ++j) fRows.push_back(src + j);
                        if (src == segSrc + segLen) segLen += st;
                        else { close(); segSrc = src; segDst = row; segLen = st; }
                        row += st;
                    }
                    close();
                    out.binSegCnt[(size_t)bI] = nSeg;
                    out.binSegBytes[(size_t)bI] = bytes;
                } else {
                    for (int key : keyOrder) {
                        startOf[key] = row;
                        int n = rowsOf[key];
                        for (int j = 0; j < n; ++j) fRows.push_back(out.keyPairBase[(size_t)E.keyOff + key] + j);
                        row += n;
                    }
                }
                for (int q = patOff[(size_t)bI]; q < patOff[(size_t)bI + 1]; ++q) {
                    int i = pats[(size_t)q];
                    for (int k = 0; k < P; ++k) {
                        if (!pps[(size_t)i].mx[(size_t)e * P + k]) continue;
                        int key = out.keyId[((size_t)e * nPat + i) * P + k];
                        rBase[rBaseOff + (size_t)q * P + k] = (uint16_t)startOf[key];
                    }
                }
            }
        };
        // (1) bins spanning all epochs (self-contained fast kernel)
        // (one sequential first-fit pass over all patterns; nothing below reads its result before the summary, so it runs
        // beside the per-epoch phases)
        out.fBinPatOff.assign(1, 0);
        out.fBinPats.clear();
        struct Joiner {
            std::thread t;
            ~Joiner() { if (t.joinable()) t.join(); }
        } allEpochBins;
        auto pack_all = [&]() {
            pack_range(0, nEpochs - 1, maxPatFast, 0, false, order, out.fBinPatOff, out.fBinPats);
            out.nFastBins = (int)out.fBinPatOff.size() - 1;
        };
        if (nThreads > 1) allEpochBins.t = std::thread(pack_all);
        else pack_all();
        lap("all-epoch bins");
        // (2) per-epoch bins (keyed kernel).  kBinPatOff holds, epoch after epoch, nKBins[e]+1 offsets into kBinPats.
        // Two bin groups per epoch (index 2e + g): g = 0 the patterns small enough for one thread each (pattern mode), g = 1
        // the rest in row mode (all boxes with rows long enough) or state mode.
        out.kBinBase.assign((size_t)nEpochs * RC_NGROUP, 0);
        out.nKBins.assign((size_t)nEpochs * RC_NGROUP, 0);
        out.epochMode.assign((size_t)nEpochs * RC_NGROUP, 0);
        out.rowOff.assign((size_t)nEpochs * (nPat + 1), 0);
        out.rowEpochBase.assign((size_t)nEpochs, 0);
        out.patRow.assign((size_t)nEpochs * nPat, 0u);
        out.patRowStride.assign((size_t)nEpochs * nPat, 0);
        const char* rowEnv = std::getenv("RC_ROWMODE");
        const bool patEnv = patternMode != 0 && (std::getenv("RC_PATMODE") ? std::atoi(std::getenv("RC_PATMODE")) != 0 : true);
        static const bool patLEnv = std::getenv("RC_PATLARGE") ? std::atoi(std::getenv("RC_PATLARGE")) != 0 : true;
        std::vector<std::vector<int>> groupIds((size_t)nEpochs * RC_NGROUP);
        static const double minLen = std::getenv("RC_ROWMODE_MIN") ? std::atof(std::getenv("RC_ROWMODE_MIN")) : 2.0;
        // one task per epoch; the row arrays of an epoch are collected locally and put one behind the other afterwards
        std::vector<std::vector<uint16_t>> eRef((size_t)nEpochs);
        std::vector<std::vector<uint32_t>> eWords((size_t)nEpochs);
        std::vector<std::vector<uint8_t>> eBr((size_t)nEpochs), ePm((size_t)nEpochs);
        auto do_epoch = [&](int e) {
            bool anyPat = false, use = false;
            int maxNO = 0;
            for (int i : orderK) {
                const PatPlan& pp = pps[(size_t)i];
                // a box of all live branches, or (not in the last epoch, whose shortcut and epilogue assume the former) a live set
                // that runs on its bounding box (pat_box)
                const bool full = pp.rowOK[(size_t)e] && pp.fullBox[(size_t)e];
                const bool box = patEnv && (full || (out.zeroRow && e + 1 < nEpochs && pp.live[(size_t)e] > 0));
                int g = RC_NGROUP - 1;                               // one state per thread
                if (box)
                    for (int q = 0; q < RC_NPATG; ++q) {
                        if (q > 0 && !patLEnv && rc_pat_cap(RC_MODE_PAT + q) > RC_PAT_MAXS) break;
                        PatBox pb;
                        if (pat_box(&pp.mx[(size_t)e * P], &pp.mn[(size_t)e * P], full, pp.rowK[(size_t)e], P, rc_pat_cap(RC_MODE_PAT + q), pb)) { g = q; break; }
                    }
                if (g == RC_NGROUP - 1) {
                    // row mode: a big pattern always (that is why it is here); a small one when its rows are long enough to pay
                    // for themselves and it fits one row bin
                    int rr = 0;
                    for (int k = 0; k < P; ++k) rr += pp.mx[(size_t)e * P + k];
                    const bool fits = pp.rowOK[(size_t)e] && pp.rowCnt[(size_t)e] > 0 && pp.rowCnt[(size_t)e] <= fastBlock &&
                                      pp.rowPad[(size_t)e] <= RC_ROW_BCAP && rr <= rowCap;
                    bool rowsWanted = fits && (double)pp.live[(size_t)e] / (double)pp.rowCnt[(size_t)e] >= minLen;
                    if (rowEnv) rowsWanted = fits && std::atoi(rowEnv) != 0;
                    if (isKeyedBig[(size_t)i] || rowsWanted) {
                        g = RC_NPATG;
                        use = true;
                        maxNO = std::max(maxNO, (int)pp.rowNO[(size_t)e]);
                    }
                }
                groupIds[(size_t)e * RC_NGROUP + g].push_back(i);
                anyPat = anyPat || g < RC_NPATG;
            }
            for (int q = 0; q < RC_NPATG; ++q) out.epochMode[(size_t)e * RC_NGROUP + q] = RC_MODE_PAT + q;
            out.epochMode[(size_t)e * RC_NGROUP + RC_NPATG] = maxNO <= 3 ? 1 : 2;     // row kernel for <= 3 / <= RC_ROW_NO other branches
            out.epochMode[(size_t)e * RC_NGROUP + RC_NGROUP - 1] = 0;
            if (use || anyPat) {
                int* ro = &out.rowOff[(size_t)e * (nPat + 1)];
                long long run = 0;
                for (int i = 0; i < nPat; ++i) {
                    const PatPlan& pp = pps[(size_t)i];
                    ro[i] = (int)run;
                    size_t r0 = 0;
                    for (int q = 0; q < e; ++q) r0 += (size_t)pp.rowCnt[(size_t)q];
                    const int n = pp.rowOK[(size_t)e] ? pp.rowCnt[(size_t)e] : 0;
                    if (n > 0) {
                        // (a pattern's rows of an epoch are contiguous in its own arrays)
                        eRef[(size_t)e].insert(eRef[(size_t)e].end(), pp.rowRefs.begin() + (long long)(r0 * RC_ROW_MAXM), pp.rowRefs.begin() + (long long)((r0 + (size_t)n) * RC_ROW_MAXM));
                        eWords[(size_t)e].insert(eWords[(size_t)e].end(), pp.rowWords.begin() + (long long)(r0 * RC_ROW_NO), pp.rowWords.begin() + (long long)((r0 + (size_t)n) * RC_ROW_NO));
                        eBr[(size_t)e].insert(eBr[(size_t)e].end(), pp.rowBr.begin() + (long long)r0, pp.rowBr.begin() + (long long)(r0 + (size_t)n));
                        ePm[(size_t)e].insert(ePm[(size_t)e].end(), pp.rowPm.begin() + (long long)r0, pp.rowPm.begin() + (long long)(r0 + (size_t)n));
                    }
                    run += n;
                    out.patRow[(size_t)e * nPat + i] = (uint32_t)pp.rowM[(size_t)e] | ((uint32_t)pp.rowK[(size_t)e] << 8) |
                                                      ((uint32_t)pp.rowNO[(size_t)e] << 16);
                    out.patRowStride[(size_t)e * nPat + i] = pp.rowStride[(size_t)e];
                }
                ro[nPat] = (int)run;
            }
        };
        {
            std::atomic<int> nextE(0);
            auto worker = [&]() { for (;;) { const int e = nextE.fetch_add(1); if (e >= nEpochs) break; do_epoch(e); } };
            const int nT = nPat < 2048 ? 1 : std::min(nThreads, nEpochs);
            if (nT <= 1) worker();
            else {
                std::vector<std::thread> pool;
                for (int t = 0; t < nT; ++t) pool.emplace_back(worker);
                for (auto& th : pool) th.join();
            }
            for (int e = 0; e < nEpochs; ++e) {
                out.rowEpochBase[(size_t)e] = (long long)out.rowRefAll.size() / RC_ROW_MAXM;
                out.rowRefAll.insert(out.rowRefAll.end(), eRef[(size_t)e].begin(), eRef[(size_t)e].end());
                out.rowWordsAll.insert(out.rowWordsAll.end(), eWords[(size_t)e].begin(), eWords[(size_t)e].end());
                out.rowBrAll.insert(out.rowBrAll.end(), eBr[(size_t)e].begin(), eBr[(size_t)e].end());
                out.rowPmAll.insert(out.rowPmAll.end(), ePm[(size_t)e].begin(), ePm[(size_t)e].end());
            }
        }
        lap("groups and row arrays");
        // ---- executed FP64 instructions per grid step (rc_kernels.cu, the update of each tier as written) ----
        out.epochFp64.assign((size_t)nEpochs, 0);
        out.epochFp64Short.assign((size_t)nEpochs, 0);
        std::vector<long long> vFp((size_t)nEpochs * RC_NGROUP, 0), vFpS((size_t)nEpochs * RC_NGROUP, 0);
        auto count_group = [&](int v) {
            const int e = v / RC_NGROUP, mode = out.epochMode[(size_t)v];
            bool direct = mode >= RC_MODE_PAT && !groupIds[(size_t)v].empty();
            for (int i : groupIds[(size_t)v]) direct = direct && pps[(size_t)i].rowNO[(size_t)e] == 0;
            for (int i : groupIds[(size_t)v]) {
                const PatPlan& pp = pps[(size_t)i];
                long long n = 0;
                if (mode >= RC_MODE_PAT) {
                    // rc_pat_run: D1 tree over the m = 1 branches, one DMUL per prefix of the shaped branches, per state one
                    // DMUL / DFMA per existing neighbour and the combining DFMA; updateD for a single-branch pattern
                    PatBox pb;
                    pat_box(&pp.mx[(size_t)e * P], &pp.mn[(size_t)e * P], pp.rowOK[(size_t)e] && pp.fullBox[(size_t)e], pp.rowK[(size_t)e], P, rc_pat_cap(mode), pb);
                    const int nS = pb.nS, nU = pb.nU;
                    const uint32_t sh = pb.shape;
                    int M[5];
                    for (int q = 0; q < 5; ++q) M[q] = (int)((sh >> (4 * q)) & 0xFu);
                    long long pre = M[0], states = M[0];
                    // (counted for the patterns that have such branches; a bin that mixes both kinds executes them for all)
                    n += nU ? 6 - nS : 0;
                    n += nU ? pre : 0;                                      // dA = D1 * gA
                    for (int q = 1; q <= nS; ++q) { pre *= M[q]; n += pre; states *= M[q]; }
                    n += states;                                            // b = fma(b, d, t2)
                    for (int q = 0; q <= nS; ++q) n += states / M[q] * (M[q] - 1);   // neighbours along each shaped branch
                    if (pb.unit) n += 1;                                    // updateD
                    if (direct) n = 2LL * M[0] - 1 + 1;
                } else if (mode) {
                    // rc_row_run: per row NO DMUL (D), per state DMUL (D * G_r), DMUL (row neighbour), NO DFMA, DFMA
                    size_t r0 = 0;
                    for (int q = 0; q < e; ++q) r0 += (size_t)pp.rowCnt[(size_t)q];
                    for (int r = 0; r < pp.rowCnt[(size_t)e]; ++r) {
                        long long Mr = 0, NO = 0;
                        while (Mr < RC_ROW_MAXM && pp.rowRefs[(r0 + (size_t)r) * RC_ROW_MAXM + (size_t)Mr] != 0xFFFF) ++Mr;
                        while (NO < RC_ROW_NO && ((pp.rowWords[(r0 + (size_t)r) * RC_ROW_NO + (size_t)NO] >> 8) & 0xFFu) != 0) ++NO;
                        n += NO + Mr * (NO + 2) + (Mr - 1) + (NO == 0 ? 1 : 0);
                    }
                } else {
                    // rc_key_step: per state and non-zero branch one DMUL into the decay factor (the first is a move) and one
                    // DMUL / DFMA of the neighbour, then the combining DFMA
                    size_t m0 = 0;
                    for (int q = 0; q < e; ++q) m0 += (size_t)pp.live[(size_t)q];
                    for (int sI = 0; sI < pp.live[(size_t)e]; ++sI) n += 2LL * (long long)(pp.meta[m0 + (size_t)sI] & 0xFFu);
                }
                vFp[(size_t)v] += n;
                if (pp.shortcutOK) vFpS[(size_t)v] += n;
            }
        };
        {
            const int nV = nEpochs * RC_NGROUP;
            std::atomic<int> nextV(0);
            auto worker = [&]() { for (;;) { const int v = nextV.fetch_add(1); if (v >= nV) break; count_group(v); } };
            const int nT = nPat < 2048 ? 1 : std::min(nThreads, nV);
            if (nT <= 1) worker();
            else {
                std::vector<std::thread> pool;
                for (int t = 0; t < nT; ++t) pool.emplace_back(worker);
                for (auto& th : pool) th.join();
            }
            for (int v = 0; v < nV; ++v) {
                out.epochFp64[(size_t)(v / RC_NGROUP)] += vFp[(size_t)v];
                out.epochFp64Short[(size_t)(v / RC_NGROUP)] += vFpS[(size_t)v];
            }
        }
        lap("instruction counts");
        out.grpDirect.assign((size_t)nEpochs * RC_NGROUP, 0);
        static const bool directEnv = std::getenv("RC_PAT_DIRECT") ? std::atoi(std::getenv("RC_PAT_DIRECT")) != 0 : true;
        {
            // the bin groups are packed independently of each other: one task per (epoch, group) over a few host threads, the
            // results appended in order
            const int nV = nEpochs * RC_NGROUP;
            std::vector<std::vector<int>> offV((size_t)nV), patsV((size_t)nV);
            for (int v = 0; v < nV; ++v) {
                const int e = v / RC_NGROUP, mode = out.epochMode[(size_t)v];
                // a pattern group in which every pattern has one live branch (the root epoch of a tree)
                bool direct = directEnv && mode >= RC_MODE_PAT && !groupIds[(size_t)v].empty();
                for (int i : groupIds[(size_t)v]) direct = direct && pps[(size_t)i].rowNO[(size_t)e] == 0;
                out.grpDirect[(size_t)v] = direct ? 1 : 0;
            }
            std::atomic<int> nextV(0);
            auto packer = [&]() {
                for (;;) {
                    const int v = nextV.fetch_add(1);
                    if (v >= nV) break;
                    if (groupIds[(size_t)v].empty()) continue;
                    const int e = v / RC_NGROUP, mode = out.epochMode[(size_t)v];
                    const bool direct = out.grpDirect[(size_t)v] != 0;
                    if (mode >= 3) pack_range(e, e, rc_pat_lanes(mode), mode, true, groupIds[(size_t)v], offV[(size_t)v], patsV[(size_t)v], direct);
                    else if (mode) pack_range(e, e, RC_NPBR, 1, true, groupIds[(size_t)v], offV[(size_t)v], patsV[(size_t)v], direct);
                    else pack_range(e, e, RC_NPB, 0, true, groupIds[(size_t)v], offV[(size_t)v], patsV[(size_t)v], direct);
                }
            };
            const int nT = nPat < 2048 ? 1 : std::min(nThreads, nV);
            if (nT <= 1) packer();
            else {
                std::vector<std::thread> pool;
                for (int t = 0; t < nT; ++t) pool.emplace_back(packer);
                for (auto& th : pool) th.join();
            }
            for (int v = 0; v < nV; ++v) {
                out.kBinBase[(size_t)v] = (int)out.kBinPatOff.size();
                const int base = (int)out.kBinPats.size();
                out.kBinPatOff.push_back(base);
                for (int o : offV[(size_t)v]) out.kBinPatOff.push_back(base + o);       // cumulative counts after every bin
                out.kBinPats.insert(out.kBinPats.end(), patsV[(size_t)v].begin(), patsV[(size_t)v].end());
                out.nKBins[(size_t)v] = (int)out.kBinPatOff.size() - 1 - out.kBinBase[(size_t)v];
            }
        }
        lap("per-epoch bins");
        out.kRowBase.assign(out.kBinPats.size() * (size_t)P, 0xFFFF);
        out.kSegs.clear();
        out.binSegOff.assign(out.kBinPatOff.size(), 0);
        out.binSegCnt.assign(out.kBinPatOff.size(), 0);
        out.binSegBytes.assign(out.kBinPatOff.size(), 0);
        for (int v = 0; v < nEpochs * RC_NGROUP; ++v) {
            const int e = v / RC_NGROUP, b0 = out.kBinBase[(size_t)v];
            // fetch offsets are stored at the same index as the bin offsets (one extra entry per group)
            std::vector<int> fOff;
            if (out.grpDirect[(size_t)v]) fOff.assign((size_t)out.nKBins[(size_t)v], (int)out.kFetchRows.size());     // nothing to fetch
            else {
                fetch_lists(e, out.kBinPatOff, out.kBinPats, b0, b0 + out.nKBins[(size_t)v], fOff, out.kFetchRows, out.kRowBase, 0,
                            out.epochMode[(size_t)v] >= RC_MODE_PAT);
            }
            fOff.push_back((int)out.kFetchRows.size());
            out.kFetchOff.insert(out.kFetchOff.end(), fOff.begin(), fOff.end());
        }
        lap("fetch lists");
        if (std::getenv("RC_PLAN_DEBUG")) {
            for (int v = 0; v < nEpochs * RC_NGROUP; ++v) {
                const int e = v / RC_NGROUP, m = out.epochMode[(size_t)v];
                const PlanHost::Epoch& E = out.ep[(size_t)e];
                int b0 = out.kBinBase[(size_t)v], nb = out.nKBins[(size_t)v];
                if (!nb) continue;
                int f0 = out.kFetchOff[(size_t)b0], f1 = out.kFetchOff[(size_t)(b0 + nb)];
                std::fprintf(stderr, "[rc plan] epoch %d: keys %d threads %d rowPairs %d mode %s patterns %d bins %d fetch rows/bin %.1f\n", e,
                             E.nKeys, E.nThr, E.rowPairs, m == 3 ? "patterns<12>" : m == 4 ? "patterns<25>" : m == 5 ? "patterns<36>" : m ? "rows" : "states", (int)groupIds[(size_t)v].size(), nb,
                             (double)(f1 - f0) / nb);
            }
            if (allEpochBins.t.joinable()) allEpochBins.t.join();
            std::fprintf(stderr, "[rc plan] all-epoch fast bins %d (block %d); big patterns %d, of which %d in row mode on the key tier\n",
                         out.nFastBins, fastBlock, (int)bigIds.size(), (int)keyedBig.size());
        }
    }
    {
        // big-tier bins (self-contained general kernel): first those of the patterns that the trajectory-key tier runs in row
        // mode — a model on that tier skips them (rc_propagate_kernel) —, then the patterns only this tier can run
        std::vector<int> rest, off2, pats2;
        for (int i : bigIds) if (!isKeyedBig[(size_t)i]) rest.push_back(i);
        out.nBinsKB = keyedBig.empty() ? 0 : pack(keyedBig, cap, RC_NPB, 1 << 30, out.binPatOff, out.binPats);
        if (keyedBig.empty()) { out.binPatOff.assign(1, 0); out.binPats.clear(); }
        const int nRest = rest.empty() ? 0 : pack(rest, cap, RC_NPB, 1 << 30, off2, pats2);
        for (int b = 0; b < nRest; ++b) {
            for (int q = off2[(size_t)b]; q < off2[(size_t)b + 1]; ++q) out.binPats.push_back(pats2[(size_t)q]);
            out.binPatOff.push_back((int)out.binPats.size());
        }
        out.nBins = out.nBinsKB + nRest;
    }
    out.minX.assign((size_t)nEpochs * nPat * P, 0);
    out.patFull.assign((size_t)nEpochs * nPat, 0);
    out.maxX.assign((size_t)nEpochs * nPat * P, 0);
    par_for(nPat, 256, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
            const PatPlan& pp = pps[(size_t)i];
            for (int e = 0; e < nEpochs; ++e) {
                for (int k = 0; k < P; ++k) {
                    out.minX[((size_t)e * nPat + i) * P + k] = pp.mn[(size_t)e * P + k];
                    out.maxX[((size_t)e * nPat + i) * P + k] = pp.mx[(size_t)e * P + k];
                }
                out.patFull[(size_t)e * nPat + i] = pp.rowOK[(size_t)e] && pp.fullBox[(size_t)e];
            }
        }
    });

    // concatenate per-epoch arrays
    out.slotOff.assign((size_t)nEpochs * (nPat + 1), 0);
    out.epochBase.assign((size_t)nEpochs, 0);
    out.epochSlots.assign((size_t)nEpochs, 0);
    out.epochSlotsShort.assign((size_t)nEpochs, 0);
    out.shortcutOK.resize((size_t)nPat);
    // (live counts gathered per pattern side by side, then one prefix sum per epoch)
    std::vector<int> liveAll((size_t)nEpochs * nPat, 0);
    par_for(nPat, 256, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i)
            for (int e = 0; e < nEpochs; ++e) liveAll[(size_t)e * nPat + i] = pps[(size_t)i].live[(size_t)e];
    });
    long long total = 0;
    for (int e = 0; e < nEpochs; ++e) {
        out.epochBase[(size_t)e] = total;
        int* so = &out.slotOff[(size_t)e * (nPat + 1)];
        const int* lv = &liveAll[(size_t)e * nPat];
        long long tot = 0, totS = 0;
        for (int i = 0; i < nPat; ++i) {
            so[i] = (int)tot;
            tot += lv[i];
            if (pps[(size_t)i].shortcutOK) totS += lv[i];
        }
        if (tot > 0x7FFFFFFFLL) return "too many live states in one epoch";
        so[nPat] = (int)tot;
        out.epochSlots[(size_t)e] = tot;
        out.epochSlotsShort[(size_t)e] = totS;
        total += tot;
    }
    out.totalSlots = total;
    for (int e = 0; e < nEpochs; ++e) out.maxEpochSlots = std::max(out.maxEpochSlots, out.epochSlots[(size_t)e]);
    out.words.assign((size_t)total * nnzw, 0u);
    out.meta.assign((size_t)total, 0u);
    par_for(nPat, 256, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
            const PatPlan& pp = pps[(size_t)i];
            out.shortcutOK[(size_t)i] = pp.shortcutOK ? 1 : 0;
            size_t wp = 0, mp = 0;
            for (int e = 0; e < nEpochs; ++e) {
                long long g0 = out.epochBase[(size_t)e] + out.slotOff[(size_t)e * (nPat + 1) + i];
                int n = pp.live[(size_t)e];
                for (int s = 0; s < n; ++s) {
                    for (int d = 0; d < nnzw; ++d) out.words[(size_t)(g0 + s) * nnzw + d] = pp.words[wp + (size_t)s * c.stride + d];
                    out.meta[(size_t)(g0 + s)] = pp.meta[mp + s];
                }
                wp += (size_t)n * c.stride;
                mp += (size_t)n;
            }
        }
    });
    // transitions
    out.trBase.assign((size_t)nEpochs, 0);
    out.trEntBase.assign((size_t)nEpochs, 0);
    // pass 1 (patterns side by side): entries of every pattern in every epoch; pass 2 after a prefix sum per epoch
    const int nTr = nEpochs > 0 ? nEpochs - 1 : 0;
    std::vector<long long> patEnt((size_t)nTr * ((size_t)nPat + 1), 0);
    par_for(nPat, 256, [&](int i0, int i1) { for (int i = i0; i < i1; ++i) {
        const PatPlan& pp = pps[(size_t)i];
        size_t cp = 0;
        for (int e = 0; e < nTr; ++e) {
            int n = pp.live[(size_t)e + 1];
            long long c = 0;
            for (int s = 0; s < n; ++s) c += pp.trCnt[cp + s];
            cp += (size_t)n;
            patEnt[(size_t)e * (nPat + 1) + i + 1] = c;
        }
    } });
    long long tb = 0, te = 0;
    for (int e = 0; e < nTr; ++e) {
        long long* row = &patEnt[(size_t)e * (nPat + 1)];
        for (int i = 0; i < nPat; ++i) row[i + 1] += row[i];
        if (row[nPat] > 0x7FFFFFFFLL) return "too many transition entries in one epoch";
        out.trBase[(size_t)e] = tb;
        out.trEntBase[(size_t)e] = te;
        tb += out.epochSlots[(size_t)e + 1] + 1;
        te += row[nPat];
    }
    out.trOff.assign((size_t)tb, 0);
    out.trEnt.resize((size_t)te);
    for (int e = 0; e < nTr; ++e)
        out.trOff[(size_t)(out.trBase[(size_t)e] + out.epochSlots[(size_t)e + 1])] = (int)patEnt[(size_t)e * (nPat + 1) + nPat];
    par_for(nPat, 256, [&](int i0, int i1) { for (int i = i0; i < i1; ++i) {
        const PatPlan& pp = pps[(size_t)i];
        size_t cp = 0, ep = 0;
        for (int e = 0; e < nTr; ++e) {
            int n = pp.live[(size_t)e + 1];
            long long s0 = out.trBase[(size_t)e] + out.slotOff[(size_t)(e + 1) * (nPat + 1) + i];
            long long run = patEnt[(size_t)e * (nPat + 1) + i];
            long long dst = out.trEntBase[(size_t)e] + run;
            for (int s = 0; s < n; ++s) {
                out.trOff[(size_t)(s0 + s)] = (int)run;
                int cn = pp.trCnt[cp + s];
                for (int q = 0; q < cn; ++q) out.trEnt[(size_t)dst++] = pp.trEnt[ep++];
                run += cn;
            }
            cp += (size_t)n;
        }
    } });
    lap("concatenation");
    // the per-pattern plans (several hundred thousand small vectors) are released while the bin images are built
    std::thread reaper([dead = std::move(pps)]() mutable { std::vector<PatPlan>().swap(dead); });
    std::string r = build_bin_images(out, sev);
    reaper.join();
    lap("bin images");
    return r;
}


// Structural self-check of a plan: every index a kernel dereferences is inside its array and every bin respects the
// shared-memory limits of the kernel that runs it.  Returns "" or a description of the first violation.  (The pool's
// compute-sanitizer is closed, so the bounds are verified on the host — tests/test_host_cpu.py runs this on every
// workload shape.)
// FNV-1a over everything a kernel reads from the plan (tests: the plan is a function of its inputs, not of the host threads
// that built it or the order in which they finished).
uint64_t plan_checksum(const PlanHost& h) {
    uint64_t x = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { x ^= b[i]; x *= 1099511628211ull; }
    };
    auto vec = [&](const auto& v) {
        const uint64_t n = v.size();
        mix(&n, sizeof n);
        if (!v.empty()) mix(v.data(), v.size() * sizeof(v[0]));
    };
    for (const auto& e : h.ep) {
        const int f[6] = {e.nKeys, e.nThr, e.rowPairs, e.keyOff, e.thrOff, e.elBase};
        mix(f, sizeof f);
    }
    vec(h.keyBranch); vec(h.keyMaxJ); vec(h.keyOp); vec(h.keyPa); vec(h.keyPb); vec(h.keyInit); vec(h.keyPairBase);
    vec(h.keyThrBase); vec(h.thrKey); vec(h.lastKeyId); vec(h.kBinBase); vec(h.nKBins); vec(h.kBinPatOff); vec(h.kBinPats);
    vec(h.kFetchOff); vec(h.kFetchRows); vec(h.kRowBase); vec(h.epochMode); vec(h.rowOff); vec(h.rowEpochBase);
    vec(h.rowRefAll); vec(h.rowWordsAll); vec(h.patRow); vec(h.kBinHdr); vec(h.kPatRec); vec(h.grpDirect); vec(h.grpMaxFetch);
    vec(h.grpMaxB); vec(h.kLane); vec(h.kElem); vec(h.kSegs); vec(h.lastPairBase); vec(h.grpDirIn); vec(h.slotOff);
    vec(h.epochBase); vec(h.words); vec(h.meta); vec(h.trBase); vec(h.trEntBase); vec(h.trOff); vec(h.trEnt); vec(h.shortcutOK);
    vec(h.fBinPatOff); vec(h.fBinPats); vec(h.binPatOff); vec(h.binPats); vec(h.epochFp64); vec(h.epochFp64Short);
    vec(h.maxX); vec(h.minX); vec(h.patFull); vec(h.epochSlots); vec(h.epochSlotsShort); vec(h.binSegOff); vec(h.binSegCnt);
    vec(h.binSegBytes); vec(h.patRowStride);
    const long long g[4] = {h.maxEpochSlots, h.dirOff, h.dirSize, h.totalSlots};
    mix(g, sizeof g);
    return x;
}

std::string check_plan(const PlanHost& h) {
    const int P = h.P, nPat = h.nPat, nE = h.nEpochs;
    auto fail = [](const std::string& m) { return std::string("plan check: ") + m; };
    if ((int)h.slotOff.size() != nE * (nPat + 1)) return fail("slotOff size");
    for (int e = 0; e < nE; ++e) {
        const int* so = &h.slotOff[(size_t)e * (nPat + 1)];
        for (int i = 0; i < nPat; ++i) {
            const int live = so[i + 1] - so[i];
            if (live < 1) return fail("empty live set");
            const long long g0 = h.epochBase[(size_t)e] + so[i];
            for (int s = 0; s < live; ++s) {
                const uint32_t mt = h.meta[(size_t)(g0 + s)];
                const int nnz = (int)(mt & 0xFFu);
                if (nnz < 1 || nnz > h.nnzw) return fail("nnz out of range");
                for (int d = 0; d < nnz; ++d) {
                    const uint32_t w = h.words[(size_t)(g0 + s) * h.nnzw + d];
                    const int k = (int)(w & 0xFFu), x = (int)((w >> 8) & 0xFFu);
                    const uint32_t up = w >> 16;
                    if (k >= P || x < 1 || x > h.maxAf) return fail("word k/x out of range");
                    if (x > h.maxX[((size_t)e * nPat + i) * P + k]) return fail("x above maxX");
                    if (up != RC_NO_UP && (int)up >= live) return fail("neighbour slot out of range");
                }
            }
            if (e > 0) {
                const int livePrev = h.slotOff[(size_t)(e - 1) * (nPat + 1) + i + 1] - h.slotOff[(size_t)(e - 1) * (nPat + 1) + i];
                const int* trOff = &h.trOff[(size_t)h.trBase[(size_t)e - 1]];
                for (int s = 0; s < live; ++s) {
                    const int a = trOff[so[i] + s], b = trOff[so[i] + s + 1];
                    if (a > b) return fail("transition offsets not monotone");
                    for (int q = a; q < b; ++q) {
                        const uint32_t en = h.trEnt[(size_t)(h.trEntBase[(size_t)e - 1] + q)];
                        if ((int)(en & 0xFFFFu) >= livePrev) return fail("transition source out of range");
                        const uint32_t widx = en >> 16;
                        if (widx != RC_W_ONE && (int)widx >= (h.maxAf + 1) * (h.maxAf + 1)) return fail("weight index out of range");
                    }
                }
            }
        }
    }
    // keyed per-epoch bins
    if (!h.kBinBase.empty()) {
        std::vector<char> seen;
        for (int v = 0; v < nE * RC_NGROUP; ++v) {
            const int e = v / RC_NGROUP;
            const PlanHost::Epoch& E = h.ep[(size_t)e];
            const bool rows = h.epochMode[(size_t)v] != 0, patM = h.epochMode[(size_t)v] >= 3;
            if (v % RC_NGROUP == 0) seen.assign((size_t)nPat, 0);
            const int b0 = h.kBinBase[(size_t)v];
            for (int bI = b0; bI < b0 + h.nKBins[(size_t)v]; ++bI) {
                const int p0 = h.kBinPatOff[(size_t)bI], p1 = h.kBinPatOff[(size_t)bI + 1];
                if (p1 - p0 < 1 || p1 - p0 > (patM ? rc_pat_lanes(h.epochMode[(size_t)v]) : rows ? RC_NPBR : RC_NPB)) return fail("patterns per keyed bin");
                const int f0 = h.kFetchOff[(size_t)bI], f1 = h.kFetchOff[(size_t)bI + 1];
                const bool direct = !h.grpDirect.empty() && h.grpDirect[(size_t)v] != 0;
                if (direct ? f1 != f0 : (f1 - f0 < 1 || f1 - f0 > RC_KROWS - 1)) return fail("pair rows per keyed bin");
                for (int q = f0; q < f1; ++q)
                    if (h.kFetchRows[(size_t)q] < 0 || h.kFetchRows[(size_t)q] >= E.rowPairs) return fail("fetch row out of range");
                long long lanes = 0, padded = 0;
                for (int q = p0; q < p1; ++q) {
                    const int i = h.kBinPats[(size_t)q];
                    if (i < 0 || i >= nPat || seen[(size_t)i]) return fail("pattern missing or repeated in keyed bins");
                    seen[(size_t)i] = 1;
                    const int live = h.slotOff[(size_t)e * (nPat + 1) + i + 1] - h.slotOff[(size_t)e * (nPat + 1) + i];
                    for (int k = 0; k < P; ++k) {
                        const int mxk = h.maxX[((size_t)e * nPat + i) * P + k];
                        if (!mxk) continue;
                        const int rb = h.kRowBase[(size_t)q * P + k];
                        if (!direct && (rb < 1 || rb + mxk - 1 > f1 - f0 - 1)) return fail("pair-row base out of range");
                    }
                    if (rows) {
                        const uint32_t pr = h.patRow[(size_t)e * nPat + i];
                        const int M = (int)(pr & 0xFFu), rk = (int)((pr >> 8) & 0xFFu), NO = (int)((pr >> 16) & 0xFFu);
                        const int nr = h.rowOff[(size_t)e * (nPat + 1) + i + 1] - h.rowOff[(size_t)e * (nPat + 1) + i];
                        if (M < 1 || M > RC_ROW_MAXM || NO > RC_ROW_NO || rk >= P) return fail("row shape");
                        lanes += patM ? 1 : nr;
                        for (int r = 0; r < nr && !patM; ++r)
                            padded += h.rowPmAll[(size_t)(h.rowEpochBase[(size_t)e] + h.rowOff[(size_t)e * (nPat + 1) + i] + r)] | 1;
                        long long covered = 0;
                        for (int r = 0; r < nr; ++r) {
                            const long long gr = h.rowEpochBase[(size_t)e] + h.rowOff[(size_t)e * (nPat + 1) + i] + r;
                            int Mr = 0, NOr = 0;         // length and other branches of this row (<= the pattern's maxima)
                            while (Mr < RC_ROW_MAXM && h.rowRefAll[(size_t)gr * RC_ROW_MAXM + Mr] != 0xFFFF) ++Mr;
                            while (NOr < RC_ROW_NO && ((h.rowWordsAll[(size_t)gr * RC_ROW_NO + NOr] >> 8) & 0xFFu) != 0) ++NOr;
                            if (Mr < 1 || Mr > M || NOr > NO || (int)h.rowBrAll[(size_t)gr] >= P) return fail("row shape (row)");
                            covered += Mr;
                            for (int j = 0; j < Mr; ++j)
                                if ((int)h.rowRefAll[(size_t)gr * RC_ROW_MAXM + j] >= live) return fail("row state index out of range");
                            for (int d = 0; d < NOr; ++d) {
                                const uint32_t w = h.rowWordsAll[(size_t)gr * RC_ROW_NO + d];
                                const uint32_t up = w >> 16;
                                if ((int)(w & 0xFFu) >= P || (up != RC_NO_UP && (int)up >= nr)) return fail("row word out of range");
                            }
                        }
                        if (covered != live) return fail("rows do not cover the live set");
                    } else lanes += live;
                }
                if (lanes > RC_BLOCK) return fail("lanes per keyed bin");
                if (padded > RC_ROW_BCAP) return fail("padded states per row bin");
            }
        }
    }
    // bin images
    if (!h.kBinBase.empty()) {
        if (h.kBinHdr.size() != h.kBinPatOff.size() * (size_t)RC_HDR_N || h.kPatRec.size() != h.kBinPats.size() * 2)
            return fail("bin image sizes");
        const long long nLaneQ = (long long)(h.kLane.size() / 4), nElemRec = (long long)(h.kElem.size() / 2);
        auto gather_ok = [&](int e, uint32_t r0, uint32_t r1) {
            if (e == 0) return (r0 == 0u || r0 == RC_IMG_SEED) && r1 == 0u;
            const long long cnt = r1 >> RC_IMG_CNT_SHIFT, prev = r1 & ((1u << RC_IMG_CNT_SHIFT) - 1u);
            if (cnt == 0) return true;                       // no sources: the value is 0 (padding of a row, or a dead state)
            const long long lo = h.trEntBase[(size_t)e - 1];
            const long long hi = e < nE - 1 ? h.trEntBase[(size_t)e] : (long long)h.trEnt.size();
            return (long long)r0 >= lo && (long long)r0 + cnt <= hi && prev < h.epochSlots[(size_t)e - 1];
        };
        for (int v = 0; v < nE * RC_NGROUP; ++v) {
            const int e = v / RC_NGROUP;
            const bool rows = h.epochMode[(size_t)v] != 0, patM = h.epochMode[(size_t)v] >= 3;
            const int patMax = rc_pat_cap(h.epochMode[(size_t)v]), patMin = rc_pat_floor(h.epochMode[(size_t)v]);
            const bool directG = !h.grpDirect.empty() && h.grpDirect[(size_t)v] != 0;
            const int b0 = h.kBinBase[(size_t)v];
            for (int bI = b0; bI < b0 + h.nKBins[(size_t)v]; ++bI) {
                const int* hdr = &h.kBinHdr[(size_t)bI * RC_HDR_N];
                const int np = hdr[1], lanes = hdr[2], nB = RC_HDR_NB(hdr[3]), nFetch = hdr[5];
                if ((hdr[3] & RC_HDR_SEED) && (!patM || e != 0)) return fail("image header seed flag");
                const bool dirIn = (hdr[3] & RC_HDR_DIRIN) != 0;
                if (dirIn && (!patM || e == 0 || h.grpDirIn.empty() || !h.grpDirIn[(size_t)v] || hdr[12] < 0 ||
                              (long long)hdr[12] + (long long)RC_DIR_LANES * (np ? nB / np : 0) > h.dirSize))
                    return fail("image header direct hand-over");
                if (hdr[0] != h.kBinPatOff[(size_t)bI] || np != h.kBinPatOff[(size_t)bI + 1] - hdr[0]) return fail("image header patterns");
                if (hdr[4] != h.kFetchOff[(size_t)bI] || nFetch != h.kFetchOff[(size_t)bI + 1] - hdr[4]) return fail("image header fetch");
                if (patM && !directG) {
                    // the segments cover the ring rows of the bin exactly once and agree with the fetch list row by row
                    if (hdr[8] < 0 || hdr[9] < 1 || (size_t)(hdr[8] + hdr[9]) * 2 > h.kSegs.size()) return fail("image segments range");
                    std::vector<char> cov((size_t)nFetch, 0);
                    int bytes = 0;
                    for (int q = hdr[8]; q < hdr[8] + hdr[9]; ++q) {
                        const int src = (int)h.kSegs[(size_t)q * 2], dst = (int)(h.kSegs[(size_t)q * 2 + 1] & 0xFFFFu), len = (int)(h.kSegs[(size_t)q * 2 + 1] >> 16);
                        if (len < 1 || dst + len > nFetch || src < 0 || src + len > h.ep[(size_t)e].rowPairs) return fail("image segment bounds");
                        for (int j = 0; j < len; ++j) {
                            if (cov[(size_t)(dst + j)] || h.kFetchRows[(size_t)(hdr[4] + dst + j)] != src + j) return fail("image segment rows");
                            cov[(size_t)(dst + j)] = 1;
                        }
                        bytes += len * 16;
                    }
                    for (int r = 0; r < nFetch; ++r) if (!cov[(size_t)r]) return fail("image segment cover");
                    if (bytes != hdr[10]) return fail("image segment bytes");
                }
                if (lanes < 1 || lanes > (patM ? rc_pat_lanes(h.epochMode[(size_t)v]) : RC_BLOCK) || nB < 0 || nB > (patM ? rc_pat_lanes(h.epochMode[(size_t)v]) * patMax : RC_ROW_BCAP))
                    return fail("image header lanes");
                if (hdr[6] < 0 || hdr[6] + (long long)lanes * (patM ? RC_PLANE_N / 4 : rows ? RC_RLANE_N / 4 : RC_LANE_N / 4) > nLaneQ || hdr[7] < 0 || hdr[7] + nB > nElemRec)
                    return fail("image record range");
                for (int q = 0; q < np; ++q) {
                    const int rec = h.kPatRec[(size_t)(hdr[0] + q) * 2 + 1];
                    if (h.kPatRec[(size_t)(hdr[0] + q) * 2] != h.kBinPats[(size_t)(hdr[0] + q)]) return fail("image pattern id");
                    if ((rec & 0xFFFF) >= lanes || (rec >> 16) > nB) return fail("image pattern offsets");
                }
                for (int t = 0; t < lanes; ++t) {
                    const uint32_t* r = &h.kLane[(size_t)hdr[6] * 4 + (size_t)t * (patM ? RC_PLANE_N : rows ? RC_RLANE_N : RC_LANE_N)];
                    if (patM) {
                        int M[5];
                        for (int q = 0; q < 5; ++q) M[q] = (int)((r[2] >> (4 * q)) & 0xFu);
                        const int nU = (int)((r[2] >> 20) & 0xFu), eOff = (int)r[5];
                        int n = M[0], nS = 0;
                        for (int q = 1; q < 5; ++q) if (M[q]) { n *= M[q]; ++nS; if (M[q] < 2 || M[q] > M[q - 1] || nS != q) return fail("image pattern shape order"); }
                        if (M[0] < 1 || M[0] > RC_ROW_MAXM || n > patMax || 1 + nS + nU > 8) return fail("image pattern shape");
                        if ((uint32_t)hdr[11] != (r[2] & 0xFFFFFu)) return fail("image bin shape");
                        if (n <= patMin) return fail("pattern in the group of a larger kernel instantiation");
                        if (directG && (nS + nU != 0 || (int)r[15] < 1 || (int)r[15] + M[0] > h.ep[(size_t)e].rowPairs)) return fail("image direct pair base");
                        for (int q = 0; q < 8; ++q) {
                            const int row = (int)((r[q >> 2] >> (8 * (q & 3))) & 0xFFu);
                            const int len = q <= nS ? M[q] : 1;
                            if (directG) { if (row != RC_KROWS - 1) return fail("image pattern pair rows (direct)"); }
                            else if (q < 1 + nS + nU ? (row < 1 || row + len - 1 >= nFetch + (h.zeroRow ? 1 : 0)) : row != 0) return fail("image pattern pair rows");
                        }
                        if ((int)r[4] >= nPat || eOff + n > nB) return fail("image pattern hand-over");
                        // n: states of the (bounding) box; live of them carry a state, the others are marked 0xFF
                        const int live = h.slotOff[(size_t)e * (nPat + 1) + r[4] + 1] - h.slotOff[(size_t)e * (nPat + 1) + r[4]];
                        if (live > n || (live != n && (!h.zeroRow || e + 1 >= nE))) return fail("image pattern states");
                        if (r[2] & RC_PLANE_DIROUT) {
                            // the lane applies the next Join itself: next-epoch state index per state, lane-ordered destination
                            const int nOut = (int)((r[2] >> 24) & 0x3Fu);
                            if (e + 1 >= nE || nOut < 1 || nOut > RC_PAT_MAXL || (long long)r[3] + (long long)(nOut - 1) * RC_DIR_LANES >= h.dirSize) return fail("image direct hand-over");
                            if (nOut < h.slotOff[(size_t)(e + 1) * (nPat + 1) + r[4] + 1] - h.slotOff[(size_t)(e + 1) * (nPat + 1) + r[4]]) return fail("image direct hand-over states");
                            int mapped = 0;
                            for (int q = 0; q < n; ++q) {
                                const int y = (int)((r[6 + (q >> 2)] >> (8 * (q & 3))) & 0xFFu);
                                if (y == 0xFF) continue;
                                if (y >= nOut) return fail("image direct hand-over map");
                                ++mapped;
                            }
                            if (mapped != live) return fail("image direct hand-over map size");
                        } else {
                            if ((long long)r[3] + live > h.epochSlots[(size_t)e] || (int)r[3] != h.slotOff[(size_t)e * (nPat + 1) + r[4]]) return fail("image pattern states");
                            std::vector<char> hit((size_t)n, 0);
                            int found = 0;
                            for (int q = 0; q < n; ++q) {
                                const int ref = (int)((r[6 + (q >> 2)] >> (8 * (q & 3))) & 0xFFu);
                                if (ref == 0xFF && live != n) continue;
                                if (ref >= live || hit[(size_t)ref]) return fail("image pattern state order");
                                hit[(size_t)ref] = 1;
                                ++found;
                            }
                            if (found != live) return fail("image pattern state cover");
                        }
                    } else if (!rows) {
                        for (int d = 0; d < 8; ++d) {
                            const int row = (int)((r[d >> 2] >> (8 * (d & 3))) & 0xFFu), up = (int)((r[2 + (d >> 2)] >> (8 * (d & 3))) & 0xFFu);
                            if (row != RC_KROWS - 1 && (row < 1 || row >= nFetch)) return fail("image pair row");
                            if (up >= lanes) return fail("image neighbour lane");
                        }
                        if ((long long)r[5] >= h.epochSlots[(size_t)e]) return fail("image carry slot");
                        if (!gather_ok(e, r[6], r[7])) return fail("image gather record");
                    } else {
                        const int M = (int)(r[2] & 0xFFu), NO = (int)((r[2] >> 8) & 0xFFu);
                        if (M < 1 || M > RC_ROW_MAXM || NO > RC_ROW_NO || (NO > 3 && NO + M > RC_ROW_SUM)) return fail("image row shape");
                        const int rr = (int)(r[1] >> 24);
                        if (rr < 1 || rr + M - 1 >= nFetch) return fail("image row-branch pair rows");
                        for (int d = 0; d < RC_ROW_NO; ++d) {
                            const int row = (int)((r[d >> 2] >> (8 * (d & 3))) & 0xFFu);
                            if (row != RC_KROWS - 1 && (row < 1 || row >= nFetch)) return fail("image pair row");
                            if (d >= NO && row != RC_KROWS - 1) return fail("image unused pair row");
                        }
                        for (int q = 0; q < RC_ROW_NO + 1; ++q) {
                            const int off = (int)((r[3 + (q >> 1)] >> (16 * (q & 1))) & 0xFFFFu);
                            if (q == 0 ? off + M > nB : (off != RC_ROW_BCAP && off + M > nB)) return fail("image row offsets");
                        }
                        if ((long long)r[7] >= h.epochSlots[(size_t)e] || (long long)(r[8] + 1) * RC_ROW_MAXM > (long long)h.rowRefAll.size() ||
                            (int)r[9] >= nPat)
                            return fail("image row hand-over");
                    }
                }
                for (int q = 0; q < nB; ++q)
                    if (!gather_ok(e, h.kElem[(size_t)(hdr[7] + q) * 2], h.kElem[(size_t)(hdr[7] + q) * 2 + 1])) return fail("image element record");
            }
        }
    }
    // all-epoch fast bins and big bins
    for (int bI = 0; bI < h.nFastBins; ++bI) {
        const int p0 = h.fBinPatOff[(size_t)bI], p1 = h.fBinPatOff[(size_t)bI + 1];
        if (p1 - p0 < 1 || (p1 - p0) * P > h.fastBlock || p1 - p0 > RC_NPB) return fail("patterns per fast bin");
        for (int e = 0; e < nE; ++e) {
            long long slots = 0, rr = 0;
            for (int q = p0; q < p1; ++q) {
                const int i = h.fBinPats[(size_t)q];
                slots += h.slotOff[(size_t)e * (nPat + 1) + i + 1] - h.slotOff[(size_t)e * (nPat + 1) + i];
                for (int k = 0; k < P; ++k) rr += h.maxX[((size_t)e * nPat + i) * P + k];
            }
            if (slots > h.fastBlock || rr > 2 * h.fastBlock - 2) return fail("fast bin over capacity");
        }
    }
    for (int bI = 0; bI < h.nBins; ++bI) {
        const int p0 = h.binPatOff[(size_t)bI], p1 = h.binPatOff[(size_t)bI + 1];
        if (p1 - p0 < 1 || p1 - p0 > RC_NPB) return fail("patterns per big bin");
        for (int e = 0; e < nE; ++e) {
            long long slots = 0;
            for (int q = p0; q < p1; ++q) {
                const int i = h.binPats[(size_t)q];
                slots += h.slotOff[(size_t)e * (nPat + 1) + i + 1] - h.slotOff[(size_t)e * (nPat + 1) + i];
            }
            if (slots > h.cap) return fail("big bin over capacity");
        }
    }
    return "";
}

}  // namespace rc
