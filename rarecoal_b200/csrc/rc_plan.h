// rc_plan.h — host-side planner: enumerations, JointStateSpace mirror, and the compiler that turns
// (patterns, ordered Join/Split structure) into per-epoch live-state lists, neighbour tables and
// Join/Split transition tables for the kernels ("uploaded once per model structure").
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "rc_types.h"

namespace rc {

// ---- enumeration / binomials (Utils.hs:100-127,208-215) -----------------------------------------
std::vector<std::vector<int>> all_configs(int maxAf, const std::vector<int>& nVec);
double choose(int n, int k);

// ---- JointStateSpace (StateSpace.hs:20-28,73-124,145-160) ----------------------------------------
struct StateSpace {
    int P = 0, maxAf = 0, nrStates = 0;
    std::vector<uint8_t> states;                 // [nrStates][P]
    std::unordered_map<uint64_t, int> toId;      // mixed-radix key -> id
    std::vector<int> x1up;                       // [nrStates][P]
    std::vector<int> x1;                         // [P]
    uint64_t key(const int32_t* x) const;
    int state_to_id(const int32_t* x) const;
    std::vector<int> fill_up(const std::vector<int>& ids) const;
};
StateSpace* make_state_space(int P, int maxAf);
bool key_fits(int P, int maxAf);

// ---- structural signature of a model --------------------------------------------------------------
struct StructEv {
    int type;     // RC_EV_JOIN_D / RC_EV_SPLIT_D
    int k, l;
    int mclass;   // split fraction class: 0 -> m == 0, 1 -> m == 1, 2 -> 0 < m < 1 (Join: 2)
};
// Update steps executed before structural event g (after event g-1), capped at maxAf, with the
// frozen mask of each — needed only to track which b(x) are still exactly zero when a Split applies
// its `b > 0` filter (Core.hs:220).
struct GapInfo {
    std::vector<unsigned> masks;
};

struct PlanHost {
    int P = 0, maxAf = 0, nPat = 0, nEpochs = 0, nnzw = 1, nBins = 0, nBinsKB = 0, cap = RC_BLOCK;
    int bigGlobal = 0;                     // the largest live set does not fit the big tier's shared memory: global scratch instead
    std::vector<int> binPatOff, binPats;   // big tier bins
    int nFastBins = 0;
    int fastBlock = RC_BLOCK;              // 64, 128 or 256 state slots per fast-tier block
    std::vector<int> fBinPatOff, fBinPats; // fast tier bins
    std::vector<uint8_t> maxX;             // [nEpochs][nPat][P] largest x_k over the live states
    std::vector<uint8_t> minX;             // [nEpochs][nPat][P] 1 if every live state has x_k >= 1 (host only)
    std::vector<uint8_t> patFull;          // [nEpochs][nPat] the live set is the box of all its live branches (host only)
    int zeroRow = 0;                       // 1: every trajectory key is led by the pair of x = 0, {1, 0} (plans with a Split)
    // Trajectory keys.  The ancestral count a_k(t) of a pattern (updateA, Core.hs:229-240, and the a-part of
    // Join/Split, Core.hs:165-170,195-200) is a function of the model and of the pattern's counts m on the leaves
    // merged into branch k so far.  One key per distinct (epoch, branch, sub-tuple); the trajectory kernel computes
    // each key once and all patterns sharing it read the result.
    struct Epoch {
        int nKeys = 0, nThr = 0, rowPairs = 0, keyOff = 0, thrOff = 0;
        int elBase = 0;                    // first element record (kElem) of the epoch: origin of its lane-ordered b region
    };
    std::vector<Epoch> ep;                 // [nEpochs]
    std::vector<uint8_t> keyBranch, keyMaxJ, keyOp;       // concatenated over epochs (Epoch::keyOff)
    std::vector<int> keyPa, keyPb, keyInit, keyPairBase, keyThrBase;
    std::vector<int> thrKey;               // trajectory-kernel thread -> key (Epoch::thrOff)
    std::vector<int> keyId;                // [nEpochs][nPat][P] (host only)
    std::vector<int> lastKeyId;            // [nPat][P] keys of the final epoch (shortcut reads a from them)
    int totalKeys = 0;
    // per-epoch bins of the keyed kernel
    std::vector<int> kBinBase, nKBins;     // [RC_NGROUP * nEpochs] first entry of the bin group in kBinPatOff / number of bins
    std::vector<int> kBinPatOff, kBinPats; // nKBins[e]+1 offsets per epoch into kBinPats (shard-local pattern ids)
    std::vector<int> kFetchOff;            // same indexing as kBinPatOff -> kFetchRows
    std::vector<int> kFetchRows;           // pair offsets inside a step row, one per shared-memory pair row of a bin
    std::vector<uint16_t> kRowBase;        // [position in kBinPats][P] first shared-memory pair row of (pattern, branch)
    // row mode (per epoch): one thread per row of states along the widest branch of a box-shaped live set
    std::vector<int> epochMode;            // [nEpochs] 0: one state per thread; indexed RC_NGROUP * e + g: 0 one state per thread; one row per thread with <= 3 (1) or <= RC_ROW_NO (2) other branches; one pattern per thread with <= RC_PAT_MAXT (3), <= RC_PAT_MAXS (4) or <= RC_PAT_MAXL (5) states
    std::vector<int> rowOff;               // [nEpochs][nPat+1] rows before a pattern in the epoch's row arrays
    std::vector<long long> rowEpochBase;   // [nEpochs] first row of the epoch in rowRefAll / rowWordsAll
    std::vector<uint16_t> rowRefAll;       // per row RC_ROW_MAXM reference-order indices (inside the pattern) of x_r = 1..M
    std::vector<uint32_t> rowWordsAll;     // per row RC_ROW_NO words: k | x_k << 8 | neighbour row << 16
    std::vector<uint8_t> rowPmAll;         // per row: longest row of its part (the stride its part's rows are stored with)
    std::vector<uint8_t> rowBrAll;         // per row: its row branch (host only; a live set left by a Split has parts with different ones)
    std::vector<uint32_t> patRow;          // [nEpochs][nPat] row length | row branch << 8 | other branches << 16
    std::vector<int> patRowStride;         // [nEpochs][nPat] reference-order stride of the row branch
    // bin images (layouts in rc_types.h): the set-up data of every keyed bin in lane order, so that a block reads
    // header -> lane record -> transition entries instead of chasing pattern / slot / word / offset tables
    std::vector<int> kBinHdr, kPatRec;
    std::vector<int> grpDirect;            // per bin group: 1 if every pattern has a single live branch — such patterns share no
                                           // trajectory keys, so their pairs are read straight from the table (no ring, no row limit)
    std::vector<int> grpMaxFetch, grpMaxB;   // per bin group: most pair rows / most b doubles of one bin (sizes the pattern kernel's shared memory)
    std::vector<uint32_t> kLane, kElem;
    std::vector<uint32_t> kSegs;           // ring segments of the pattern-mode bins, 2 words each
    std::vector<int> binSegOff, binSegCnt, binSegBytes;   // per bin (same indexing as kBinPatOff): its segments / bytes per grid step
    long long maxEpochSlots = 0;           // largest number of live states in one epoch; after the bin images are built: stride
                                           // of the b carry buffers = that (slot order) + the lane-ordered region behind it
    long long dirOff = 0, dirSize = 0;     // lane-ordered region of the carry buffers (direct hand-over between pattern-mode epochs)
    std::vector<int> lastPairBase;         // [nPat] first pair (in a step row of the last epoch) of the pattern's only live branch there, -1 if it has several
    bool lastSingle = false;               // every pattern has one live branch in the last epoch
    std::vector<int> grpDirIn;             // per bin group: 1 if its blocks read b in lane order (written by the previous epoch's lanes)
    std::vector<int> slotOff;              // [nEpochs][nPat+1]
    std::vector<long long> epochBase;      // [nEpochs]
    std::vector<uint32_t> words, meta;
    std::vector<long long> trBase, trEntBase;
    std::vector<int> trOff;
    std::vector<uint32_t> trEnt;
    std::vector<uint8_t> shortcutOK;       // [nPat]
    std::vector<long long> epochSlots;     // [nEpochs] total live states per epoch
    std::vector<long long> epochSlotsShort;// [nEpochs] same, restricted to patterns that take the shortcut
    // FP64 instructions (DMUL / DFMA, one issue slot of the FP64 pipe each) the keyed kernels execute per grid step of
    // epoch e, summed over the patterns (all / those that take the shortcut): the numerator of the pipe-level roofline
    std::vector<long long> epochFp64, epochFp64Short;
    int maxLive = 0;
    long long totalSlots = 0;
};

// Builds the plan for `patterns` ([nPat][P]) with initial ancestral counts n_k - m_k.  Returns "" on
// success, else an error message (RC_ERR_UNSUPPORTED).
std::string build_plan(int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec,
                       const std::vector<StructEv>& sev, const std::vector<GapInfo>& gaps, bool hasSplit,
                       int smemBudgetBytes, PlanHost& out, int patternMode = 1);
// patternMode: 1 (default) boxes of at most RC_PAT_MAXL states run one pattern per thread; 0 every pattern goes to the row /
// state bins (the "pattern_mode" option of rc_set_option; the parity tests cover both)

// Bounds / capacity self-check of a built plan ("" if consistent).
std::string check_plan(const PlanHost& h);
uint64_t plan_checksum(const PlanHost& h);

}  // namespace rc
