// rc_types.h — POD layouts shared by the host planner and the sm_100a kernels.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

#define RC_MAX_P 32          // branches (k is stored in 8 bits; tables are P*(maxAf+1) doubles per step)
#define RC_MAX_AF 62         // maxAf (x stored in 8 bits; weight index (maxAf+1)^2 must fit 16 bits minus sentinel)
#define RC_BLOCK 256         // threads per propagation block
// event type codes (same values as RC_EV_JOIN / RC_EV_SPLIT in include/rarecoal_b200.h)
#define RC_EV_JOIN_D 0
#define RC_EV_SPLIT_D 1
// row mode of the keyed tier: one thread per row of states along the widest branch of a box-shaped live set
#define RC_ROW_NO 7          // other non-zero branches of a row at most
#ifndef RC_PBLOCK
#define RC_PBLOCK 64         // pattern mode: threads (= patterns) per block
#endif
#ifndef RC_PAT_MAXT
#define RC_PAT_MAXT 12       // pattern mode: three instantiations of the kernel by the live states of a pattern (all in the
#endif
#define RC_PAT_MAXS 25       //   registers of one thread) — the fewer states, the fewer registers and the more resident blocks
#define RC_PAT_MAXL 36
#define RC_NGROUP 5          // bin groups per epoch: tiny, small, large patterns (one thread each), rows (one thread per row), states
#define RC_NPATG 3           // ... of which the first RC_NPATG are the pattern groups
// epochMode of a pattern group = RC_MODE_PAT + index of its capacity in {RC_PAT_MAXT, RC_PAT_MAXS, RC_PAT_MAXL}
#define RC_MODE_PAT 3
static inline __host__ __device__ int rc_pat_cap(int mode) { return mode == RC_MODE_PAT ? RC_PAT_MAXT : mode == RC_MODE_PAT + 1 ? RC_PAT_MAXS : RC_PAT_MAXL; }
// patterns per thread of the smallest pattern group (rc_pat_kernel.cuh) — its bins hold RC_PBLOCK * RC_PAT12_NP patterns — and
// lanes per state in the lane-ordered region of the carry buffers (one stride for the bins of every group).  Two patterns per
// thread (sharing the per-step ring copies, barrier and slot arithmetic) measured slower on C3: 7.73 ms against 7.61 with 6
// resident blocks, 7.9 with 5 or 4 — the tier is bound by its shared-memory loads, which two patterns per thread do not reduce.
#ifndef RC_PAT12_NP
#define RC_PAT12_NP 1
#endif
#define RC_DIR_LANES (RC_PBLOCK * RC_PAT12_NP)
static inline __host__ __device__ int rc_pat_lanes(int mode) { return mode == RC_MODE_PAT ? RC_PBLOCK * RC_PAT12_NP : RC_PBLOCK; }
static inline __host__ __device__ int rc_pat_floor(int mode) { return mode == RC_MODE_PAT ? 0 : mode == RC_MODE_PAT + 1 ? RC_PAT_MAXT : RC_PAT_MAXS; }
#define RC_PLANE_N 16        // words of a pattern-mode lane record
#define RC_ROW_SUM 11        // with more than 3 other branches: other branches + row length <= RC_ROW_SUM (bounds the kernel variants)
#define RC_ROW_MAXM 10       // longest row
#define RC_ROW_BCAP 1536     // states per row-mode block at most (rows are stored with an odd stride: no bank conflicts)
#define RC_NPBR 128          // patterns per row-mode bin at most
// keyed fast tier: shared-memory pair rows per parity (row 0 = {dt,0}, last row = {1,0} padding)
#define RC_KROWS 256
// how a trajectory key obtains its starting a from the previous epoch (popJoinA / popSplitA, Core.hs:165-170,195-200)
#define RC_KOP_INIT 0        // a = n_k - m_k
#define RC_KOP_COPY 1        // a = a[pa]
#define RC_KOP_JOIN 2        // a = a[pb] + a[pa]           (pa: this branch, pb: the branch joined in)
#define RC_KOP_ZERO 3        // a = 0                       (branch joined away)
#define RC_KOP_SPLIT_K 4     // a = m * a[pb] + a[pa]
#define RC_KOP_SPLIT_L 5     // a = (1 - m) * a[pa]
#define RC_NPB 64            // patterns per bin (block) at most

// Packed per-slot dimension entry: one per non-zero component x_k of a live state x.
//   bits  0..7   k        branch index
//   bits  8..15  x_k      (>= 1)
//   bits 16..31  up       pattern-relative slot of x + e_k, RC_NO_UP if that state is not live
#define RC_NO_UP 0xFFFFu
#define RC_W_ONE 0xFFFFu     // transition weight index meaning "1.0" (Join)

// Per-slot meta word:  bits 0..7 nnz, bit 8 isUnit (x == e_k), bits 16..23 |x| = sum x
#define RC_META_UNIT 0x100u

// One transition entry (epoch e -> e+1): new_b[slot] += old_b[src] * W[widx]
//   bits 0..15 src (pattern-relative slot in epoch e), bits 16..31 widx (x_l*(maxAf+1)+s, or RC_W_ONE)

// Device view of a compiled plan (one per problem x event-structure signature).
struct RcPlanDev {
    int P, maxAf, A1;            // A1 = maxAf + 1
    int nPat;                    // patterns in this shard
    int nEpochs;                 // structural epochs = fired Join/Split events + 1
    int nnzw;                    // words per slot
    int nBins;
    int nBinsKB;                 // the first nBinsKB big-tier bins hold patterns that models on the trajectory-key tier run in row mode
    int cap;                     // slots per block
    int zeroRow;                 // 1: every trajectory key is led by the pair of x = 0, {1, 0}, at keyPairBase - 1
    int bigGlobal;               // 1: the big tier keeps its per-slot arrays in a global scratch buffer (live sets beyond shared memory)
    const int* binPatOff;        // [nBins+1]          big-tier bins (several states per thread)
    const int* binPats;          // shard-local pattern ids, bin after bin
    int nFastBins;
    int fastBlock;               // threads (= state slots) per fast-tier block: 64, 128 or 256
    const int* fBinPatOff;       // [nFastBins+1]      fast-tier bins (one state per thread)
    const int* fBinPats;
    const uint8_t* maxX;         // [nEpochs][nPat][P] largest x_k over the live states of a pattern
    const int* slotOff;          // [nEpochs][nPat+1]  slot offsets inside the epoch's arrays
    const long long* epochBase;  // [nEpochs]  base of epoch e inside words/meta (in slots)
    const uint32_t* words;       // [totalSlots][nnzw]
    const uint32_t* meta;        // [totalSlots]
    const long long* trBase;     // [nEpochs]   base of transition e (into epoch e+1 slots) in trOff space
    const int* trOff;            // per new slot: [.. +1] offsets into trEnt, relative to trEntBase[e]
    const long long* trEntBase;  // [nEpochs]
    const uint32_t* trEnt;
    const uint8_t* shortcutOK;   // [nPat] structural+positivity shortcut test (Core.hs:80-89) at the shortcut step
    const int32_t* aInit;        // [nPat][P]  n_k - m_k
    const uint8_t* aEvtPos;      // unused in v1
    const double* combFac;       // [nPat]  prod choose(n_k, m_k)   (Core.hs:49)
    const uint8_t* support;      // [nPat][P] 1 where m_k > 0        (Core.hs:50-52)
    const int32_t* patGlobal;    // [nPat] index in the caller's pattern order
    // ---- trajectory keys (keyed fast tier) ----
    const struct RcEpochDev* ep; // [nEpochs]
    int totalKeys, nKeysLast;
    const uint8_t* keyBranch;    // per key (concatenated over epochs, RcEpochDev::keyOff)
    const uint8_t* keyMaxJ;      // pairs of the key = largest x_k any pattern with this key reaches in the epoch
    const uint8_t* keyOp;        // RC_KOP_*
    const int* keyPa;            // parent keys in the previous epoch (-1: none)
    const int* keyPb;
    const int* keyInit;          // n_k - m_k for RC_KOP_INIT
    const int* keyPairBase;      // first pair of the key inside a step row (pair 0 of every row is {dt, 0})
    const int* keyThrBase;       // first trajectory-kernel thread of the key
    const int* thrKey;           // trajectory-kernel thread -> key (RcEpochDev::thrOff)
    const int* lastKeyId;        // [nPat][P] keys of the final epoch
    const int* lastPairBase;     // [nPat] first pair of the pattern's only live branch in a step row of the final epoch (-1: several)
    const int* kBinBase;         // [nEpochs] first entry of the epoch in kBinPatOff / kFetchOff
    const int* nKBins;           // [nEpochs] keyed bins of the epoch
    const int* kBinPatOff;       // per epoch nKBins[e]+1 offsets into kBinPats
    const int* kBinPats;         // shard-local pattern ids, epoch after epoch, bin after bin
    const int* kFetchOff;        // same indexing as kBinPatOff -> kFetchRows
    const int* kFetchRows;       // pair offset inside a step row for every shared-memory pair row of a bin
    const uint16_t* kRowBase;    // [position in kBinPats][P] first shared-memory pair row of (pattern, branch)
    long long maxEpochSlots;     // stride (doubles per model) of the b carry buffers
    long long dirOff;            // ... whose doubles [dirOff, maxEpochSlots) are the lane-ordered region (direct hand-over)
    // ---- row mode ----
    const int* epochMode;        // [nEpochs] 0: one state per thread; one row per thread with <= 3 (1) or <= RC_ROW_NO (2) other branches; one pattern per thread with <= RC_PAT_MAXT (3), <= RC_PAT_MAXS (4) or <= RC_PAT_MAXL (5) states
    const int* rowOff;           // [nEpochs][nPat+1] rows before a pattern
    const long long* rowEpochBase; // [nEpochs] first row of the epoch in rowRefAll / rowWordsAll
    const uint16_t* rowRefAll;   // per row RC_ROW_MAXM reference-order indices of its states
    const uint32_t* rowWordsAll; // per row RC_ROW_NO words: k | x_k << 8 | neighbour row << 16
    const uint32_t* patRow;      // [nEpochs][nPat] row length | row branch << 8 | other branches << 16
    // ---- bin images: what a block's set-up needs, already resolved to lane order (rc_plan.h, build_bin_images) ----
    const int* kBinHdr;          // [bin][RC_HDR_N], same bin indexing as kBinPatOff
    const int* kPatRec;          // [position in kBinPats][2]: pattern id, first lane | first b double << 16
    const uint32_t* kLane;       // lane records, RC_LANE_N (state mode) or RC_RLANE_N (row mode) words each
    const uint32_t* kElem;       // [element][2], row mode: gather record of every b double of a bin
    const uint32_t* kSegs;       // [segment][2], pattern mode: runs of pairs a bin copies per grid step (RC_HDR_N comment)
};

// Bin header (ints): 0 first entry in kBinPats/kPatRec, 1 patterns, 2 lanes (states or rows), 3 b doubles (row mode, padded
// rows; pattern mode: | RC_HDR_DIRIN if the bin's b arrives in lane order, see [12]), 4 first entry in kFetchRows, 5 pair rows to fetch, 6 first lane record (in 16-byte units), 7 first
// element record.
#define RC_HDR_DIRIN 0x40000000
#define RC_HDR_SEED 0x20000000      // first epoch, pattern mode: b is 1 at the last state of every lane (the pattern itself), else 0
#define RC_HDR_NB(x) ((x) & 0x1FFFFFFF)
#define RC_E_FUSE 0x40000000          // epoch argument of the pattern kernel: finish the patterns here, the last epoch is not launched
#define RC_PLANE_DIROUT 0x80000000u
#define RC_PLANE_UNIT 0x40000000u     // lane record [2]: the state at local index 0 is e_A (updateD reads it)
#define RC_HDR_N 16
// Pattern mode, further header ints: 8 first ring segment of the bin in kSegs, 9 segments, 10 bytes they bring in per grid step,
// 11 the shape of the bin's patterns (one per bin: MA | MB << 4 | ...), 12 (RC_HDR_DIRIN) first double of the bin in the
// lane-ordered region of the carry buffers, laid out [state of the box][lane], RC_PBLOCK lanes.
// A segment (2 x u32: first pair inside a step row of the trajectory table; first ring row | pair rows << 16) is a run of
// pairs that is contiguous both in the table and in the ring: one bulk copy per segment and grid step fills a ring slot.
// To make the runs long, a key's pairs are stored with the same odd stride in the table and in the ring and the ring rows of a
// bin are sorted by their place in the table.
static inline __host__ __device__ int rc_pair_stride(int n) { return (n >= 2 && !(n & 1)) ? n + 1 : n; }
// Lane record (8 x u32).
//   state mode: [0..1] pair row of component d (8 x u8; RC_KROWS-1 = the {1,0} padding pair)
//               [2..3] lane of the neighbour x + e_k (8 x u8; own lane = no such state)
//               [4] meta word  [5] slot inside the epoch (carry buffer)  [6..7] gather record
//   row mode (RC_RLANE_N = 12 x u32):
//               [0..1] pair rows of the other branches (7 x u8) | pair row of (row branch, j = 1) << 24 in [1]
//               [2] M | NO << 8
//               [3..6] b offsets in doubles (8 x u16): own row, then the neighbour row along each other branch
//                      (RC_ROW_BCAP = the always-zero row)
//               [7] first slot of the pattern in the epoch  [8] row index in rowRefAll  [9] pattern id  [10..11] 0
//   pattern mode (RC_PLANE_N words; one lane = one pattern, its states at the mixed-radix position
//               ((((x_A - 1) MB' + x_B - 1) MC' + x_C - 1) MD' + x_D - 1) ME' + x_E - 1, M' = max(M, 1)):
//               [0..1] pair row of x = 1 of every non-zero branch (8 x u8): A, the shaped branches B.., then the m = 1 branches
//               [2] MA | MB << 4 | MC << 8 | MD << 12 | ME << 16 | number of m = 1 branches << 20
//               [3] first slot of the pattern in the epoch  [4] pattern id  [5] first b double of the pattern in the bin
//               [6..14] reference-order index of every state (36 x u8)
//               Direct hand-over (bit 31 of [2]): the next event is a Join and the pattern's next bin reads b in lane order.
//               The lane then applies the Join itself: [6..14] hold the next epoch's lane-local index of every state, bits
//               24..29 of [2] the number of states there and [3] the pattern's first double in the lane-ordered region (its
//               states follow at stride RC_PBLOCK).
//   The header's first-lane-record field counts 16-byte units (2 per state lane, 3 per row lane, 4 per pattern lane).
// Gather record (2 x u32): first transition entry (absolute index in trEnt; RC_IMG_SEED: the value is 1, the seed state)
// and (first slot of the pattern in the previous epoch) | entries << RC_IMG_CNT_SHIFT.
#define RC_LANE_N 8
#define RC_RLANE_N 12
#define RC_IMG_SEED 0xFFFFFFFFu
#define RC_IMG_CNT_SHIFT 26

struct RcEpochDev {
    int nKeys, nThr, rowPairs, keyOff, thrOff;
    int elBase;                  // first element record of the epoch
};

// Per (model, epoch): grid steps [beg, end) executed in that epoch and where its first step row starts in the
// trajectory table (in pairs).
struct RcModelEpochDev {
    int beg, end;
    long long tabBase;
};

// Structural event as executed by the kernel.
struct RcSEvDev {
    int step;      // grid step at which it fires (before that step's update, Core.hs:120-135)
    int type;      // RC_EV_JOIN / RC_EV_SPLIT
    int k, l;
    double m;      // split fraction
    int wOff;      // offset (in doubles) of this event's weight table W[x_l][s] in the model pool
    int pad;
};

// Per-model header.
struct RcModelDev {
    int nSteps;        // grid steps to run at most (T)
    int sShort;        // step at which the shortcut test applies, -1 = never
    int nSEv;
    int sevOff;        // first structural event of this model in the pooled RcSEvDev array
    int rowsAvail;     // table rows computed for this model
    int pad;           // caller's model index
    int mode;          // 0: self-contained tiers only; 1: fast-tier patterns use the keyed kernel
    int mepOff;        // first RcModelEpochDev of this model
    long long aEndOff; // offset (doubles) of this model's per-key end-of-epoch a values
    long long aShortOff; // offset (doubles) of this model's per-key a at the shortcut step (final epoch keys)
    long long tabOff;  // offset (doubles) of row 0 in the table pool
    long long wPool;   // offset (doubles) of split weight tables in the weight pool
    double theta;
    double disc[RC_MAX_P];
    double Nshort[RC_MAX_P];   // population sizes at the shortcut step
};

// Piecewise-constant population sizes and freeze flags used by the table kernel: per (model, branch) the grid steps at
// which the size changes (SetPopSize, Core.hs:152-153; many entries under SetGrowthRate), per model the steps at which the
// frozen mask changes (SetFreeze, Core.hs:154-157).  Both lists start with an entry for step 0 (N = 1, nothing frozen,
// Core.hs:296-297) and are sorted by step.
struct RcSizeStepDev {
    int sBegin;
    int pad;
    double N;
};
struct RcMaskStepDev {
    int sBegin;
    unsigned frozenMask;
};

// Table row layout (doubles), one row per (model, grid step):
//   [0]                      dt
//   [1]                      frozen mask (as double)
//   [2 + 2*(k*A1 + j)]       16-byte pairs, k < P, j < A1:
//                              j == 0 : { EA[k], invN[k] }
//                                       EA[k]   = exp(-0.5*dt/N_k)   (Core.hs:240); exactly 1.0 if the branch is
//                                                 frozen or dt <= 0 (updateA is then the identity)
//                                       invN[k] = 1/N_k, 0 if frozen (so a_k/N_k drops out of t1)
//                              j >= 1 : { CN[k][j], E2[k][j] }
//                                       CN = j(j-1)/2 * (1/N_k)                       (Core.hs:270), 0 if frozen
//                                       E2 = 1 - exp(-(j(j+1)/2) * (1/N_k) * dt)      (Core.hs:279), 0 if frozen
//   [2 + 2*P*A1 + k]         N[k]
// Rows are padded to an even number of doubles so that every pair stays 16-byte aligned.
#define RC_ROW_PAIRS 2
static inline __host__ __device__ int rc_row_n(int P, int A1) { return RC_ROW_PAIRS + 2 * P * A1; }
static inline __host__ __device__ int rc_row_len(int P, int A1) { return (RC_ROW_PAIRS + 2 * P * A1 + P + 1) & ~1; }
#define RC_ROWCAP 256        // fast tier: table rows of at most this many doubles (two-slot ring in shared memory)
