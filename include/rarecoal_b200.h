/* rarecoal_b200.h — C ABI of librarecoal_b200.so
 *
 * Drop-in boundary for ONE path of stschiff/rarecoal: the probability of every allele-sharing
 * pattern of a histogram under a ModelSpec and the multinomial log-likelihood built on it.
 * The reference has no FFI of its own; the seam this library replaces is the per-pattern fan-out
 *
 *     src/RarecoalLib/MaxUtils.hs:107-110
 *         let jointStateSpace = makeJointStateSpace nrPops maxAf
 *             coreFunc        = C1.getProb modelSpec jointStateSpace nVecModelMapped
 *         patternProbs <- sequence $ parMap rdeepseq coreFunc standardOrderModelMapped
 *
 * plus the reduction at MaxUtils.hs:85-91 (computeLogLikelihoodFromSpec).  A Haskell maintainer
 * binds these entry points with `foreign import ccall` (see INTEGRATION.md and
 * haskell/RarecoalB200.hs).  Plain pointers and sizes only; the caller owns every host array, the
 * library owns all device memory inside rc_ctx.  One rc_ctx per OS thread / per GPU; a context is
 * not re-entrant.  There is NO CPU fallback: every entry point that computes fails with
 * RC_ERR_CUDA when no CUDA device is usable.
 */
#ifndef RARECOAL_B200_H
#define RARECOAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rc_ctx rc_ctx;

/* Return codes.  Positive codes mirror the reference's exception constructors
 * (src/RarecoalLib/Utils.hs:47-52); negative codes are library-level failures. */
enum {
    RC_OK = 0,
    RC_ERR_MODEL = 1,        /* RarecoalModelException  (validateModel, StateSpace.hs:163-197)      */
    RC_ERR_COMP = 2,         /* RarecoalCompException   (Core.hs:42,53-56; MaxUtils.hs:72-75)        */
    RC_ERR_TEMPLATE = 3,     /* RarecoalModeltemplateException (ModelTemplate.hs:65,226)              */
    RC_ERR_HIST = 4,         /* RarecoalHistogramException (Utils.hs:133,143,153,162,203)             */
    RC_ERR_ARG = -1,         /* bad argument / call order                                            */
    RC_ERR_CUDA = -2,        /* CUDA runtime failure or no device                                    */
    RC_ERR_UNSUPPORTED = -3  /* problem exceeds a compiled-in limit (see rc_last_error)              */
};

/* ModelEvent / EventType (src/RarecoalLib/StateSpace.hs:30-40).
 *   type 0  Join k l        : branch l merges into k                       (v unused)
 *   type 1  Split k l m     : fraction v=m of branch l moves to k
 *   type 2  SetPopSize k p  : v = p                                        (l unused)
 *   type 3  SetFreeze k b   : v != 0 -> frozen                             (l unused)
 *   type 4  SetGrowthRate k r : v = r                                      (l unused)
 *           NOT a reference constructor (StateSpace.hs:36-40 has four); BASELINE.json's north_star names it.  It is sugar:
 *           from its time t0 on, N_k(t) = N_k(t0) exp(-r (t - t0)) (t = time into the past) until the next SetPopSize /
 *           SetGrowthRate on k or the Join that removes k, realised as one `SetPopSize k N_k(t_s)` at every grid point
 *           t0 < t_s < tStop (rc_desugar_growth returns that list).  r = 0 ends a growth phase at the current size.
 *           It contributes nothing to the regularisation penalty.                                                  */
enum { RC_EV_JOIN = 0, RC_EV_SPLIT = 1, RC_EV_POPSIZE = 2, RC_EV_FREEZE = 3, RC_EV_GROWTH = 4 };
typedef struct {
    double t;
    int32_t type;
    int32_t k;
    int32_t l;
    double v;
} rc_event;

/* Counters of the last rc_eval / rc_eval_partial on this context. */
typedef struct {
    int64_t state_steps;     /* executed b(x) updates over all models (SURVEY.md §8d unit, Core.hs:248) */
    int64_t slots;           /* live states resident at t=0 over all models                            */
    double kernel_ms;        /* CUDA-event span of the propagation launches on the eval stream (a call cut into chunks: of */
                             /* its last chunk; likewise traj_ms and tables_ms)                                            */
    double traj_ms;          /* trajectory kernels (first plan group): elapsed on their side stream,     */
                             /* concurrent with the propagation launches                                 */
    double tables_ms;        /* CUDA-event time of the per-model table kernel                           */
    double total_ms;         /* CUDA-event time first launch -> last launch of the call                 */
    int32_t launches;        /* kernels launched by the call                                            */
    int32_t plan_cache_hit;  /* 1 if every model reused a cached transition plan                        */
    int32_t n_epochs;        /* structural epochs of the last model                                     */
    int32_t n_steps;         /* grid steps executed by the last model                                   */
    int64_t fp64_instr;      /* FP64 instructions (DMUL / DFMA, thread level) the propagation kernels of the trajectory-key */
                             /* tier executed for these models: numerator of the pipe-level roofline      */
    double plan_build_ms;    /* host time spent building and uploading transition plans in this call                     */
    int32_t plan_misses;     /* event structures of this call that were not in the plan cache                            */
    int32_t plan_evictions;  /* plans evicted (least recently used first) to make room                                   */
    double ktab_mb;          /* MiB of trajectory tables written by the call (all chunks)                                */
    int32_t chunks;          /* chunks the batch was cut into to respect "ktab_total_mb"                                 */
    int32_t models_selfcontained; /* models whose table exceeded the budget and ran the self-contained kernels          */
    int32_t graph_replay;    /* 1 if the launch chain of the (last chunk of the) call was replayed as a CUDA graph        */
    int32_t pad_;
    double host_ms;          /* host time of the call up to its last launch: model resolution (validateModel, events to   */
                             /* grid steps), plan lookup, packing and enqueueing                                         */
} rc_stats;

/* ---- context ---------------------------------------------------------------------------------- */

/* Creates a context bound to CUDA device `device` (cudaSetDevice).  Fails with RC_ERR_CUDA when the
 * device cannot be used.                                                                           */
int rc_create(int device, rc_ctx** out);
void rc_destroy(rc_ctx* ctx);
/* Message of the last failing call on this context ("" if none).  Valid until the next call.       */
const char* rc_last_error(rc_ctx* ctx);

/* ---- problem = what computeFrequencySpectrum closes over (MaxUtils.hs:97-107) ------------------- */

/* patterns: [nPat][P] in model-branch order, normally computeStandardOrder (Utils.hs:169-177) mapped
 * through turnHistPatternIntoModelPattern (Utils.hs:114-119); nVec likewise; counts[i] = histogram
 * count of pattern i (0 if unseen, MaxUtils.hs:112); totalSites = raTotalNrSites.  Every pattern
 * must satisfy 1 <= sum <= maxAf (the JointStateSpace of StateSpace.hs:73-87 has no other states). */
int rc_set_problem(rc_ctx* ctx, int P, int maxAf, int nPat, const int32_t* patterns,
                   const int32_t* nVec, const int64_t* counts, int64_t totalSites);

/* Pattern sharding for the multi-GPU path: this context evaluates only the patterns dealt to `rank`
 * of `world` (contiguous chunks of equal work in lexicographic pattern order, so that a rank's patterns share
 * their trajectory keys).  Call after rc_set_problem; default (0,1).                                */
int rc_set_shard(rc_ctx* ctx, int rank, int world);

/* mTimeSteps (StateSpace.hs:44; Utils.hs:64-73).  t[0..T) strictly increasing, > 0.               */
int rc_set_times(rc_ctx* ctx, int T, const double* t);

/* getTimeSteps n0 lingen tMax (Utils.hs:64-73) computed by the library; returns the number of grid
 * points and fills out[0..min(n,cap)).                                                             */
int rc_time_steps(int n0, int lingen, double tMax, double* out, int cap);

/* ---- evaluation ------------------------------------------------------------------------------- */

/* Evaluates B models (ModelSpecs sharing the time grid) on the current problem.
 *   ev / evOff : events of model b are ev[evOff[b] .. evOff[b+1])  (mEvents, any order; sorted
 *                stably by time inside, Core.hs:292-294)
 *   theta[b]   : mTheta;  disc[b*P .. b*P+P) : mDiscoveryRates;  noShortcut : mNoShortcut
 *   siteRed    : effective-site reduction factor of computeLogLikelihoodFromSpec (MaxUtils.hs:85-91)
 *   outProbs   : NULL or [B][nPat] — getProb of every pattern (Core.hs:38-56)
 *   outLogl    : [B] — siteRed * (sum cnt_i log p_i + (total - sum cnt_i) log(1 - sum p_i))
 *   outStatus  : [B] — RC_OK, RC_ERR_MODEL (validateModel failed) or RC_ERR_COMP (binomial overflow,
 *                pattern outside nVec); outLogl[b] is NaN for a failed model
 * Returns RC_OK when the call itself ran (individual models may still carry a status).  With a
 * shard set, outProbs holds zeros for patterns owned by other ranks and outLogl is computed from
 * this rank's partial sums only — use rc_eval_partial + rc_finalize for the sharded path.          */
int rc_eval(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta,
            const double* disc, int noShortcut, double siteRed, double* outProbs, double* outLogl,
            int32_t* outStatus);

/* rc_eval with mNoShortcut per model (StateSpace.hs:49 makes it a field of every ModelSpec): noShortcut[b] != 0 disables
 * the closed-form tail (Core.hs:80-118) for model b only.                                          */
int rc_eval_models(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta,
                   const double* disc, const int32_t* noShortcut /*[B]*/, double siteRed, double* outProbs,
                   double* outLogl, int32_t* outStatus);

/* Sharded path: like rc_eval, but leaves this rank's partial sums on the DEVICE.
 *   dPartial   : device pointer, [2*B] doubles: (sum cnt_i log p_i, sum p_i) per model, this shard
 *   dProbs     : device pointer or NULL, [B][nPat]; entries of patterns owned by other ranks are 0
 *   stream     : cudaStream_t to launch on (NULL = default stream); the call does not synchronise
 * The caller all-reduces (sum) dPartial (and dProbs) across ranks — torch.distributed / NCCL — and
 * then calls rc_finalize.                                                                          */
int rc_eval_partial(rc_ctx* ctx, int B, const rc_event* ev, const int32_t* evOff, const double* theta,
                    const double* disc, int noShortcut, double* dPartial, double* dProbs, void* stream,
                    int32_t* outStatus);

/* outLogl[b] = siteRed * (ll_b + (totalSites - sum counts) * log(1 - sp_b)) from reduced partials
 * (device pointer, [2*B]); synchronises `stream` and writes host outLogl[B].                       */
int rc_finalize(rc_ctx* ctx, int B, const double* dPartialReduced, double siteRed, void* stream,
                double* outLogl);

/* ---- multi-GPU (the reference's fan-out over one machine's cores, MaxUtils.hs:108-110, taken across GPUs) ---------- */

/* One context that drives nDevices GPUs of this process (devices == NULL: 0 .. nDevices-1).  Every call of this header
 * works on it; rc_eval / rc_eval_models shard the batch over the devices themselves:
 *   - B >= nDevices: models sharded — device d evaluates a contiguous slice of the models on all patterns, no collective;
 *   - B <  nDevices: patterns sharded — every device evaluates its pattern shard of all models, the 2*B partial sums
 *     (and the probabilities when outProbs != NULL) are summed with ncclAllReduce over NVLink, device 0 finalises.
 * rc_set_option(ctx, "multi_shard", 0 auto | 1 models | 2 patterns) forces one.  NCCL is bound at run time (libnccl.so.2);
 * without it the pattern-sharded path fails with RC_ERR_CUDA — there is no substitute path.  rc_eval_partial /
 * rc_finalize / rc_set_shard return RC_ERR_ARG on such a context (their device pointers belong to one device).          */
int rc_create_multi(int nDevices, const int32_t* devices, rc_ctx** out);
/* Devices behind a context (1 for rc_create).                                                       */
int rc_device_count(rc_ctx* ctx);

/* One process per GPU (MPI / torchrun style): rank 0 obtains a 128-byte NCCL unique id (ncclGetUniqueId), the launcher
 * hands it to every rank, and each rank joins its context to the communicator.  rc_comm_init also fixes the pattern shard
 * (rc_set_shard(rank, nRanks)); from then on rc_eval / rc_eval_models evaluate the rank's shard, all-reduce the partial
 * sums (and probabilities) with ncclAllReduce on the context's stream, and return the full result on EVERY rank — all
 * ranks must make the same calls with the same models.  rc_comm_allreduce is the same sum for callers of rc_eval_partial
 * (in place, device pointer, `stream` as in rc_eval_partial).                                                         */
int rc_comm_unique_id(void* out128);
int rc_comm_init(rc_ctx* ctx, int nRanks, int rank, const void* uniqueId128);
int rc_comm_allreduce(rc_ctx* ctx, double* dBuf, int64_t n, void* stream);

/* getProb for ONE pattern (Core.hs:38-56; the `rarecoal prob` command, Prob.hs:28): builds a
 * one-pattern problem on a scratch context state.  nVec/config have length P.                      */
int rc_get_prob(rc_ctx* ctx, int P, int maxAf, const int32_t* nVec, const int32_t* config,
                const rc_event* ev, int nEv, double theta, const double* disc, int noShortcut,
                double* outProb);

/* Tuning switches (defaults in parentheses):
 *   "keyed" (1)            patterns of the fast tier go through the trajectory-key kernels when the table fits;
 *                          0 forces the self-contained kernels (used by the parity tests to cover both)
 *   "pattern_mode" (1)     box-shaped live sets of at most 36 states run one pattern per thread; 0 sends every pattern
 *                          to the row / state kernels (used by the parity tests to cover those kernels too)
 *   "ktab_model_mb" (1024) largest trajectory table of one model; a model above it uses the self-contained kernels
 *   "ktab_total_mb" (12288) largest trajectory table memory of one call
 *   "graphs" (1)           the launch chain of a call is captured as a CUDA graph the second time the same launch signature
 *                          (batch size, event structures, buffers) is seen and replayed from then on; 0 = plain launches
 *   "plan_cache_max" (64)  transition plans (one per event structure) kept per context, least recently used evicted first
 *   "plan_cache_mb" (49152) device memory the cached plans may hold
 *   "fuse_last" (1)        a last epoch that consists of the grid step of the last Join and the shortcut (a tree model) is
 *                          finished by the launches of the epoch before it and not launched itself; same numbers bit for bit
 *   "pattern_embed" (0)    1: live sets that are not boxes (what a Split leaves) also run one pattern per thread, on their
 *                          bounding boxes of at most 36 states; measured slower than the row kernel on the example data
 *   "big_smem_kb" (0)      > 0: shared-memory budget of the big tier (patterns with more than 256 live states).  A live
 *                          set that does not fit runs the same kernel on a global scratch buffer (any size up to 65 534
 *                          states per pattern); the option exists so that tests can force that mode on a small problem */
int rc_set_option(rc_ctx* ctx, const char* name, double value);

int rc_last_stats(rc_ctx* ctx, rc_stats* out);

/* Measures the FP64 FMA rate of the device with a register-resident DFMA chain kernel (the
 * roofline denominator; MEASURED_PEAKS.json carries none for FP64).  Returns TFLOP/s in *out.      */
int rc_fp64_peak(rc_ctx* ctx, double* outTflops);

/* ---- host-side mirrors of small reference functions (no device work) --------------------------- */

/* computeAllConfigs maxAf nVec (Utils.hs:100-104): returns n, fills out[n][P] up to cap rows.      */
int rc_all_configs(int maxAf, int P, const int32_t* nVec, int32_t* out, int cap);
/* validateModel (StateSpace.hs:163-197): RC_OK or RC_ERR_MODEL with a message in err.              */
int rc_validate_model(int P, const double* disc, const rc_event* ev, int nEv, char* err, int errCap);
/* getRegularizationPenalty (StateSpace.hs:199-221).                                                */
double rc_reg_penalty(int P, double reg, const rc_event* ev, int nEv);
/* choose n k (Utils.hs:208-210).                                                                   */
double rc_choose(int n, int k);
/* The model's events with every SetGrowthRate (RC_EV_GROWTH) rewritten into per-grid-step SetPopSize events, stably
 * sorted by time — what rc_eval evaluates.  Returns the number of events, fills out up to cap.                     */
int rc_desugar_growth(int T, const double* times, const rc_event* ev, int nEv, rc_event* out, int cap);

/* The pattern indices rc_set_shard(rank, world) assigns to `rank` (host only): returns the count, fills out up to cap.
 * Shards of all ranks are disjoint and cover every pattern.                                                     */
int rc_shard_patterns(int P, int nPat, const int32_t* patterns, int rank, int world, int32_t* out, int cap);

/* Runs the host planner only (no device): executed state-steps (one unit = one iteration of the
 * updateB loop, Core.hs:248), live states at t=0, largest live set of a pattern, structural epochs and
 * kernel bins for this problem/model, after a bounds / capacity self-check of the compiled plan (RC_ERR_UNSUPPORTED
 * if it fails).  Used for roofline accounting and to check the planner on CPU.                       */
int rc_plan_stats(int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec, int T,
                  const double* times, const rc_event* ev, int nEv, int noShortcut, int64_t* stateSteps,
                  int64_t* slots0, int32_t* maxLive, int32_t* nEpochs, int32_t* nBins);
/* FNV-1a checksum of the plan the calling thread's last successful rc_plan_stats built (every array a kernel reads): a plan
 * is a function of the problem and the event structure, not of the host threads that built it (RC_PLAN_THREADS).   */
uint64_t rc_plan_last_checksum(void);

/* JointStateSpace (StateSpace.hs:20-28,73-124): ids are positions in computeAllConfigs.            */
typedef struct rc_statespace rc_statespace;
rc_statespace* rc_ss_create(int P, int maxAf);
void rc_ss_destroy(rc_statespace* ss);
int rc_ss_nr_states(const rc_statespace* ss);
int rc_ss_state_to_id(const rc_statespace* ss, const int32_t* x);            /* -1 if not a state   */
int rc_ss_id_to_state(const rc_statespace* ss, int id, int32_t* x);
int rc_ss_x1up(const rc_statespace* ss, int id, int32_t* out);               /* -1 where sum>maxAf  */
int rc_ss_x1(const rc_statespace* ss, int k);
/* fillUpStateSpace (StateSpace.hs:145-160): returns n, fills out up to cap.                        */
int rc_ss_fill_up(const rc_statespace* ss, const int32_t* ids, int n, int32_t* out, int cap);

/* ---- formats either side of the path (host only) ------------------------- */

/* Model template DSL (ModelTemplate.hs:78-145).  Writes a JSON description of the parsed template
 * ({"branches", "events":[{"cmd","t","a","b","v","flag"}], "constraints", "params"}; "params" = getParamNames,
 * ModelTemplate.hs:147-174) into out.  Returns 0, -RC_ERR_TEMPLATE (message in err), or the required buffer size
 * when cap is too small.                                                                            */
int rc_template_to_json(const char* text, char* out, int cap, char* err, int errCap);
/* instantiateModel (ModelTemplate.hs:176-241): events of the ModelSpec for a parameter dictionary.  Returns the
 * number of events (K = Join then SetPopSize), or -RC_ERR_TEMPLATE (parse error / missing parameter) /
 * -RC_ERR_MODEL (unknown branch, violated constraint) with a message in err.                        */
int rc_template_instantiate(const char* text, int nParams, const char* const* names, const double* values,
                            rc_event* out, int cap, char* err, int errCap);
/* The same for B parameter vectors (values[B][nParams]), the template parsed once: events of the models one after the other in
 * `out` (evOff[b] .. evOff[b+1], ready for rc_eval), per model a status (0, RC_ERR_MODEL, RC_ERR_TEMPLATE: no events) and, if
 * `penalty` is given, getRegularizationPenalty.  Returns the number of events or a negative code.                              */
int rc_template_instantiate_batch(const char* text, int nParams, const char* const* names, int B, const double* values, double reg,
                                  rc_event* out, int cap, int32_t* evOff, int32_t* status, double* penalty, char* err, int errCap);
/* makeParameterDict (ModelTemplate.hs:243-255): text of a maxl ("name value") or mcmc ('#' header, 3 lines skipped,
 * MaxL column) output file.  namesOut receives the names separated by '\n'.  Returns the number of pairs.          */
int rc_params_parse(const char* text, char* namesOut, int namesCap, double* values, int cap);

/* Rare-allele histogram file (NAMES= / N= / MAX_M= / [MIN_M=] / TOTAL_SITES= header, then "m1,m2,.. count" rows;
 * exampleData/fourPop.histogram.txt:1-5).                                                                       */
typedef struct rc_hist rc_hist;
rc_hist* rc_hist_read(const char* path, char* err, int errCap);
void rc_hist_free(rc_hist* h);
int rc_hist_info(const rc_hist* h, int32_t* nPops, int32_t* nRows, int32_t* minM, int32_t* maxM, int64_t* totalSites,
                 int32_t* nVec, char* names, int namesCap);
/* loadHistogram filters (Utils.hs:130-167), computeStandardOrder (Utils.hs:169-177) and the mapping of patterns and
 * nVec to model-branch order (Utils.hs:114-119): the arguments of rc_set_problem.  maxAf == 0 keeps the file's
 * MAX_M.  Returns the number of patterns (outputs filled up to cap rows) or -RC_ERR_HIST with a message in err.    */
int rc_hist_problem(const rc_hist* h, int maxAf, int minAf, const int32_t* conditionOn, int nCond,
                    const int32_t* exclude, int nExclude, int nModelBranches, const char* const* modelBranches,
                    int32_t* patternsOut, int64_t* countsOut, int32_t* nVecOut, int cap, char* err, int errCap);

const char* rc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RARECOAL_B200_H */
