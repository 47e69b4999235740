// rc_multi.h — single-process multi-GPU context (rc_create_multi): N per-device contexts behind one rc_ctx.
// Built on the public C ABI only (rc_create / rc_set_problem / rc_comm_init / rc_eval_models ... on its children).
#pragma once
#include <stdint.h>

#include "../../include/rarecoal_b200.h"

struct RcMulti;

int rc_multi_create(int nDevices, const int32_t* devices, RcMulti** out);
void rc_multi_destroy(RcMulti* m);
const char* rc_multi_last_error(RcMulti* m);
int rc_multi_count(RcMulti* m);
rc_ctx* rc_multi_first(RcMulti* m);
int rc_multi_set_problem(RcMulti* m, int P, int maxAf, int nPat, const int32_t* patterns, const int32_t* nVec,
                         const int64_t* counts, int64_t totalSites);
int rc_multi_set_times(RcMulti* m, int T, const double* t);
int rc_multi_set_option(RcMulti* m, const char* name, double value);
int rc_multi_eval_models(RcMulti* m, int B, const rc_event* ev, const int32_t* evOff, const double* theta, const double* disc,
                         const int32_t* noShortcut, double siteRed, double* outProbs, double* outLogl, int32_t* outStatus);
int rc_multi_get_prob(RcMulti* m, int P, int maxAf, const int32_t* nVec, const int32_t* config, const rc_event* ev, int nEv,
                      double theta, const double* disc, int noShortcut, double* outProb);
int rc_multi_last_stats(RcMulti* m, rc_stats* out);
