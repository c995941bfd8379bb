// extern "C" entry points for the coarse tracker (include/tandem_b200.h).
#include "../../include/tandem_b200.h"
#include "capi_common.h"
#include "common.cuh"
#include "tracker.h"

struct tdm_tracker {
  tdm::TrackerIface* impl;
};

extern "C" {

int tdm_tracker_create(int w, int h, float huber, float cutoff, int n_max, int device, tdm_tracker** out) {
  TDM_API_BEGIN
  TDM_CHECK(out, "null argument");
  *out = new tdm_tracker{tdm::make_tracker(w, h, huber, cutoff, n_max, device)};
  return TDM_OK;
  TDM_API_END
}
void tdm_tracker_destroy(tdm_tracker* t) {
  if (!t) return;
  try { delete t->impl; } catch (...) {}
  delete t;
}
int tdm_tracker_set_k(tdm_tracker* t, int w, int h, float fx, float fy, float cx, float cy) {
  TDM_API_BEGIN
  TDM_CHECK(t, "null handle");
  t->impl->set_k(w, h, fx, fy, cx, cy);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_set_reference(tdm_tracker* t, int n, const float* pc_u, const float* pc_v, const float* pc_idepth,
                              const float* pc_color, float ref_exposure, const double ref_aff_g2l[2]) {
  TDM_API_BEGIN
  TDM_CHECK(t && ref_aff_g2l && (n == 0 || (pc_u && pc_v && pc_idepth && pc_color)), "null argument");
  t->impl->set_reference(n, pc_u, pc_v, pc_idepth, pc_color, ref_exposure, ref_aff_g2l);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_set_new(tdm_tracker* t, const float* dInew) {
  TDM_API_BEGIN
  TDM_CHECK(t && dInew, "null argument");
  t->impl->set_new(dInew);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_calc_res(tdm_tracker* t, const double* refToNew, float new_exposure, const double aff_g2l[2],
                         float cutoffTH, double res6[6]) {
  TDM_API_BEGIN
  TDM_CHECK(t && refToNew && aff_g2l && res6, "null argument");
  t->impl->calc_res(refToNew, new_exposure, aff_g2l, cutoffTH, res6);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_calc_g(tdm_tracker* t, float new_exposure, const double aff_g2l[2], double H[64], double b[8]) {
  TDM_API_BEGIN
  TDM_CHECK(t && aff_g2l && H && b, "null argument");
  t->impl->calc_g(new_exposure, aff_g2l, H, b);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_calc_res_g(tdm_tracker* t, const double* refToNew, float new_exposure, const double aff_g2l[2],
                           float cutoffTH, double res6[6], double H[64], double b[8]) {
  TDM_API_BEGIN
  TDM_CHECK(t && refToNew && aff_g2l && res6 && H && b, "null argument");
  t->impl->calc_res_g(refToNew, new_exposure, aff_g2l, cutoffTH, res6, H, b);
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_synchronize(tdm_tracker* t) {
  TDM_API_BEGIN
  TDM_CHECK(t, "null handle");
  t->impl->synchronize();
  return TDM_OK;
  TDM_API_END
}
int tdm_tracker_run_resident(tdm_tracker* t, int iters, float* ms_total) {
  TDM_API_BEGIN
  TDM_CHECK(t && ms_total, "null argument");
  t->impl->run_resident(iters, ms_total);
  return TDM_OK;
  TDM_API_END
}

}  // extern "C"
