// Type-erased interface of the level-0 photometric tracker (tracker.cu).
#pragma once
#include "common.cuh"

namespace tdm {

class TrackerIface {
 public:
  virtual ~TrackerIface() = default;
  virtual void set_k(int w, int h, float fx, float fy, float cx, float cy) = 0;
  virtual void set_reference(int n, const float* u, const float* v, const float* idepth, const float* color,
                             float ref_exposure, const double ref_aff[2]) = 0;
  virtual void set_new(const float* dI) = 0;
  virtual void calc_res(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH, double res6[6]) = 0;
  virtual void calc_g(float new_exposure, const double aff[2], double H[64], double b[8]) = 0;
  virtual void calc_res_g(const double* refToNew, float new_exposure, const double aff[2], float cutoffTH, double res6[6],
                          double H[64], double b[8]) = 0;
  virtual void synchronize() = 0;
  virtual void run_resident(int iters, float* ms) = 0;
};

TrackerIface* make_tracker(int w, int h, float huber, float cutoff, int n_max, int device);

}  // namespace tdm
