// Drop-in replacement header for tandem/libdr/cuda_coarse_tracker/include/public/cuda_coarse_tracker.h (:9-82)
// implemented over the tandem_b200 C ABI.  Caller: tandem/src/FullSystem/CoarseTracker.cpp:103-106,144,732,777-795,
// 861-887.  The class owns only an opaque handle; callers `new` it, so the layout is ours to define.
#ifndef PBA_CUDA_COARSE_TRACKER_H
#define PBA_CUDA_COARSE_TRACKER_H

#include <Eigen/Dense>

#include <vector>

struct tdm_tracker;   // include/tandem_b200.h
struct tdm_pyramid;
struct tdm_fusion;

class CudaCoarseTracker {
public:
  CudaCoarseTracker(int w, int h, float setting_huberTH, float setting_coarseCutoffTH);

  void setK(int w, int h, float fx, float fy, float cx, float cy);

  ~CudaCoarseTracker();

  void init(int n_max_in = 0);

  void free();

  void setReference(int n_in, float const *pc_u_in, float const *pc_v_in, float const *pc_idepth_in,
                    float const *pc_color_in, float ref_exposure_in, Eigen::Vector2d const &ref_aff_g2l_in);

  void setNew(float const *dInew_in);

  Eigen::Matrix<double, 6, 1> calcRes(Eigen::Matrix<double, 4, 4> const &refToNew, float new_exposure,
                                      Eigen::Vector2d const &aff_g2l, float cutoffTH);

  void calcG(Eigen::Matrix<double, 8, 8> &H_out, Eigen::Matrix<double, 8, 1> &b_out, const float new_exposure,
             const Eigen::Vector2d &aff_g2l);

  // Extension (not in the reference): calcRes + calcG in ONE launch, no warped buffers.
  Eigen::Matrix<double, 6, 1> calcResAndG(Eigen::Matrix<double, 4, 4> const &refToNew, float new_exposure,
                                          Eigen::Vector2d const &aff_g2l, float cutoffTH,
                                          Eigen::Matrix<double, 8, 8> &H_out, Eigen::Matrix<double, 8, 1> &b_out);

  // Extension: calcRes of several motion hypotheses (FullSystem::trackNewCoarse tries up to 31, FullSystem.cpp:437-530) in ONE
  // launch and one synchronisation; element k equals calcRes(refToNew[k], new_exposure, aff_g2l[k], cutoffTH) bit for bit.
  std::vector<Eigen::Matrix<double, 6, 1>> calcResBatch(std::vector<Eigen::Matrix<double, 4, 4>> const &refToNew, float new_exposure,
                                                        std::vector<Eigen::Vector2d> const &aff_g2l, float cutoffTH);

  // Extensions for the stages either side of the evaluation:
  //  * setNewFromPyramid: setNew from a device-resident FrameHessian::makeImages replacement (tdm_pyramid);
  //  * setReferenceDense: the dense part of CoarseTracker::setCoarseTrackingRef + setReference; the depth map is the
  //    render_index-th map of the last DrFusion::RenderAsync and never leaves the device;
  //  * track: one pyramid level of CoarseTracker::trackNewestCoarse without host round trips.
  void setNewFromPyramid(tdm_pyramid *pyramid, int level);

  int setReferenceDense(tdm_fusion *fusion, int render_index, Eigen::Matrix<double, 4, 4> const &T_depth_to_ref,
                        int tracking_step, bool dense_only, int n_sparse, float const *pc_u, float const *pc_v,
                        float const *pc_idepth, float const *pc_color, float const *idepth0, tdm_pyramid *ref_pyramid,
                        float ref_exposure, Eigen::Vector2d const &ref_aff_g2l);

  struct TrackResult {
    Eigen::Matrix<double, 4, 4> refToNew;
    Eigen::Vector2d aff_g2l;
    Eigen::Matrix<double, 6, 1> res;
    int iterations;
    float levelCutoffRepeat;
  };
  TrackResult track(Eigen::Matrix<double, 4, 4> const &refToNew, Eigen::Vector2d const &aff_g2l, float new_exposure,
                    float coarseCutoffTH, int maxIterations, float lambdaExtrapolationLimit, bool fix_a, bool fix_b);

  void synchronize();

  void startTiming();

  float endTimingMilliseconds();

private:
  tdm_tracker *handle_ = nullptr;
  const int w = 0, h = 0;
  const float setting_huberTH, setting_coarseCutoffTH;
  bool timing_ = false;
  double t_start_ = 0;
};

#endif  // PBA_CUDA_COARSE_TRACKER_H
