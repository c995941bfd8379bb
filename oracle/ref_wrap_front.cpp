// Test infrastructure (never linked into the product): harness around two function bodies of the REFERENCE,
//   FrameHessian::makeImages                 tandem/src/FullSystem/HessianBlocks.cpp:128-191
//   CoarseTracker::setCoarseTrackingRef      tandem/src/FullSystem/CoarseTracker.cpp:655-725 (the dense-depth block)
// which oracle/ref_build.mk cuts out of the files where they lie under /root/reference (sed line ranges -> oracle/_ref/gen/*.inc,
// git-ignored build output) and compiles here, unmodified, against a stand-in for Eigen / Sophus / the DSO globals they touch.
// Purpose: pin oracle/front_oracle.c against the reference's own statements.
// What is the reference's: every statement inside the two bodies (loop bounds, flat-index gradient range, rounding, the OOB
// test, nearest-depth rule, raster-order append and the ++pc_n quirk).  What is the stand-in's: 3x3 float products (summed
// k = 0,1,2 like Eigen's coefficient-wise product) and the rigid-transform inverse / product in double.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include <Eigen/Dense>   // tests/cpp/eigen_stub

#define PYR_LEVELS 6
typedef Eigen::Matrix<float, 3, 3> Mat33f;
typedef Eigen::Matrix<float, 3, 1> Vec3f;
typedef Eigen::Matrix<double, 3, 3> Mat33;
typedef Eigen::Matrix<double, 3, 1> Vec3;
typedef Eigen::Matrix<double, 4, 4> Mat44;

// ---- stand-in for Sophus::SE3d: exactly the members the dense block uses
class SE3 {
 public:
  Mat33 R;
  Vec3 t;
  SE3() { R = Mat33::Identity(); t.setZero(); }
  explicit SE3(const Mat44& m) {
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R(r, c) = m(r, c); t(r) = m(r, 3); }
  }
  SE3 inverse() const { SE3 o; o.R = R.transpose(); o.t = -(o.R * t); return o; }
  SE3 operator*(const SE3& b) const { SE3 o; o.R = R * b.R; o.t = R * b.t + t; return o; }
  const Mat33& rotationMatrix() const { return R; }
  const Vec3& translation() const { return t; }
};

// ---- globals / classes of DSO that the two bodies name
static int wG[PYR_LEVELS], hG[PYR_LEVELS];
static int pyrLevelsUsed = 0;
static int setting_gammaWeightsPixelSelect = 1;   // settings.cpp default; irrelevant with HCalib == 0
static int setting_tracking_step = 1;
static bool dense_tracking_with_dense_depth_only = false;
static bool dr_timing = false;

struct CalibHessian { float getBGradOnly(float) { return 1.f; } };
struct FrameShell { SE3 camToWorld; };
struct FrameHessian {
  Eigen::Vector3f* dI = nullptr;
  Eigen::Vector3f* dIp[PYR_LEVELS] = {};
  float* absSquaredGrad[PYR_LEVELS] = {};
  FrameShell* shell = nullptr;
  void makeImages(float* color, CalibHessian* HCalib);
};

#include "_ref/gen/make_images.inc"   // HessianBlocks.cpp:128-191, verbatim

struct TandemCoarseTrackingDepthMap { bool is_valid = false; float cam_to_world[16]; float* depth = nullptr; };

extern "C" {

// dI_out: levels concatenated, wl*hl float3 each; absgrad_out likewise.  Rows 0 and h-1 of (dx, dy, absgrad) are left
// uninitialised by the reference (`new Eigen::Vector3f[]`): the harness pre-fills the fresh arrays with 0 via placement so
// that a comparison is defined there (documented in front_oracle.c as well).
int ref_front_make_images(const float* color, int w, int h, int levels, float* dI_out, float* absgrad_out) {
  if (levels > PYR_LEVELS) return -1;
  pyrLevelsUsed = levels;
  for (int l = 0; l < levels; ++l) { wG[l] = w >> l; hG[l] = h >> l; }
  FrameHessian fh;
  std::vector<float> c(color, color + (size_t)w * h);
  fh.makeImages(c.data(), nullptr);
  size_t off = 0;
  for (int l = 0; l < levels; ++l) {
    const size_t n = (size_t)wG[l] * hG[l];
    for (size_t i = 0; i < n; ++i) {
      const bool border = i < (size_t)wG[l] || i >= (size_t)wG[l] * (hG[l] - 1);   // never written by the reference
      dI_out[3 * (off + i)] = fh.dIp[l][i][0];
      dI_out[3 * (off + i) + 1] = border ? 0.f : fh.dIp[l][i][1];
      dI_out[3 * (off + i) + 2] = border ? 0.f : fh.dIp[l][i][2];
      absgrad_out[off + i] = border ? 0.f : fh.absSquaredGrad[l][i];
    }
    off += n;
    delete[] fh.dIp[l];
    delete[] fh.absSquaredGrad[l];
  }
  return 0;
}

// depth: the dense depth map rendered at pose c2w_depth (row-major 4x4 float, like TandemCoarseTrackingDepthMap);
// c2w_ref: camToWorld of the reference keyframe (row-major 4x4 double).  pc_* hold n_before+1 valid entries on entry
// (slot n_before = the stale slot of CoarseTracker.cpp:717-722).  Returns the new pc_n[0]; T_out = T_dense_depth_to_last.
int ref_front_dense_reference(const float* depth_in, int width_in, int height_in, int step, const float* c2w_depth,
                              const double* c2w_ref, float fx, float fy, float cx, float cy, int dense_only, int n_before,
                              const float* idepth0, const float* ref_gray, float* pc_u0, float* pc_v0, float* pc_idepth0,
                              float* pc_color0, double* T_out) {
  int w[PYR_LEVELS] = {width_in}, h[PYR_LEVELS] = {height_in};
  Mat33f K[PYR_LEVELS], Ki[PYR_LEVELS];
  K[0].setZero(); Ki[0].setZero();
  // CoarseTracker::makeK (CoarseTracker.cpp:130-146 region): K[0] from fx, fy, cx, cy; Ki[0] = K[0].inverse()
  K[0](0, 0) = fx; K[0](1, 1) = fy; K[0](0, 2) = cx; K[0](1, 2) = cy; K[0](2, 2) = 1;
  Ki[0](0, 0) = 1.0f / fx; Ki[0](1, 1) = 1.0f / fy; Ki[0](0, 2) = -cx / fx; Ki[0](1, 2) = -cy / fy; Ki[0](2, 2) = 1;
  setting_tracking_step = step;
  dense_tracking_with_dense_depth_only = dense_only != 0;
  int pc_n[PYR_LEVELS] = {n_before};
  float* pc_u[PYR_LEVELS] = {pc_u0};
  float* pc_v[PYR_LEVELS] = {pc_v0};
  float* pc_idepth[PYR_LEVELS] = {pc_idepth0};
  float* pc_color[PYR_LEVELS] = {pc_color0};
  std::vector<float> idz((size_t)width_in * height_in, 0.f);
  if (idepth0) idz.assign(idepth0, idepth0 + (size_t)width_in * height_in);
  float* idepth[PYR_LEVELS] = {idz.data()};
  std::vector<Eigen::Vector3f> gray((size_t)width_in * height_in);
  for (size_t i = 0; i < gray.size(); ++i) { gray[i][0] = ref_gray[i]; gray[i][1] = gray[i][2] = 0; }
  FrameShell shell;
  Mat44 m;
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = c2w_ref[4 * r + c];
  shell.camToWorld = SE3(m);
  FrameHessian ref_frame;
  ref_frame.shell = &shell;
  ref_frame.dIp[0] = gray.data();
  FrameHessian* lastRef = &ref_frame;
  TandemCoarseTrackingDepthMap dd;
  dd.is_valid = true;
  for (int i = 0; i < 16; ++i) dd.cam_to_world[i] = c2w_depth[i];
  std::vector<float> dcopy(depth_in, depth_in + (size_t)width_in * height_in);
  dd.depth = dcopy.data();
  const TandemCoarseTrackingDepthMap* dense_depth = &dd;
  (void)dr_timing;

#include "_ref/gen/dense_ref.inc"   // CoarseTracker.cpp:655-725, verbatim

  if (T_out) {
    Mat44 c2w_mat;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) c2w_mat(r, c) = dense_depth->cam_to_world[4 * r + c];
    const SE3 T = lastRef->shell->camToWorld.inverse() * SE3(c2w_mat);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T_out[4 * r + c] = T.R(r, c); T_out[4 * r + 3] = T.t(r); }
    T_out[12] = T_out[13] = T_out[14] = 0; T_out[15] = 1;
  }
  return pc_n[0];
}

}  // extern "C"
