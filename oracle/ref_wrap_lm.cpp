// Test infrastructure (never linked into the product): harness around the Levenberg-Marquardt level loop of the REFERENCE,
//   CoarseTracker::trackNewestCoarse      tandem/src/FullSystem/CoarseTracker.cpp:750-916
// cut out of the file where it lies by oracle/ref_build.mk (-> oracle/_ref/gen/lm_loop.inc, git-ignored build output) and
// compiled here, unmodified, against a stand-in for Eigen / Sophus (tests/cpp/eigen_stub) and the members it touches.
// Purpose: pin oracle/lm_driver.py.  The loop runs at lvl = 0 with `cudaCoarseTracker` set, so every residual /
// normal-equation evaluation goes through two C callbacks (calcRes, calcG) - the test plugs the SAME tracker object into the
// reference loop and into lm_driver.track_level0 and compares poses, affine parameters, iteration counts and cutoff repeats.
// What is the reference's: the control flow (cutoff doubling, lambda schedule, extrapolation, SCALE_*, accept / reject, the
// break on |inc|, the repeat-level rule).  What is the stand-in's: the 8x8 / 7x7 / 6x6 LDLT solve and SE3::exp.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <Eigen/Dense>   // tests/cpp/eigen_stub

#define PYR_LEVELS 6
#define SCALE_XI_ROT 1.0f      /* HessianBlocks.h:60-66 */
#define SCALE_XI_TRANS 0.5f
#define SCALE_A 10.0f
#define SCALE_B 1000.0f
typedef Eigen::Matrix<double, 8, 8> Mat88;
typedef Eigen::Matrix<double, 8, 1> Vec8;
typedef Eigen::Matrix<double, 7, 1> Vec7;
typedef Eigen::Matrix<double, 6, 1> Vec6;
typedef Eigen::Matrix<double, 5, 1> Vec5;
typedef Eigen::Matrix<double, 3, 1> Vec3;
typedef Eigen::Matrix<double, 2, 1> Vec2;
typedef Eigen::Matrix<float, 2, 1> Vec2f;
typedef Eigen::Matrix<double, 3, 3> Mat33;
typedef Eigen::Matrix<double, 4, 4> Mat44;

// ---- stand-in for Sophus::SE3d (tangent = (upsilon, omega), left-multiplicative update as the loop uses it)
class SE3 {
 public:
  Mat33 R;
  Vec3 t;
  SE3() { R = Mat33::Identity(); t.setZero(); }
  static SE3 exp(const Vec6& xi) {
    const Vec3 ups(xi[0], xi[1], xi[2]), om(xi[3], xi[4], xi[5]);
    const double th = om.norm();
    Mat33 Om; Om.setZero();
    Om(0, 1) = -om[2]; Om(0, 2) = om[1]; Om(1, 0) = om[2]; Om(1, 2) = -om[0]; Om(2, 0) = -om[1]; Om(2, 1) = om[0];
    double A, B, C;
    if (th < 1e-8) { A = 1.0 - th * th / 6; B = 0.5 - th * th / 24; C = 1.0 / 6 - th * th / 120; }
    else { A = std::sin(th) / th; B = (1 - std::cos(th)) / (th * th); C = (th - std::sin(th)) / (th * th * th); }
    const Mat33 Om2 = Om * Om;
    SE3 o;
    Mat33 V;
    for (int i = 0; i < 9; ++i) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      o.R.d[i] = I + A * Om.d[i] + B * Om2.d[i];
      V.d[i] = I + B * Om.d[i] + C * Om2.d[i];
    }
    o.t = V * ups;
    return o;
  }
  SE3 operator*(const SE3& b) const { SE3 o; o.R = R * b.R; o.t = R * b.t + t; return o; }
  Mat44 matrix() const {
    Mat44 m; m.setZero();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m(r, c) = R(r, c); m(r, 3) = t(r); }
    m(3, 3) = 1;
    return m;
  }
  Vec6 log() const { Vec6 v; v.setZero(); return v; }   // only inside `if (debugPrint)` printouts
};

struct AffLight {   // util/NumType.h:166-192
  AffLight(double a_, double b_) : a(a_), b(b_) {}
  AffLight() : a(0), b(0) {}
  double a, b;
  static Vec2 fromToVecExposure(float exposureF, float exposureT, AffLight g2F, AffLight g2T) {
    if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
    const double a = std::exp(g2T.a - g2F.a) * exposureT / exposureF;
    const double b = g2T.b - a * g2F.b;
    return Vec2(a, b);
  }
  Vec2 vec() { return Vec2(a, b); }
};

typedef void (*CalcResFn)(void* user, const double* refToNew16_rowmajor, float new_exposure, const double* aff2, float cutoffTH, double* res6);
typedef void (*CalcGFn)(void* user, float new_exposure, const double* aff2, double* H64_rowmajor, double* b8);

struct CudaCoarseTrackerStub {   // the three members of CudaCoarseTracker the loop calls (cuda_coarse_tracker.h:27-38)
  void* user; CalcResFn res; CalcGFn g; int n_res = 0, n_g = 0;
  void setNew(float const*) {}
  Vec6 calcRes(const Mat44& T, float new_exposure, Vec2 aff, float cutoffTH) {
    double Tm[16], a[2] = {aff[0], aff[1]}, r[6];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Tm[4 * i + j] = T(i, j);
    res(user, Tm, new_exposure, a, cutoffTH, r);
    ++n_res;
    Vec6 v; for (int i = 0; i < 6; ++i) v[i] = r[i];
    return v;
  }
  void calcG(Mat88& H, Vec8& b, float new_exposure, Vec2 aff) {
    double Hm[64], bm[8], a[2] = {aff[0], aff[1]};
    g(user, new_exposure, a, Hm, bm);
    ++n_g;
    for (int i = 0; i < 8; ++i) { for (int j = 0; j < 8; ++j) H(i, j) = Hm[8 * i + j]; b[i] = bm[i]; }
  }
};

struct FrameHessian { float ab_exposure = 1; Eigen::Vector3f* dIp[PYR_LEVELS] = {}; };

// ---- settings / members of CoarseTracker the loop names
static float setting_coarseCutoffTH = 20;
static int setting_affineOptModeA = 0, setting_affineOptModeB = 0;
static bool setting_debugout_runquiet = true;
static bool debugPrint = false;

namespace {
struct Loop {
  CudaCoarseTrackerStub* cudaCoarseTracker = nullptr;
  FrameHessian* newFrame = nullptr;
  FrameHessian* lastRef = nullptr;
  AffLight lastRef_aff_g2l;
  Vec5 lastResiduals;
  Vec3 lastFlowIndicators;
  int iterations_lvl0 = 0;
  float cutoff_repeat_lvl0 = 1;
  // never taken with cudaCoarseTracker != 0 at lvl 0 (CoarseTracker.cpp:775,781,792,861,877), but the branches must compile
  Vec6 calcRes(int, const SE3&, AffLight, float) { std::abort(); }
  void calcGSSSE(int, Mat88&, Vec8&, const SE3&, AffLight) { std::abort(); }

  bool run(FrameHessian* newFrameHessian, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, Vec5 minResForAbort) {
    lastResiduals.setConstant(NAN);
    lastFlowIndicators.setConstant(1000);
#include "_ref/gen/lm_loop.inc"   // CoarseTracker.cpp:750-916, verbatim
    lastToNew_out = refToNew_current;      // :919-920
    aff_g2l_out = aff_g2l_current;
    (void)flag_save;
    return true;
  }
};
}  // namespace

extern "C" {
// T16 (row-major 4x4, in/out), aff2 (in/out).  res6_out = lastResiduals[0], flow indicators (3), -, -.  Returns 1 if the loop
// finished (0 = aborted by the minResForAbort rule); counts[0] = calcRes calls, counts[1] = calcG calls.
int ref_lm_track_level0(void* user, CalcResFn res, CalcGFn g, double* T16, double* aff2, float new_exposure, float ref_exposure,
                        const double* ref_aff2, float coarse_cutoff, int fix_a, int fix_b, double* out5, int* counts) {
  CudaCoarseTrackerStub trk{user, res, g};
  FrameHessian nf, rf;
  nf.ab_exposure = new_exposure; rf.ab_exposure = ref_exposure;
  Loop L;
  L.cudaCoarseTracker = &trk; L.lastRef = &rf; L.lastRef_aff_g2l = AffLight(ref_aff2[0], ref_aff2[1]);
  setting_coarseCutoffTH = coarse_cutoff;
  setting_affineOptModeA = fix_a ? -1 : 0;    // settings.h: < 0 = fixed
  setting_affineOptModeB = fix_b ? -1 : 0;
  SE3 T;
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.R(r, c) = T16[4 * r + c]; T.t(r) = T16[4 * r + 3]; }
  AffLight aff(aff2[0], aff2[1]);
  Vec5 minRes; minRes.setConstant(NAN);       // NaN: the abort comparison is always false (FullSystem passes NaN on the first try, FullSystem.cpp:455)
  const bool ok = L.run(&nf, T, aff, 0, minRes);
  const Mat44 M = T.matrix();
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T16[4 * r + c] = M(r, c);
  aff2[0] = aff.a; aff2[1] = aff.b;
  out5[0] = L.lastResiduals[0];
  for (int i = 0; i < 3; ++i) out5[1 + i] = L.lastFlowIndicators[i];
  out5[4] = 0;
  counts[0] = trk.n_res; counts[1] = trk.n_g;
  return ok ? 1 : 0;
}
}
