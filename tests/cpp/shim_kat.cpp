// Links the header-compatible C++ shims (include/dr_mvsnet, include/dr_fusion) against libtandem_b200.so and runs
// (1) the reference's known-answer test through the DrMvsnet class surface, exactly as FullSystem::initDr does
//     (FullSystem.cpp:284-285: new DrMvsnet(path); test_dr_mvsnet(*mvsnet, sample_inputs, print, 4)),
// (2) the DrFusion call order of tandem_backend.cpp:166-190,
// (3) CudaCoarseTracker (shims/cuda_coarse_tracker.cpp) against the stand-in <Eigen/Dense> of tests/cpp/eigen_stub.
#include <cmath>
#include <cstdio>
#include <vector>

#include <stdexcept>

#include "cuda_coarse_tracker/cuda_coarse_tracker.h"   // needs <Eigen/Dense>: tests/cpp/eigen_stub here, the real one in tandem
#include "dr_fusion/dr_fusion.h"
#include "dr_mvsnet/dr_mvsnet.h"

// (3) CudaCoarseTracker through the C++ class exactly as CoarseTracker.cpp:103-106,144,732,777-795,861-887 drives it:
//     construct, init, setK, setReference, setNew, calcRes, calcG (+ the fused extension), error convention = exceptions.
static bool tracker_kat() {
  const int w = 160, h = 120;
  std::vector<float> dI((size_t)3 * w * h);
  auto I = [](float x, float y) { return 100.f + 50.f * std::sin(0.1f * x) + 30.f * std::cos(0.13f * y); };
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float* p = &dI[3 * (size_t)(x + y * w)];
      p[0] = I((float)x, (float)y);
      p[1] = 0.5f * (I((float)x + 1, (float)y) - I((float)x - 1, (float)y));
      p[2] = 0.5f * (I((float)x, (float)y + 1) - I((float)x, (float)y - 1));
    }
  std::vector<float> u, v, id, col;
  for (int y = 6; y < h - 6; y += 2)
    for (int x = 6; x < w - 6; x += 2) { u.push_back((float)x); v.push_back((float)y); id.push_back(0.5f); col.push_back(I((float)x, (float)y)); }
  const int n = (int)u.size();
  CudaCoarseTracker trk(w, h, 9.f, 20.f);
  trk.init();
  bool threw = false;
  try { trk.setK(w + 1, h, 80, 80, 79.5f, 59.5f); } catch (std::runtime_error const&) { threw = true; }   // cpp:359
  trk.setK(w, h, 80, 80, 79.5f, 59.5f);
  trk.setReference(n, u.data(), v.data(), id.data(), col.data(), 1.f, Eigen::Vector2d(0, 0));
  trk.setNew(dI.data());
  Eigen::Matrix<double, 4, 4> T = Eigen::Matrix<double, 4, 4>::Identity();
  Eigen::Matrix<double, 6, 1> r0 = trk.calcRes(T, 1.f, Eigen::Vector2d(0, 0), 20.f);
  T(0, 3) = 0.02;
  Eigen::Matrix<double, 6, 1> r1 = trk.calcRes(T, 1.f, Eigen::Vector2d(0, 0), 20.f);
  Eigen::Matrix<double, 8, 8> H, H2;
  Eigen::Matrix<double, 8, 1> b, b2;
  trk.calcG(H, b, 1.f, Eigen::Vector2d(0, 0));
  Eigen::Matrix<double, 6, 1> r2 = trk.calcResAndG(T, 1.f, Eigen::Vector2d(0, 0), 20.f, H2, b2);
  double asym = 0, dfused = 0, hmax = 0;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      asym = std::fmax(asym, std::fabs(H(i, j) - H(j, i)));
      dfused = std::fmax(dfused, std::fabs(H(i, j) - H2(i, j)));
      hmax = std::fmax(hmax, std::fabs(H(i, j)));
    }
  printf("CudaCoarseTracker: n=%d  E(identity)=%.3e terms=%.0f | E(shift)=%.3e terms=%.0f | |H|max=%.3e asym=%.2e fused-vs-two-call=%.2e  setK mismatch threw=%d\n",
         n, r0(0), r0(1), r1(0), r1(1), hmax, asym, dfused, (int)threw);
  return threw && r0(1) == (double)n && r0(0) < 1e-3 * n && r1(0) > r0(0) && r1(1) > 0.9 * n && hmax > 0 && asym <= 1e-9 * hmax &&
         dfused <= 1e-6 * hmax && std::fabs(r2(0) - r1(0)) <= 1e-6 * r1(0);
}

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage: shim_kat <model.pt|weights.tdmw> <sample_inputs.bin>\n"); return 2; }
  DrMvsnet mvsnet(argv[1]);
  const bool ok = test_dr_mvsnet(mvsnet, argv[2], true, 2);
  printf("test_dr_mvsnet: %s\n", ok ? "PASS" : "FAIL");

  DrFusionOptions o{};
  o.voxel_size = 0.01f; o.num_buckets = 50021; o.bucket_size = 10; o.num_blocks = 40000; o.block_size = 8;
  o.max_sdf_weight = 64; o.truncation_distance = 0.04f; o.max_sensor_depth = 10.f; o.min_sensor_depth = 0.1f;
  o.num_render_streams = 1; o.fx = 80; o.fy = 80; o.cx = 79.5f; o.cy = 59.5f; o.height = 120; o.width = 160;
  DrFusion fusion(o);
  std::vector<unsigned char> bgr(120 * 160 * 3, 128);
  std::vector<float> depth(120 * 160, 1.0f);  // fronto-parallel plane at 1 m
  const float pose[16] = {1, 0, 0, 5, 0, 1, 0, 5, 0, 0, 1, 5, 0, 0, 0, 1};
  fusion.IntegrateScanAsync(bgr.data(), depth.data(), pose);
  fusion.RenderAsync({pose});
  std::vector<unsigned char*> rb;
  std::vector<float*> rd;
  fusion.GetRenderResult(rb, rd);
  double err = 0; int n = 0;
  for (int i = 0; i < 120 * 160; ++i) if (rd[0][i] > 0) { err += std::fabs(rd[0][i] - 1.0f); ++n; }
  printf("DrFusion plane render: %d hits, mean |depth-1| = %.4f\n", n, n ? err / n : -1.0);
  const bool ok2 = n > 120 * 160 * 0.9 && err / n < 0.01;
  const bool ok3 = tracker_kat();
  printf("CudaCoarseTracker shim: %s\n", ok3 ? "PASS" : "FAIL");
  return (ok && ok2 && ok3) ? 0 : 1;
}
