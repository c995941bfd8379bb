// Links the header-compatible C++ shims (include/dr_mvsnet, include/dr_fusion) against libtandem_b200.so and runs
// (1) the reference's known-answer test through the DrMvsnet class surface, exactly as FullSystem::initDr does
//     (FullSystem.cpp:284-285: new DrMvsnet(path); test_dr_mvsnet(*mvsnet, sample_inputs, print, 4)),
// (2) the DrFusion call order of tandem_backend.cpp:166-190.
#include <cmath>
#include <cstdio>
#include <vector>

#include "dr_fusion/dr_fusion.h"
#include "dr_mvsnet/dr_mvsnet.h"

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
  return (ok && ok2) ? 0 : 1;
}
