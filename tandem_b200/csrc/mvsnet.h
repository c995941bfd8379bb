// Type-erased interface of the CVA-MVSNet engine (implemented in mvsnet.cu for fp32 and bf16 storage).
#pragma once
#include <string>

namespace tdm {

class MvsnetIface {
 public:
  virtual ~MvsnetIface() = default;
  virtual void set_option(const std::string& key, int value) = 0;
  // K3x3x3: three row-major 3x3 intrinsics, stage1..stage3.
  virtual void call_async(int H, int W, int V, int ref_index, unsigned char* const* bgrs, const float* K3x3x3,
                          float* const* c2ws, float dmin, float dmax, float discard) = 0;
  virtual bool ready() = 0;
  virtual void wait() = 0;
  virtual void get_result(float* depth, float* conf, float* depth_dense, float* conf_dense) = 0;
  virtual void stage_output(int stage, const std::string& which, float* out, size_t cap) = 0;
  virtual long long debug_tensor(const std::string& name, float* out, size_t cap, int* dims4) = 0;
  virtual void run_resident(int iters, float* ms_total, int* launches) = 0;
  virtual std::string profile() = 0;
  // building blocks of tdm_mvsnet_run_resident_multi (several engines of one device timed together)
  virtual void* resident_stream() = 0;                 // cudaStream_t
  virtual int resident_device() const = 0;
  virtual int resident_launch(int iters) = 0;          // enqueue iters forwards, no synchronisation; returns launches per forward
};

MvsnetIface* make_mvsnet(const std::string& weights_path, int precision, int device);
void debug_homography(const float* K3x3, const float* c2w_ref, const float* c2w_src, float* rot9, float* trans3);
void debug_conv_plan(int cin, int npad, int kd, int D, int H, int W, int pd, int mode, int smem_kb, long long* out12);

}  // namespace tdm
