// TDMW weight container reader (layout documented in tandem_b200/weights_io.py) and BatchNorm folding.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace tdm {

struct HostTensor {
  std::vector<int> dims;
  std::vector<float> data;
  size_t numel() const { return data.size(); }
};

struct WeightFile {
  std::map<std::string, HostTensor> t;
  int depth_num[3] = {0, 0, 0};
  bool view_aggregation = false;

  const HostTensor& get(const std::string& n) const {
    auto it = t.find(n);
    if (it == t.end()) throw Error("weight tensor missing: " + n);
    return it->second;
  }
  bool has(const std::string& n) const { return t.count(n) != 0; }
};

inline std::string resolve_weight_path(const std::string& p) {
  // DrMvsnet(filename) receives ".../model.pt" in the reference (FullSystem.cpp:284); the weights of the
  // from-scratch runtime live beside it as ".../model.tdmw".
  auto ends_with = [&](const char* s) {
    size_t n = std::strlen(s);
    return p.size() >= n && p.compare(p.size() - n, n, s) == 0;
  };
  if (ends_with(".tdmw")) return p;
  size_t dot = p.find_last_of('.');
  size_t slash = p.find_last_of('/');
  if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) return p.substr(0, dot) + ".tdmw";
  return p + ".tdmw";
}

inline WeightFile load_tdmw(const std::string& path_in) {
  const std::string path = resolve_weight_path(path_in);
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error("cannot open weight file " + path);
  std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.size() < 32 || std::memcmp(buf.data(), "TDMW0001", 8) != 0) throw Error(path + ": not a TDMW0001 file");
  size_t p = 8;
  auto rd32 = [&]() { uint32_t v; if (p + 4 > buf.size()) throw Error("TDMW truncated"); std::memcpy(&v, &buf[p], 4); p += 4; return v; };
  auto rd64 = [&]() { uint64_t v; if (p + 8 > buf.size()) throw Error("TDMW truncated"); std::memcpy(&v, &buf[p], 8); p += 8; return v; };
  WeightFile wf;
  uint32_t n = rd32();
  for (int i = 0; i < 3; ++i) wf.depth_num[i] = (int)rd32();
  wf.view_aggregation = rd32() != 0;
  struct Meta { std::string name; std::vector<int> dims; uint64_t off; };
  std::vector<Meta> metas;
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t ln = rd32();
    if (p + ln > buf.size()) throw Error("TDMW truncated");
    Meta m;
    m.name.assign(&buf[p], ln); p += ln;
    uint32_t nd = rd32();
    for (uint32_t d = 0; d < nd; ++d) m.dims.push_back((int)rd32());
    m.off = rd64();
    metas.push_back(std::move(m));
  }
  uint64_t nf = rd64();
  if (p + nf * 4 > buf.size()) throw Error("TDMW data truncated");
  const char* data = &buf[p];
  for (auto& m : metas) {
    size_t cnt = 1;
    for (int d : m.dims) cnt *= (size_t)d;
    if (m.off + cnt > nf) throw Error("TDMW tensor out of range: " + m.name);
    HostTensor ht;
    ht.dims = m.dims;
    ht.data.resize(cnt);
    std::memcpy(ht.data.data(), data + m.off * 4, cnt * 4);
    wf.t[m.name] = std::move(ht);
  }
  return wf;
}

// A convolution with BatchNorm folded in, rearranged to [tap][cin_padded][cout] fp32.
struct FoldedConv {
  int cin = 0, cout = 0, kd = 1, kh = 1, kw = 1;
  bool transposed = false;
  std::vector<float> w;     // [kd*kh*kw][cin][cout]
  std::vector<float> bias;  // [cout] (empty if none)
};

// conv_w: PyTorch layout (cout,cin,k..) or for transposed (cin,cout,k..). bn_prefix empty -> no BN.
inline FoldedConv fold_conv(const WeightFile& wf, const std::string& conv_w, const std::string& conv_b,
                            const std::string& bn_prefix, bool transposed, int cin_pad = 0) {
  const HostTensor& W = wf.get(conv_w);
  FoldedConv fc;
  fc.transposed = transposed;
  const int nd = (int)W.dims.size();
  const int a = W.dims[0], b = W.dims[1];
  const int cout = transposed ? b : a, cin = transposed ? a : b;
  if (nd == 4) { fc.kd = 1; fc.kh = W.dims[2]; fc.kw = W.dims[3]; }
  else if (nd == 5) { fc.kd = W.dims[2]; fc.kh = W.dims[3]; fc.kw = W.dims[4]; }
  else throw Error("unexpected conv weight rank: " + conv_w);
  const int taps = fc.kd * fc.kh * fc.kw;
  fc.cout = cout;
  fc.cin = cin_pad > cin ? cin_pad : cin;
  std::vector<double> scale(cout, 1.0), shift(cout, 0.0);
  bool has_bias = false;
  if (!conv_b.empty()) {
    const HostTensor& B = wf.get(conv_b);
    for (int c = 0; c < cout; ++c) shift[c] = B.data[c];
    has_bias = true;
  }
  if (!bn_prefix.empty()) {
    const auto& g = wf.get(bn_prefix + ".weight").data;
    const auto& be = wf.get(bn_prefix + ".bias").data;
    const auto& mu = wf.get(bn_prefix + ".running_mean").data;
    const auto& var = wf.get(bn_prefix + ".running_var").data;
    for (int c = 0; c < cout; ++c) {
      const double s = (double)g[c] / std::sqrt((double)var[c] + 1e-5);
      scale[c] = s;
      shift[c] = (shift[c] - (double)mu[c]) * s + (double)be[c];
    }
    has_bias = true;
  }
  fc.w.assign((size_t)taps * fc.cin * cout, 0.f);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        const size_t src = transposed ? ((size_t)ci * cout + co) * taps + t : ((size_t)co * cin + ci) * taps + t;
        fc.w[((size_t)t * fc.cin + ci) * cout + co] = (float)((double)W.data[src] * scale[co]);
      }
  if (has_bias) {
    fc.bias.resize(cout);
    for (int c = 0; c < cout; ++c) fc.bias[c] = (float)shift[c];
  }
  return fc;
}

}  // namespace tdm
