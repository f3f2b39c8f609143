// Weight intake: takes the reference state_dict tensors unchanged (SURVEY.md 8b "Weights"), folds
// BatchNorm where the op order allows it (conv -> BN), keeps BN as a scale/shift prologue where it
// does not (BN -> ReLU -> conv in the D-TDNN layers, xvector.py:266-275), re-lays the LSTM matrices
// for the recurrent kernel, and builds the DFT / mel tables of the two front-ends in float64.
#include <cmath>
#include <cstring>

#include "cbx_internal.h"
#include "tc.cuh"

namespace cbx {
namespace {

using TensorMap = std::map<std::string, std::pair<const float*, int64_t>>;

struct Packer {
  std::vector<float> host;
  std::vector<std::pair<const float**, size_t>> fix;
  void add(const float** slot, const std::vector<float>& v) {
    size_t off = (host.size() + 63) & ~size_t(63);   // 256-byte alignment of every tensor
    host.resize(off + v.size());
    std::memcpy(host.data() + off, v.data(), v.size() * sizeof(float));
    fix.push_back({slot, off});
  }
  int upload(cbx_ctx* c, float** blob) {
    if (*blob) { cudaFree(*blob); *blob = nullptr; }
    CBX_CUDA_OK(c, cudaMalloc((void**)blob, host.size() * sizeof(float)));
    CBX_CUDA_OK(c, cudaMemcpy(*blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (auto& f : fix) *f.first = *blob + f.second;
    return CBX_OK;
  }
};

// fp32 -> bf16 (round to nearest even), two per float slot of the blob
static std::vector<float> to_bf16_words(const float* w, size_t n) {
  std::vector<float> out((n + 1) / 2, 0.f);
  uint16_t* h = reinterpret_cast<uint16_t*>(out.data());
  for (size_t i = 0; i < n; ++i) {
    uint32_t u; std::memcpy(&u, w + i, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    h[i] = (uint16_t)(u >> 16);
  }
  return out;
}

struct Getter {
  const TensorMap& t; cbx_ctx* c; bool ok = true;
  const float* get(const std::string& name, int64_t numel) {
    auto it = t.find(name);
    if (it == t.end()) { if (ok) c->err = "missing tensor: " + name; ok = false; return nullptr; }
    if (it->second.second != numel) {
      if (ok) c->err = "bad size for " + name + ": got " + std::to_string(it->second.second) + " want " + std::to_string(numel);
      ok = false; return nullptr;
    }
    return it->second.first;
  }
};

// fp32 -> tf32 (10-bit mantissa), round to nearest even, kept in an fp32 container
float round_tf32(float x) {
  uint32_t b;
  std::memcpy(&b, &x, 4);
  if ((b & 0x7f800000u) != 0x7f800000u) { b += 0xfffu + ((b >> 13) & 1u); b &= ~0x1fffu; }
  std::memcpy(&x, &b, 4);
  return x;
}

struct Bn { std::vector<float> scale, shift; };

Bn fold_bn(Getter& g, const std::string& p, int c, bool affine = true) {
  Bn r; r.scale.assign(c, 1.f); r.shift.assign(c, 0.f);
  const float* mean = g.get(p + ".running_mean", c);
  const float* var = g.get(p + ".running_var", c);
  const float* w = affine ? g.get(p + ".weight", c) : nullptr;
  const float* b = affine ? g.get(p + ".bias", c) : nullptr;
  if (!g.ok) return r;
  for (int i = 0; i < c; ++i) {
    double s = 1.0 / std::sqrt((double)var[i] + (double)kBnEps);
    if (affine) s *= (double)w[i];
    r.scale[i] = (float)s;
    r.shift[i] = (float)((affine ? (double)b[i] : 0.0) - (double)mean[i] * s);
  }
  return r;
}

// conv2d weight [co][ci][3][3] -> [co][(kh*3+kw)*ci_n + ci] scaled by BN scale[co]
void pack_conv3x3(std::vector<float>& out, int kstride, const float* w, const Bn& bn, int ci_n) {
  for (int co = 0; co < kFcmC; ++co)
    for (int ci = 0; ci < ci_n; ++ci)
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          out[(size_t)co * kstride + (kh * 3 + kw) * ci_n + ci] = w[((co * ci_n + ci) * 3 + kh) * 3 + kw] * bn.scale[co];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int load_ve(cbx_ctx* c, const TensorMap& t) {
  Getter g{t, c};
  Packer pk;
  VeWeights& W = c->ve;
  const int H = kVeHidden, G = kVeGates;
  for (int l = 0; l < 3; ++l) {
    const int in = l == 0 ? kVeMels : H;
    std::string s = std::to_string(l);
    const float* wih = g.get("lstm.weight_ih_l" + s, (int64_t)G * in);
    const float* whh = g.get("lstm.weight_hh_l" + s, (int64_t)G * H);
    const float* bih = g.get("lstm.bias_ih_l" + s, G);
    const float* bhh = g.get("lstm.bias_hh_l" + s, G);
    if (!g.ok) return CBX_ERR_ARG;
    pk.add(l == 0 ? &W.wih0 : &W.wih[l], std::vector<float>(wih, wih + (size_t)G * in));
    std::vector<float> tr((size_t)H * G), bias(G);
    for (int n = 0; n < G; ++n) {
      for (int k = 0; k < H; ++k) tr[(size_t)k * G + n] = whh[(size_t)n * H + k];
      bias[n] = bih[n] + bhh[n];
    }
    pk.add(&W.whhT[l], tr);
    pk.add(&W.bias[l], bias);
    // permuted copies for the tensor-core recurrence: row 128 j + 4 u + g  <-  gate g of unit 32 j + u
    std::vector<float> wih_p((size_t)G * in), whh_p((size_t)G * H), bias_p(G);
    for (int n = 0; n < G; ++n) {
      const int jj = n >> 7, uu = (n >> 2) & 31, gg = n & 3;
      const int src = gg * H + jj * 32 + uu;
      std::memcpy(&wih_p[(size_t)n * in], wih + (size_t)src * in, sizeof(float) * in);
      for (int k = 0; k < H; ++k) whh_p[(size_t)n * H + k] = round_tf32(whh[(size_t)src * H + k]);
      bias_p[n] = bias[src];
    }
    pk.add(&W.wih_p[l], wih_p);
    pk.add(&W.whh_p[l], whh_p);
    pk.add(&W.bias_p[l], bias_p);
  }
  const float* wp = g.get("proj.weight", (int64_t)kVeEmbed * H);
  const float* bp = g.get("proj.bias", kVeEmbed);
  if (!g.ok) return CBX_ERR_ARG;
  std::vector<float> wpT((size_t)H * kVeEmbed);
  for (int n = 0; n < kVeEmbed; ++n)
    for (int k = 0; k < H; ++k) wpT[(size_t)k * kVeEmbed + n] = wp[(size_t)n * H + k];
  pk.add(&W.wpT, wpT);
  pk.add(&W.bp, std::vector<float>(bp, bp + kVeEmbed));
  W.wih[0] = nullptr;
  int rc = pk.upload(c, &W.blob);
  if (rc) return rc;
  W.wih[0] = W.wih0;
  for (int l = 0; l < 3; ++l) {
    const int in = l == 0 ? kVeMels : H;
    W.tm_wih[l] = tc::make_map_2d(W.wih[l], G, in, in, 128, true);
    W.tm_wih_p[l] = tc::make_map_2d(W.wih_p[l], G, in, in, 128, true);
    W.tm_wih_p256[l] = tc::make_map_2d(W.wih_p[l], G, in, in, 256, true);
  }
  W.loaded = true;
  return CBX_OK;
}

// ------------------------------------------------------------------------------------------------
int load_xv(cbx_ctx* c, const TensorMap& t) {
  Getter g{t, c};
  Packer pk;
  XvWeights& W = c->xv;
  static const int kBlockLayers[3] = {12, 24, 16};

  {  // head.conv1 (1 -> 32) + bn1
    const float* w = g.get("head.conv1.weight", 32 * 9);
    Bn bn = fold_bn(g, "head.bn1", 32);
    if (!g.ok) return CBX_ERR_ARG;
    std::vector<float> o(32 * 9);
    for (int co = 0; co < 32; ++co)
      for (int k = 0; k < 9; ++k) o[co * 9 + k] = w[co * 9 + k] * bn.scale[co];
    pk.add(&W.conv1_w, o);
    pk.add(&W.conv1_b, bn.shift);
  }
  for (int layer = 0; layer < 2; ++layer)
    for (int blk = 0; blk < 2; ++blk) {
      std::string p = "head.layer" + std::to_string(layer + 1) + "." + std::to_string(blk);
      const float* w1 = g.get(p + ".conv1.weight", 32 * 32 * 9);
      Bn b1 = fold_bn(g, p + ".bn1", 32);
      const float* w2 = g.get(p + ".conv2.weight", 32 * 32 * 9);
      Bn b2 = fold_bn(g, p + ".bn2", 32);
      if (!g.ok) return CBX_ERR_ARG;
      std::vector<float> o1(32 * 288);
      pack_conv3x3(o1, 288, w1, b1, 32);
      pk.add(&W.res[layer][blk][0].w, o1);
      pk.add(&W.res[layer][blk][0].bias, b1.shift);
      W.res[layer][blk][0].K = 288;
      const int K2 = blk == 0 ? 320 : 288;     // block 0: 1x1 stride-2 shortcut conv appended as 32 more K columns
      std::vector<float> o2((size_t)32 * K2), bias2 = b2.shift;
      pack_conv3x3(o2, K2, w2, b2, 32);
      if (blk == 0) {
        const float* ws = g.get(p + ".shortcut.0.weight", 32 * 32);
        Bn bs = fold_bn(g, p + ".shortcut.1", 32);
        if (!g.ok) return CBX_ERR_ARG;
        for (int co = 0; co < 32; ++co) {
          for (int ci = 0; ci < 32; ++ci) o2[(size_t)co * K2 + 288 + ci] = ws[co * 32 + ci] * bs.scale[co];
          bias2[co] += bs.shift[co];
        }
      }
      pk.add(&W.res[layer][blk][1].w, o2);
      pk.add(&W.res[layer][blk][1].bias, bias2);
      W.res[layer][blk][1].K = K2;
    }
  {
    const float* w = g.get("head.conv2.weight", 32 * 32 * 9);
    Bn bn = fold_bn(g, "head.bn2", 32);
    if (!g.ok) return CBX_ERR_ARG;
    std::vector<float> o(32 * 288);
    pack_conv3x3(o, 288, w, bn, 32);
    pk.add(&W.head_conv2.w, o);
    pk.add(&W.head_conv2.bias, bn.shift);
    W.head_conv2.K = 288;
  }
  {  // xvector.tdnn: Conv1d(320->128,k5,s2,p2)+BN+ReLU.  Reference channel = c*10+f (xvector.py:125-126);
     // our FCM output row is [f][c], so k = tap*320 + f*32 + c.
    const float* w = g.get("xvector.tdnn.linear.weight", 128 * 320 * 5);
    Bn bn = fold_bn(g, "xvector.tdnn.nonlinear.batchnorm", 128);
    if (!g.ok) return CBX_ERR_ARG;
    std::vector<float> o((size_t)128 * 1600);
    for (int co = 0; co < 128; ++co)
      for (int ch = 0; ch < 32; ++ch)
        for (int f = 0; f < 10; ++f)
          for (int tap = 0; tap < 5; ++tap)
            o[(size_t)co * 1600 + tap * 320 + f * 32 + ch] = w[((size_t)co * 320 + ch * 10 + f) * 5 + tap] * bn.scale[co];
    pk.add(&W.tdnn.w, o);
    pk.add(&W.tdnn.bias, bn.shift);
    W.tdnn.K = 1600;
  }
  int li = 0, ch = kTdnnC;
  for (int b = 0; b < 3; ++b) {
    for (int i = 0; i < kBlockLayers[b]; ++i, ++li) {
      std::string p = "xvector.block" + std::to_string(b + 1) + ".tdnnd" + std::to_string(i + 1);
      DenseLayerW& L = W.dense[li];
      const int cin = ch + kGrowth * i;
      L.cin = cin;
      Bn bn1 = fold_bn(g, p + ".nonlinear1.batchnorm", cin);
      const float* w1 = g.get(p + ".linear1.weight", (int64_t)kBnC * cin);
      Bn bn2 = fold_bn(g, p + ".nonlinear2.batchnorm", kBnC);
      const float* wl = g.get(p + ".cam_layer.linear_local.weight", kGrowth * kBnC * 3);
      const float* wc1 = g.get(p + ".cam_layer.linear1.weight", kCamHid * kBnC);
      const float* bc1 = g.get(p + ".cam_layer.linear1.bias", kCamHid);
      const float* wc2 = g.get(p + ".cam_layer.linear2.weight", kGrowth * kCamHid);
      const float* bc2 = g.get(p + ".cam_layer.linear2.bias", kGrowth);
      if (!g.ok) return CBX_ERR_ARG;
      pk.add(&L.a1, bn1.scale);
      pk.add(&L.b1, bn1.shift);
      std::vector<float> o1((size_t)kBnC * cin);
      for (int n = 0; n < kBnC; ++n)
        for (int k = 0; k < cin; ++k) o1[(size_t)n * cin + k] = w1[(size_t)n * cin + k] * bn2.scale[n];
      pk.add(&L.w1, o1);
      pk.add(&L.w1h, to_bf16_words(o1.data(), o1.size()));
      pk.add(&L.t2, bn2.shift);
      std::vector<float> ol((size_t)kGrowth * 3 * kBnC);
      for (int n = 0; n < kGrowth; ++n)
        for (int cc = 0; cc < kBnC; ++cc)
          for (int tap = 0; tap < 3; ++tap) ol[(size_t)n * 384 + tap * kBnC + cc] = wl[((size_t)n * kBnC + cc) * 3 + tap];
      pk.add(&L.wl, ol);
      pk.add(&L.wlh, to_bf16_words(ol.data(), ol.size()));      // bf16 copy (bf16 mode: local convolution on bf16 operands)
      pk.add(&L.wc1, std::vector<float>(wc1, wc1 + kCamHid * kBnC));
      {   // transposed copies: lanes along the output index read consecutive addresses in the gate kernel
        std::vector<float> t1((size_t)kBnC * kCamHid), t2((size_t)kCamHid * kGrowth);
        for (int o = 0; o < kCamHid; ++o) for (int k = 0; k < kBnC; ++k) t1[(size_t)k * kCamHid + o] = wc1[(size_t)o * kBnC + k];
        for (int o = 0; o < kGrowth; ++o) for (int k = 0; k < kCamHid; ++k) t2[(size_t)k * kGrowth + o] = wc2[(size_t)o * kCamHid + k];
        pk.add(&L.wc1T, t1);
        pk.add(&L.wc2T, t2);
      }
      pk.add(&L.bc1, std::vector<float>(bc1, bc1 + kCamHid));
      pk.add(&L.wc2, std::vector<float>(wc2, wc2 + kGrowth * kCamHid));
      pk.add(&L.bc2, std::vector<float>(bc2, bc2 + kGrowth));
    }
    ch += kGrowth * kBlockLayers[b];
    std::string p = "xvector.transit" + std::to_string(b + 1);
    TransitW& T = W.transit[b];
    T.cin = ch; T.cout = ch / 2;
    Bn bn = fold_bn(g, p + ".nonlinear.batchnorm", ch);
    const float* w = g.get(p + ".linear.weight", (int64_t)(ch / 2) * ch);
    if (!g.ok) return CBX_ERR_ARG;
    pk.add(&T.a, bn.scale);
    pk.add(&T.b, bn.shift);
    pk.add(&T.w, std::vector<float>(w, w + (size_t)(ch / 2) * ch));
    pk.add(&T.wh, to_bf16_words(w, (size_t)(ch / 2) * ch));
    ch /= 2;
  }
  {
    Bn bn = fold_bn(g, "xvector.out_nonlinear.batchnorm", kStatsC);
    const float* w = g.get("xvector.dense.linear.weight", (int64_t)kXvEmbed * 2 * kStatsC);
    Bn bf = fold_bn(g, "xvector.dense.nonlinear.batchnorm", kXvEmbed, /*affine=*/false);
    if (!g.ok) return CBX_ERR_ARG;
    pk.add(&W.out_a, bn.scale);
    pk.add(&W.out_b, bn.shift);
    std::vector<float> o((size_t)kXvEmbed * 2 * kStatsC);
    for (int n = 0; n < kXvEmbed; ++n)
      for (int k = 0; k < 2 * kStatsC; ++k) o[(size_t)n * 2 * kStatsC + k] = w[(size_t)n * 2 * kStatsC + k] * bf.scale[n];
    pk.add(&W.fin_w, o);
    pk.add(&W.fin_b, bf.shift);
  }
  int rc = pk.upload(c, &W.blob);
  if (rc) return rc;
  W.tm_tdnn = tc::make_map_2d(W.tdnn.w, kTdnnC, W.tdnn.K, W.tdnn.K, 128, true);
  for (int i = 0; i < 52; ++i) {
    W.tm_w1[i] = tc::make_map_2d(W.dense[i].w1, kBnC, W.dense[i].cin, W.dense[i].cin, 128, true);
    W.tm_w1h[i] = tc::make_map_2d_bf16(W.dense[i].w1h, kBnC, W.dense[i].cin, W.dense[i].cin, 128);
    W.tm_wl[i] = tc::make_map_2d(W.dense[i].wl, kGrowth, 3 * kBnC, 3 * kBnC, 32, true);
    W.tm_wlh[i] = tc::make_map_2d_bf16(W.dense[i].wlh, kGrowth, 3 * kBnC, 3 * kBnC, 32);
  }
  for (int l = 0; l < 2; ++l)
    for (int b = 0; b < 2; ++b)
      for (int k = 0; k < 2; ++k) W.tm_res[l][b][k] = tc::make_map_2d(W.res[l][b][k].w, kFcmC, W.res[l][b][k].K, W.res[l][b][k].K, 32, true);
  W.tm_head2 = tc::make_map_2d(W.head_conv2.w, kFcmC, 288, 288, 32, true);
  for (int b = 0; b < 3; ++b) W.tm_tr[b] = tc::make_map_2d(W.transit[b].w, W.transit[b].cout, W.transit[b].cin, W.transit[b].cin, 128, true);
  for (int b = 0; b < 3; ++b) W.tm_tr256[b] = tc::make_map_2d(W.transit[b].w, W.transit[b].cout, W.transit[b].cin, W.transit[b].cin, 256, true);
  for (int b = 0; b < 3; ++b) W.tm_trh[b] = tc::make_map_2d_bf16(W.transit[b].wh, W.transit[b].cout, W.transit[b].cin, W.transit[b].cin, 128);
  W.loaded = true;
  return CBX_OK;
}

// ------------------------------------------------------------------------------------------------
// Front-end tables (float64 on the host, rounded once to fp32).
namespace {

double slaney_mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : m * f_sp;
}
double slaney_hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}

// Tensor-core front-end tables.  Rows 2b, 2b+1 = re, im of DFT bin (first_bin + b), b < nbins - 1 (the last bin pair is
// zero padding); each value v is stored as hi = tf32(v), lo = tf32(v - hi) for the 3xTF32 product (SURVEY.md 8d hazard 3).
void split_dft(Packer& pk, const std::vector<double>& dd, int K, int first_bin, int nbins, const float** hi_slot, const float** lo_slot) {
  std::vector<float> hi((size_t)2 * nbins * K, 0.f), lo((size_t)2 * nbins * K, 0.f);
  for (int b = 0; b < nbins - 1; ++b)
    for (int part = 0; part < 2; ++part)
      for (int n = 0; n < K; ++n) {
        const double v = dd[(size_t)(2 * (first_bin + b) + part) * K + n];
        const float h = round_tf32((float)v);
        hi[(size_t)(2 * b + part) * K + n] = h;
        lo[(size_t)(2 * b + part) * K + n] = round_tf32((float)(v - (double)h));
      }
  pk.add(hi_slot, hi);
  pk.add(lo_slot, lo);
}
// Per DFT bin, for the epilogue's running mel accumulation: triangular banks are 2-sparse per bin and the two filters a bin
// feeds are adjacent, so a thread keeps TWO accumulators in registers -- filter `cur` (A) and `cur + 1` (B) -- and walks the bins
// in order.  Entry b = {wA, wB, bits(pre)}: first write out and retire `pre` finished filters (A -> out[cur], A = B, B = 0, ++cur),
// then A += wA * power, B += wB * power.  Every filter receives its bins' contributions in ascending bin order starting from 0,
// exactly like a per-filter sum.  (The previous table {w0, w1, m0, m1} drove read-modify-write accumulators in shared memory: a
// dependent LDS -> FFMA -> STS chain per bin.)
void bin_table(Packer& pk, const std::vector<float>& bank, int nmel, int nbank_bins, int first_bin, int nbins, const float** slot) {
  std::vector<float> t((size_t)4 * nbins, 0.f);
  int cur = 0;
  for (int b = 0; b < nbins; ++b) {
    int m[2] = {0, 0}; float w[2] = {0.f, 0.f}; int cnt = 0;
    const int k = first_bin + b;
    if (b < nbins - 1 && k < nbank_bins)
      for (int mm = 0; mm < nmel; ++mm) {
        const float v = bank[(size_t)mm * nbank_bins + k];
        if (v != 0.f) {
          if (cnt < 2) { m[cnt] = mm; w[cnt] = v; }
          ++cnt;
        }
      }
    int pre = 0;
    float wA = 0.f, wB = 0.f;
    bool ok = cnt <= 2 && (cnt < 2 || m[1] == m[0] + 1) && (cnt == 0 || m[0] >= cur);
    if (ok && cnt > 0) {
      if (cnt == 2 || m[0] > cur + 1) { pre = m[0] - cur; cur = m[0]; }      // a lone filter one ahead goes into B without retiring A
      if (cnt == 2) { wA = w[0]; wB = w[1]; }
      else if (m[0] == cur) wA = w[0];
      else wB = w[0];
    }
    if (!ok) { fprintf(stderr, "libcbx: mel bank is not a chain of adjacent triangles at bin %d (cnt %d)\n", k, cnt); }
    t[4 * b] = wA; t[4 * b + 1] = wB;
    std::memcpy(&t[4 * b + 2], &pre, 4);
  }
  pk.add(slot, t);
}

}  // namespace

int build_frontend_tables(cbx_ctx* c) {
  Packer pk;
  FrontendTables& F = c->ft;
  const double PI = 3.14159265358979323846;
  {  // VoiceEncoder: periodic Hann folded into the 400-point DFT (melspec.py:57-64); rows 2k = re, 2k+1 = im
    std::vector<float> d((size_t)kVeSpecN * kVeNfft);
    std::vector<double> dd((size_t)kVeSpecN * kVeNfft);
    for (int k = 0; k < kVeBins; ++k)
      for (int n = 0; n < kVeNfft; ++n) {
        double w = 0.5 - 0.5 * std::cos(2.0 * PI * n / kVeNfft);
        double ang = 2.0 * PI * (double)((k * n) % kVeNfft) / kVeNfft;
        dd[(size_t)(2 * k) * kVeNfft + n] = w * std::cos(ang);
        dd[(size_t)(2 * k + 1) * kVeNfft + n] = -w * std::sin(ang);
      }
    for (size_t i = 0; i < d.size(); ++i) d[i] = (float)dd[i];
    pk.add(&F.ve_dft, d);
    split_dft(pk, dd, kVeNfft, 1, kVeTcBins, &F.ve_dft_hi, &F.ve_dft_lo);
    {
      // Even / odd form.  The periodic Hann window is symmetric about j = 200 (w[j] = w[400 - j], w[0] = 0, w[200] = 1), so
      //   re X[k] = sum_{j=1..199} w[j] cos(2 pi k j / 400) (s[j] + s[400-j])  +  (-1)^k s[200]
      //   im X[k] = - sum_{j=1..199} w[j] sin(2 pi k j / 400) (s[j] - s[400-j])
      // : two K = 200 products instead of one K = 400 product -- half the MMAs, half the DFT-matrix bytes per tile.
      // Column c of the matrix is j = c + 1 (c = 0..199; c = 199 is j = 200: e[200] = s[200], o[200] = 0), padded to 224.
      constexpr int EO_ROWS = 208, EO_K = 224;
      std::vector<float> hi((size_t)2 * EO_ROWS * EO_K, 0.f), lo((size_t)2 * EO_ROWS * EO_K, 0.f);
      for (int b = 0; b < kVeTcBins - 1; ++b) {
        const int k = 1 + b;
        for (int c = 0; c < 200; ++c) {
          const int j = c + 1;
          const double wj = 0.5 - 0.5 * std::cos(2.0 * PI * j / kVeNfft);
          const double ang = 2.0 * PI * (double)((k * j) % kVeNfft) / kVeNfft;
          const double vr = j < 200 ? wj * std::cos(ang) : ((k & 1) ? -1.0 : 1.0);
          const double vi = j < 200 ? -wj * std::sin(ang) : 0.0;
          const float hr = round_tf32((float)vr), hi_i = round_tf32((float)vi);
          hi[(size_t)b * EO_K + c] = hr;                       lo[(size_t)b * EO_K + c] = round_tf32((float)(vr - (double)hr));
          hi[(size_t)(EO_ROWS + b) * EO_K + c] = hi_i;          lo[(size_t)(EO_ROWS + b) * EO_K + c] = round_tf32((float)(vi - (double)hi_i));
        }
      }
      pk.add(&F.ve_eo_hi, hi);
      pk.add(&F.ve_eo_lo, lo);
    }
    // librosa.filters.mel(sr=16000,n_fft=400,n_mels,fmin=0,fmax=8000): Slaney scale + area norm (melspec.py:11-16)
    auto slaney_bank = [](int nmels) {
      std::vector<double> edges(nmels + 2);
      const double m_lo = slaney_hz_to_mel(0.0), m_hi = slaney_hz_to_mel(8000.0);
      for (int i = 0; i < nmels + 2; ++i) edges[i] = slaney_mel_to_hz(m_lo + (m_hi - m_lo) * i / (nmels + 1));
      std::vector<float> mel((size_t)nmels * kVeBins);
      for (int m = 0; m < nmels; ++m)
        for (int k = 0; k < kVeBins; ++k) {
          double f = (double)kSR / 2.0 * k / (kVeBins - 1);
          double up = (f - edges[m]) / (edges[m + 1] - edges[m]);
          double dn = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
          float tri = (float)std::fmax(0.0, std::fmin(up, dn));
          mel[(size_t)m * kVeBins + k] = (float)((double)tri * (2.0 / (edges[m + 2] - edges[m])));
        }
      return mel;
    };
    std::vector<float> mel = slaney_bank(kVeMels);
    pk.add(&F.ve_mel, mel);
    bin_table(pk, mel, kVeMels, kVeBins, 1, kVeTcBins, &F.ve_bins);
    // S3Tokenizer: same STFT, 128 mels (s3tokenizer.py:39-47)
    bin_table(pk, slaney_bank(kS3Mels), kS3Mels, kVeBins, 1, kVeTcBins, &F.s3_bins);
  }
  {  // Kaldi: DC removal, pre-emphasis 0.97 (replicate-left), Povey window and zero-pad to 512 folded into
     // one [514][400] matrix (torchaudio kaldi.py:183-211).
    std::vector<double> pov(kKWin);
    for (int n = 0; n < kKWin; ++n) pov[n] = std::pow(0.5 - 0.5 * std::cos(2.0 * PI * n / (kKWin - 1)), 0.85);
    std::vector<float> d((size_t)kKSpecN * kKWin);
    std::vector<double> dd((size_t)kKSpecN * kKWin);
    std::vector<double> y(kKWin + 1), z(kKWin);
    for (int col = 0; col < kKSpecN; ++col) {
      const int k = col >> 1;
      for (int n = 0; n < kKWin; ++n) {
        double ang = 2.0 * PI * (double)((k * n) % kKPad) / kKPad;
        y[n] = pov[n] * ((col & 1) ? -std::sin(ang) : std::cos(ang));
      }
      y[kKWin] = 0.0;
      double zbar = 0.0;
      for (int j = 0; j < kKWin; ++j) {
        z[j] = y[j] - 0.97 * y[j + 1];
        if (j == 0) z[j] -= 0.97 * y[0];
        zbar += z[j];
      }
      zbar /= kKWin;
      for (int j = 0; j < kKWin; ++j) { dd[(size_t)col * kKWin + j] = z[j] - zbar; d[(size_t)col * kKWin + j] = (float)(z[j] - zbar); }
    }
    pk.add(&F.k_dft, d);
    split_dft(pk, dd, kKWin, 1, kKTcBins, &F.k_dft_hi, &F.k_dft_lo);
    // 80 HTK-mel triangles 20 Hz..8 kHz on the 512-point grid, slopes in mel, Nyquist bin weight 0 (kaldi.py:436-511)
    auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
    const double lo = mel(20.0), hi = mel(8000.0), delta = (hi - lo) / (kKMels + 1);
    std::vector<float> bank((size_t)kKMels * kKBins, 0.f);
    for (int m = 0; m < kKMels; ++m) {
      // torchaudio evaluates this in float32; keep float arithmetic for the slopes to stay close to it
      float left = (float)(lo + m * delta), center = (float)(lo + (m + 1) * delta), right = (float)(lo + (m + 2) * delta);
      for (int k = 0; k < kKBins - 1; ++k) {
        float mk = 1127.0f * std::log(1.0f + ((float)kSR / kKPad * (float)k) / 700.0f);
        float up = (mk - left) / (center - left), dn = (right - mk) / (right - center);
        bank[(size_t)m * kKBins + k] = std::fmax(0.f, std::fmin(up, dn));
      }
    }
    pk.add(&F.k_mel, bank);
    bin_table(pk, bank, kKMels, kKBins, 1, kKTcBins, &F.k_bins);
  }
  int rc = pk.upload(c, &F.blob);
  if (rc) return rc;
  // TMA maps of the split DFT matrices (B operands of the front-end GEMM): boxes of 32 k x (256 | rest) rows
  F.tm_ve_hi[0] = tc::make_map_2d(F.ve_dft_hi, 2 * kVeTcBins, kVeNfft, kVeNfft, 256, false);
  F.tm_ve_lo[0] = tc::make_map_2d(F.ve_dft_lo, 2 * kVeTcBins, kVeNfft, kVeNfft, 256, false);
  F.tm_ve_hi[1] = tc::make_map_2d(F.ve_dft_hi, 2 * kVeTcBins, kVeNfft, kVeNfft, 2 * kVeTcBins - 256, false);
  F.tm_ve_lo[1] = tc::make_map_2d(F.ve_dft_lo, 2 * kVeTcBins, kVeNfft, kVeNfft, 2 * kVeTcBins - 256, false);
  F.tm_ve_eo_hi = tc::make_map_2d(F.ve_eo_hi, 416, 224, 224, 208, false);
  F.tm_ve_eo_lo = tc::make_map_2d(F.ve_eo_lo, 416, 224, 224, 208, false);
  F.tm_k_hi = tc::make_map_2d(F.k_dft_hi, 2 * kKTcBins, kKWin, kKWin, 256, false);
  F.tm_k_lo = tc::make_map_2d(F.k_dft_lo, 2 * kKTcBins, kKWin, kKWin, 256, false);
  return CBX_OK;
}

}  // namespace cbx
