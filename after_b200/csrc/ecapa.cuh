// ECAPATDNN.forward -- the timbre encoder (after/diffusion/networks/ecapa_encoder.py:458-624): TDNN (reflect-padded
// conv -> ReLU -> eval BatchNorm), two SE-Res2Net blocks, multi-layer feature aggregation, attentive statistics
// pooling, BatchNorm, 1x1 conv -> (B, out_dim).  It runs once per chunk (2.5 GFLOP per stream against 1.1 TFLOP of
// sampling), so everything here is plain fp32 FFMA: one register-tiled conv kernel that understands channel slices,
// a second summed input (Res2Net's x_i + y_{i-1}), reflect padding and the ReLU/BN/tanh epilogue, plus a handful of
// small reduction kernels.  Activations are frame-major (B, T, C) like the rest of the library.
#pragma once
#include <cmath>
#include "codec_kernels.cuh"
#include "context.cuh"

namespace after {

struct EcapaConvArgs {
  const float* in1 = nullptr; int ld1 = 0, off1 = 0;  // input channels [off1, off1 + Cin) of rows of ld1 floats
  const float* in2 = nullptr; int ld2 = 0, off2 = 0;  // optional second input, added element-wise
  const float* w = nullptr;                           // [Cout][k * Cin], tap-major
  const float* bias = nullptr; int bias_ld = 0;       // bias[b * bias_ld + co]  (bias_ld = 0: shared)
  int Cin = 0, Cout = 0, k = 1, dilation = 1, T = 0;
  int relu = 0;
  const float* post_s = nullptr; const float* post_b = nullptr;  // y = y * s[co] + b[co]  (folded BatchNorm)
  int tanh_out = 0;
  float* out = nullptr; int ldo = 0, offo = 0;
};

// 64 frames x 64 output channels per 256-thread block (lane = frame, warp = 8 output channels), K staged 16 at a time.
// Reflect padding: frame index -1 -> 1, T -> T-2 (Conv1dSamePaddingReflect, ecapa_encoder.py:12-82).
__global__ void __launch_bounds__(256)
ecapa_conv_kernel(EcapaConvArgs a) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Ws[16][64 + 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, t0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int K = a.k * a.Cin;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const bool w_ok = (n0 + lrow) < a.Cout;
  const float* Wp = a.w + (size_t)(n0 + lrow) * K + lk;
  const int half = (a.k - 1) / 2;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int tap = 0; tap < a.k; ++tap) {
    int tt = t0 + lrow + (tap - half) * a.dilation;
    const bool in_tile = (t0 + lrow) < a.T;
    if (tt < 0) tt = -tt;
    if (tt >= a.T) tt = 2 * (a.T - 1) - tt;
    tt = max(0, min(tt, a.T - 1));
    const float* p1 = a.in1 + ((size_t)b * a.T + tt) * a.ld1 + a.off1 + lk;
    const float* p2 = a.in2 ? a.in2 + ((size_t)b * a.T + tt) * a.ld2 + a.off2 + lk : nullptr;
    for (int c0 = 0; c0 < a.Cin; c0 += 16) {
      float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_tile) {
        ra = *reinterpret_cast<const float4*>(p1 + c0);
        if (p2) {
          const float4 r2 = *reinterpret_cast<const float4*>(p2 + c0);
          ra.x += r2.x; ra.y += r2.y; ra.z += r2.z; ra.w += r2.w;
        }
      }
      const float4 rw = w_ok ? *reinterpret_cast<const float4*>(Wp + (size_t)tap * a.Cin + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
      As[lk + 0][lrow] = ra.x; As[lk + 1][lrow] = ra.y; As[lk + 2][lrow] = ra.z; As[lk + 3][lrow] = ra.w;
      Ws[lk + 0][lrow] = rw.x; Ws[lk + 1][lrow] = rw.y; Ws[lk + 2][lrow] = rw.z; Ws[lk + 3][lrow] = rw.w;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float a0 = As[kk][lane], a1 = As[kk][lane + 32];
        const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][warp * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Ws[kk][warp * 8 + 4]);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fmaf(a0, bb[j], acc[0][j]);
          acc[1][j] = fmaf(a1, bb[j], acc[1][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int t = t0 + lane + 32 * i;
    if (t >= a.T) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = n0 + warp * 8 + j;
      if (co >= a.Cout) continue;
      float y = acc[i][j] + (a.bias ? a.bias[(size_t)b * a.bias_ld + co] : 0.f);
      if (a.relu) y = fmaxf(y, 0.f);
      if (a.post_s) y = fmaf(y, a.post_s[co], a.post_b[co]);
      if (a.tanh_out) y = tanhf(y);
      a.out[((size_t)b * a.T + t) * a.ldo + a.offo + co] = y;
    }
  }
}

// out[b, co] = act(bias[co] + sum_ci W[co, ci] in[b, ci]);  act: 0 none, 1 ReLU, 2 sigmoid, 3 SiLU.  One warp per output.
__global__ void vec_linear_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                  float* __restrict__ out, int B, int Cin, int Cout, int act) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= B * Cout) return;
  const int lane = threadIdx.x & 31;
  const int b = o / Cout, co = o - b * Cout;
  float s = 0.f;
  for (int c = lane; c < Cin; c += 32) s = fmaf(W[(size_t)co * Cin + c], in[(size_t)b * Cin + c], s);
  s = warp_sum(s);
  if (lane == 0) {
    s += bias ? bias[co] : 0.f;
    if (act == 1) s = fmaxf(s, 0.f);
    else if (act == 2) s = 1.0f / (1.0f + expf(-s));
    else if (act == 3) s = s / (1.0f + expf(-s));
    out[o] = s;
  }
}

// mean over frames (SEBlock squeeze, ecapa_encoder.py:268) and, optionally, the population std used by the global
// context of the pooling layer (sqrt(clamp(mean((x - mean)^2), eps)), ecapa_encoder.py:402-412).
__global__ void frame_stats_kernel(const float* __restrict__ x, float* __restrict__ mean_out, float* __restrict__ std_out,
                                   int T, int C, int out_ld, int std_off) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  const float* p = x + (size_t)b * T * C + c;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += p[(size_t)t * C];
  const float m = s / (float)T;
  mean_out[(size_t)b * out_ld + c] = m;
  if (std_out) {
    float q = 0.f;
    for (int t = 0; t < T; ++t) { const float d = p[(size_t)t * C] - m; q = fmaf(d, d, q); }
    std_out[(size_t)b * out_ld + std_off + c] = sqrtf(fmaxf(q / (float)T, 1e-12f));
  }
}

// SE excitation + residual: out[b,t,c] = s[b,c] * y[b,t,c] + res[b,t,c]   (ecapa_encoder.py:271-272, 360-362)
__global__ void se_apply_kernel(const float* __restrict__ y, const float* __restrict__ s, const float* __restrict__ res,
                                int res_ld, int res_off, float* __restrict__ out, int ldo, int offo, int B, int T, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T * C) return;
  const int c = (int)(i % C);
  const size_t bt = i / C;
  const int b = (int)(bt / T);
  out[bt * ldo + offo + c] = fmaf(s[(size_t)b * C + c], y[i], res[bt * res_ld + res_off + c]);
}

// Attentive statistics (ecapa_encoder.py:414-455): w = softmax_t(logits[b, :, c]);  mean = sum_t w x ;
// std = sqrt(clamp(sum_t w (x - mean)^2, eps)).  out[b] = [mean (C) | std (C)].
__global__ void attentive_pool_kernel(const float* __restrict__ logits, const float* __restrict__ x, float* __restrict__ out,
                                      int T, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  const float* lp = logits + (size_t)b * T * C + c;
  const float* xp = x + (size_t)b * T * C + c;
  float m = -INFINITY;
  for (int t = 0; t < T; ++t) m = fmaxf(m, lp[(size_t)t * C]);
  float den = 0.f, num = 0.f;
  for (int t = 0; t < T; ++t) {
    const float e = expf(lp[(size_t)t * C] - m);
    den += e;
    num = fmaf(e, xp[(size_t)t * C], num);
  }
  const float mean = num / den;
  float var = 0.f;
  for (int t = 0; t < T; ++t) {
    const float e = expf(lp[(size_t)t * C] - m);
    const float d = xp[(size_t)t * C] - mean;
    var = fmaf(e, d * d, var);
  }
  out[(size_t)b * 2 * C + c] = mean;
  out[(size_t)b * 2 * C + C + c] = sqrtf(fmaxf(var / den, 1e-12f));
}

struct TimbreEncoder {
  after_config cfg{};
  Arena* arena = nullptr;
  const TensorMap* sd = nullptr;
  int nb = 0, maxB = 0, maxT = 0, scale = 8, se_c = 128, att_c = 128, out_dim = 6;
  std::vector<int> ch, ks, dil;

  struct Tdnn { float *w, *b, *ps, *pb; int cin, cout, k, d; };
  struct SeRes2 {
    Tdnn tdnn1, tdnn2;
    std::vector<Tdnn> sub;
    float *se_w1, *se_b1, *se_w2, *se_b2;
    bool has_shortcut = false;
    float *sc_w = nullptr, *sc_b = nullptr;
    int cin, cout;
  };
  Tdnn first, mfa, asp_tdnn;
  std::vector<SeRes2> blocks;
  float *asp_wx = nullptr, *asp_wctx = nullptr;  // asp.tdnn weight split: frames part [att, C], context part [att, 2C]
  float *asp_conv_w = nullptr, *asp_conv_b = nullptr;
  float *fc_w = nullptr, *fc_b = nullptr;        // asp_bn folded in
  // workspace
  float *zf = nullptr, *x0 = nullptr, *y = nullptr, *y2 = nullptr, *res = nullptr, *feats = nullptr, *xm = nullptr, *att = nullptr,
        *logits = nullptr;
  float *vec_a = nullptr, *vec_b = nullptr, *vec_c = nullptr, *ctx = nullptr, *pooled = nullptr;
  int feat_c = 0, cmax = 0;

  const HostTensor& get(const std::string& key) const {
    auto it = sd->find(key);
    if (it == sd->end()) throw Error(AFTER_EMISSING, "missing tensor '" + key + "'");
    return it->second;
  }
  // (Cout, Cin, k) -> [Cout][k*Cin]
  float* conv_weight(const std::string& key, int cout, int cin, int k) {
    const HostTensor& w = get(key);
    AFTER_REQUIRE(w.shape.size() == 3 && w.shape[0] == cout && w.shape[1] == cin && w.shape[2] == k, AFTER_ESHAPE,
                  "tensor '" + key + "' has an unexpected shape");
    std::vector<float> m((size_t)cout * k * cin);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int kk = 0; kk < k; ++kk) m[((size_t)o * k + kk) * cin + c] = w.data[((size_t)o * cin + c) * k + kk];
    return arena->upload(m);
  }
  float* vec(const std::string& key, int n) {
    const HostTensor& t = get(key);
    AFTER_REQUIRE(t.numel() == n, AFTER_ESHAPE, "tensor '" + key + "' has an unexpected shape");
    return arena->upload(t.data);
  }
  void bn_fold(const std::string& prefix, int C, std::vector<float>& s, std::vector<float>& b) {
    const HostTensor &w = get(prefix + ".weight"), &bi = get(prefix + ".bias"), &rm = get(prefix + ".running_mean"),
                     &rv = get(prefix + ".running_var");
    AFTER_REQUIRE(w.numel() == C && bi.numel() == C && rm.numel() == C && rv.numel() == C, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight' has an unexpected shape");
    s.resize(C); b.resize(C);
    for (int c = 0; c < C; ++c) {
      s[c] = w.data[c] / std::sqrt(rv.data[c] + 1e-5f);
      b[c] = bi.data[c] - rm.data[c] * s[c];
    }
  }
  void make_tdnn(Tdnn& t, const std::string& prefix, int cin, int cout, int k, int d) {
    AFTER_REQUIRE(cin % 16 == 0 && k % 2 == 1, AFTER_EINVAL, "timbre encoder: channels must be multiples of 16, kernels odd");
    t.cin = cin; t.cout = cout; t.k = k; t.d = d;
    t.w = conv_weight(prefix + ".conv.conv.weight", cout, cin, k);
    t.b = vec(prefix + ".conv.conv.bias", cout);
    std::vector<float> s, b;
    bn_fold(prefix + ".norm", cout, s, b);
    t.ps = arena->upload(s);
    t.pb = arena->upload(b);
  }

  void finalize(const after_config& c, const TensorMap& tensors, Arena* ar) {
    cfg = c; arena = ar; sd = &tensors;
    nb = c.te_n_blocks;
    AFTER_REQUIRE(nb >= 3 && nb <= AFTER_MAX_STAGES, AFTER_EINVAL, "timbre encoder needs 3..8 channel stages");
    AFTER_REQUIRE(c.te_global_context == 1, AFTER_EINVAL, "timbre encoder: only global_context = True is supported (every shipped config)");
    ch.assign(c.te_channels, c.te_channels + nb);
    ks.assign(c.te_kernel_sizes, c.te_kernel_sizes + nb);
    dil.assign(c.te_dilations, c.te_dilations + nb);
    scale = c.te_res2net_scale; se_c = c.te_se_channels; att_c = c.te_attention_channels; out_dim = c.te_out_dim;
    maxB = c.max_batch; maxT = c.seq_len;
    make_tdnn(first, "blocks.0", c.te_in_size, ch[0], ks[0], dil[0]);
    blocks.resize(nb - 2);
    feat_c = 0;
    cmax = std::max(c.te_in_size, ch[0]);
    for (int i = 1; i < nb - 1; ++i) {
      SeRes2& s = blocks[i - 1];
      const std::string p = "blocks." + std::to_string(i);
      s.cin = ch[i - 1]; s.cout = ch[i];
      AFTER_REQUIRE(s.cout % scale == 0 && (s.cout / scale) % 16 == 0, AFTER_EINVAL, "timbre encoder: channels / res2net_scale must be a multiple of 16");
      make_tdnn(s.tdnn1, p + ".tdnn1", s.cin, s.cout, 1, 1);
      s.sub.resize(scale - 1);
      for (int j = 0; j < scale - 1; ++j)
        make_tdnn(s.sub[j], p + ".res2net_block.blocks." + std::to_string(j), s.cout / scale, s.cout / scale, ks[i], dil[i]);
      make_tdnn(s.tdnn2, p + ".tdnn2", s.cout, s.cout, 1, 1);
      s.se_w1 = conv_weight(p + ".se_block.conv1.conv.weight", se_c, s.cout, 1);
      s.se_b1 = vec(p + ".se_block.conv1.conv.bias", se_c);
      s.se_w2 = conv_weight(p + ".se_block.conv2.conv.weight", s.cout, se_c, 1);
      s.se_b2 = vec(p + ".se_block.conv2.conv.bias", s.cout);
      s.has_shortcut = s.cin != s.cout;
      if (s.has_shortcut) {
        s.sc_w = conv_weight(p + ".shortcut.conv.weight", s.cout, s.cin, 1);
        s.sc_b = vec(p + ".shortcut.conv.bias", s.cout);
      }
      feat_c += s.cout;
      cmax = std::max(cmax, s.cout);
    }
    const int CL = ch[nb - 1];
    AFTER_REQUIRE(feat_c == CL, AFTER_EINVAL, "timbre encoder: sum of the SE-Res2Net widths must equal the last channel count");
    make_tdnn(mfa, "mfa", CL, CL, ks[nb - 1], dil[nb - 1]);
    {  // asp.tdnn on cat([x, mean, std]): split the 1x1 weight into the per-frame part and the per-stream context part
      const HostTensor& w = get("asp.tdnn.conv.conv.weight");
      AFTER_REQUIRE(w.shape.size() == 3 && w.shape[0] == att_c && w.shape[1] == 3 * CL && w.shape[2] == 1, AFTER_ESHAPE,
                    "tensor 'asp.tdnn.conv.conv.weight' has an unexpected shape");
      std::vector<float> wx((size_t)att_c * CL), wc((size_t)att_c * 2 * CL);
      for (int o = 0; o < att_c; ++o) {
        std::copy(w.data.begin() + (size_t)o * 3 * CL, w.data.begin() + (size_t)o * 3 * CL + CL, wx.begin() + (size_t)o * CL);
        std::copy(w.data.begin() + (size_t)o * 3 * CL + CL, w.data.begin() + (size_t)(o + 1) * 3 * CL, wc.begin() + (size_t)o * 2 * CL);
      }
      asp_wx = arena->upload(wx);
      asp_wctx = arena->upload(wc);
      asp_tdnn.cin = CL; asp_tdnn.cout = att_c; asp_tdnn.k = 1; asp_tdnn.d = 1;
      asp_tdnn.w = asp_wx;
      asp_tdnn.b = vec("asp.tdnn.conv.conv.bias", att_c);
      std::vector<float> s, b;
      bn_fold("asp.tdnn.norm", att_c, s, b);
      asp_tdnn.ps = arena->upload(s);
      asp_tdnn.pb = arena->upload(b);
    }
    asp_conv_w = conv_weight("asp.conv.conv.weight", CL, att_c, 1);
    asp_conv_b = vec("asp.conv.conv.bias", CL);
    {  // fc(asp_bn(v)) = (W diag(s)) v + (W b + c)
      std::vector<float> s, b;
      bn_fold("asp_bn", 2 * CL, s, b);
      const HostTensor& w = get("fc.conv.weight");
      const HostTensor& fb = get("fc.conv.bias");
      AFTER_REQUIRE(w.shape.size() == 3 && w.shape[0] == out_dim && w.shape[1] == 2 * CL && w.shape[2] == 1 && fb.numel() == out_dim,
                    AFTER_ESHAPE, "tensor 'fc.conv.weight' has an unexpected shape");
      std::vector<float> fw((size_t)out_dim * 2 * CL), fbias(out_dim);
      for (int o = 0; o < out_dim; ++o) {
        double acc = fb.data[o];
        for (int cc = 0; cc < 2 * CL; ++cc) {
          fw[(size_t)o * 2 * CL + cc] = w.data[(size_t)o * 2 * CL + cc] * s[cc];
          acc += (double)w.data[(size_t)o * 2 * CL + cc] * b[cc];
        }
        fbias[o] = (float)acc;
      }
      fc_w = arena->upload(fw);
      fc_b = arena->upload(fbias);
    }
    cmax = std::max(cmax, CL);
    const size_t rows = (size_t)maxB * maxT;
    zf = arena->alloc<float>(rows * c.te_in_size);
    x0 = arena->alloc<float>(rows * cmax);
    y = arena->alloc<float>(rows * cmax);
    y2 = arena->alloc<float>(rows * cmax);
    res = arena->alloc<float>(rows * cmax);
    feats = arena->alloc<float>(rows * CL);
    xm = arena->alloc<float>(rows * CL);
    att = arena->alloc<float>(rows * att_c);
    logits = arena->alloc<float>(rows * CL);
    vec_a = arena->alloc<float>((size_t)maxB * cmax);
    vec_b = arena->alloc<float>((size_t)maxB * std::max(se_c, att_c));
    vec_c = arena->alloc<float>((size_t)maxB * cmax);
    ctx = arena->alloc<float>((size_t)maxB * 2 * CL);
    pooled = arena->alloc<float>((size_t)maxB * 2 * CL);
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    sd = nullptr;
  }

  void conv(const EcapaConvArgs& a, int B, cudaStream_t st) {
    dim3 grid(ceil_div(a.Cout, 64), ceil_div(a.T, 64), B);
    ecapa_conv_kernel<<<grid, 256, 0, st>>>(a);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }
  void tdnn(const Tdnn& t, const float* in1, int ld1, int off1, const float* in2, int ld2, int off2, float* out, int ldo, int offo,
            int B, int T, cudaStream_t st, const float* batch_bias = nullptr, int tanh_out = 0) {
    EcapaConvArgs a;
    a.in1 = in1; a.ld1 = ld1; a.off1 = off1; a.in2 = in2; a.ld2 = ld2; a.off2 = off2;
    a.w = t.w; a.bias = batch_bias ? batch_bias : t.b; a.bias_ld = batch_bias ? t.cout : 0;
    a.Cin = t.cin; a.Cout = t.cout; a.k = t.k; a.dilation = t.d; a.T = T;
    a.relu = 1; a.post_s = t.ps; a.post_b = t.pb; a.tanh_out = tanh_out;
    a.out = out; a.ldo = ldo; a.offo = offo;
    conv(a, B, st);
  }
  void vec_linear(const float* in, const float* W, const float* bias, float* out, int B, int Cin, int Cout, int act, cudaStream_t st) {
    vec_linear_kernel<<<ceil_div(B * Cout, 8), 256, 0, st>>>(in, W, bias, out, B, Cin, Cout, act);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  // z dev (B, in_size, T) channel-first -> cond dev (B, out_dim)
  void forward(const float* z, float* cond, int B, int T, cudaStream_t st) {
    AFTER_REQUIRE(B >= 1 && B <= maxB, AFTER_EINVAL, "batch exceeds max_batch given at after_create");
    AFTER_REQUIRE(T >= 2 && T <= maxT, AFTER_EINVAL, "T must be in [2, seq_len given at after_create]");
    const int Cin = cfg.te_in_size, CL = ch[nb - 1];
    {
      dim3 grid(ceil_div(T, 32), ceil_div(Cin, 32), B);
      channels_to_frames_kernel<<<grid, 256, 0, st>>>(z, zf, Cin, T);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    tdnn(first, zf, Cin, 0, nullptr, 0, 0, x0, ch[0], 0, B, T, st);
    const float* xin = x0;
    int xin_ld = ch[0], xin_off = 0, foff = 0;
    for (size_t bi = 0; bi < blocks.size(); ++bi) {
      const SeRes2& s = blocks[bi];
      const int C = s.cout, sub = C / scale;
      const float* rs = xin; int rs_ld = xin_ld, rs_off = xin_off;
      if (s.has_shortcut) {
        EcapaConvArgs a;
        a.in1 = xin; a.ld1 = xin_ld; a.off1 = xin_off; a.w = s.sc_w; a.bias = s.sc_b; a.Cin = s.cin; a.Cout = C; a.T = T;
        a.out = res; a.ldo = C;
        conv(a, B, st);
        rs = res; rs_ld = C; rs_off = 0;
      }
      tdnn(s.tdnn1, xin, xin_ld, xin_off, nullptr, 0, 0, y, C, 0, B, T, st);
      // Res2Net: slice 0 passes through; slice j+1 = TDNN(y_{j+1} + out_j)   (ecapa_encoder.py:203-223)
      AFTER_CUDA_CHECK(cudaMemcpy2DAsync(y2, (size_t)C * 4, y, (size_t)C * 4, (size_t)sub * 4, (size_t)B * T, cudaMemcpyDeviceToDevice, st));
      for (int j = 0; j < scale - 1; ++j)
        tdnn(s.sub[j], y, C, (j + 1) * sub, j == 0 ? nullptr : y2, C, j * sub, y2, C, (j + 1) * sub, B, T, st);
      tdnn(s.tdnn2, y2, C, 0, nullptr, 0, 0, y, C, 0, B, T, st);
      {  // squeeze-excitation
        dim3 grid(ceil_div(C, 128), B);
        frame_stats_kernel<<<grid, 128, 0, st>>>(y, vec_a, nullptr, T, C, C, 0);
        AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
        vec_linear(vec_a, s.se_w1, s.se_b1, vec_b, B, C, se_c, 1, st);
        vec_linear(vec_b, s.se_w2, s.se_b2, vec_c, B, se_c, C, 2, st);
        const size_t n = (size_t)B * T * C;
        se_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, vec_c, rs, rs_ld, rs_off, feats, CL, foff, B, T, C);
        AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
      }
      xin = feats; xin_ld = CL; xin_off = foff;
      foff += C;
    }
    tdnn(mfa, feats, CL, 0, nullptr, 0, 0, xm, CL, 0, B, T, st);
    // attentive statistics pooling with global context
    {
      dim3 grid(ceil_div(CL, 128), B);
      frame_stats_kernel<<<grid, 128, 0, st>>>(xm, ctx, ctx, T, CL, 2 * CL, CL);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    vec_linear(ctx, asp_wctx, asp_tdnn.b, vec_b, B, 2 * CL, att_c, 0, st);  // per-stream bias: W_ctx [mean; std] + b
    tdnn(asp_tdnn, xm, CL, 0, nullptr, 0, 0, att, att_c, 0, B, T, st, vec_b, 1);
    {
      EcapaConvArgs a;
      a.in1 = att; a.ld1 = att_c; a.w = asp_conv_w; a.bias = asp_conv_b; a.Cin = att_c; a.Cout = CL; a.T = T;
      a.out = logits; a.ldo = CL;
      conv(a, B, st);
      dim3 grid(ceil_div(CL, 128), B);
      attentive_pool_kernel<<<grid, 128, 0, st>>>(logits, xm, pooled, T, CL);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    vec_linear(pooled, fc_w, fc_b, cond, B, 2 * CL, out_dim, 0, st);
    if (cfg.te_use_tanh) {
      tanh_kernel<<<ceil_div(B * out_dim, 256), 256, 0, st>>>(cond, (size_t)B * out_dim);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
  }
};

}  // namespace after
