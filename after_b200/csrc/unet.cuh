// UNET1D conv denoiser (after/diffusion/networks/unet1d.py:30-429) with SelfAttention1d (blocks.py:201-243) on one
// B200: the GroupNorm/SiLU conv blocks with time / condition scale-shift MLPs that RectifiedFlow can bind as `net`
// instead of DenoiserV2.
//
// Same machinery as the codec (codec.cuh): activations are frame-major fp32 (B, T, C), every convolution is
//     act_operand_kernel (GroupNorm(min(16, C/4)) or GroupNorm(1) + SiLU)  ->  tap-GEMM (+bias, +residual)
// with strided pools reading the input as (B, T/r, r, C) phases and the nearest-neighbour upsample + conv3 folded into one
// r-phase 3-tap conv (the taps of an output phase that hit the same source frame are summed at load time).
// The channel concatenation [x | skip | time_cond] in front of gn1 is materialised once (its GroupNorm groups straddle
// the parts), statistics come from gn_stats_kernel, the time / condition modulation is one element-wise pass.
#pragma once
#include <cmath>
#include "codec.cuh"
#include "denoiser.cuh"
#include "ecapa.cuh"

namespace after {

// out (B, T, Ca + Cb + Cc) = [a | b | c] along channels; b, c optional
__global__ void __launch_bounds__(256)
concat3_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb, const float* __restrict__ c, int Cc,
               float* __restrict__ out, size_t rows) {
  const int Ct = Ca + Cb + Cc;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Ct) return;
  const size_t r = i / Ct;
  const int ch = (int)(i - r * Ct);
  float v;
  if (ch < Ca) v = a[r * Ca + ch];
  else if (ch < Ca + Cb) v = b[r * Cb + (ch - Ca)];
  else v = c[r * Cc + (ch - Ca - Cb)];
  out[i] = v;
}

// x[b, t, c] <- (x * t_mult[b, c] + t_add[b, c]) * c_mult[b, c] + c_add[b, c]     (unet1d.py:99-108)
// tm (B, 2C) = [t_mult | t_add]; cm (B, 2C) = [c_mult | c_add] or null
__global__ void __launch_bounds__(256)
modulate_kernel(float* __restrict__ x, const float* __restrict__ tm, const float* __restrict__ cm, int T, int C, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const size_t b = i / ((size_t)T * C);
  float v = x[i] * tm[b * 2 * C + c] + tm[b * 2 * C + C + c];
  if (cm) v = v * cm[b * 2 * C + c] + cm[b * 2 * C + C + c];
  x[i] = v;
}

__global__ void silu_kernel(float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = silu(x[i]);
}

// SPE (unet1d.py:7-25): out[b] = [sin(w_f * scale * t_b) | cos(...)], f < dim/2
__global__ void spe_kernel(const float* __restrict__ t, const float* __restrict__ w, float* __restrict__ out, int B, int half,
                           float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, f = i - b * half;
  const float a = w[f] * (t[b] * scale);
  out[(size_t)b * 2 * half + f] = sinf(a);
  out[(size_t)b * 2 * half + half + f] = cosf(a);
}

// rows [dst0, dst0 + n) of a (rows, C) table <- value (the dropped condition of classifier-free guidance)
__global__ void fill_kernel(float* __restrict__ x, float v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// Dense softmax attention of SelfAttention1d (blocks.py:222-241): qkv (B, T, 3C) = [q | k | v], heads are contiguous
// dh-channel slices; out (B, T, C).  scale = dh^-1/4 on q and on k.  One warp per query: lanes split the keys of a
// 32-key tile (online softmax across tiles), then every lane accumulates its output dims (d = lane + 32 i).
// T <= a few hundred at the levels that carry attention (the deepest ones), so this is a latency-sized kernel.
template <int MAXD>  // dh <= 32 * MAXD
__global__ void __launch_bounds__(128)
dense_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int C, int heads) {
  const int dh = C / heads;
  const int b = blockIdx.z, hd = blockIdx.y;
  const int q_idx = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q_idx >= T) return;
  const float scale = rsqrtf(sqrtf((float)dh));  // dh^-1/4
  const float* base = qkv + (size_t)b * T * 3 * C;
  const float* qp = base + (size_t)q_idx * 3 * C + hd * dh;
  float m = -INFINITY, l = 0.f;
  float acc[MAXD];
#pragma unroll
  for (int i = 0; i < MAXD; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < T; k0 += 32) {
    const int kj = k0 + lane;
    float s = -INFINITY;
    if (kj < T) {
      const float* kp = base + (size_t)kj * 3 * C + C + hd * dh;
      float d = 0.f;
      for (int e = 0; e < dh; ++e) d = fmaf(qp[e] * scale, kp[e] * scale, d);
      s = d;
    }
    const float mt = fmaxf(m, warp_max(s));
    const float corr = expf(m - mt);  // m = -inf on the first tile: exp(-inf) = 0
    const float p = kj < T ? expf(s - mt) : 0.f;
    l = l * corr + warp_sum(p);
#pragma unroll
    for (int i = 0; i < MAXD; ++i) acc[i] *= corr;
    const int nk = min(32, T - k0);
    for (int j = 0; j < nk; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const float* vp = base + (size_t)(k0 + j) * 3 * C + 2 * C + hd * dh;
#pragma unroll
      for (int i = 0; i < MAXD; ++i) {
        const int d = lane + 32 * i;
        if (d < dh) acc[i] = fmaf(pj, vp[d], acc[i]);
      }
    }
    m = mt;
  }
  float* op = out + ((size_t)b * T + q_idx) * C + hd * dh;
  const float inv = 1.0f / l;
#pragma unroll
  for (int i = 0; i < MAXD; ++i) {
    const int d = lane + 32 * i;
    if (d < dh) op[d] = acc[i] * inv;
  }
}

// out[t'] = sum_k w_k x[r t' + k - pad]: tap k reads phase (k - pad) mod r of input row t' + floor((k - pad) / r)
inline TapTable strided_taps_general(int k, int r, int pad) {
  AFTER_REQUIRE(k >= 1 && k <= MAX_TAPS, AFTER_EINVAL, "conv kernel size not supported");
  TapTable t;
  t.ntaps = k;
  for (int i = 0; i < k; ++i) {
    const int j = i - pad;
    const int ph = ((j % r) + r) % r;
    t.phase[0][i] = (int8_t)ph;
    t.shift[0][i] = (int16_t)((j - ph) / r);
  }
  return t;
}

struct UNet : ConvNet {
  after_config cfg{};
  int in_size = 0, out_size = 0, n = 0, k = 0, time_ch = 0, tci = 0, tcc = 0, cond_ch = 0, n_attn = 0, in0 = 0;
  bool res_last = false;
  std::vector<int> ch, ratios, ins;  // ratios carries the leading 1 (unet1d.py:287)
  int maxT = 0, maxN = 0;

  struct Mlp { float *w0 = nullptr, *b0 = nullptr, *w2 = nullptr, *b2 = nullptr; int in = 0, out = 0; };
  struct Block {
    NormAct gn1, gn2;
    ConvLayer conv1, conv2, to_out;
    Mlp time_mlp, cond_mlp;
    bool has_to_out = false, has_cond = false, res = true;
    int in_c = 0, out_c = 0, skip_c = 0, cat_c = 0;
  };
  struct Attn { bool on = false; NormAct norm; ConvLayer qkv, out; int C = 0, heads = 0; };
  struct Down { Block blk; Attn attn; ConvLayer pool; int ratio = 1; };
  struct Up { bool has_up = false; ConvLayer up; int ratio = 1; Block blk; Attn attn; };
  std::vector<Down> downs;
  std::vector<Up> ups;
  Block mid; Attn mid_attn;
  std::vector<ConvLayer> cond_emb;  // n + 1 convs when tcc > 0
  float* spe_w = nullptr;

  // workspace (frame-major)
  std::vector<float*> skip_buf, tc_lvl;  // per level
  float *xa = nullptr, *xb = nullptr, *y1 = nullptr, *cat = nullptr, *qkv = nullptr, *att = nullptr, *tc_in = nullptr, *tc_mid = nullptr;
  float *temb = nullptr, *hid = nullptr, *tmod = nullptr, *cmod = nullptr, *cond_dev = nullptr, *time_dev = nullptr;
  double* st_slot = nullptr;
  // sampler state (RectifiedFlow.sample over this net): channel-first public tensors
  float *x_state = nullptr, *x3 = nullptr, *tc3 = nullptr, *cond3 = nullptr, *time3 = nullptr, *guidance = nullptr, *proj = nullptr;

  void make_gn_silu(NormAct& a, const std::string& prefix, int C, int groups, int act) {
    a.C = C; a.norm = NORM_GROUP; a.act = act; a.groups = groups;
    AFTER_REQUIRE(groups >= 1 && C % groups == 0, AFTER_EINVAL, "GroupNorm channels not divisible by groups ('" + prefix + "')");
    const HostTensor& g = get(prefix + ".weight");
    const HostTensor& b = get(prefix + ".bias");
    AFTER_REQUIRE(g.numel() == C && b.numel() == C, AFTER_ESHAPE, "tensor '" + prefix + ".weight' has an unexpected shape");
    a.gamma = arena->upload(g.data);
    a.beta = arena->upload(b.data);
  }
  void make_mlp(Mlp& m, const std::string& prefix, int in, int out) {
    const HostTensor& w0 = get(prefix + ".0.weight");
    const HostTensor& w2 = get(prefix + ".2.weight");
    AFTER_REQUIRE(w0.numel() == (int64_t)128 * in && w2.numel() == (int64_t)out * 128, AFTER_ESHAPE,
                  "tensor '" + prefix + ".0.weight' has an unexpected shape");
    m.w0 = arena->upload(w0.data); m.b0 = arena->upload(get(prefix + ".0.bias").data);
    m.w2 = arena->upload(w2.data); m.b2 = arena->upload(get(prefix + ".2.bias").data);
    m.in = in; m.out = out;
  }
  // ConvBlock1D (unet1d.py:30-118)
  void make_block(Block& b, const std::string& p, int in_c, int out_c, int skip_c, bool res) {
    b.in_c = in_c; b.out_c = out_c; b.skip_c = skip_c; b.cat_c = in_c + skip_c + tcc; b.res = res;
    AFTER_REQUIRE(b.cat_c >= 4 && out_c >= 4, AFTER_EINVAL, "UNET1D block with fewer than 4 channels (GroupNorm(min(16, C // 4)) undefined)");
    make_gn_silu(b.gn1, p + ".gn1", b.cat_c, std::min(16, b.cat_c / 4), ACT_SILU);
    make_conv_plain(b.conv1, p + ".conv1", b.cat_c, out_c, k, conv_taps(k, 1, false), 1);
    make_gn_silu(b.gn2, p + ".gn2", out_c, std::min(16, out_c / 4), ACT_SILU);
    make_conv_plain(b.conv2, p + ".conv2", out_c, out_c, k, conv_taps(k, 1, false), 1);
    make_mlp(b.time_mlp, p + ".time_mlp", time_ch, 2 * out_c);
    b.has_cond = cond_ch > 0;
    if (b.has_cond) make_mlp(b.cond_mlp, p + ".cond_mlp", cond_ch, 2 * out_c);
    b.has_to_out = skip_c > 0;
    if (b.has_to_out) make_conv_plain(b.to_out, p + ".to_out", in_c, out_c, 1, conv_taps(1, 1, false), 1);
    else AFTER_REQUIRE(!res || in_c == out_c, AFTER_EINVAL, "UNET1D residual block with in_c != out_c and no to_out");
  }
  void make_attn(Attn& a, const std::string& p, int C, int heads) {
    a.on = true; a.C = C; a.heads = heads;
    AFTER_REQUIRE(heads >= 1 && C % heads == 0 && C / heads <= 128, AFTER_EINVAL, "SelfAttention1d head size not supported");
    make_gn_silu(a.norm, p + ".norm", C, 1, ACT_NONE);
    make_conv_plain(a.qkv, p + ".qkv_proj", C, 3 * C, 1, conv_taps(1, 1, false), 1);
    make_conv_plain(a.out, p + ".out_proj", C, C, 1, conv_taps(1, 1, false), 1);
  }
  // nn.Upsample(nearest, r) -> Conv1d(k = 3, same)  as one r-phase conv over source offsets {-1, 0, +1}   (unet1d.py:219-223)
  void make_upsample_conv(ConvLayer& L, const std::string& prefix, int cin, int cout, int r) {
    AFTER_REQUIRE(r >= 2 && r <= MAX_PHASES, AFTER_EINVAL, "UNET1D upsampling ratio not supported");
    const HostTensor& w = get(prefix + ".weight");
    AFTER_REQUIRE(w.shape.size() == 3 && w.shape[0] == cout && w.shape[1] == cin && w.shape[2] == 3, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight' has an unexpected shape");
    const HostTensor& b = get(prefix + ".bias");
    TapTable t;
    t.ntaps = 3;
    t.n_per_phase = cout;
    std::vector<float> m((size_t)r * cout * 3 * cin, 0.f);
    for (int p = 0; p < r; ++p) {
      for (int tap = 0; tap < 3; ++tap) { t.phase[p][tap] = 0; t.shift[p][tap] = (int16_t)(tap - 1); }
      for (int kk = 0; kk < 3; ++kk) {
        const int j = p + kk - 1;
        const int off = j >= 0 ? j / r : -1;  // floor(j / r) for j >= -1
        for (int o = 0; o < cout; ++o)
          for (int c = 0; c < cin; ++c) m[(((size_t)p * cout + o) * 3 + (off + 1)) * cin + c] += w.data[((size_t)o * cin + c) * 3 + kk];
      }
    }
    std::vector<float> bias((size_t)r * cout);
    for (int p = 0; p < r; ++p) std::copy(b.data.begin(), b.data.end(), bias.begin() + (size_t)p * cout);
    build_gemm_weight(L.w, *arena, m, bias.data(), r * cout, cin, t, tc_mode());
    L.cin = cin; L.cout = cout; L.in_phases = 1; L.out_phases = r;
  }

  void finalize(const after_config& c, const TensorMap& tensors, int prec, Arena* ar) {
    cfg = c; precision = prec; arena = ar; sd = &tensors;
    in_size = c.un_in_size; out_size = c.un_out_size > 0 ? c.un_out_size : c.un_in_size;
    n = c.un_n_levels; k = c.un_kernel_size; time_ch = c.un_time_channels; tci = c.un_time_cond_in_channels;
    tcc = c.un_time_cond_channels; cond_ch = c.un_cond_channels; n_attn = c.un_n_attn_layers; res_last = c.un_use_res_last != 0;
    AFTER_REQUIRE(n >= 1 && n <= AFTER_MAX_STAGES, AFTER_EINVAL, "bad UNET1D level count");
    AFTER_REQUIRE(k % 2 == 1 && k <= MAX_TAPS, AFTER_EINVAL, "UNET1D kernel_size must be odd and <= 7");
    AFTER_REQUIRE(time_ch >= 2 && time_ch % 2 == 0, AFTER_EINVAL, "UNET1D time_channels must be even and >= 2");
    AFTER_REQUIRE(tcc == 0 || tci > 0, AFTER_EINVAL, "UNET1D time_cond_channels > 0 needs time_cond_in_channels > 0");
    ch.assign(c.un_channels, c.un_channels + n);
    ratios.assign(1, 1);
    for (int i = 0; i + 1 < n; ++i) ratios.push_back(c.un_ratios[i]);
    in0 = in_size + (tcc ? 0 : tci);
    ins.assign(1, in0);
    for (int i = 0; i + 1 < n; ++i) ins.push_back(ch[i]);
    maxT = c.seq_len; maxN = 3 * c.max_batch;
    int total = 1;
    for (int r : ratios) { AFTER_REQUIRE(r >= 1 && r <= MAX_PHASES, AFTER_EINVAL, "UNET1D ratio not supported"); total *= r; }
    AFTER_REQUIRE(maxT % total == 0, AFTER_EINVAL, "seq_len must be a multiple of the product of the UNET1D ratios");

    if (tcc) {
      cond_emb.resize(n + 1);
      make_conv_plain(cond_emb[0], "cond_emb_time.0.0", tci, tcc, k, conv_taps(k, 1, false), 1);
      for (int i = 0; i < n; ++i) {
        const int r = ratios[i];
        make_conv_plain(cond_emb[i + 1], "cond_emb_time." + std::to_string(i + 1) + ".0", tcc, tcc, k,
                        r == 1 ? conv_taps(k, 1, false) : strided_taps_general(k, r, k / 2), r);
      }
    }
    downs.resize(n);
    for (int i = 0; i < n; ++i) {
      const std::string p = "down_layers." + std::to_string(i);
      Down& d = downs[i];
      d.ratio = ratios[i];
      make_block(d.blk, p + ".conv", ins[i], ins[i], 0, true);
      if (i >= 1 && i >= n - n_attn) make_attn(d.attn, p + ".self_attn", ins[i], 4);
      make_conv_plain(d.pool, p + ".pool", ins[i], ch[i], k, d.ratio == 1 ? conv_taps(k, 1, false) : strided_taps_general(k, d.ratio, k / 2),
                      d.ratio);
    }
    make_block(mid, "middle_block.conv", ch[n - 1], ch[n - 1], 0, true);
    if (n_attn > 0) make_attn(mid_attn, "middle_block.self_attn", ch[n - 1], ch[n - 1] / 32);
    ups.resize(n);
    for (int i = 1; i <= n; ++i) {
      const std::string p = "up_layers." + std::to_string(i - 1);
      Up& u = ups[i - 1];
      const bool last = i == n;
      const int ic = ch[n - i], oc = last ? out_size : ch[n - i - 1];
      u.ratio = ratios[n - i];
      if (u.ratio == 1) {
        u.has_up = ic != oc;
        if (u.has_up) make_conv_plain(u.up, p + ".up", ic, oc, 3, conv_taps(3, 1, false), 1);
      } else {
        u.has_up = true;
        make_upsample_conv(u.up, p + ".up.1", ic, oc, u.ratio);
      }
      make_block(u.blk, p + ".conv", oc, oc, last ? in0 : oc, last ? res_last : true);
      if (!last && i <= n_attn) make_attn(u.attn, p + ".self_attn", oc, 4);
    }
    {
      const int half = time_ch / 2;
      std::vector<float> w(half);
      for (int f = 0; f < half; ++f) w[f] = powf(1.0f / 10000.0f, 2.0f * (float)f / (float)time_ch);
      spe_w = arena->upload(w);
    }

    // ---- workspace: widest tensor at any level is max(cat_c, 3C of an attention) x T_level
    size_t mx = 0, mxp = 0;
    {
      int T = maxT;
      auto see = [&](int C, int Tl) { mx = std::max(mx, (size_t)C * Tl); mxp = std::max(mxp, (size_t)std::max(C, 64) * Tl); };
      for (int i = 0; i < n; ++i) {
        see(downs[i].blk.cat_c, T); see(3 * ins[i], T); see(ch[i], T);
        T /= ratios[i];
      }
      see(mid.cat_c, T); see(3 * ch[n - 1], T);
      for (int i = 1; i <= n; ++i) {
        T *= ratios[n - i];
        see(ups[i - 1].blk.cat_c, T); see(3 * ups[i - 1].blk.out_c, T); see(ch[n - i], T);
      }
    }
    alloc_workspace(mx, mxp, maxN, 1);
    const size_t el = mx * (size_t)maxN;
    xa = buf[0]; xb = buf[1]; y1 = buf[2];
    cat = arena->alloc<float>(el); qkv = arena->alloc<float>(el); att = arena->alloc<float>(el);
    skip_buf.resize(n); tc_lvl.resize(n);
    {
      int T = maxT;
      for (int i = 0; i < n; ++i) {
        skip_buf[i] = arena->alloc<float>((size_t)maxN * T * ins[i]);
        tc_lvl[i] = tcc ? arena->alloc<float>((size_t)maxN * T * tcc) : nullptr;
        T /= ratios[i];
      }
      tc_mid = tcc ? arena->alloc<float>((size_t)maxN * T * tcc) : nullptr;
    }
    tc_in = tci ? arena->alloc<float>((size_t)maxN * maxT * tci) : nullptr;
    int max_c = out_size;
    for (int c2 : ch) max_c = std::max(max_c, c2);
    max_c = std::max(max_c, in0);
    temb = arena->alloc<float>((size_t)maxN * time_ch);
    hid = arena->alloc<float>((size_t)maxN * 128);
    tmod = arena->alloc<float>((size_t)maxN * 2 * max_c);
    cmod = arena->alloc<float>((size_t)maxN * 2 * max_c);
    cond_dev = arena->alloc<float>((size_t)maxN * std::max(cond_ch, 1));
    time_dev = arena->alloc<float>((size_t)std::max(maxN, (int)c.max_steps));
    st_slot = arena->alloc<double>((size_t)maxN * 16 * 2);
    x_state = arena->alloc<float>((size_t)maxN * std::max(in_size, out_size) * maxT);
    x3 = arena->alloc<float>((size_t)maxN * in_size * maxT);
    tc3 = arena->alloc<float>((size_t)maxN * std::max(tci, 1) * maxT);
    cond3 = arena->alloc<float>((size_t)maxN * std::max(cond_ch, 1));
    time3 = arena->alloc<float>((size_t)maxN);
    guidance = arena->alloc<float>(4);
    proj = arena->alloc<float>((size_t)maxN * maxT * out_size);
    time_grid_cap = (size_t)c.max_steps * maxN;
    time_grid_dev = arena->alloc<float>(time_grid_cap);
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    sd = nullptr;
  }
  void destroy() { graphs.destroy(); }

  // ---- run-time pieces ----------------------------------------------------------------------------------------
  void stats_of(const float* x, int B, int T, int C, int groups, cudaStream_t st) {
    AFTER_CUDA_CHECK(cudaMemsetAsync(st_slot, 0, (size_t)B * groups * 2 * sizeof(double), st));
    gn_stats_launch(x, st_slot, B, T, C, groups, st);
  }
  void mlp(const Mlp& m, const float* in, float* out, int B, cudaStream_t st) {
    vec_linear_kernel<<<ceil_div(B * 128, 8), 256, 0, st>>>(in, m.w0, m.b0, hid, B, m.in, 128, 3);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    vec_linear_kernel<<<ceil_div(B * m.out, 8), 256, 0, st>>>(hid, m.w2, m.b2, out, B, 128, m.out, 0);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }
  void ident_conv(const ConvLayer& L, const float* x, int C, int B, int T, float* out, const float* res, cudaStream_t st) {
    NormAct ident;
    produce(x, ident, nullptr, L, B, T, C, st);
    conv(L, B, T / L.in_phases, out, res, nullptr, 1, st);
  }
  // ConvBlock1D.forward (unet1d.py:84-118): x (B, T, in_c) [+ skip + tc] -> out (B, T, out_c); out != x
  void run_block(const Block& b, const float* x, const float* skip, const float* tc, float* out, int B, int T, cudaStream_t st) {
    const float* h = x;
    if (b.skip_c || tcc) {
      const size_t rows = (size_t)B * T;
      const size_t tot = rows * b.cat_c;
      concat3_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, b.in_c, skip, b.skip_c, tc, tcc, cat, rows);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
      h = cat;
    }
    stats_of(h, B, T, b.cat_c, b.gn1.groups, st);
    produce(h, b.gn1, st_slot, b.conv1, B, T, b.cat_c, st);
    conv(b.conv1, B, T, y1, nullptr, nullptr, 1, st);
    mlp(b.time_mlp, temb, tmod, B, st);
    if (b.has_cond) mlp(b.cond_mlp, cond_dev, cmod, B, st);
    {
      const size_t nel = (size_t)B * T * b.out_c;
      modulate_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(y1, tmod, b.has_cond ? cmod : nullptr, T, b.out_c, nel);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    const float* res = nullptr;
    if (b.res) {
      res = x;
      if (b.has_to_out) {  // to_out(x_in): the residual sees the input before the concatenation
        ident_conv(b.to_out, x, b.in_c, B, T, out, nullptr, st);
        res = out;
      }
    }
    stats_of(y1, B, T, b.out_c, b.gn2.groups, st);
    produce(y1, b.gn2, st_slot, b.conv2, B, T, b.out_c, st);
    conv(b.conv2, B, T, out, res, nullptr, 1, st);
  }
  // SelfAttention1d.forward (blocks.py:222-243), in place on x (B, T, C)
  void run_attn(const Attn& a, float* x, int B, int T, cudaStream_t st) {
    stats_of(x, B, T, a.C, 1, st);
    produce(x, a.norm, st_slot, a.qkv, B, T, a.C, st);
    conv(a.qkv, B, T, qkv, nullptr, nullptr, 1, st);
    dim3 grid(ceil_div(T, 4), a.heads, B);
    const int dh = a.C / a.heads;
    if (dh <= 32) dense_attn_kernel<1><<<grid, 128, 0, st>>>(qkv, att, T, a.C, a.heads);
    else if (dh <= 64) dense_attn_kernel<2><<<grid, 128, 0, st>>>(qkv, att, T, a.C, a.heads);
    else dense_attn_kernel<4><<<grid, 128, 0, st>>>(qkv, att, T, a.C, a.heads);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    ident_conv(a.out, att, a.C, B, T, x, x, st);  // + residual, in place (each element read then written by one thread)
  }
  void silu_inplace(float* x, size_t nel, cudaStream_t st) {
    silu_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(x, nel);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  // UNET1D.forward (unet1d.py:376-429) on device buffers: x_cf (N, in_size, T) channel-first, time_dev (N), cond_dev (N, cond),
  // tcond_cf (N, tci, T) channel-first -> out_frames (N, T, out_size) frame-major
  void forward_frames(const float* x_cf, const float* tcond_cf, float* out_frames, int N, int T, cudaStream_t st) {
    {
      spe_kernel<<<ceil_div(N * (time_ch / 2), 256), 256, 0, st>>>(time_dev, spe_w, temb, N, time_ch / 2, 32.0f);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
      dim3 grid(ceil_div(T, 32), ceil_div(in_size, 32), N);
      channels_to_frames_kernel<<<grid, 256, 0, st>>>(x_cf, tci && !tcc ? xb : xa, in_size, T);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
      if (tci) {
        dim3 g2(ceil_div(T, 32), ceil_div(tci, 32), N);
        channels_to_frames_kernel<<<g2, 256, 0, st>>>(tcond_cf, tc_in, tci, T);
        AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
        if (!tcc) {  // time_cond concatenated to the input (unet1d.py:413-414)
          const size_t rows = (size_t)N * T, tot = rows * in0;
          concat3_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(xb, in_size, tc_in, tci, nullptr, 0, xa, rows);
          AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
        }
      }
    }
    float *x = xa, *o = xb;
    int Tl = T;
    const float* tc_prev = tc_in;
    int tc_prev_c = tci;
    for (int i = 0; i < n; ++i) {
      const Down& d = downs[i];
      if (tcc) {  // cond_emb_time[i]: conv (stride ratios[i-1] for i >= 1) + SiLU   (unet1d.py:305-323, 386-389)
        ident_conv(cond_emb[i], tc_prev, tc_prev_c, N, i == 0 ? Tl : Tl * ratios[i - 1], tc_lvl[i], nullptr, st);
        silu_inplace(tc_lvl[i], (size_t)N * Tl * tcc, st);
        tc_prev = tc_lvl[i]; tc_prev_c = tcc;
      }
      run_block(d.blk, x, nullptr, tcc ? tc_lvl[i] : nullptr, skip_buf[i], N, Tl, st);
      if (d.attn.on) run_attn(d.attn, skip_buf[i], N, Tl, st);
      ident_conv(d.pool, skip_buf[i], ins[i], N, Tl, o, nullptr, st);
      std::swap(x, o);
      Tl /= d.ratio;
    }
    if (tcc) {
      ident_conv(cond_emb[n], tc_prev, tcc, N, Tl * ratios[n - 1], tc_mid, nullptr, st);
      silu_inplace(tc_mid, (size_t)N * Tl * tcc, st);
    }
    run_block(mid, x, nullptr, tcc ? tc_mid : nullptr, o, N, Tl, st);
    std::swap(x, o);
    if (mid_attn.on) run_attn(mid_attn, x, N, Tl, st);
    for (int i = 1; i <= n; ++i) {
      const Up& u = ups[i - 1];
      const int lvl = n - i;
      if (u.has_up) {
        ident_conv(u.up, x, ch[n - i], N, Tl, o, nullptr, st);
        std::swap(x, o);
      }
      Tl *= u.ratio;
      float* dst = i == n ? out_frames : o;
      run_block(u.blk, x, skip_buf[lvl], tcc ? tc_lvl[lvl] : nullptr, dst, N, Tl, st);
      if (i < n) std::swap(x, o);
      if (u.attn.on) run_attn(u.attn, x, N, Tl, st);
    }
  }

  void check_shape(int N, int T) {
    AFTER_REQUIRE(N >= 1 && N <= maxN, AFTER_EINVAL, "batch exceeds 3*max_batch given at after_create");
    int total = 1;
    for (int r : ratios) total *= r;
    AFTER_REQUIRE(T >= total && T <= maxT && T % total == 0, AFTER_EINVAL,
                  "T must be a multiple of the product of the UNET1D ratios and <= seq_len given at after_create");
  }

  // UNET1D.forward with the reference's public layouts (channel-first)
  void forward(const float* x, const float* time, const float* cond, const float* time_cond, float* out, int N, int T, cudaStream_t st) {
    check_shape(N, T);
    AFTER_REQUIRE(cond_ch == 0 || cond != nullptr, AFTER_EINVAL, "this UNET1D takes a global condition: cond must not be null");
    AFTER_REQUIRE(tci == 0 || time_cond != nullptr, AFTER_EINVAL, "this UNET1D takes a time condition: time_cond must not be null");
    AFTER_CUDA_CHECK(cudaMemcpyAsync(time_dev, time, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
    if (cond_ch) AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_dev, cond, (size_t)N * cond_ch * 4, cudaMemcpyDeviceToDevice, st));
    forward_frames(x, time_cond, proj, N, T, st);
    dim3 grid(ceil_div(T, 32), ceil_div(out_size, 32), N);
    tokens_to_channels_kernel<<<grid, 256, 0, st>>>(proj, out, out_size, T);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  // CFG batch of RectifiedFlow.model_forward (model.py:730-743; MIDI layout export_midi.py:322-345) from x_state / cond / time_cond
  void build_cfg_inputs(const float* cond, const float* time_cond, int B, int T, int variant, cudaStream_t st) {
    const size_t nc = (size_t)B * cond_ch, nt = (size_t)B * tci * T;
    auto fill = [&](float* p, size_t cnt) {
      if (!cnt) return;
      fill_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(p, cfg.drop_value, cnt);
      AFTER_CUDA_CHECK(cudaGetLastError());
    };
    for (int grp = 0; grp < 3; ++grp) {
      const bool tc_on = variant == AFTER_CFG_AUDIO ? grp <= 1 : grp == 0;
      const bool c_on = variant == AFTER_CFG_AUDIO ? grp == 0 : grp <= 1;
      if (nc) {
        if (c_on) AFTER_CUDA_CHECK(cudaMemcpyAsync(cond3 + grp * nc, cond, nc * 4, cudaMemcpyDeviceToDevice, st));
        else fill(cond3 + grp * nc, nc);
      }
      if (nt) {
        if (tc_on) AFTER_CUDA_CHECK(cudaMemcpyAsync(tc3 + grp * nt, time_cond, nt * 4, cudaMemcpyDeviceToDevice, st));
        else fill(tc3 + grp * nt, nt);
      }
    }
  }
  // one velocity evaluation at time_dev[0..3B): proj <- net([x; x; x]) frame-major
  void cfg_network(int B, int T, cudaStream_t st) {
    const size_t nx = (size_t)B * in_size * T;
    for (int grp = 0; grp < 3; ++grp)
      AFTER_CUDA_CHECK(cudaMemcpyAsync(x3 + grp * nx, x_state, nx * 4, cudaMemcpyDeviceToDevice, st));
    if (cond_ch) AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_dev, cond3, (size_t)3 * B * cond_ch * 4, cudaMemcpyDeviceToDevice, st));
    forward_frames(x3, tc3, proj, 3 * B, T, st);
  }
  void set_guidance(float g_timbre, float g_structure, int variant, float clamp, float dt, cudaStream_t st) {
    const float total = 0.5f * (g_structure + g_timbre);
    const float first = variant == AFTER_CFG_AUDIO ? g_timbre : g_structure;
    const float second = variant == AFTER_CFG_AUDIO ? g_structure : g_timbre;
    float hbuf[4] = {total, first / std::max(second, clamp), dt, 0.f};
    AFTER_CUDA_CHECK(cudaMemcpyAsync(guidance, hbuf, sizeof(hbuf), cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  void combine(int B, int T, float* xout, int euler, cudaStream_t st) {
    dim3 grid(ceil_div(T, 32), ceil_div(out_size, 32), B);
    cfg_combine_kernel<<<grid, 256, 0, st>>>(proj, guidance, x_state, xout, B, out_size, T, euler);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  // RectifiedFlow.model_forward (model.py:721-761) over this net
  void model_forward(const float* x, const float* time, const float* cond, const float* time_cond, float* out, int B, int T,
                     float g_t, float g_s, int variant, float clamp, cudaStream_t st) {
    check_shape(3 * B, T);
    AFTER_REQUIRE(in_size == out_size, AFTER_EINVAL, "sampling needs out_size == in_size");
    AFTER_REQUIRE(variant == AFTER_CFG_AUDIO || variant == AFTER_CFG_MIDI, AFTER_EINVAL, "unknown cfg_variant");
    AFTER_REQUIRE((cond_ch == 0 || cond) && (tci == 0 || time_cond), AFTER_EINVAL, "null condition tensor");
    AFTER_CUDA_CHECK(cudaMemcpyAsync(x_state, x, (size_t)B * in_size * T * 4, cudaMemcpyDeviceToDevice, st));
    for (int grp = 0; grp < 3; ++grp)
      AFTER_CUDA_CHECK(cudaMemcpyAsync(time_dev + grp * B, time, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    build_cfg_inputs(cond, time_cond, B, T, variant, st);
    set_guidance(g_t, g_s, variant, clamp, 1.0f, st);
    cfg_network(B, T, st);
    combine(B, T, out, 0, st);
  }

  // RectifiedFlow.sample (model.py:763-785) over this net: the Euler loop is one CUDA graph per (B, T, steps)
  void sample(const float* x0, const float* cond, const float* time_cond, float* out, int B, int T, int nb_steps, float g_t,
              float g_s, int variant, float clamp, cudaStream_t st) {
    check_shape(3 * B, T);
    AFTER_REQUIRE(in_size == out_size, AFTER_EINVAL, "sampling needs out_size == in_size");
    AFTER_REQUIRE(nb_steps >= 1 && nb_steps <= cfg.max_steps, AFTER_EINVAL, "nb_steps exceeds max_steps given at after_create");
    AFTER_REQUIRE(variant == AFTER_CFG_AUDIO || variant == AFTER_CFG_MIDI, AFTER_EINVAL, "unknown cfg_variant");
    AFTER_REQUIRE((cond_ch == 0 || cond) && (tci == 0 || time_cond), AFTER_EINVAL, "null condition tensor");
    const size_t xb_ = (size_t)B * in_size * T * 4;
    AFTER_CUDA_CHECK(cudaMemcpyAsync(x_state, x0, xb_, cudaMemcpyDeviceToDevice, st));
    build_cfg_inputs(cond, time_cond, B, T, variant, st);
    std::vector<float> tg = Denoiser::time_grid(nb_steps);
    std::vector<float> tt((size_t)nb_steps * 3 * B);
    for (int s = 0; s < nb_steps; ++s)
      for (int i = 0; i < 3 * B; ++i) tt[(size_t)s * 3 * B + i] = tg[s];
    AFTER_REQUIRE((size_t)nb_steps * 3 * B <= time_grid_cap, AFTER_EINVAL, "time grid exceeds its buffer");
    AFTER_CUDA_CHECK(cudaMemcpyAsync(time_grid_dev, tt.data(), tt.size() * 4, cudaMemcpyHostToDevice, st));
    set_guidance(g_t, g_s, variant, clamp, 1.0f / (float)nb_steps, st);  // synchronises: tt stays alive until here
    graphs.run({2, B, T, nb_steps}, st, [&] {
      for (int s = 0; s < nb_steps; ++s) {
        AFTER_CUDA_CHECK(cudaMemcpyAsync(time_dev, time_grid_dev + (size_t)s * 3 * B, (size_t)3 * B * 4, cudaMemcpyDeviceToDevice, st));
        cfg_network(B, T, st);
        combine(B, T, x_state, 1, st);
      }
    });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(out, x_state, xb_, cudaMemcpyDeviceToDevice, st));
  }
  float* time_grid_dev = nullptr;
  size_t time_grid_cap = 0;
};

}  // namespace after
