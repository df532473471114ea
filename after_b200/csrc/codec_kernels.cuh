// Element-wise / filter-bank kernels of the codec (after/autoencoder/networks/SimpleNetsStream.py, pqmf.py,
// after/autoencoder/core.py) and of the structure encoder (after/diffusion/networks/encoder.py).
// All activations are frame-major fp32 (B, T, C); convolutions themselves are tap-GEMMs (gemm.cuh).
#pragma once
#include "common.cuh"

namespace after {

enum NormMode { NORM_NONE = 0, NORM_GROUP = 1, NORM_AFFINE = 2 };
enum ActMode { ACT_NONE = 0, ACT_SNAKE = 1, ACT_SILU = 2 };

struct ActParams {
  int norm = NORM_NONE;
  int act = ACT_NONE;
  // NORM_GROUP: GroupNorm(groups, C, eps=1e-5, affine) over (C/groups x T) per stream, biased variance
  //             (SimpleNetsStream.py:95-147 offline branch; statistics come from the producer's epilogue)
  const double* stats = nullptr;  // [B][groups][2] = sum, sum of squares
  int groups = 1;
  int stat_frames = 0;            // frames the statistics cover (0: the T frames of this call; streaming: pad + T)
  const float* gamma = nullptr;   // [C]
  const float* beta = nullptr;    // [C]
  // NORM_AFFINE: y = (x - mu[c]) * rs[c] + be[c]   (eval-mode BatchNorm, encoder.py:39-48; folded at load)
  const float* mu = nullptr;
  const float* rs = nullptr;
  const float* be = nullptr;
  // ACT_SNAKE: y + sin^2(alpha y) * inv_beta,  inv_beta = 1 / (beta + 1e-9)   (core.py:217-218)
  const float* alpha = nullptr;
  const float* inv_beta = nullptr;
};

struct OperandOut {
  float* f32 = nullptr;
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
};

// sin^2 has period pi: reduce r = x - k*pi (Cody-Waite, two-term pi) into [-pi/2, pi/2] and evaluate the odd Taylor
// polynomial through r^11 there (|error| < 6e-8): fp32-accurate for |x| up to ~1e5, ~14 instructions instead of the
// ~30 of sinf's general path.  (core.py:217-218 evaluates sin(alpha*x)**2 on un-normalised activations.)
__device__ __forceinline__ float sin_squared(float x) {
  const float k = rintf(x * 0.31830988618379067154f);
  float r = fmaf(k, -3.14159274101257324219f, x);   // pi rounded to fp32
  r = fmaf(k, 8.74227765734758577309e-8f, r);       // pi - fp32(pi) = -8.742e-8  (so subtracting k*pi adds +k*8.742e-8)
  const float r2 = r * r;
  float p = fmaf(r2, -2.50521083854417187751e-8f, 2.75573192239858906526e-6f);
  p = fmaf(p, r2, -1.98412698412698412698e-4f);
  p = fmaf(p, r2, 8.33333333333333333333e-3f);
  p = fmaf(p, r2, -1.66666666666666666667e-1f);
  p = fmaf(p, r2, 1.0f);
  const float s = r * p;
  return s * s;
}
__device__ __forceinline__ float snake_beta(float y, float a, float ib) { return fmaf(sin_squared(y * a), ib, y); }
__device__ __forceinline__ float silu(float y) { return y / (1.0f + expf(-y)); }

// x (B, T, C) fp32 -> operand (B, T, C): act(norm(x)).  grid = (ceil(T / frames_per_block), B), 256 threads.
// C % 4 == 0 is required for the vector path (C4 = C / 4 float4 per frame); a scalar variant handles the rest.
template <int VEC>
__global__ void __launch_bounds__(256)
act_operand_kernel(const float* __restrict__ x, OperandOut out, ActParams p, int T, int C, int Cp, int frames_per_block,
                   int out_T, int out_t0) {
  // out_T / out_t0: the output holds out_T frames per stream and this call's frames start at out_t0 (offline: T and 0;
  // streaming: the conv's persistent operand, whose first frames are the cached left context)
  pdl_wait();
  pdl_trigger();
  // Cp >= C: output channels [C, Cp) are written as zeros (operands of the few 16/32-channel layers are padded to the
  // 64-channel K granule of the tensor-core path)
  extern __shared__ float sm[];  // mu[C], rs[C], be[C], al[C], ib[C]
  float* s_mu = sm;
  float* s_rs = sm + C;
  float* s_be = sm + 2 * C;
  float* s_al = sm + 3 * C;
  float* s_ib = sm + 4 * C;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mu = 0.f, rs = 1.f, be = 0.f;
    if (p.norm == NORM_GROUP) {
      const int cpg = C / p.groups;
      const int g = c / cpg;
      const double n = (double)cpg * (double)(p.stat_frames > 0 ? p.stat_frames : T);
      const double s = p.stats[((size_t)b * p.groups + g) * 2];
      const double q = p.stats[((size_t)b * p.groups + g) * 2 + 1];
      const double mean = s / n;
      double var = q / n - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      mu = (float)mean;
      rs = (float)(1.0 / sqrt(var + 1e-5)) * p.gamma[c];
      be = p.beta[c];
    } else if (p.norm == NORM_AFFINE) {
      mu = p.mu[c]; rs = p.rs[c]; be = p.be[c];
    }
    s_mu[c] = mu; s_rs[c] = rs; s_be[c] = be;
    s_al[c] = p.act == ACT_SNAKE ? p.alpha[c] : 0.f;
    s_ib[c] = p.act == ACT_SNAKE ? p.inv_beta[c] : 0.f;
  }
  __syncthreads();
  const int t0 = blockIdx.x * frames_per_block;
  const int nt = min(frames_per_block, T - t0);
  const int per_frame = Cp / VEC;
  const int total = nt * per_frame;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int f = i / per_frame;
    const int c = (i - f * per_frame) * VEC;
    const size_t off = ((size_t)b * out_T + out_t0 + t0 + f) * Cp + c;
    float v[VEC];
    if (c < C) {
      const size_t in_off = ((size_t)b * T + t0 + f) * C + c;
      if (VEC == 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x + in_off);
        v[0] = xv.x; v[1 % VEC] = xv.y; v[2 % VEC] = xv.z; v[3 % VEC] = xv.w;
      } else {
        v[0] = x[in_off];
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float y = (v[j] - s_mu[c + j]) * s_rs[c + j] + s_be[c + j];
        if (p.act == ACT_SNAKE) y = snake_beta(y, s_al[c + j], s_ib[c + j]);
        else if (p.act == ACT_SILU) y = silu(y);
        v[j] = y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = 0.f;
    }
    if (VEC == 4) {
      if (out.f32) *reinterpret_cast<float4*>(out.f32 + off) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
      if (out.hi) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16(v[j % VEC], h[j], l[j]);
        *reinterpret_cast<uint2*>(out.hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (out.lo) *reinterpret_cast<uint2*>(out.lo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
    } else {
      if (out.f32) out.f32[off] = v[0];
      if (out.hi) {
        __nv_bfloat16 h, l;
        split_bf16(v[0], h, l);
        out.hi[off] = h;
        if (out.lo) out.lo[off] = l;
      }
    }
  }
}

// Fast path of the operand pass (C % 4 == 0): the block is (Cp / 4) channel quads x R frame rows, so a thread's
// channel quad never changes: its folded norm (y = x * sc + sh) and Snake parameters live in registers, the frame loop is
// load float4 -> 2 FMA + Snake -> bf16 split -> store, with no shared memory and no index division.  (ncu on the generic
// kernel above, baseAE encode at 8 chunks: 49 instructions per element at 71 % issue utilisation -- the index
// division and 20 shared-memory parameter loads per 4 elements were half of them; profiles/r02a_ncu_codec_full.json.)
// grid = (ceil(T / frames_per_block), B), block = (Cp / 4) * R threads (a multiple of 32, <= 256).
__global__ void __launch_bounds__(256)
act_operand_rows_kernel(const float* __restrict__ x, OperandOut out, ActParams p, int T, int C, int Cp, int frames_per_block,
                        int out_T, int out_t0, int R) {
  pdl_wait();
  pdl_trigger();
  const int q = Cp >> 2;               // channel quads per frame
  const int cq = threadIdx.x % q;      // one division per thread, outside the loop
  const int row = threadIdx.x / q;
  const int c = cq * 4;
  const int b = blockIdx.y;
  const bool live = c < C;             // quads in [C, Cp) are zero padding
  // y = (x - mu) * rs + be, in exactly that form (the generic kernel and the oracle round the same way)
  float mu4[4] = {0.f, 0.f, 0.f, 0.f}, rs4[4] = {1.f, 1.f, 1.f, 1.f}, be4[4] = {0.f, 0.f, 0.f, 0.f};
  float al[4] = {0.f, 0.f, 0.f, 0.f}, ib[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p.norm == NORM_GROUP) {
        const int cpg = C / p.groups;
        const int g = (c + j) / cpg;
        const double n = (double)cpg * (double)(p.stat_frames > 0 ? p.stat_frames : T);
        const double s = p.stats[((size_t)b * p.groups + g) * 2];
        const double qq = p.stats[((size_t)b * p.groups + g) * 2 + 1];
        const double mean = s / n;
        double var = qq / n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        mu4[j] = (float)mean;
        rs4[j] = (float)(1.0 / sqrt(var + 1e-5)) * p.gamma[c + j];
        be4[j] = p.beta[c + j];
      } else if (p.norm == NORM_AFFINE) {
        mu4[j] = p.mu[c + j]; rs4[j] = p.rs[c + j]; be4[j] = p.be[c + j];
      }
      if (p.act == ACT_SNAKE) { al[j] = p.alpha[c + j]; ib[j] = p.inv_beta[c + j]; }
    }
  }
  const int t0 = blockIdx.x * frames_per_block;
  const int nt = min(frames_per_block, T - t0);
  const float* xin = x + ((size_t)b * T + t0) * C + c;
  const size_t obase = ((size_t)b * out_T + out_t0 + t0) * Cp + c;
#pragma unroll 4
  for (int f = row; f < nt; f += R) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float4 xv = *reinterpret_cast<const float4*>(xin + (size_t)f * C);
      v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y = (v[j] - mu4[j]) * rs4[j] + be4[j];
        if (p.act == ACT_SNAKE) y = snake_beta(y, al[j], ib[j]);
        else if (p.act == ACT_SILU) y = silu(y);
        v[j] = y;
      }
    }
    const size_t off = obase + (size_t)f * Cp;
    if (out.f32) *reinterpret_cast<float4*>(out.f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
    if (out.hi) {
      __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(out.hi + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
      if (out.lo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
        *reinterpret_cast<uint2*>(out.lo + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
      }
    }
  }
}

// Stand-alone GroupNorm statistics of x (B, T, C): stats[b][g] += {sum, sum of squares}.  Used where the producer
// is not a tap-GEMM epilogue (PQMF output, naive-kernel layers).  grid = (ceil(T / 256), B), 256 threads.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int T, int C, int groups) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 256;
  const int nt = min(256, T - t0);
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = x + ((size_t)b * T + t0) * C + c;
    float s = 0.f, q = 0.f;
    for (int t = 0; t < nt; ++t) {
      const float v = p[(size_t)t * C];
      s += v;
      q = fmaf(v, v, q);
    }
    double* d = stats + ((size_t)b * groups + c / cpg) * 2;
    atomicAdd(d, (double)s);
    atomicAdd(d + 1, (double)q);
  }
}

// channel-first (B, C, T) <-> frame-major (B, T, C) fp32 (the reference's public layouts are channel-first)
__global__ void __launch_bounds__(256)
channels_to_frames_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T) {
  __shared__ float tile[32][33];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? in[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (c < C && t < T) out[((size_t)b * T + t) * C + c] = tile[tx][i];
  }
}

// -------------------------------------------------------------------------------------------
// PQMF analysis (pqmf.py:263-271, 286-290): out[b, t, m] = sum_k hk[m, k] * x[b, M t + k - pad_l], then negate
// odd bands at even frames (reverse_half, pqmf.py:16-20).  wT: [K][M] transposed filter bank.
// block = 64 frames x M(=16) bands, 256 threads: thread -> band (tid & 15), 4 frames.
// -------------------------------------------------------------------------------------------
constexpr int PQ_FRAMES = 64;

__global__ void __launch_bounds__(256)
pqmf_analysis_kernel(const float* __restrict__ audio, const float* __restrict__ wT, float* __restrict__ out, int T,
                     int K, int pad_l) {
  constexpr int M = 16;
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  float* s_w = sm;                  // [K][M]
  float* s_x = sm + (size_t)K * M;  // [PQ_FRAMES * M + K]
  const int b = blockIdx.y, t0 = blockIdx.x * PQ_FRAMES;
  const int64_t S = (int64_t)T * M;
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) s_w[i] = wT[i];
  const int seg = PQ_FRAMES * M + K;
  const int64_t s0 = (int64_t)t0 * M - pad_l;
  for (int i = threadIdx.x; i < seg; i += blockDim.x) {
    const int64_t s = s0 + i;
    s_x[i] = (s >= 0 && s < S) ? audio[(size_t)b * S + s] : 0.f;
  }
  __syncthreads();
  const int m = threadIdx.x & 15, fg = threadIdx.x >> 4;  // fg: 0..15 -> frames fg*4..fg*4+3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* xs = s_x + fg * 4 * M;
  for (int k = 0; k < K; ++k) {
    const float w = s_w[k * M + m];
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[f] = fmaf(w, xs[f * M + k], acc[f]);
  }
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int t = t0 + fg * 4 + f;
    if (t < T) {
      const float v = ((m & 1) && !(t & 1)) ? -acc[f] : acc[f];
      out[((size_t)b * T + t) * M + m] = v;
    }
  }
}

// -------------------------------------------------------------------------------------------
// Loudness gate + PQMF synthesis (SimpleNetsStream.py:644-646, pqmf.py:292-301):
//   u (B, T, 2M) -> g[c] = u[c] * sigmoid(u[M + c])            (use_loudness; else u is (B, T, M))
//   xs[c, t] = g[c, t] * (-1 if c odd and t even)
//   y[m, t] = M * sum_{c,k} w[m, c, k] xs[c, t + k - pad_l] ;  audio[b, t*M + (M-1-m)] = y[m, t]
// wT: [K][M(c)][M(m)].  block = 128 frames, 256 threads: thread -> band m (tid & 15), 8 frames.
// -------------------------------------------------------------------------------------------
constexpr int PS_FRAMES = 128;

__global__ void __launch_bounds__(256)
pqmf_synthesis_kernel(const float* __restrict__ u, const float* __restrict__ wT, float* __restrict__ audio, int T,
                      int K, int pad_l, int loud) {
  constexpr int M = 16;
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  float* s_w = sm;                      // [K][M][M]
  float* s_x = sm + (size_t)K * M * M;  // [(PS_FRAMES + K - 1)][M]
  const int b = blockIdx.y, t0 = blockIdx.x * PS_FRAMES;
  for (int i = threadIdx.x; i < K * M * M; i += blockDim.x) s_w[i] = wT[i];
  const int rows = PS_FRAMES + K - 1;
  const int ld = loud ? 2 * M : M;
  for (int i = threadIdx.x; i < rows * M; i += blockDim.x) {
    const int r = i / M, c = i % M;
    const int t = t0 + r - pad_l;
    float v = 0.f;
    if (t >= 0 && t < T) {
      const float* row = u + ((size_t)b * T + t) * ld;
      v = row[c];
      if (loud) v *= 1.0f / (1.0f + expf(-row[M + c]));
      if ((c & 1) && !(t & 1)) v = -v;
    }
    s_x[i] = v;
  }
  __syncthreads();
  const int m = threadIdx.x & 15, fg = threadIdx.x >> 4;  // frames fg*8 .. fg*8+7
  float acc[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) acc[f] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float* xr = s_x + (size_t)(fg * 8 + k) * M;
#pragma unroll
    for (int c = 0; c < M; ++c) {
      const float w = s_w[(k * M + c) * M + m];
#pragma unroll
      for (int f = 0; f < 8; ++f) acc[f] = fmaf(w, xr[f * M + c], acc[f]);
    }
  }
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    const int t = t0 + fg * 8 + f;
    if (t < T) audio[(size_t)b * T * M + (size_t)t * M + (M - 1 - m)] = acc[f] * (float)M;
  }
}

// -------------------------------------------------------------------------------------------
// Streaming kernels (cached_conv / CachedGroupNorm stream branch, see codec.cuh "streaming")
// -------------------------------------------------------------------------------------------
// CachedGroupNorm stream branch (SimpleNetsStream.py:134-144): statistics over [pad ; x], pad = the previous P frames
// (zeros before the start).  Per-frame per-group {sum, sumsq} of every frame seen live in a ring of cap = P + Tmax
// entries per stream; entry of absolute frame a is at a % cap.  `seen` (device): absolute index of this call's first
// frame = frames_seen_lat * scale.  The window [seen - P, seen + T) is reduced into stats[b][g] (pre-zeroed, fp64
// atomics); entries of the new frames are computed from x and stored, older ones are read back.
// grid = (ceil((P + T) / GN_WIN_ENTRIES), B), 256 threads, groups | 256.
constexpr int GN_WIN_ENTRIES = 256;
__global__ void __launch_bounds__(256)
gn_window_kernel(const float* __restrict__ x, double* __restrict__ hist, double* __restrict__ stats,
                 const long long* __restrict__ seen_lat, int scale, int P, int cap, int T, int C, int groups) {
  __shared__ double red[256][2];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int g = threadIdx.x % groups;
  const int epi = 256 / groups;  // entries per iteration
  const long long seen = seen_lat[0] * (long long)scale;
  double s = 0.0, q = 0.0;
  const int w0 = blockIdx.x * GN_WIN_ENTRIES;
  for (int w = w0 + threadIdx.x / groups; w < min(w0 + GN_WIN_ENTRIES, P + T); w += epi) {
    const long long a = seen - P + w;  // absolute frame
    if (a < 0) continue;               // before the start: zeros
    double* e = hist + (((size_t)b * cap + (size_t)(a % cap)) * groups + g) * 2;
    if (w < P) {
      s += e[0]; q += e[1];
    } else {
      const float* xp = x + ((size_t)b * T + (w - P)) * C + g * cpg;
      float fs = 0.f, fq = 0.f;
      for (int c = 0; c < cpg; ++c) { const float v = xp[c]; fs += v; fq = fmaf(v, v, fq); }
      e[0] = (double)fs; e[1] = (double)fq;
      s += (double)fs; q += (double)fq;
    }
  }
  red[threadIdx.x][0] = s; red[threadIdx.x][1] = q;
  __syncthreads();
  if (threadIdx.x < groups) {
    double ts = 0.0, tq = 0.0;
    for (int i = threadIdx.x; i < 256; i += groups) { ts += red[i][0]; tq += red[i][1]; }
    double* d = stats + ((size_t)b * groups + threadIdx.x) * 2;
    atomicAdd(d, ts);
    atomicAdd(d + 1, tq);
  }
}

// AlignBranches delay of a residual branch (cached_conv): xd[t] = X[t - d] over the stream X; state = last d frames.
// grid = (B), 256 threads; d * C <= 256 * RD_MAX.
constexpr int RD_MAX = 32;
__global__ void __launch_bounds__(256)
res_delay_kernel(const float* __restrict__ x, float* __restrict__ state, float* __restrict__ xd, int d, int T, int C) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x;
  const float* xb = x + (size_t)b * T * C;
  float* sb = state + (size_t)b * d * C;
  float* ob = xd + (size_t)b * T * C;
  for (int i = threadIdx.x; i < T * C; i += 256) {
    const int t = i / C;
    ob[i] = t < d ? sb[i] : xb[i - d * C];
  }
  // new state = last d frames of [state ; x]: read everything this thread will write before anyone writes
  float keep[RD_MAX];
#pragma unroll
  for (int j = 0; j < RD_MAX; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < d * C) {
      const int src = i + T * C;  // index into [state ; x]
      keep[j] = src < d * C ? sb[src] : xb[src - d * C];
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RD_MAX; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < d * C) sb[i] = keep[j];
  }
}

// End of a streaming call: every cached conv keeps the last `ns` frames of its operand in front of the next call's
// frames (CachedPadding1d): slab[0, ns) <- slab[T, T + ns) for all layers at once.  One block per (layer, stream);
// the copy moves data to lower addresses, chunk by chunk in ascending order with a barrier between a chunk's reads and
// writes, so overlapping source / destination ranges (ns > T) are safe.  Block (0, 0) also advances the frame counter.
struct RollDesc {
  void* a;        // bf16 hi or fp32 slab base
  void* b;        // bf16 lo slab base or null
  int elem_bytes; // 2 | 4
  int ns;         // frames kept (aligned count)
  int Cp;         // channels per frame
  int slab;       // frames per stream slab
  int t_scale;    // frames of this layer's input per latent frame
};
__global__ void __launch_bounds__(256)
stream_roll_kernel(const RollDesc* __restrict__ descs, int T_lat, long long* __restrict__ seen_lat) {
  pdl_wait();
  pdl_trigger();
  const RollDesc d = descs[blockIdx.x];
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) seen_lat[0] += T_lat;
  const int T = T_lat * d.t_scale;
  const size_t row = (size_t)d.Cp * d.elem_bytes / 16;  // uint4 per frame (Cp * elem_bytes is a multiple of 16)
  const size_t n = (size_t)d.ns * row, shift = (size_t)T * row;
  for (int arr = 0; arr < 2; ++arr) {
    uint4* base = reinterpret_cast<uint4*>(arr == 0 ? d.a : d.b);
    if (!base) continue;
    base += (size_t)b * d.slab * row;
    for (size_t c0 = 0; c0 < n; c0 += 256 * 4) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t i = c0 + threadIdx.x + 256 * j;
        if (i < n) v[j] = base[i + shift];
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t i = c0 + threadIdx.x + 256 * j;
        if (i < n) base[i] = v[j];
      }
      __syncthreads();
    }
  }
}

__global__ void stream_advance_kernel(int T_lat, long long* __restrict__ seen_lat) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) seen_lat[0] += T_lat;
}

// AE_notcausal.decode (export_autoencoder.py:128-153), audio side: y (B, (T + nf) * r) decoded from [z_buffer ; z];
//   y[:nf r] <- (1 - alpha) out_buffer + alpha y[:nf r], alpha = linspace(0, 1, nf r); out_buffer <- y[-nf r:];
//   out (B, T r) <- y[:-nf r]
__global__ void __launch_bounds__(256)
overlap_add_kernel(const float* __restrict__ y, float* __restrict__ out_buffer, float* __restrict__ out, int n_out, int n_fade) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float* yb = y + (size_t)b * (n_out + n_fade);
  float* ob = out_buffer + (size_t)b * n_fade;
  // requires n_out >= n_fade (host checks): thread i < n_fade is the only reader and the only writer of out_buffer[i]
  if (i < n_out) {
    float v = yb[i];
    if (i < n_fade) {
      // torch.linspace(0, 1, n_fade)[i] the way ATen fills it (symmetric halves)
      const float step = 1.0f / (float)(n_fade - 1);
      const float alpha = i < n_fade / 2 ? step * (float)i : 1.0f - step * (float)(n_fade - 1 - i);
      v = (1.0f - alpha) * ob[i] + alpha * v;
    }
    out[(size_t)b * n_out + i] = v;
  }
  if (i < n_fade) ob[i] = yb[n_out + i];  // last n_fade samples of y: beyond the faded head because n_out >= n_fade
}

// tanh epilogue of the structure encoder when use_tanh is set (encoder.py:296-297)
__global__ void tanh_kernel(float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = tanhf(x[i]);
}

}  // namespace after
