// Tap-GEMM: the one contraction kernel behind every linear layer of the denoiser and every 1-D convolution of
// the codec / structure encoder.
//
//   out[b, t, n] = sum_{tap} sum_{c}  A[b, t + shift(tap), phase(tap), c] * W[n, tap * Cin + c]   (+ epilogue)
//
// A is a frame-major activation (B, T, P, Cin) -- P "phases" per frame (P = stride of a strided conv, else 1) --
// rows outside [0, T) read as zero (the conv padding).  A linear layer is the 1-tap case; a dilated conv has
// shift = k*dilation - pad_left; a strided conv (k = 2f, stride f) reads phase/shift = (k - pad) mod/div f; a
// transposed conv (k = 2f, stride f) is f output phases (column blocks of n) of 2 taps each.
//
// Two implementations share one epilogue:
//   * tap_gemm_tc_kernel   -- tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM); operand tiles are
//                             fetched by TMA (SWIZZLE_128B; a 4-D map whose out-of-bounds zero fill *is* the conv
//                             padding) through an mbarrier ring.  fp32 accuracy comes from a split
//                             x = hi + lo (both bf16):  A*W ~= Alo*Whi + Ahi*Wlo + Ahi*Whi  (three MMAs into the same
//                             accumulator; the dropped lo*lo term is ~2^-18 relative).  bf16 mode issues hi*hi only.
//   * tap_gemm_simt_kernel -- fp32 FFMA register-tiled (validation mode, and the few layers with Cin % 64 != 0).
//
// Epilogue options: +bias | exact GELU | rotary embedding on q,k (transformerv2.py:267, rotary_embedding.py:143-173)
// | +residual | fp32 and/or bf16 hi/lo output | GroupNorm sum / sum-of-squares of the result (for the *next*
// layer's norm, SimpleNetsStream.py:165-167) accumulated with fp64 atomics.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace after {

constexpr int MAX_TAPS = 8;
constexpr int MAX_PHASES = 4;

struct TapTable {
  int ntaps = 1;        // taps per output phase; K = ntaps * Cin
  int n_per_phase = 0;  // output columns per output phase (0: one phase)
  int8_t phase[MAX_PHASES][MAX_TAPS] = {};
  int16_t shift[MAX_PHASES][MAX_TAPS] = {};
};

struct GemmEpi {
  float* out_f32 = nullptr;          // [B*T, ldo] fp32 result (optional)
  __nv_bfloat16* out_hi = nullptr;   // [B*T, ldo] bf16 split of the result (optional)
  __nv_bfloat16* out_lo = nullptr;
  int ldo = 0;
  const float* bias = nullptr;       // [N]; with bias_mod > 0 indexed by (n % bias_mod)
  int bias_mod = 0;
  const float* res = nullptr;        // [B*T, ldo] residual added after bias/activation
  int gelu = 0;
  // rotary embedding on the q and k thirds of a QKV projection
  int rope = 0;
  int D = 0;                         // embed dim: column / D = 0:q 1:k 2:v
  int rot_half = 16;                 // rotary_dim / 2
  const float2* rope_tab = nullptr;  // [T][rot_half] (cos, sin)
  // GroupNorm statistics of the result: stats[(b * groups + g) * 2 + {0,1}] += {sum, sum of squares}
  double* stats = nullptr;
  int stat_groups = 0;
  int stat_cpg = 1;                  // channels per group
  int stat_cmod = 0;                 // channel = n % stat_cmod (transposed conv: several phases share channels)
};

// vals: NV consecutive columns [col0, col0+NV) of frame t of batch b; col0 % 4 == 0, NV % 4 == 0.
// All lanes of a warp call this with the same (b, col0) and different t; `valid` masks rows beyond T.
template <int NV>
__device__ __forceinline__ void epi_store(const GemmEpi& e, int b, int t, int T, int col0, float* v, bool valid) {
  const int row = b * T + t;
  if (e.bias) {
    const int bc = e.bias_mod > 0 ? col0 % e.bias_mod : col0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 bb = *reinterpret_cast<const float4*>(e.bias + bc + i);
      v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
    }
  }
  if (e.gelu) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = gelu_erf(v[i]);
  }
  if (e.rope) {
    int part = col0 / e.D;
    int d0 = col0 & 63;  // head_dim = 64
    if (part < 2 && d0 < 2 * e.rot_half && valid) {
      const float2* tab = e.rope_tab + (size_t)t * e.rot_half + (d0 >> 1);
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        if (d0 + i < 2 * e.rot_half) {
          float2 cs = tab[i >> 1];
          float a = v[i], bq = v[i + 1];
          v[i] = a * cs.x - bq * cs.y;
          v[i + 1] = bq * cs.x + a * cs.y;
        }
      }
    }
  }
  const size_t off = (size_t)row * e.ldo + col0;
  if (e.res && valid) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 r = *reinterpret_cast<const float4*>(e.res + off + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (e.stats) {
    // warp-uniform walk over the groups these NV columns touch
    const int c0 = e.stat_cmod > 0 ? col0 % e.stat_cmod : col0;
    int g = c0 / e.stat_cpg;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int gi = (c0 + i) / e.stat_cpg;
      if (gi != g) {
        s = warp_sum(s); q = warp_sum(q);
        if ((threadIdx.x & 31) == 0) {
          double* p = e.stats + ((size_t)b * e.stat_groups + g) * 2;
          atomicAdd(p, (double)s); atomicAdd(p + 1, (double)q);
        }
        g = gi; s = 0.f; q = 0.f;
      }
      const float x = valid ? v[i] : 0.f;
      s += x; q = fmaf(x, x, q);
    }
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) {
      double* p = e.stats + ((size_t)b * e.stat_groups + g) * 2;
      atomicAdd(p, (double)s); atomicAdd(p + 1, (double)q);
    }
  }
  if (!valid) return;
  if (e.out_f32) {
#pragma unroll
    for (int i = 0; i < NV; i += 4)
      *reinterpret_cast<float4*>(e.out_f32 + off + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (e.out_hi) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(v[i + j], h[j], l[j]);
      *reinterpret_cast<uint2*>(e.out_hi + off + i) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      if (e.out_lo)
        *reinterpret_cast<uint2*>(e.out_lo + off + i) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
  }
}

// =====================================================================================
// fp32 SIMT tap-GEMM: 64x64 output tile per 256-thread block (4x4 register micro-tiles), K staged 16 at a time.
// Requires Cin % 16 == 0 and N % 4 == 0; T, N otherwise arbitrary (predicated).  grid = (N/64, T/64, B).
// The tile is kept small because the warp-uniform statistics walk in epi_store needs all 32 lanes of a warp on
// the same column group: lanes of a warp = 32 consecutive frames, each thread 4 columns x (2 x 2 frames)...
// so here a thread owns frames {ty, ty+32} x 2 and columns tx*4..+3 -- see the index math below.
// =====================================================================================
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256)
tap_gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W, GemmEpi epi, TapTable taps, int T, int P,
                     int Cin, int N) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int oph = taps.n_per_phase > 0 ? n0 / taps.n_per_phase : 0;
  const int K = taps.ntaps * Cin;
  // loader: 64 rows x 16 k = 256 float4 -> one float4 per thread per operand
  const int lrow = tid >> 2;
  const int lk = (tid & 3) * 4;
  const bool w_ok = (n0 + lrow) < N;
  const float* Wp = W + (size_t)(n0 + lrow) * K + lk;
  // compute mapping: warp w (0..7), lane l: frames fr = l + 32*{0,1}, columns (w*8 .. w*8+7)
  const int warp = tid >> 5, lane = tid & 31;

  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < taps.ntaps; ++tap) {
    const int tt = t0 + lrow + taps.shift[oph][tap];
    const bool a_ok = tt >= 0 && tt < T;
    const float* Ap = A + (((size_t)b * T + (a_ok ? tt : 0)) * P + taps.phase[oph][tap]) * Cin + lk;
    for (int c0 = 0; c0 < Cin; c0 += SG_BK) {
      float4 ra = a_ok ? *reinterpret_cast<const float4*>(Ap + c0) : make_float4(0, 0, 0, 0);
      float4 rw = w_ok ? *reinterpret_cast<const float4*>(Wp + (size_t)tap * Cin + c0) : make_float4(0, 0, 0, 0);
      __syncthreads();
      As[lk + 0][lrow] = ra.x; As[lk + 1][lrow] = ra.y; As[lk + 2][lrow] = ra.z; As[lk + 3][lrow] = ra.w;
      Ws[lk + 0][lrow] = rw.x; Ws[lk + 1][lrow] = rw.y; Ws[lk + 2][lrow] = rw.z; Ws[lk + 3][lrow] = rw.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SG_BK; ++k) {
        const float a0 = As[k][lane], a1 = As[k][lane + 32];
        const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][warp * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Ws[k][warp * 8 + 4]);
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
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + warp * 8 + jh * 4;
      if (col >= N) continue;  // warp-uniform
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      epi_store<4>(epi, b, t, T, col, v, t < T);
    }
  }
}

// Anything-goes fallback (Cin or N not a multiple of 4/16; a handful of 12-channel layers): one thread per output.
__global__ void tap_gemm_naive_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                      const float* __restrict__ bias, const float* __restrict__ res,
                                      float* __restrict__ out, TapTable taps, int B, int T, int P, int Cin, int N,
                                      int gelu) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * T * N) return;
  const int n = (int)(idx % N);
  const int t = (int)((idx / N) % T);
  const int b = (int)(idx / ((size_t)N * T));
  const int oph = taps.n_per_phase > 0 ? n / taps.n_per_phase : 0;
  const int K = taps.ntaps * Cin;
  float acc = 0.f;
  for (int tap = 0; tap < taps.ntaps; ++tap) {
    const int tt = t + taps.shift[oph][tap];
    if (tt < 0 || tt >= T) continue;
    const float* a = A + (((size_t)b * T + tt) * P + taps.phase[oph][tap]) * Cin;
    const float* w = W + (size_t)n * K + (size_t)tap * Cin;
    for (int c = 0; c < Cin; ++c) acc = fmaf(a[c], w[c], acc);
  }
  if (bias) acc += bias[n];
  if (gelu) acc = gelu_erf(acc);
  if (res) acc += res[idx];
  out[idx] = acc;
}

// =====================================================================================
// tcgen05 tap-GEMM
// =====================================================================================
namespace tc {

constexpr int BM = 128;   // UMMA M (cta_group::1): 128 frames of one stream
constexpr int BK = 64;    // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2-5: epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      printf("after_b200: mbarrier wait timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
  static constexpr int B_BYTES = BN * BK * 2;
  __host__ __device__ static constexpr int stage_bytes(int nprod) { return (nprod > 1 ? 2 : 1) * (A_BYTES + B_BYTES); }
  __host__ __device__ static constexpr int stages(int nprod) {
    int s = (200 * 1024) / stage_bytes(nprod);
    return s > 8 ? 8 : s;
  }
  __host__ __device__ static constexpr int total(int nprod) { return stages(nprod) * stage_bytes(nprod) + 1024 /*align*/ + 256 /*barriers*/; }
};

// grid = (N / BN, ceil(T / 128), B).  tmA_*: 4-D bf16 map over (Cin, P, T, B), box (64, 1, 128, 1);
// tmB_*: 2-D bf16 map over W [N, ntaps*Cin], box (64, BN).
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tap_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   GemmEpi epi, const __grid_constant__ TapTable taps, int T, int Cin, int nprod) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int A_BYTES = Smem<BN>::A_BYTES, B_BYTES = Smem<BN>::B_BYTES;
  const int n_ops = nprod > 1 ? 2 : 1;  // hi only, or hi + lo
  const int stage_bytes = n_ops * (A_BYTES + B_BYTES);
  const int n_stages = nprod > 1 ? Smem<BN>::stages(3) : Smem<BN>::stages(1);

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)n_stages * stage_bytes);
  uint64_t* full = bars;              // [n_stages]
  uint64_t* empty = bars + 8;         // [n_stages]
  uint64_t* tmem_full = bars + 16;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int cblocks = Cin / BK;
  const int nkb = taps.ntaps * cblocks;
  const int oph = taps.n_per_phase > 0 ? n0 / taps.n_per_phase : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    if (nprod > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN < 32 ? 32 : BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int kb = 0;
      for (int tap = 0; tap < taps.ntaps; ++tap) {
        const int ta = t0 + taps.shift[oph][tap];
        const int ph = taps.phase[oph][tap];
        for (int cb = 0; cb < cblocks; ++cb, ++kb) {
          const int s = kb % n_stages;
          const uint32_t par = (kb / n_stages) & 1;
          mbar_wait(&empty[s], par ^ 1);
          uint8_t* st = tiles + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
          tma_load_4d(&tmA_hi, &full[s], st, cb * BK, ph, ta, b);
          tma_load_2d(&tmB_hi, &full[s], st + A_BYTES, kb * BK, n0);
          if (nprod > 1) {
            tma_load_4d(&tmA_lo, &full[s], st + A_BYTES + B_BYTES, cb * BK, ph, ta, b);
            tma_load_2d(&tmB_lo, &full[s], st + 2 * A_BYTES + B_BYTES, kb * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % n_stages;
        const uint32_t par = (kb / n_stages) & 1;
        mbar_wait(&full[s], par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + A_BYTES);
        const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
        uint32_t accum = kb > 0;
        if (nprod > 1) {
          // small cross terms first, dominant hi*hi last
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) { umma_bf16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
        }
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) { umma_bf16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
    const int quad = warp & 3;
    const int t = t0 + quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(taddr + c, v);
      epi_store<16>(epi, b, t, T, n0 + c, v, t < T);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN < 32 ? 32 : BN) : "memory");
  }
}

}  // namespace tc
}  // namespace after
