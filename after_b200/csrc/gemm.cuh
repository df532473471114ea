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
  int debug_skip = 0;                // -DAFTER_DEBUG builds only: 1 = drop the epilogue's global traffic, 2 = no operand loads, 4 = no MMAs
  unsigned long long* debug_ts = nullptr;  // -DAFTER_DEBUG builds only: %globaltimer marks of CTA 0
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
                     int Cin, int N, int Tin, size_t a_bstride) {
  // Tin: input rows per stream (offline: T; a streaming conv reads [cached context ; new frames]); a_bstride: elements
  // between the streams of A
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
    const bool a_ok = tt >= 0 && tt < Tin;
    const float* Ap = A + (size_t)b * a_bstride + ((size_t)(a_ok ? tt : 0) * P + taps.phase[oph][tap]) * Cin + lk;
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
                                      int gelu, int Tin, size_t a_bstride) {
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
    if (tt < 0 || tt >= Tin) continue;
    const float* a = A + (size_t)b * a_bstride + ((size_t)tt * P + taps.phase[oph][tap]) * Cin;
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue now, wait later (tcgen05.wait::ld covers every outstanding load of the warp)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Row-coalesced epilogue of one 32 (frames) x 32 (columns) accumulator block.  On entry lane l holds row l of the
// block in v[32] (the TMEM 32x32b layout); the block is transposed through a per-warp shared-memory tile so that
// afterwards lane l owns COLUMN l and the warp walks the rows: every global access (residual, fp32 / bf16 outputs,
// RoPE table) is then one contiguous 128-byte (64-byte for bf16) row segment.  Frames >= T are skipped (uniform).

// Epilogue flavours of the pair kernel, resolved at compile time so that each instantiation carries only its own code:
//   EPI_PLAIN : (+bias) (+residual) -> fp32 (and/or bf16 hi/lo), optional GroupNorm statistics
//   EPI_ROPE  : rotary embedding on the q / k thirds -> fp32                       (QKV projection, no bias)
//   EPI_GELU  : +bias, exact GELU -> bf16 hi/lo (and/or fp32)                      (MLP up-projection)
//   EPI_ROPE_BF16 : EPI_ROPE with a bf16 result (out_hi): the QKV buffer of the bf16 mode -- the GEMM epilogues run at
//               the chip's write bandwidth (profiles/EXPERIMENTS.md), so half the bytes is half the epilogue
enum EpiMode { EPI_PLAIN = 0, EPI_ROPE = 1, EPI_GELU = 2, EPI_ROPE_BF16 = 3 };
__host__ __device__ constexpr bool epi_is_rope(int mode) { return mode == EPI_ROPE || mode == EPI_ROPE_BF16; }

// Global operands of one half-block's epilogue (residual rows or RoPE cos/sin rows), requested before the TMEM load.
// Half-block = 32 frames x 16 columns; transposed ownership: lane -> 4 consecutive columns (lane & 3) * 4 of rows
// (lane >> 2) + 8 i, i = 0..3.
struct EpiAux {
  float4 a[4];
};
template <int MODE>
__device__ __forceinline__ void epi_prefetch(const GemmEpi& e, EpiAux& aux, int b, int t_base, int T, int col0, int lane) {
  const int c4 = (lane & 3) * 4, rsub = lane >> 2;
  const int col = col0 + c4;
  const int nvalid = T - t_base;
  if (MODE == EPI_PLAIN) {
    if (e.res) {
      const float* p = e.res + ((size_t)b * T + t_base + rsub) * e.ldo + col;
      const size_t step = (size_t)8 * e.ldo;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        aux.a[i] = (i * 8 + rsub) < nvalid ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        p += step;
      }
    }
  } else if (epi_is_rope(MODE)) {
    const bool rot = (col0 / e.D) < 2 && (col0 & 63) < 2 * e.rot_half;  // warp-uniform
    if (rot) {
      const float2* p = e.rope_tab + (size_t)(t_base + rsub) * e.rot_half + ((col & 63) >> 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        aux.a[i] = (i * 8 + rsub) < nvalid ? *reinterpret_cast<const float4*>(p) : make_float4(1.f, 0.f, 1.f, 0.f);
        p += 8 * e.rot_half;
      }
    }
  }
}

// hi/lo bf16 split of four values with the packed converter: 2 cvt.rn.bf16x2 + 4 subtractions + 2 cvt
__device__ __forceinline__ void split4_bf16(const float4& x, uint2& hi, uint2& lo) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y), l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
  hi = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
}

// Row-coalesced epilogue of one 32 (frames) x 16 (columns) accumulator half-block.  On entry lane l holds row l in
// v[16] (the TMEM 32x32b layout); the half-block is transposed through a per-warp 2 KB shared-memory tile (16-byte
// chunks XOR-swizzled by (row >> 1) & 3: conflict-free both ways) so that afterwards every global access (residual,
// fp32 / bf16 outputs, RoPE table) is a contiguous 64-byte (32-byte for bf16) row segment = whole 32-byte sectors.
// Sixteen epilogue warps (4 per scheduler) rather than eight: the epilogue is latency-bound per warp
// (TMEM load -> smem transpose -> arithmetic -> store is one dependent chain), so resident warps are what scales it.
constexpr int EPI_STAGE_BYTES = 32 * 16 * 4;  // per warp
constexpr int EPI_WARPS = 16;                 // warps 2..17: TMEM lane quarter = warp & 3, column quarter = (warp - 2) >> 2

// OUT: which outputs exist, when the caller knows (bit 0 fp32, bit 1 bf16 hi, bit 2 bf16 lo; -1 = look at the
// pointers at run time); FULL: all 32 rows are valid.  Both only remove per-row pointer tests / row predicates: at 16
// epilogue warps per SM this code is issue-bound (SASS: 535 instructions per half-block in the GELU flavour before,
// ~40 % of them tests, selects and address arithmetic), so instructions per element are what sets its duration.
template <int MODE, int OUT = -1, bool FULL = false>
__device__ __forceinline__ void epi_block(const GemmEpi& e, float* stg, const uint32_t* v, const EpiAux& aux, float4 bias4, int b,
                                          int t_base, int T, int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(stg + lane * 16 + ((i ^ ((lane >> 1) & 3)) << 2)) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  __syncwarp();
  const int cq = lane & 3;
  const int rsub = lane >> 2;
  const int col = col0 + cq * 4;
  const int nvalid = T - t_base;
  const bool rot = epi_is_rope(MODE) && (col0 / e.D) < 2 && (col0 & 63) < 2 * e.rot_half;  // warp-uniform
  const bool has_res = MODE == EPI_PLAIN && e.res != nullptr;
  const size_t off0 = ((size_t)b * T + t_base + rsub) * e.ldo + col;
  const size_t step = (size_t)8 * e.ldo;
  const bool has_f = OUT < 0 ? e.out_f32 != nullptr : (OUT & 1) != 0;
  const bool has_h = OUT < 0 ? e.out_hi != nullptr : (OUT & 2) != 0;
  const bool has_l = OUT < 0 ? e.out_lo != nullptr : (OUT & 4) != 0;
  float* pf = has_f ? e.out_f32 + off0 : nullptr;
  __nv_bfloat16* ph = has_h ? e.out_hi + off0 : nullptr;
  __nv_bfloat16* pl = has_l ? e.out_lo + off0 : nullptr;
  float ssum = 0.f, ssq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + rsub;
    float4 x = *reinterpret_cast<const float4*>(stg + r * 16 + ((cq ^ ((r >> 1) & 3)) << 2));
    x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
    if (MODE == EPI_GELU) {
      const float2 g0 = gelu_erf_fast2(make_float2(x.x, x.y)), g1 = gelu_erf_fast2(make_float2(x.z, x.w));
      x = make_float4(g0.x, g0.y, g1.x, g1.y);
    }
    if (rot) {
      const float4 cs = aux.a[i];
      const float a0 = x.x, b0 = x.y, a1 = x.z, b1 = x.w;
      x.x = a0 * cs.x - b0 * cs.y; x.y = b0 * cs.x + a0 * cs.y;
      x.z = a1 * cs.z - b1 * cs.w; x.w = b1 * cs.z + a1 * cs.w;
    }
    if (has_res) { x.x += aux.a[i].x; x.y += aux.a[i].y; x.z += aux.a[i].z; x.w += aux.a[i].w; }
    if ((FULL || r < nvalid) && !(kDebugBuild && (e.debug_skip & 1))) {
      if (MODE == EPI_PLAIN) {
        ssum += (x.x + x.y) + (x.z + x.w);
        ssq = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, ssq))));
      }
      if (has_f) *reinterpret_cast<float4*>(pf) = x;
      if (has_h) {
        uint2 hi, lo;
        split4_bf16(x, hi, lo);
        *reinterpret_cast<uint2*>(ph) = hi;
        if (has_l) *reinterpret_cast<uint2*>(pl) = lo;
      }
    }
    if (has_f) pf += step;
    if (has_h) ph += step;
    if (has_l) pl += step;
  }
  if (MODE == EPI_PLAIN && e.stats) {
    // fold the 8 row sub-lanes (xor 4, 8, 16) and the neighbouring 4-column lane (xor 1): lanes 0 and 2 then hold the
    // sums of 8 aligned columns, which always share a GroupNorm group here (channels per group % 8 == 0)
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);  ssq += __shfl_xor_sync(0xffffffffu, ssq, 4);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 8);  ssq += __shfl_xor_sync(0xffffffffu, ssq, 8);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 16); ssq += __shfl_xor_sync(0xffffffffu, ssq, 16);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);  ssq += __shfl_xor_sync(0xffffffffu, ssq, 1);
    if ((lane & 0x1D) == 0) {
      const int c = e.stat_cmod > 0 ? col % e.stat_cmod : col;
      double* p = e.stats + ((size_t)b * e.stat_groups + c / e.stat_cpg) * 2;
      atomicAdd(p, (double)ssum);
      atomicAdd(p + 1, (double)ssq);
    }
  }
  __syncwarp();
}

// Up-projection epilogue with TMA stores: lane l holds row l of a 32 x 16 half-block (the TMEM 32x32b layout); bias and
// GELU are applied in place, the bf16 hi (and lo) rows -- 32 bytes each -- go row-major into the warp's 2 KB staging tile
// and ONE lane issues a 3-D cp.async.bulk.tensor store per array (frames beyond T are clipped by the TMA unit).  Against
// epi_block this drops the shared-memory transposition (4 st + 4 ld.shared.v4), the per-row address arithmetic and 8
// global store instructions per lane; the staging tile is reused as soon as the previous store has READ it
// (cp.async.bulk.wait_group.read), which the next half-block's TMEM load and GELU cover.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void epi_block_gelu_tma(const CUtensorMap* o_hi, const CUtensorMap* o_lo, bool want_lo, float* stg,
                                                   const uint32_t* v, const float* __restrict__ bias, int b, int t_base,
                                                   int col0, int lane) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + col0 + 4 * i);  // same address in every lane: one broadcast
    const float2 g0 = gelu_erf_fast2(make_float2(__uint_as_float(v[4 * i]) + bb.x, __uint_as_float(v[4 * i + 1]) + bb.y));
    const float2 g1 = gelu_erf_fast2(make_float2(__uint_as_float(v[4 * i + 2]) + bb.z, __uint_as_float(v[4 * i + 3]) + bb.w));
    uint2 h2, l2;
    split4_bf16(make_float4(g0.x, g0.y, g1.x, g1.y), h2, l2);
    hi[2 * i] = h2.x; hi[2 * i + 1] = h2.y;
    lo[2 * i] = l2.x; lo[2 * i + 1] = l2.y;
  }
  // the previous half-block's stores must have read the staging tile before it is overwritten
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  uint4* sh = reinterpret_cast<uint4*>(stg) + lane * 2;         // hi tile: [32 rows][32 B]
  uint4* sl = reinterpret_cast<uint4*>(stg) + 64 + lane * 2;    // lo tile behind it
  sh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  sh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  if (want_lo) {
    sl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    sl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(o_hi, stg, col0, t_base, b);
    if (want_lo) tma_store_3d(o_lo, reinterpret_cast<const uint4*>(stg) + 64, col0, t_base, b);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
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
  pdl_wait();     // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail (common.cuh)
  pdl_trigger();

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


// =====================================================================================
// tcgen05 tap-GEMM, CTA-pair version (cta_group::2): the main kernel.
//
// A cluster of two CTAs (one TPC) owns a 256 x BN output tile: CTA r holds frames [t0 + 128 r, +128) of A and
// rows [n0 + r BN/2, +BN/2) of W in its own shared memory, the leader CTA issues one M=256 MMA per K=16 slice and
// each CTA receives its 128 x BN half of the accumulator in its own TMEM.  Halving the W bytes each SM pulls from
// L2 is what matters: the 1-CTA 128x128 kernel needs ~85 B/clk/SM of operands (bf16x3 mode) against the ~42 B/clk/SM
// the L2 delivers chip-wide, this one needs ~43.
//
// Persistent: grid = 2 x min(#tiles, #SM pairs); each cluster walks tiles (n fastest) round-robin.  The TMEM
// accumulator is double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 0 : TMA producer (both CTAs; transaction bytes of both land on the leader's `full` barrier)
//   warp 1 : TMEM alloc/dealloc (both CTAs), MMA issue (leader only; tcgen05.commit multicast frees the stage in
//            both CTAs and publishes the accumulator to both epilogues)
//   warps 2-5 : epilogue (tcgen05.ld -> epi_store); their arrival on the leader's `tmem_empty` recycles the buffer
// =====================================================================================
__device__ __forceinline__ void ts_mark(const GemmEpi& e, int slot) {
  if (kDebugBuild && e.debug_ts && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    e.debug_ts[slot] = t;
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t mbar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint32_t mbar_cluster, void* dst, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc2(int n) {  // M = 256 across the CTA pair
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN>
struct Smem2 {
  static constexpr int A_BYTES = BM * BK * 2;         // 16 KiB: this CTA's 128 frames
  static constexpr int B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the W tile
  __host__ __device__ static constexpr int stage_bytes(int nprod) { return (nprod > 1 ? 2 : 1) * (A_BYTES + B_BYTES); }
  __host__ __device__ static constexpr int stages(int nprod) {
    int s = (192 * 1024) / stage_bytes(nprod);
    return s > 8 ? 8 : s;
  }
  __host__ __device__ static constexpr int total(int nprod) { return stages(nprod) * stage_bytes(nprod) + 1024 + 512 + EPI_WARPS * EPI_STAGE_BYTES; }
};

constexpr int NUM_THREADS2 = 64 + 32 * EPI_WARPS;

template <int BN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
tap_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    GemmEpi epi, const __grid_constant__ TapTable taps, int T, int Cin, int nprod, int n_tiles_n,
                    int m_tiles_per_b, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int A_BYTES = Smem2<BN>::A_BYTES, B_BYTES = Smem2<BN>::B_BYTES;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  const int n_ops = nprod > 1 ? 2 : 1;
  const int stage_bytes = n_ops * (A_BYTES + B_BYTES);
  const int n_stages = nprod > 1 ? Smem2<BN>::stages(3) : Smem2<BN>::stages(1);

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)n_stages * stage_bytes);
  uint64_t* full = bars;             // [8]  (the leader's copies are the live ones)
  uint64_t* empty = bars + 8;        // [8]
  uint64_t* tmem_full = bars + 16;   // [2]
  uint64_t* tmem_empty = bars + 18;  // [2]  (leader's copies)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* epi_stage = reinterpret_cast<float*>(bars + 64);  // EPI_WARPS x [32][16], swizzled

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) ts_mark(epi, 0);
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int cblocks = Cin / BK;
  const int nkb = taps.ntaps * cblocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
    if (nprod > 1) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);  // epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail (common.cuh)
  pdl_trigger();
  if (threadIdx.x == 0) ts_mark(epi, 1);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int nt = tile % n_tiles_n, mt = tile / n_tiles_n;
        const int b = mt / m_tiles_per_b;
        const int t0 = (mt % m_tiles_per_b) * (2 * BM) + (int)rank * BM;
        const int n0 = nt * BN;
        const int oph = taps.n_per_phase > 0 ? n0 / taps.n_per_phase : 0;
        const int nrow = n0 + (int)rank * (BN / 2);
        for (int tap = 0; tap < taps.ntaps; ++tap) {
          const int ta = t0 + taps.shift[oph][tap];
          const int ph = taps.phase[oph][tap];
          for (int cb = 0; cb < cblocks; ++cb, ++it) {
            const int s = it % n_stages;
            const uint32_t par = (it / n_stages) & 1;
            mbar_wait(&empty[s], par ^ 1);
            uint8_t* st = tiles + (size_t)s * stage_bytes;
            const uint32_t fb = mapa(smem_u32(&full[s]), 0);
            if (kDebugBuild && (epi.debug_skip & 2)) {  // experiment: no operand traffic at all
              if (rank == 0) mbar_expect_tx(&full[s], 0);
              continue;
            }
            if (rank == 0) mbar_expect_tx(&full[s], 2 * stage_bytes);
            const int kcol = (tap * cblocks + cb) * BK;
            tma2_load_4d(&tmA_hi, fb, st, cb * BK, ph, ta, b);
            tma2_load_2d(&tmB_hi, fb, st + A_BYTES, kcol, nrow);
            if (nprod > 1) {
              tma2_load_4d(&tmA_lo, fb, st + A_BYTES + B_BYTES, cb * BK, ph, ta, b);
              tma2_load_2d(&tmB_lo, fb, st + 2 * A_BYTES + B_BYTES, kcol, nrow);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc2(BN);
      int it = 0, ti = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++ti) {
        const int buf = ti & 1;
        mbar_wait(&tmem_empty[buf], ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t par = (it / n_stages) & 1;
          mbar_wait(&full[s], par);
          if (it < 4) ts_mark(epi, 2 + it);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(tiles + (size_t)s * stage_bytes);
          const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + A_BYTES);
          const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
          uint32_t accum = kb > 0;
          if (kDebugBuild && (epi.debug_skip & 4)) {  // experiment: no MMAs
            umma2_commit_mc(&empty[s]);
            continue;
          }
          if (nprod > 1) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) { umma2_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma2_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          }
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) { umma2_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
          umma2_commit_mc(&empty[s]);
        }
        umma2_commit_mc(&tmem_full[buf]);
        if (ti < 4) ts_mark(epi, 8 + ti);
      }
    }
  } else {
    const int quad = warp & 3;
    const uint32_t te0 = mapa(smem_u32(&tmem_empty[0]), 0), te1 = mapa(smem_u32(&tmem_empty[1]), 0);
    int ti = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++ti) {
      const int nt = tile % n_tiles_n, mt = tile / n_tiles_n;
      const int b = mt / m_tiles_per_b;
      const int t_base = (mt % m_tiles_per_b) * (2 * BM) + (int)rank * BM + quad * 32;
      const int n0 = nt * BN;
      const int buf = ti & 1;
      mbar_wait(&tmem_full[buf], (ti >> 1) & 1);
      if (warp == 2 && lane == 0 && ti < 4) ts_mark(epi, 12 + 2 * ti);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      float* stg = epi_stage + (warp - 2) * (32 * 16);
      if (t_base < T) {
        constexpr int QCOLS = BN / 4;  // this warp's column range: [cq * QCOLS, +QCOLS)
        const int cbeg = ((warp - 2) >> 2) * QCOLS;
#pragma unroll 1
        for (int c = cbeg; c < cbeg + QCOLS; c += 16) {
          EpiAux aux;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(kDebugBuild && (epi.debug_skip & 1))) {
            epi_prefetch<MODE>(epi, aux, b, t_base, T, n0 + c, lane);  // in flight during the TMEM load
            if (epi.bias) bias4 = *reinterpret_cast<const float4*>(epi.bias + n0 + c + (lane & 3) * 4);
          }
          uint32_t v[16];
          tmem_ld16_issue(taddr + c, v);
          tmem_ld_wait();
          constexpr int OUTS = MODE == EPI_ROPE ? 1 : MODE == EPI_ROPE_BF16 ? 2 : -1;  // the RoPE flavours write the fp32 / the bf16 QKV buffer only
          if (T - t_base >= 32) epi_block<MODE, OUTS, true>(epi, stg, v, aux, bias4, b, t_base, T, n0 + c, lane);
          else epi_block<MODE, OUTS, false>(epi, stg, v, aux, bias4, b, t_base, T, n0 + c, lane);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (warp == 2 && lane == 0 && ti < 4) ts_mark(epi, 13 + 2 * ti);
      if (lane == 0) mbar_arrive_cluster(buf ? te1 : te0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) ts_mark(epi, 20);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if (threadIdx.x == 32) ts_mark(epi, 21);
}


// =====================================================================================
// Fused MLP: hidden = GELU(a W0^T + b0) and out = hidden W2^T + b2 + residual in ONE persistent launch of the pair
// kernel.  Every cluster first works through its share of the up-projection tiles (problem 0), then through its
// share of the down-projection tiles (problem 1); a down tile of row block mt may start once all up tiles of mt
// have been stored, which the epilogues publish through a per-row-block counter in global memory
// (threadfence + red.release  ->  ld.acquire + fence.proxy.async before the TMA reads the hidden activations).
// What this buys over two launches: the down-projection mainloops overlap the un-hidden last epilogue of the
// up-projection (write-bandwidth bound, ~5 us), one prologue/teardown/launch gap disappears, and the 48 down tiles
// land on the clusters with the least up-projection work.  Deadlock-free because problem-0 tiles never wait and
// precede problem-1 tiles in every cluster's list, and grid <= #SM pairs (all clusters co-resident).
// =====================================================================================
struct LinearProblem {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  CUtensorMap o_hi, o_lo;  // TMA-store views (N, T, B) of the bf16 hi / lo outputs (up projection)
  GemmEpi epi;
  int Cin;        // K
  int n_tiles_n;  // N / BN
  int n_tiles;    // n_tiles_n * (#row blocks)
  // K-split of the down projection (problem 1 only): with ksplit == 2 every output tile is two work items, K halves
  // [0, K/2) and [K/2, K).  The first half's epilogue is `epi` (bias + residual -> h), the second half's is `epi2`
  // (plain -> a partial buffer); the consumer of h adds the partial, always in that order, so results stay
  // deterministic.  Why: 48 down tiles of 24 K-blocks each leave 26 of the 74 clusters idle and put a 20 us mainloop
  // behind the up-projection's last epilogue; 96 half items spread over all clusters and halve that tail.
  int ksplit;
  GemmEpi epi2;
};

// Work list of cluster `cluster_id`: its up-projection tiles (prob 0) first, then down-projection items (prob 1 =
// first / only K part, prob 2 = second K half).
// groups == 1: item j of problem 1 is output tile j / ksplit, so both halves of a tile, and the tiles of the row blocks
// that complete first, come first; a down item waits for ALL up tiles of its row block.
// groups == 2 (ksplit == 2, an even number of up N tiles): the up tiles are walked N-half-major -- every row block's tiles
// of the first half of the hidden columns, then every row block's tiles of the second half -- and the down items K-half-
// major; a down item of K half g waits only for the up tiles of N half g of its row block (counter flags[2 mt + g]).  After
// the first round of up tiles EVERY first-half down item is ready, instead of the items of the first half of the row
// blocks, so fewer clusters sit out the last up epilogues: +1 % (fp32 mode) / +2.8 % (bf16 mode) steps/s, bitwise identical
// results (profiles/r02x_ab_mlp_groups.jsonl).  (K thirds with three groups balance the clusters -- every one ends 45-48 us
// after the start instead of 44-51 -- but the third partial tensor costs what that gains: profiles/r02y_ab_mlp_k3.jsonl.)
__device__ __forceinline__ bool fused_item(int i, int cluster_id, int n_clusters, int n0, int n1, int ksplit, int groups,
                                           int& prob, int& tile) {
  const int mine0 = cluster_id < n0 ? (n0 - cluster_id + n_clusters - 1) / n_clusters : 0;
  if (i < mine0) {
    prob = 0;
    tile = cluster_id + i * n_clusters;
    return true;
  }
  const int j = (n_clusters - 1 - cluster_id) + (i - mine0) * n_clusters;  // lightest clusters take problem 1 first
  if (j < n1 * ksplit) {
    if (groups == 2) {
      prob = 1 + j / n1;
      tile = j % n1;
    } else {
      prob = 1 + (j % ksplit);
      tile = j / ksplit;
    }
    return true;
  }
  return false;
}
// (row block, N tile) of up-projection tile index `tile`; with 2 groups the index runs N-half-major (see fused_item)
__device__ __forceinline__ void up_tile_coords(int tile, int n_tiles_n, int n_tiles, int groups, int& mt, int& nt) {
  if (groups == 2) {
    const int h = n_tiles_n >> 1, per_group = n_tiles >> 1;  // N tiles per half and row block; tiles per half
    const int g = tile >= per_group ? 1 : 0, r = tile - g * per_group;
    mt = r / h;
    nt = g * h + (r - mt * h);
  } else {
    nt = tile % n_tiles_n;
    mt = tile / n_tiles_n;
  }
}

// One more warp than the pair kernel: warp 2 + EPI_WARPS publishes finished up-projection tiles (see below).
constexpr int NUM_THREADS_MLP = NUM_THREADS2 + 32;

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS_MLP, 1)
mlp_fused_tc2_kernel(const __grid_constant__ LinearProblem p0, const __grid_constant__ LinearProblem p1, int T, int nprod,
                     int m_tiles_per_b, int* __restrict__ flags, int flag_need, int flag_groups, unsigned long long* dbg) {
  auto mark = [&](int slot) {  // profiling experiments only: per-cluster %globaltimer marks of the leader CTA
    if (kDebugBuild && dbg && (blockIdx.x & 1) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[(blockIdx.x >> 1) * 16 + slot] = t;
    }
  };
  extern __shared__ uint8_t smem_raw[];
  constexpr int A_BYTES = Smem2<BN>::A_BYTES, B_BYTES = Smem2<BN>::B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;
  const int n_ops = nprod > 1 ? 2 : 1;
  const int stage_bytes = n_ops * (A_BYTES + B_BYTES);
  const int n_stages = nprod > 1 ? Smem2<BN>::stages(3) : Smem2<BN>::stages(1);

  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)n_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* tmem_full = bars + 16;
  uint64_t* tmem_empty = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* pub_bar = bars + 24;  // [2]: "every epilogue warp of this CTA has stored its part of up tile ti" (parity ti & 1)
  float* epi_stage = reinterpret_cast<float*>(bars + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p0.a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p0.b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p1.a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p1.b_hi) : "memory");
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EPI_WARPS);
      mbar_init(&pub_bar[i], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail (common.cuh)
  pdl_trigger();
  if (threadIdx.x == 0) mark(0);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, prob, tile;
      for (int i = 0; fused_item(i, cluster_id, n_clusters, p0.n_tiles, p1.n_tiles, p1.ksplit, flag_groups, prob, tile); ++i) {
        const LinearProblem& p = prob ? p1 : p0;
        int nt, mt;
        if (prob == 0) up_tile_coords(tile, p0.n_tiles_n, p0.n_tiles, flag_groups, mt, nt);
        else { nt = tile % p1.n_tiles_n; mt = tile / p1.n_tiles_n; }
        const int b = mt / m_tiles_per_b;
        const int t0 = (mt % m_tiles_per_b) * (2 * BM) + (int)rank * BM;
        const int nrow = nt * BN + (int)rank * (BN / 2);
        if (prob >= 1) {  // the hidden activations of row block mt must be complete and visible to the async proxy
          const int* flag = flags + (flag_groups == 2 ? 2 * mt + (prob - 1) : mt);
          mark(1);
          int seen;
          long long t_start = clock64();
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
            if (seen < flag_need && clock64() - t_start > 4000000000LL) {
              printf("after_b200: fused MLP dependency wait timeout (cluster %d, row block %d: %d of %d)\n", cluster_id, mt, seen, flag_need);
              __trap();
            }
          } while (seen < flag_need);
          asm volatile("fence.proxy.async;" ::: "memory");
          mark(2);
        }
        const int cblocks = prob ? p.Cin / BK / p1.ksplit : p.Cin / BK;
        const int cb0 = prob == 2 ? cblocks : 0;
        for (int cb = cb0; cb < cb0 + cblocks; ++cb, ++it) {
          const int s = it % n_stages;
          const uint32_t par = (it / n_stages) & 1;
          mbar_wait(&empty[s], par ^ 1);
          uint8_t* st = tiles + (size_t)s * stage_bytes;
          const uint32_t fb = mapa(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * stage_bytes);
          tma2_load_4d(&p.a_hi, fb, st, cb * BK, 0, t0, b);
          tma2_load_2d(&p.b_hi, fb, st + A_BYTES, cb * BK, nrow);
          if (nprod > 1) {
            tma2_load_4d(&p.a_lo, fb, st + A_BYTES + B_BYTES, cb * BK, 0, t0, b);
            tma2_load_2d(&p.b_lo, fb, st + 2 * A_BYTES + B_BYTES, cb * BK, nrow);
          }
        }
        mark(prob ? 4 : 3);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc2(BN);
      int it = 0, prob, tile;
      for (int ti = 0; fused_item(ti, cluster_id, n_clusters, p0.n_tiles, p1.n_tiles, p1.ksplit, flag_groups, prob, tile); ++ti) {
        const int nkb = prob ? p1.Cin / BK / p1.ksplit : p0.Cin / BK;
        const int buf = ti & 1;
        mbar_wait(&tmem_empty[buf], ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t par = (it / n_stages) & 1;
          mbar_wait(&full[s], par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(tiles + (size_t)s * stage_bytes);
          const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + A_BYTES);
          const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
          uint32_t accum = kb > 0;
          if (nprod > 1) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) { umma2_bf16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma2_bf16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          }
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) { umma2_bf16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, accum); accum = 1; }
          umma2_commit_mc(&empty[s]);
        }
        umma2_commit_mc(&tmem_full[buf]);
      }
    }
  } else if (warp >= 2 + EPI_WARPS) {
    // Publisher: when all epilogue warps of this CTA have stored their part of an up-projection tile (pub_bar, CTA-scope
    // release/acquire), one thread makes those stores visible at gpu scope and bumps the row block's counter.  The
    // epilogue warps never wait for that fence: in the in-kernel trace the former "every thread __threadfence(),
    // bar.sync, red.release" sequence held all 16 warps for ~2.5 us after each up tile, on the path every
    // down-projection item waits for (now ~1.2 us until the next accumulator is picked up).  Causality: stores -po->
    // mbarrier.arrive(release.cta) -sw-> try_wait(acquire.cta) -po-> fence.acq_rel.gpu; red.release.gpu -sw-> the
    // consumer's ld.acquire.gpu.
    if (lane == 0) {
      int prob, tile;
      for (int ti = 0; fused_item(ti, cluster_id, n_clusters, p0.n_tiles, p1.n_tiles, p1.ksplit, flag_groups, prob, tile) && prob == 0; ++ti) {
        int mt, nt;
        up_tile_coords(tile, p0.n_tiles_n, p0.n_tiles, flag_groups, mt, nt);
        int* flag = flags + (flag_groups == 2 ? 2 * mt + (nt >= (p0.n_tiles_n >> 1) ? 1 : 0) : mt);
        mbar_wait(&pub_bar[ti & 1], (ti >> 1) & 1);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flag) : "memory");
      }
    }
  } else {
    const int quad = warp & 3;
    const uint32_t te0 = mapa(smem_u32(&tmem_empty[0]), 0), te1 = mapa(smem_u32(&tmem_empty[1]), 0);
    int prob, tile;
    for (int ti = 0; fused_item(ti, cluster_id, n_clusters, p0.n_tiles, p1.n_tiles, p1.ksplit, flag_groups, prob, tile); ++ti) {
      const GemmEpi& pe1 = prob == 2 ? p1.epi2 : p1.epi;
      int nt, mt;
      if (prob == 0) up_tile_coords(tile, p0.n_tiles_n, p0.n_tiles, flag_groups, mt, nt);
      else { nt = tile % p1.n_tiles_n; mt = tile / p1.n_tiles_n; }
      const int b = mt / m_tiles_per_b;
      const int t_base = (mt % m_tiles_per_b) * (2 * BM) + (int)rank * BM + quad * 32;
      const int n0 = nt * BN;
      const int buf = ti & 1;
      mbar_wait(&tmem_full[buf], (ti >> 1) & 1);
      const bool tr = warp == 2 && lane == 0 && ti < 2;
      if (tr) mark(8 + 4 * ti);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      float* stg = epi_stage + (warp - 2) * (32 * 16);
      if (t_base < T) {
        constexpr int QCOLS = BN / 4;
        const int cbeg = ((warp - 2) >> 2) * QCOLS;
#pragma unroll 1
        for (int c = cbeg; c < cbeg + QCOLS; c += 16) {
          EpiAux aux;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (prob == 0) {
            if (p0.epi.bias) bias4 = *reinterpret_cast<const float4*>(p0.epi.bias + n0 + c + (lane & 3) * 4);
          } else {
            epi_prefetch<EPI_PLAIN>(pe1, aux, b, t_base, T, n0 + c, lane);
            if (pe1.bias) bias4 = *reinterpret_cast<const float4*>(pe1.bias + n0 + c + (lane & 3) * 4);
          }
          uint32_t v[16];
          tmem_ld16_issue(taddr + c, v);
          tmem_ld_wait();
          if (tr && c == cbeg) mark(9 + 4 * ti);
          const bool full = T - t_base >= 32;
          if (prob == 0) {  // hidden activations: bf16 hi (+ lo in the 3-product mode), never fp32 (launch_mlp_fused)
            epi_block_gelu_tma(&p0.o_hi, &p0.o_lo, nprod > 1, stg, v, p0.epi.bias, b, t_base, n0 + c, lane);
          } else if (full) {
            epi_block<EPI_PLAIN, -1, true>(pe1, stg, v, aux, bias4, b, t_base, T, n0 + c, lane);
          } else {
            epi_block<EPI_PLAIN>(pe1, stg, v, aux, bias4, b, t_base, T, n0 + c, lane);
          }
          if (tr && c == cbeg) mark(10 + 4 * ti);
        }
      }
      if (tr) mark(11 + 4 * ti);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        // hand the stored tile to the publisher BEFORE releasing the accumulator buffer: a warp can then reach its next
        // arrival on the same pub_bar parity (tile ti + 2, same TMEM buffer) only after every warp has arrived for ti.
        // "Stored" = this warp's TMA stores have COMPLETED (wait_group without .read), ordered before the release by a
        // proxy fence (the consumer reads them back through the async proxy after its acquire).
        if (prob == 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&pub_bar[ti & 1])) : "memory");
        }
        mbar_arrive_cluster(buf ? te1 : te0);
      }
      if (warp == 2 && lane == 0) mark(prob ? 6 : 5);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) mark(7);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace after
