// One persistent kernel for a whole live streaming block of the exported Streamer.sample (after_scripts/export.py:398-416):
// nb_steps Euler steps x (DenoiserV2 forward over 3 CFG rows of T frames against the per-step KV history, CFG combine,
// roll_cache), rows = 3 B T <= 16.
//
// Why: a 12-row step is 28 kernels of a few microseconds of work each; as separate launches (even inside a CUDA graph
// with programmatic dependent launch) a step costs ~266 us, i.e. ~9.5 us per kernel of launch / drain latency
// (profiles/r01c_launches_stream_summary.txt).  Here one CTA per SM stays resident and the 27 phases of a step are
// separated by a grid barrier (one atomic per CTA + a polled acquire load) instead of a kernel boundary.
// Measured (base, B = 1, 4 frames, 8 steps): 2.15 ms as 28 launches per step -> 1.47 ms (first persistent version) -> 1.32 ms
// (165 us per step; the tiny model, with a quarter of the weights, 127 us).  Per-barrier %globaltimer trace of CTA 0 (debug
// build, AFTER_DEBUG_TRACE_STREAM): barrier wait 1.0-1.5 us, phase work A 4.6 / B 5.5 / C 3.5 / D 4.3 us.  What moved it from
// 2.5 ms (first version) down: 512-thread CTAs, output columns interleaved over ALL CTAs (CTA-major numbering put the
// 512 down-projection columns on 32 SMs), all of a column's weights requested in one trip and BEFORE the wait for the
// phase's operand rows (cp.async.bulk copies against an mbarrier), the down projection's K range quartered over the warps
// of a quad (three quarters of the warps had no column), phase A's QKV weights and AdaLN rows requested in front of its
// LayerNorm arithmetic, phase B's LayerNorm tail on a head-less warp that has its AdaLN-c row in registers when the heads
// finish, LayerNorm parameters of all layers resident in shared memory, a release-reduction arrive in the grid barrier,
// chunked in-place roll.  Tried and dropped: requesting a phase's first weights before the BARRIER (register spills at 512
// threads, no gain); a whole 1536-long weight row in registers (12 float4 per lane: spills, 12 us per down phase).
// What is left is ~1 us of barrier and 1-2 L2 round trips per phase, 27 phases per step.
//
// Phases of a step (B = barrier):
//   embed        h0[(b,t)] = GELU(W_in x + b_in)                                          (transformerv2.py:387-391)   B
//   per layer l  A: every CTA normalises all rows itself (LN0 -> AdaLN-t -> LN1; 12 x D values -- cheaper than a phase
//                   of its own), CTA 0 publishes the modulated h, then the CTAs stream W_qkv (one warp per output
//                   column, all rows at once, exact fp32) into the layer's q|k|v slot                                   B
//                B: one CTA per token: banded attention over [history ; block] with RoPE applied on the fly, residual,
//                   LN2 -> AdaLN-c, LN3 -> MLP operand                                        (transformerv2.py:190-236) B
//                C: hidden = GELU(a W0^T + b0)                                                                          B
//                D: h = h + hidden W2^T + b2                                                                            B
//   out-proj     proj = h W_out^T + b_out                                                                               B
//   combine      x += dt * (d_none + g (d_mid + f (d_full - d_mid) - d_none)); roll the KV history of this step         B
// h ping-pongs between two buffers so that a phase never overwrites rows another CTA is still reading.
#pragma once
#include "denoiser_kernels.cuh"

namespace after {

struct StreamLayerDev {
  const float *qkv_w, *mlp0_w, *mlp0_b, *mlp2_w, *mlp2_b, *n1_g, *n1_b, *n3_g, *n3_b;
};
struct StreamNetDev {
  StreamLayerDev layer[8];
  int L, D, HID, C, chunk, window, W, maxN, maxRows, ada_ld;
  const float *pe_wt, *pe_b, *out_w, *out_b;
  const float2* rope_tab;
  const float *adaT, *adaC;
  SeqMap map;
  const float* guidance;
  float *x_state, *h0, *hA, *hB, *qkv_stream, *sk_a, *sk_hid, *proj, *kcache, *vcache;
  size_t cache_slab;        // floats of KV history per diffusion step
  size_t adaC_step_stride;  // floats of the AdaLN-c table per diffusion step
  unsigned* barrier;        // zeroed before the launch
  unsigned long long* dbg;  // -DAFTER_DEBUG builds: %globaltimer at every barrier entry / exit of CTA 0 (first 128 barriers)
};

constexpr int SS_MAXM = 16;

__device__ __forceinline__ void ss_grid_sync(unsigned* ctr, unsigned& target, unsigned nb, unsigned long long* dbg = nullptr) {
  __syncthreads();
  if (kDebugBuild && dbg && blockIdx.x == 0 && threadIdx.x == 0 && target / nb < 128) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[2 * (target / nb)] = t;
  }
  if (threadIdx.x == 0) {
    target += nb;
    // arrive = one fire-and-forget release reduction: it orders this CTA's phase (made visible to thread 0 by the bar.sync
    // above) before the count, and unlike a fence + atomicAdd it does not wait for a round trip before the polling starts
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v < target && clock64() - t0 > 4000000000LL) {  // ~2 s: a broken barrier traps instead of hanging the GPU
        printf("after_b200: streaming-block grid barrier timeout (block %d: %u of %u)\n", blockIdx.x, v, target);
        __trap();
      }
    } while (v < target);
    if (kDebugBuild && dbg && blockIdx.x == 0 && target / nb <= 128) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[2 * (target / nb - 1) + 1] = t;
    }
  }
  __syncthreads();
}

// out[m, n] = act(sum_k As[m, k] W[n, k] + bias[n]) (+ res[m, n]); As in shared memory; one warp per column n.
// NU x 512 B of a column's weights are requested per trip (one trip for K <= NU * 128), and -- through `rows_ready`, called
// once between the first column's first weight requests and their first use -- before the phase's operand rows have
// landed in shared memory, so the weight stream and the row copy overlap.  `pre` (optional): the first column's weights, requested
// by the caller even earlier (phase A asks for its QKV weights before it normalises the rows).
struct SsNoWait { __device__ __forceinline__ void operator()() const {} };
template <int NU>
__device__ __forceinline__ void ss_w_request(float4 (&w)[NU], const float* __restrict__ wr, int K, int lane) {
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int k = lane * 4 + 128 * u;
    w[u] = k < K ? *reinterpret_cast<const float4*>(wr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NU, typename RowsReady>
__device__ __forceinline__ void ss_linear(const float* As, const float* __restrict__ Wm, const float* __restrict__ bias,
                                          const float* res, float* out, int ldo, int M, int N, int K, int gelu,
                                          RowsReady rows_ready, const float4* pre = nullptr) {
  const int lane = threadIdx.x & 31;
  // column n -> CTA n % #CTAs, warp n / #CTAs: the columns (and their weight rows) are spread over ALL SMs first.  (With
  // CTA-major numbering the 512 columns of the down projection landed on the first 32 CTAs, 16 each: 9.9 us per phase in the
  // per-barrier trace of CTA 0 against 4.2 us for the up projection.)
  const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  const int nw = gridDim.x * (blockDim.x >> 5);
  bool first = true;
  for (int n = gw; n < N; n += nw) {
    float acc[SS_MAXM];
#pragma unroll
    for (int m = 0; m < SS_MAXM; ++m) acc[m] = 0.f;
    for (int kb = 0; kb < K; kb += NU * 128) {  // one trip when K <= NU * 128 (every projection but a wide down projection)
      float4 w[NU];
      if (first && pre != nullptr) {
#pragma unroll
        for (int u = 0; u < NU; ++u) w[u] = pre[u];
      } else {
        ss_w_request<NU>(w, Wm + (size_t)n * K + kb, K - kb, lane);
      }
      if (first) { rows_ready(); first = false; }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int k = kb + lane * 4 + 128 * u;
        if (k < K) {
#pragma unroll
          for (int m = 0; m < SS_MAXM; ++m) {
            if (m < M) {
              const float4 a = *reinterpret_cast<const float4*>(As + m * K + k);
              acc[m] = fmaf(a.x, w[u].x, fmaf(a.y, w[u].y, fmaf(a.z, w[u].z, fmaf(a.w, w[u].w, acc[m]))));
            }
          }
        }
      }
    }
    float v = 0.f;
#pragma unroll
    for (int m = 0; m < SS_MAXM; ++m) {
      const float t = warp_sum(acc[m]);
      if (lane == m) v = t;
    }
    if (lane < M) {
      if (bias) v += bias[n];
      if (gelu) v = gelu_erf(v);
      const size_t o = (size_t)lane * ldo + n;
      if (res) v += res[o];
      out[o] = v;
    }
  }
  if (first) rows_ready();  // warps without a column still observe the row copy (every thread polls its barrier phase)
}

// Down projection of a 16-warp CTA: only N = D (512) columns for ~2400 warps, each K = HID (1536) long -- one warp per
// column left three quarters of the warps idle and the busy ones with three L2 round trips.  Here a column's K range is
// split over the 4 consecutive warps of a quad (column = CTA + #CTAs * quad, so the columns still spread over all SMs): every
// warp requests its <= NUQ x 512 B of weights at once, the four partial sums of a row meet in shared memory, in fixed order.
template <int NUQ, typename RowsReady>
__device__ __forceinline__ void ss_linear_kquad(const float* As, const float* __restrict__ Wm, const float* __restrict__ bias,
                                                const float* res, float* out, int ldo, int M, int N, int K, float* red,
                                                RowsReady rows_ready) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x + gridDim.x * (warp >> 2);
  const int part = warp & 3, Kq = K >> 2;
  const bool live = n < N;
  const float* wr = Wm + (size_t)n * K + part * Kq;
  float4 w[NUQ];
#pragma unroll
  for (int u = 0; u < NUQ; ++u) {
    const int k = lane * 4 + 128 * u;
    w[u] = (live && k < Kq) ? *reinterpret_cast<const float4*>(wr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  rows_ready();
  float acc[SS_MAXM];
#pragma unroll
  for (int m = 0; m < SS_MAXM; ++m) acc[m] = 0.f;
  if (live) {
#pragma unroll
    for (int u = 0; u < NUQ; ++u) {
      const int k = lane * 4 + 128 * u;
      if (k < Kq) {
#pragma unroll
        for (int m = 0; m < SS_MAXM; ++m) {
          if (m < M) {
            const float4 a = *reinterpret_cast<const float4*>(As + m * K + part * Kq + k);
            acc[m] = fmaf(a.x, w[u].x, fmaf(a.y, w[u].y, fmaf(a.z, w[u].z, fmaf(a.w, w[u].w, acc[m]))));
          }
        }
      }
    }
  }
  float v = 0.f;
#pragma unroll
  for (int m = 0; m < SS_MAXM; ++m) {
    const float t = warp_sum(acc[m]);
    if (lane == m) v = t;
  }
  if (lane < SS_MAXM) red[warp * SS_MAXM + lane] = v;
  __syncthreads();
  if (live && part == 0 && lane < M) {
    v = ((red[warp * SS_MAXM + lane] + red[(warp + 1) * SS_MAXM + lane]) + red[(warp + 2) * SS_MAXM + lane]) + red[(warp + 3) * SS_MAXM + lane];
    if (bias) v += bias[n];
    const size_t o = (size_t)lane * ldo + n;
    if (res) v += res[o];
    out[o] = v;
  }
}

// Operand rows of a phase (up to 96 KB) into shared memory with ONE bulk copy per 16 KB against an mbarrier: a loop of
// per-thread float4 loads paid one L2 round trip per iteration (9 iterations for the 74 KB of the down projection: ~4 us
// of the 7 us that phase took in the per-barrier trace).  Split in two so that the consumer's weight requests go out between
// the issue and the wait.
__device__ __forceinline__ void ss_rows_issue(float* As, const float* __restrict__ A, int n, uint64_t* bar) {
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
  __syncthreads();  // every warp is done with the previous contents of As
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)n * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(As);
    const char* src = reinterpret_cast<const char*>(A);
    for (uint32_t off = 0; off < bytes; off += 16384) {
      const uint32_t sz = min(16384u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + off), "l"(src + off), "r"(sz), "r"(bar_s) : "memory");
    }
  }
}
struct SsRowsWait {
  uint64_t* bar;
  unsigned* parity;
  __device__ __forceinline__ void operator()() const {
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(bar_s), "r"(*parity) : "memory");
    }
    *parity ^= 1u;
  }
};

// NT threads per CTA: 512 where the attention phase's registers allow it (MAXK <= 12), so that the 3 D / 3 x D output columns
// of the big projections are at most one per warp on 148 SMs; else 256.
template <int NH, int MAXK, int NT>
__global__ void __launch_bounds__(NT, 1)
stream_block_kernel(const __grid_constant__ StreamNetDev net, int B, int T, int nb_steps) {
  constexpr int D = NH * 64;
  constexpr int NV = D / 32;
  constexpr int NWARPS = NT / 32;
  constexpr int NU_D = D / 128;   // weight slices per lane of a K = D projection
  extern __shared__ __align__(16) float ss_smem[];  // [M][max(D, HID)] operand rows; the attention phase uses its head
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ float kq_red[NWARPS * SS_MAXM];  // partial sums of the K-quartered down projection
  unsigned load_parity = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&load_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const int M = 3 * B * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned nb = gridDim.x;
  // LayerNorm affine parameters of every layer (norm1 / norm3 weight and bias: L x 4 x D floats) live in shared memory for
  // the whole block behind the operand rows: the row phases read them at shared-memory latency
  float* ln_s = ss_smem + (size_t)SS_MAXM * max(D, net.HID);
  for (int i = threadIdx.x; i < net.L * 4 * (D / 4); i += blockDim.x) {
    const int e = (i % (D / 4)) * 4, which = (i / (D / 4)) & 3, l = i / (D / 4) / 4;
    const StreamLayerDev& ly = net.layer[l];
    const float* src = which == 0 ? ly.n1_g : which == 1 ? ly.n1_b : which == 2 ? ly.n3_g : ly.n3_b;
    *reinterpret_cast<float4*>(ln_s + (size_t)(l * 4 + which) * D + e) = *reinterpret_cast<const float4*>(src + e);
  }
  __syncthreads();
  unsigned target = 0;
  const float g = net.guidance[0], fct = net.guidance[1], dt = net.guidance[2];

  for (int s = 0; s < nb_steps; ++s) {
    const float* adaC = net.adaC + (size_t)s * net.adaC_step_stride;
    float* kc_s = net.kcache + (size_t)s * net.cache_slab;
    float* vc_s = net.vcache + (size_t)s * net.cache_slab;
    // ---- embed: B T rows x D outputs, one output per thread; the B T x C inputs staged in shared memory
    {
      const int nx = B * T * net.C;
      for (int i = threadIdx.x; i < nx; i += blockDim.x) {  // ss_smem[(b, t), c]
        const int c = i % net.C, r = i / net.C;
        const int b = r / T, t = r - b * T;
        ss_smem[i] = net.x_state[((size_t)b * net.C + c) * T + t];
      }
      __syncthreads();
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * T * D; i += nb * blockDim.x) {
        const int d = i % D, r = i / D;
        float acc = net.pe_b[d];
        const float* xs = ss_smem + (size_t)r * net.C;
#pragma unroll 16
        for (int c = 0; c < net.C; ++c) acc = fmaf(xs[c], net.pe_wt[(size_t)c * D + d], acc);
        net.h0[(size_t)r * D + d] = gelu_erf(acc);
      }
    }
    ss_grid_sync(net.barrier, target, nb, net.dbg);

    for (int l = 0; l < net.L; ++l) {
      const StreamLayerDev& ly = net.layer[l];
      float* qkv_l = net.qkv_stream + (size_t)l * net.maxRows * 3 * D;
      // ---- phase A: LN0 -> AdaLN-t -> (publish h) -> LN1 -> operand rows in shared memory, then the QKV projection
      // (this warp's QKV weight row is requested before the rows are normalised, the AdaLN-t / LN1 parameters of a row
      // together with the row itself: one L2 round trip in front of the arithmetic instead of three)
      float4 wq[NU_D];
      {
        const int n0 = warp * (int)nb + (int)blockIdx.x;
        if (n0 < 3 * D) ss_w_request<NU_D>(wq, ly.qkv_w + (size_t)n0 * D, D, lane);
      }
      for (int row = warp; row < M; row += NWARPS) {
        const int n = row / T, t = row - n * T;
        const float* hp = l == 0 ? net.h0 + (size_t)(net.map.src_seq[n] * T + t) * D : net.hA + (size_t)row * D;
        const float* ap = net.adaT + (size_t)(net.map.t_row0[n] + t * net.map.t_stride[n]) * net.ada_ld + l * 2 * D;
        float x[NV];
        float4 al[NV / 4], be[NV / 4];
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const int e = (i * 32 + lane) * 4;
          const float4 v = *reinterpret_cast<const float4*>(hp + e);
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
          al[i] = *reinterpret_cast<const float4*>(ap + e);
          be[i] = *reinterpret_cast<const float4*>(ap + D + e);
        }
        float mean, rstd;
        row_stats<NV>(x, D, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const int e = (i * 32 + lane) * 4;
          x[4 * i + 0] = (x[4 * i + 0] - mean) * rstd * (1.f + al[i].x) + be[i].x;
          x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * (1.f + al[i].y) + be[i].y;
          x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * (1.f + al[i].z) + be[i].z;
          x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * (1.f + al[i].w) + be[i].w;
          if (blockIdx.x == 0)
            *reinterpret_cast<float4*>(net.hB + (size_t)row * D + e) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
        row_stats<NV>(x, D, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const int e = (i * 32 + lane) * 4;
          const float4 gg = *reinterpret_cast<const float4*>(ln_s + (size_t)(l * 4 + 0) * D + e);
          const float4 bb = *reinterpret_cast<const float4*>(ln_s + (size_t)(l * 4 + 1) * D + e);
          float4 o;
          o.x = (x[4 * i + 0] - mean) * rstd * gg.x + bb.x;
          o.y = (x[4 * i + 1] - mean) * rstd * gg.y + bb.y;
          o.z = (x[4 * i + 2] - mean) * rstd * gg.z + bb.z;
          o.w = (x[4 * i + 3] - mean) * rstd * gg.w + bb.w;
          *reinterpret_cast<float4*>(ss_smem + (size_t)row * D + e) = o;
        }
      }
      __syncthreads();
      ss_linear<NU_D>(ss_smem, ly.qkv_w, nullptr, nullptr, qkv_l, 3 * D, M, 3 * D, D, 0, SsNoWait{}, wq);
      ss_grid_sync(net.barrier, target, nb, net.dbg);

      // ---- phase B: one CTA per token (warp = head, lane = dims (2 lane, 2 lane + 1) of it); the row's LayerNorm tail runs
      // on a warp that has no head (when the CTA has one to spare), which requests the AdaLN-c / LN3 parameters while the
      // heads are still busy: the tail then starts from registers instead of two more L2 round trips
      constexpr int LN_WARP = NWARPS > NH ? NH : 0;
      for (int row = blockIdx.x; row < M; row += nb) {
        float* xs = ss_smem;  // [D]
        const int hd = warp;
        const int n = row / T, t = row - n * T;
        float4 t_al[NV / 4], t_be[NV / 4];
        if (warp == LN_WARP && LN_WARP != 0) {
          const float* ap = adaC + (size_t)net.map.c_row[n] * net.ada_ld + l * 2 * D;
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const int e = (i * 32 + lane) * 4;
            t_al[i] = *reinterpret_cast<const float4*>(ap + e);
            t_be[i] = *reinterpret_cast<const float4*>(ap + D + e);
          }
        }
        if (hd < NH) {
          const int W = net.W;
          const int p = W + t, Lk = W + T;
          const int c0 = (p / net.chunk) * net.chunk;
          const int ce = min(c0 + net.chunk, Lk);
          const int ks = min(c0, max(0, p - net.window + 1));
          const int nk = ce - ks;
          const size_t coff = ((size_t)l * net.maxN + n) * W * D + hd * 64 + 2 * lane;
          const float* qrow = qkv_l + (size_t)row * (3 * D) + hd * 64 + 2 * lane;
          const float* kcn = kc_s + coff;
          const float* vcn = vc_s + coff;
          const float* kblk = qkv_l + (size_t)n * T * (3 * D) + D + hd * 64 + 2 * lane;
          const bool rot = lane < 16;
          float2 q = *reinterpret_cast<const float2*>(qrow);
          const float2 r = *reinterpret_cast<const float2*>(net.hB + (size_t)row * D + hd * 64 + 2 * lane);  // residual, with the keys
          if (rot) {
            const float2 cq = net.rope_tab[p * 16 + lane];
            q = make_float2(q.x * cq.x - q.y * cq.y, q.y * cq.x + q.x * cq.y);
          }
          float sc[MAXK];
          float2 vv[MAXK];
#pragma unroll
          for (int j = 0; j < MAXK; ++j) {
            sc[j] = 0.f;
            vv[j] = make_float2(0.f, 0.f);
            if (j < nk) {
              const int kp = ks + j;
              const float* kr = kp < W ? kcn + (size_t)kp * D : kblk + (size_t)(kp - W) * (3 * D);
              const float* vr = kp < W ? vcn + (size_t)kp * D : kblk + D + (size_t)(kp - W) * (3 * D);
              float2 k = *reinterpret_cast<const float2*>(kr);
              vv[j] = *reinterpret_cast<const float2*>(vr);
              if (rot) {
                const float2 cs = net.rope_tab[kp * 16 + lane];
                k = make_float2(k.x * cs.x - k.y * cs.y, k.y * cs.x + k.x * cs.y);
              }
              sc[j] = fmaf(q.x, k.x, q.y * k.y);
            }
          }
#pragma unroll
          for (int j = 0; j < MAXK; ++j) sc[j] = warp_sum(sc[j]) * 0.125f;
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < MAXK; ++j) if (j < nk) m = fmaxf(m, sc[j]);
          float lsum = 0.f;
          float2 o = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < MAXK; ++j) {
            if (j < nk) {
              const float pr = expf(sc[j] - m);
              lsum += pr;
              o.x = fmaf(pr, vv[j].x, o.x);
              o.y = fmaf(pr, vv[j].y, o.y);
            }
          }
          const float inv = 1.0f / lsum;
          *reinterpret_cast<float2*>(xs + hd * 64 + 2 * lane) = make_float2(r.x + o.x * inv, r.y + o.y * inv);
        }
        __syncthreads();
        if (warp == LN_WARP) {
          float x[NV];
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
          }
          if (LN_WARP == 0) {  // no spare warp: the parameters are requested here
            const float* ap = adaC + (size_t)net.map.c_row[n] * net.ada_ld + l * 2 * D;
#pragma unroll
            for (int i = 0; i < NV / 4; ++i) {
              const int e = (i * 32 + lane) * 4;
              t_al[i] = *reinterpret_cast<const float4*>(ap + e);
              t_be[i] = *reinterpret_cast<const float4*>(ap + D + e);
            }
          }
          float mean, rstd;
          row_stats<NV>(x, D, mean, rstd);
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const int e = (i * 32 + lane) * 4;
            x[4 * i + 0] = (x[4 * i + 0] - mean) * rstd * (1.f + t_al[i].x) + t_be[i].x;
            x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * (1.f + t_al[i].y) + t_be[i].y;
            x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * (1.f + t_al[i].z) + t_be[i].z;
            x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * (1.f + t_al[i].w) + t_be[i].w;
            *reinterpret_cast<float4*>(net.hB + (size_t)row * D + e) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
          row_stats<NV>(x, D, mean, rstd);
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const int e = (i * 32 + lane) * 4;
            const float4 gg = *reinterpret_cast<const float4*>(ln_s + (size_t)(l * 4 + 2) * D + e);
            const float4 bb = *reinterpret_cast<const float4*>(ln_s + (size_t)(l * 4 + 3) * D + e);
            float4 ov;
            ov.x = (x[4 * i + 0] - mean) * rstd * gg.x + bb.x;
            ov.y = (x[4 * i + 1] - mean) * rstd * gg.y + bb.y;
            ov.z = (x[4 * i + 2] - mean) * rstd * gg.z + bb.z;
            ov.w = (x[4 * i + 3] - mean) * rstd * gg.w + bb.w;
            *reinterpret_cast<float4*>(net.sk_a + (size_t)row * D + e) = ov;
          }
        }
        __syncthreads();
      }
      ss_grid_sync(net.barrier, target, nb, net.dbg);

      // ---- phase C: MLP up projection + GELU
      ss_rows_issue(ss_smem, net.sk_a, M * D, &load_bar);
      ss_linear<NU_D>(ss_smem, ly.mlp0_w, ly.mlp0_b, nullptr, net.sk_hid, net.HID, M, net.HID, D, 1, SsRowsWait{&load_bar, &load_parity});
      ss_grid_sync(net.barrier, target, nb, net.dbg);

      // ---- phase D: MLP down projection + residual: reads hB, writes hA (the next layer's input)
      ss_rows_issue(ss_smem, net.sk_hid, M * net.HID, &load_bar);
      if (NWARPS == 16 && D <= 4 * (int)nb && (net.HID & 15) == 0)
        ss_linear_kquad<3>(ss_smem, ly.mlp2_w, ly.mlp2_b, net.hB, net.hA, D, M, D, net.HID, kq_red, SsRowsWait{&load_bar, &load_parity});
      else
        ss_linear<4>(ss_smem, ly.mlp2_w, ly.mlp2_b, net.hB, net.hA, D, M, D, net.HID, 0, SsRowsWait{&load_bar, &load_parity});
      ss_grid_sync(net.barrier, target, nb, net.dbg);
    }

    // ---- out projection
    ss_rows_issue(ss_smem, net.hA, M * D, &load_bar);
    ss_linear<NU_D>(ss_smem, net.out_w, net.out_b, nullptr, net.proj, net.C, M, net.C, D, 0, SsRowsWait{&load_bar, &load_parity});
    ss_grid_sync(net.barrier, target, nb, net.dbg);

    // ---- CFG combine + Euler update (model.py:751-759, 777-783) and roll_cache(T, s) (transformerv2.py:167-186)
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = nb * blockDim.x;
    for (int i = gt; i < B * net.C * T; i += gn) {
      const int t = i % T, c = (i / T) % net.C, b = i / (T * net.C);
      const float df = net.proj[((size_t)(b)*T + t) * net.C + c];
      const float dm = net.proj[((size_t)(B + b) * T + t) * net.C + c];
      const float dn = net.proj[((size_t)(2 * B + b) * T + t) * net.C + c];
      const float d = dn + g * (dm + fct * (df - dm) - dn);
      net.x_state[i] = net.x_state[i] + d * dt;
    }
    {
      const int W = net.W, r = min(T, W > 0 ? T : 0);
      const int N3 = 3 * B;
      for (int i = gt; i < net.L * N3 * 2 * D; i += gn) {
        const int d = i % D;
        const int which = (i / D) & 1;
        const int n = (i / (2 * D)) % N3;
        const int l = i / (2 * D * N3);
        float* c = (which ? vc_s : kc_s) + ((size_t)l * net.maxN + n) * W * D;
        const float* last = net.qkv_stream + ((size_t)l * net.maxRows + (size_t)n * T) * (3 * D) + (which ? 2 * D : D);
        // ascending chunks of 8 slots: a chunk's sources (index >= its first slot + r) are all read before any of its
        // slots is written, and earlier chunks only wrote lower slots -- in place, with 8 independent loads in flight
        for (int j0 = 0; j0 < W; j0 += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int src = j0 + u + r;
            if (j0 + u < W) v[u] = src < W ? c[(size_t)src * D + d] : last[(size_t)(src - W) * (3 * D) + d];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (j0 + u < W) c[(size_t)(j0 + u) * D + d] = v[u];
        }
      }
    }
    ss_grid_sync(net.barrier, target, nb, net.dbg);
  }
}

}  // namespace after
