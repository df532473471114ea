// One persistent kernel for a whole live streaming block of the exported Streamer.sample (after_scripts/export.py:398-416):
// nb_steps Euler steps x (DenoiserV2 forward over 3 CFG rows of T frames against the per-step KV history, CFG combine,
// roll_cache), rows = 3 B T <= 16.
//
// Why: a 12-row step is 28 kernels of a few microseconds of work each; as separate launches (even inside a CUDA graph
// with programmatic dependent launch) a step costs ~266 us, i.e. ~9.5 us per kernel of launch / drain latency
// (profiles/r01c_launches_stream_summary.txt).  Here one CTA per SM stays resident and the 27 phases of a step are
// separated by a grid barrier (one atomic per CTA + a polled acquire load) instead of a kernel boundary.
// Measured (base, B = 1, 4 frames, 8 steps): 2.15 ms as 28 launches per step -> 1.47 ms (184 us per step; the tiny model,
// with a quarter of the weights, 139 us).  Per-barrier %globaltimer trace of CTA 0 (debug build,
// AFTER_DEBUG_TRACE_STREAM): barrier wait 0.9-1.8 us, phase work A 5.3 / B 7-8 / C 3.6 / D 4.9 us.  What moved it from
// 2.5 ms (first version) down: 512-thread CTAs, output columns interleaved over ALL CTAs (CTA-major numbering put the
// 512 down-projection columns on 32 SMs), four 512-byte weight loads in flight per warp, operand rows fetched with
// cp.async.bulk instead of a loop of per-thread loads, chunked in-place roll.  Tried and dropped: requesting a phase's
// first weights before the barrier (register spills at 512 threads, no gain).
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
    __threadfence();
    atomicAdd(ctr, 1u);
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

// out[m, n] = act(sum_k As[m, k] W[n, k] + bias[n]) (+ res[m, n]); As in shared memory; one warp per column n
__device__ __forceinline__ void ss_linear(const float* As, const float* __restrict__ Wm, const float* __restrict__ bias,
                                          const float* res, float* out, int ldo, int M, int N, int K, int gelu) {
  const int lane = threadIdx.x & 31;
  // column n -> CTA n % #CTAs, warp n / #CTAs: the columns (and their weight rows) are spread over ALL SMs first.  (With
  // CTA-major numbering the 512 columns of the down projection landed on the first 32 CTAs, 16 each: 9.9 us per phase in the
  // per-barrier trace of CTA 0 against 4.2 us for the up projection.)
  const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int n = gw; n < N; n += nw) {
    float acc[SS_MAXM];
#pragma unroll
    for (int m = 0; m < SS_MAXM; ++m) acc[m] = 0.f;
    const float* wr = Wm + (size_t)n * K;
    // four 512-byte weight loads in flight per warp: with one load per iteration a warp waited a whole L2 round trip per
    // 128 weights and the 57 MB of weights of a step streamed at 0.44 TB/s chip-wide (measured: 266 -> 221 us per step)
    for (int k0 = lane * 4; k0 < K; k0 += 512) {
      float4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[u] = k0 + 128 * u < K ? *reinterpret_cast<const float4*>(wr + k0 + 128 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + 128 * u < K) {
#pragma unroll
          for (int m = 0; m < SS_MAXM; ++m) {
            if (m < M) {
              const float4 a = *reinterpret_cast<const float4*>(As + m * K + k0 + 128 * u);
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
}

// Operand rows of a phase (up to 96 KB) into shared memory with ONE bulk copy per 16 KB against an mbarrier: a loop of
// per-thread float4 loads paid one L2 round trip per iteration (9 iterations for the 74 KB of the down projection: ~4 us
// of the 7 us that phase took in the per-barrier trace).
__device__ __forceinline__ void ss_load_rows(float* As, const float* __restrict__ A, int n, uint64_t* bar, unsigned& parity) {
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
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar_s), "r"(parity) : "memory");
  }
  parity ^= 1u;
}

// NT threads per CTA: 512 where the attention phase's registers allow it (MAXK <= 12), so that the 3 D / 3 x D output columns
// of the big projections are at most one per warp on 148 SMs; else 256.
template <int NH, int MAXK, int NT>
__global__ void __launch_bounds__(NT, 1)
stream_block_kernel(const __grid_constant__ StreamNetDev net, int B, int T, int nb_steps) {
  constexpr int D = NH * 64;
  constexpr int NV = D / 32;
  constexpr int NWARPS = NT / 32;
  extern __shared__ __align__(16) float ss_smem[];  // [M][max(D, HID)] operand rows; the attention phase uses its head
  __shared__ __align__(8) uint64_t load_bar;
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
      for (int row = warp; row < M; row += NWARPS) {
        const int n = row / T, t = row - n * T;
        const float* hp = l == 0 ? net.h0 + (size_t)(net.map.src_seq[n] * T + t) * D : net.hA + (size_t)row * D;
        const float* ap = net.adaT + (size_t)(net.map.t_row0[n] + t * net.map.t_stride[n]) * net.ada_ld + l * 2 * D;
        float x[NV];
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(hp + (i * 32 + lane) * 4);
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        float mean, rstd;
        row_stats<NV>(x, D, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const int e = (i * 32 + lane) * 4;
          const float4 al = *reinterpret_cast<const float4*>(ap + e);
          const float4 be = *reinterpret_cast<const float4*>(ap + D + e);
          x[4 * i + 0] = (x[4 * i + 0] - mean) * rstd * (1.f + al.x) + be.x;
          x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * (1.f + al.y) + be.y;
          x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * (1.f + al.z) + be.z;
          x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * (1.f + al.w) + be.w;
          if (blockIdx.x == 0)
            *reinterpret_cast<float4*>(net.hB + (size_t)row * D + e) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
        row_stats<NV>(x, D, mean, rstd);
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
          const int e = (i * 32 + lane) * 4;
          const float4 gg = *reinterpret_cast<const float4*>(ly.n1_g + e);
          const float4 bb = *reinterpret_cast<const float4*>(ly.n1_b + e);
          float4 o;
          o.x = (x[4 * i + 0] - mean) * rstd * gg.x + bb.x;
          o.y = (x[4 * i + 1] - mean) * rstd * gg.y + bb.y;
          o.z = (x[4 * i + 2] - mean) * rstd * gg.z + bb.z;
          o.w = (x[4 * i + 3] - mean) * rstd * gg.w + bb.w;
          *reinterpret_cast<float4*>(ss_smem + (size_t)row * D + e) = o;
        }
      }
      __syncthreads();
      ss_linear(ss_smem, ly.qkv_w, nullptr, nullptr, qkv_l, 3 * D, M, 3 * D, D, 0);
      ss_grid_sync(net.barrier, target, nb, net.dbg);

      // ---- phase B: one CTA per token (warp = head, lane = dims (2 lane, 2 lane + 1) of it)
      for (int row = blockIdx.x; row < M; row += nb) {
        float* xs = ss_smem;  // [D]
        const int hd = warp;
        const int n = row / T, t = row - n * T;
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
          const float2 r = *reinterpret_cast<const float2*>(net.hB + (size_t)row * D + hd * 64 + 2 * lane);
          *reinterpret_cast<float2*>(xs + hd * 64 + 2 * lane) = make_float2(r.x + o.x * inv, r.y + o.y * inv);
        }
        __syncthreads();
        if (warp == 0) {
          float x[NV];
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
          }
          float mean, rstd;
          row_stats<NV>(x, D, mean, rstd);
          const float* ap = adaC + (size_t)net.map.c_row[n] * net.ada_ld + l * 2 * D;
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const int e = (i * 32 + lane) * 4;
            const float4 al = *reinterpret_cast<const float4*>(ap + e);
            const float4 be = *reinterpret_cast<const float4*>(ap + D + e);
            x[4 * i + 0] = (x[4 * i + 0] - mean) * rstd * (1.f + al.x) + be.x;
            x[4 * i + 1] = (x[4 * i + 1] - mean) * rstd * (1.f + al.y) + be.y;
            x[4 * i + 2] = (x[4 * i + 2] - mean) * rstd * (1.f + al.z) + be.z;
            x[4 * i + 3] = (x[4 * i + 3] - mean) * rstd * (1.f + al.w) + be.w;
            *reinterpret_cast<float4*>(net.hB + (size_t)row * D + e) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
          row_stats<NV>(x, D, mean, rstd);
#pragma unroll
          for (int i = 0; i < NV / 4; ++i) {
            const int e = (i * 32 + lane) * 4;
            const float4 gg = *reinterpret_cast<const float4*>(ly.n3_g + e);
            const float4 bb = *reinterpret_cast<const float4*>(ly.n3_b + e);
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
      ss_load_rows(ss_smem, net.sk_a, M * D, &load_bar, load_parity);
      ss_linear(ss_smem, ly.mlp0_w, ly.mlp0_b, nullptr, net.sk_hid, net.HID, M, net.HID, D, 1);
      ss_grid_sync(net.barrier, target, nb, net.dbg);

      // ---- phase D: MLP down projection + residual: reads hB, writes hA (the next layer's input)
      ss_load_rows(ss_smem, net.sk_hid, M * net.HID, &load_bar, load_parity);
      ss_linear(ss_smem, ly.mlp2_w, ly.mlp2_b, net.hB, net.hA, D, M, D, net.HID, 0);
      ss_grid_sync(net.barrier, target, nb, net.dbg);
    }

    // ---- out projection
    ss_load_rows(ss_smem, net.hA, M * D, &load_bar, load_parity);
    ss_linear(ss_smem, net.out_w, net.out_b, nullptr, net.proj, net.C, M, net.C, D, 0);
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
