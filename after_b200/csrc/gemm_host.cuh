// Host-side launchers for the tap-GEMMs in gemm.cuh: TMA descriptor creation, tile selection, dispatch between the
// tcgen05 / SIMT / naive kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <map>
#include <tuple>
#include "context.cuh"
#include "gemm.cuh"

namespace after {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    AFTER_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    AFTER_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, -2, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Weights: 2-D bf16 row-major [rows, cols], box = [box_rows, 64 cols], 128-byte swizzle.
inline CUtensorMap make_tmap_weight(const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AFTER_REQUIRE(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled(weight) failed (" + std::to_string((int)r) + ")");
  return m;
}

// Activations: 4-D bf16 (C, P, T, B) with C fastest; box (64, 1, 128, 1).  Frames outside [0, T) are zero-filled
// by the TMA unit, which is exactly the zero padding of the convolutions.
// bstride: elements between streams (0: densely packed, C * P * T).
inline CUtensorMap make_tmap_act(const void* ptr, int C, int P, int T, int B, size_t bstride = 0) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)P, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * P * 2, (cuuint64_t)(bstride ? bstride : (size_t)C * P * T) * 2};
  cuuint32_t box[4] = {(cuuint32_t)tc::BK, 1, (cuuint32_t)tc::BM, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AFTER_REQUIRE(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled(activation) failed (" + std::to_string((int)r) + ")");
  return m;
}

// Output of a GEMM epilogue as a TMA *store* target: 3-D bf16 (N, T, B) with N fastest, box (16 columns, 32 frames, 1), no
// swizzle: one epilogue warp's half-block.  Frames beyond T are clipped by the TMA unit (ragged last row block).
inline CUtensorMap make_tmap_store16(const void* ptr, int N, int T, int B) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)N * T * 2};
  cuuint32_t box[3] = {16, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AFTER_REQUIRE(r == CUDA_SUCCESS, -2, "cuTensorMapEncodeTiled(store) failed (" + std::to_string((int)r) + ")");
  return m;
}

inline int pick_bn(int N, int n_per_phase = 0) {
  for (int bn : {128, 64, 32})
    if (N % bn == 0 && (n_per_phase == 0 || n_per_phase % bn == 0)) return bn;
  return 32;
}

// An fp32 weight matrix [N, K] (K = ntaps * Cin, tap-major) with its optional bf16 hi/lo split + TMA maps.
struct GemmWeight {
  float* w = nullptr;     // fp32 [N, K]
  float* bias = nullptr;  // [N] or null
  __nv_bfloat16 *hi = nullptr, *lo = nullptr;
  CUtensorMap map_hi{}, map_lo{};
  int N = 0, K = 0, Cin = 0, bn = 0;
  TapTable taps;
  bool tc_ok = false;
  // CTA-pair kernel: 256 x bn2 tiles, each CTA loads bn2/2 rows of W
  CUtensorMap map2_hi{}, map2_lo{};
  int bn2 = 0;
  bool tc2_ok = false;
};

inline int pick_bn2(int N, int n_per_phase, int cap = 256) {
  static int env_cap = -1;  // AFTER_BN2_CAP (debug builds): widest CTA-pair tile for every weight, for A/B runs
  if (env_cap < 0) { const char* e = debug_env("AFTER_BN2_CAP"); env_cap = e ? atoi(e) : 256; }
  for (int bn : {256, 128, 64})
    if (bn <= cap && bn <= env_cap && N % bn == 0 && (n_per_phase == 0 || n_per_phase % bn == 0)) return bn;
  return 0;
}
inline bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = debug_env("AFTER_GEMM");
    v = (e && std::string(e) == "1cta") ? 0 : 1;
  }
  return v == 1;
}

// A GEMM A-operand buffer: fp32 (SIMT modes) or bf16 hi/lo (tcgen05 modes); the 4-D maps depend on the view
// (C, P, T, B) and are cached per view.
struct ActOperand {
  float* f32 = nullptr;
  __nv_bfloat16 *hi = nullptr, *lo = nullptr;
  size_t capacity = 0;  // elements
  struct Maps { CUtensorMap hi, lo; };
  std::map<std::tuple<int, int, int, int>, Maps> cache;
  size_t bstride = 0;  // elements between streams (0: densely packed); set once for a streaming conv's persistent operand
  std::map<std::tuple<int, int, int>, Maps> store_cache;  // TMA-store views (N, T, B) of the same buffers
  const Maps& store_maps(int N, int T, int B) {
    auto key = std::make_tuple(N, T, B);
    auto it = store_cache.find(key);
    if (it == store_cache.end()) {
      AFTER_REQUIRE((size_t)N * T * B <= capacity, AFTER_EINVAL, "activation view exceeds the operand buffer");
      Maps m;
      m.hi = make_tmap_store16(hi, N, T, B);
      m.lo = make_tmap_store16(lo ? lo : hi, N, T, B);
      it = store_cache.emplace(key, m).first;
    }
    return it->second;
  }
  const Maps& maps(int C, int P, int T, int B) {
    auto key = std::make_tuple(C, P, T, B);
    auto it = cache.find(key);
    if (it == cache.end()) {
      const size_t per = bstride ? bstride : (size_t)C * P * T;
      AFTER_REQUIRE((size_t)C * P * T <= per && per * B <= capacity, AFTER_EINVAL, "activation view exceeds the operand buffer");
      Maps m;
      m.hi = make_tmap_act(hi, C, P, T, B, bstride);
      m.lo = make_tmap_act(lo ? lo : hi, C, P, T, B, bstride);
      it = cache.emplace(key, m).first;
    }
    return it->second;
  }
};

// fp32 -> bf16 hi/lo split of a whole buffer (weights at finalize time)
__global__ void split_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 h, l;
  split_bf16(in[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

// Upload a [N, ntaps*Cin] fp32 weight (+bias) and, in the tensor-core modes, its bf16 split and TMA maps.
inline void build_gemm_weight(GemmWeight& gw, Arena& arena, const std::vector<float>& w, const float* bias_host, int N,
                              int Cin, const TapTable& taps, bool tc_mode) {
  gw.N = N;
  gw.Cin = Cin;
  gw.taps = taps;
  gw.K = taps.ntaps * Cin;
  AFTER_REQUIRE((size_t)N * gw.K == w.size(), AFTER_ESHAPE, "weight matrix size mismatch");
  gw.w = arena.upload(w);
  gw.bias = nullptr;
  if (bias_host) {
    std::vector<float> b(bias_host, bias_host + N);
    gw.bias = arena.upload(b);
  }
  gw.bn = pick_bn(N, taps.n_per_phase);
  gw.tc_ok = tc_mode && Cin % tc::BK == 0 && N % 32 == 0 && (taps.n_per_phase == 0 || taps.n_per_phase % gw.bn == 0);
  if (gw.tc_ok) {
    const size_t n = w.size();
    gw.hi = arena.alloc<__nv_bfloat16>(n);
    gw.lo = arena.alloc<__nv_bfloat16>(n);
    split_bf16_kernel<<<(unsigned)((n + 255) / 256), 256>>>(gw.w, gw.hi, gw.lo, n);
    AFTER_CUDA_CHECK(cudaGetLastError());
    gw.map_hi = make_tmap_weight(gw.hi, N, gw.K, gw.bn);
    gw.map_lo = make_tmap_weight(gw.lo, N, gw.K, gw.bn);
    gw.bn2 = pick_bn2(N, taps.n_per_phase);
    gw.tc2_ok = gw.bn2 > 0;
    if (gw.tc2_ok) {
      gw.map2_hi = make_tmap_weight(gw.hi, N, gw.K, gw.bn2 / 2);
      gw.map2_lo = make_tmap_weight(gw.lo, N, gw.K, gw.bn2 / 2);
    }
  }
}

inline void alloc_operand(ActOperand& op, Arena& arena, size_t elems, bool tc_mode, bool need_f32) {
  op.capacity = elems;
  if (tc_mode) {
    op.hi = arena.alloc<__nv_bfloat16>(elems);
    op.lo = arena.alloc<__nv_bfloat16>(elems);
  }
  if (!tc_mode || need_f32) op.f32 = arena.alloc<float>(elems);
}

template <int BN>
inline void launch_tap_gemm_tc_bn(const ActOperand::Maps& am, const GemmWeight& W, const GemmEpi& epi, int B, int T,
                                  int nprod, cudaStream_t st) {
  static bool attr_set = false;
  const int smem = tc::Smem<BN>::total(nprod > 1 ? 3 : 1);
  if (!attr_set) {
    const int mx = std::max(tc::Smem<BN>::total(3), tc::Smem<BN>::total(1));
    AFTER_CUDA_CHECK(cudaFuncSetAttribute(tc::tap_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    attr_set = true;
  }
  dim3 grid(W.N / BN, ceil_div(T, tc::BM), B);
  launch_k(tc::tap_gemm_tc_kernel<BN>, grid, dim3(tc::NUM_THREADS), (size_t)smem, st, am.hi, am.lo, W.map_hi, W.map_lo, epi, W.taps, T,
           W.Cin, nprod);
}

template <int BN, int MODE>
inline void launch_tap_gemm_tc2_bn(const ActOperand::Maps& am, const GemmWeight& W, const GemmEpi& epi, int B, int T,
                                   int nprod, cudaStream_t st) {
  static bool attr_set = false;
  static int n_pairs = 0;
  const int smem = tc::Smem2<BN>::total(nprod > 1 ? 3 : 1);
  if (!attr_set) {
    const int mx = std::max(tc::Smem2<BN>::total(3), tc::Smem2<BN>::total(1));
    AFTER_CUDA_CHECK(cudaFuncSetAttribute(tc::tap_gemm_tc2_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    int dev = 0, sms = 0;
    AFTER_CUDA_CHECK(cudaGetDevice(&dev));
    AFTER_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    n_pairs = std::max(1, sms / 2);
    attr_set = true;
  }
  const int n_tiles_n = W.N / BN;
  const int m_tiles_per_b = ceil_div(T, 2 * tc::BM);
  const int n_tiles = n_tiles_n * m_tiles_per_b * B;
  const int clusters = std::min(n_tiles, n_pairs);
  launch_k(tc::tap_gemm_tc2_kernel<BN, MODE>, dim3(2 * clusters), dim3(tc::NUM_THREADS2), (size_t)smem, st, am.hi, am.lo, W.map2_hi,
           W.map2_lo, epi, W.taps, T, W.Cin, nprod, n_tiles_n, m_tiles_per_b, n_tiles);
}

// cudaFuncSetAttribute must not first run inside a stream capture: touch every instantiation once up front.
inline void tc_init_kernels() {
  static bool done = false;
  if (done) return;
  auto set = [](const void* f, int bytes) {
    AFTER_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  };
  set((const void*)tc::tap_gemm_tc_kernel<128>, std::max(tc::Smem<128>::total(3), tc::Smem<128>::total(1)));
  set((const void*)tc::tap_gemm_tc_kernel<64>, std::max(tc::Smem<64>::total(3), tc::Smem<64>::total(1)));
  set((const void*)tc::tap_gemm_tc_kernel<32>, std::max(tc::Smem<32>::total(3), tc::Smem<32>::total(1)));
#define AFTER_SET_TC2(BN)                                                                                          \
  set((const void*)tc::tap_gemm_tc2_kernel<BN, tc::EPI_PLAIN>, std::max(tc::Smem2<BN>::total(3), tc::Smem2<BN>::total(1))); \
  set((const void*)tc::tap_gemm_tc2_kernel<BN, tc::EPI_ROPE>, std::max(tc::Smem2<BN>::total(3), tc::Smem2<BN>::total(1)));  \
  set((const void*)tc::tap_gemm_tc2_kernel<BN, tc::EPI_GELU>, std::max(tc::Smem2<BN>::total(3), tc::Smem2<BN>::total(1)));
  AFTER_SET_TC2(256)
  set((const void*)tc::tap_gemm_tc2_kernel<256, tc::EPI_ROPE_BF16>, std::max(tc::Smem2<256>::total(3), tc::Smem2<256>::total(1)));
  set((const void*)tc::mlp_fused_tc2_kernel<256>, std::max(tc::Smem2<256>::total(3), tc::Smem2<256>::total(1)));
  AFTER_SET_TC2(128)
  AFTER_SET_TC2(64)
#undef AFTER_SET_TC2
  done = true;
}

// Fused MLP launch (see mlp_fused_tc2_kernel).  Returns false when the shapes do not fit the fused kernel (the caller
// then issues the two GEMMs separately).  `flags`: one zeroed int per 256-row block.
inline bool use_fused_mlp() {
  static int v = -1;
  if (v < 0) {
    const char* e = debug_env("AFTER_FUSED_MLP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// The down projection's K range is split in two work items per tile (see LinearProblem::ksplit); AFTER_MLP_KSPLIT=1
// turns that off (+3.7 % steps/s at base B=8 with it, profiles/r01b_ab_knobs.jsonl).
inline int mlp_ksplit() {
  static int v = -1;
  if (v < 0) {
    const char* e = debug_env("AFTER_MLP_KSPLIT");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}
// `partial`: when non-null and the K-split is on, receives the second K half of the down projection ([B*T, ldo] fp32,
// no bias / residual); the caller adds it to epi1.out_f32 when it next reads that tensor.  *used_partial says whether
// it was written.
inline int mlp_flag_groups() {  // AFTER_MLP_FLAG_GROUPS=1 (debug builds): one dependency counter per row block, for A/B runs
  static int v = -1;
  if (v < 0) { const char* e = debug_env("AFTER_MLP_FLAG_GROUPS"); v = (e && e[0] == '1') ? 1 : 2; }
  return v;
}
inline bool launch_mlp_fused(ActOperand& a_in, const GemmWeight& W0, GemmEpi epi0, ActOperand& hid, const GemmWeight& W2,
                             GemmEpi epi1, int B, int T, int precision, int* flags, cudaStream_t st,
                             float* partial = nullptr, bool* used_partial = nullptr) {
  constexpr int BN = 256;
  if (used_partial) *used_partial = false;
  if (precision == AFTER_PRECISION_FP32_SIMT || !use_pair_kernel() || !use_fused_mlp()) return false;
  if (!W0.tc2_ok || !W2.tc2_ok || W0.bn2 != BN || W2.bn2 != BN || W0.taps.ntaps != 1 || W2.taps.ntaps != 1) return false;
  if (W0.N != W2.Cin || epi0.out_hi != hid.hi || epi0.out_hi == nullptr || epi0.out_f32 != nullptr) return false;
  static int n_pairs = 0;
  if (!n_pairs) {
    int dev = 0, sms = 0;
    AFTER_CUDA_CHECK(cudaGetDevice(&dev));
    AFTER_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    n_pairs = std::max(1, sms / 2);
  }
  const int nprod = precision == AFTER_PRECISION_BF16 ? 1 : 3;
  const int m_tiles_per_b = ceil_div(T, 2 * tc::BM);
  const int n_m_tiles = m_tiles_per_b * B;
  epi0.bias = W0.bias;
  epi1.bias = W2.bias;
  {
    const char* e = debug_env("AFTER_DEBUG_SKIP_EPILOGUE");
    if (e) epi0.debug_skip = epi1.debug_skip = atoi(e);
  }
  tc::LinearProblem p0, p1;
  const ActOperand::Maps& m0 = a_in.maps(W0.Cin, 1, T, B);
  const ActOperand::Maps& m1 = hid.maps(W2.Cin, 1, T, B);
  p0.a_hi = m0.hi; p0.a_lo = m0.lo; p0.b_hi = W0.map2_hi; p0.b_lo = W0.map2_lo; p0.epi = epi0; p0.Cin = W0.Cin;
  p0.n_tiles_n = W0.N / BN; p0.n_tiles = p0.n_tiles_n * n_m_tiles;
  p1.a_hi = m1.hi; p1.a_lo = m1.lo; p1.b_hi = W2.map2_hi; p1.b_lo = W2.map2_lo; p1.epi = epi1; p1.Cin = W2.Cin;
  p1.n_tiles_n = W2.N / BN; p1.n_tiles = p1.n_tiles_n * n_m_tiles;
  p0.ksplit = 1; p1.ksplit = 1;
  {  // the up projection's epilogue stores its bf16 hi / lo tiles with TMA (epi_block_gelu_tma)
    const ActOperand::Maps& sm = hid.store_maps(W0.N, T, B);
    p0.o_hi = sm.hi; p0.o_lo = sm.lo;
    p1.o_hi = sm.hi; p1.o_lo = sm.lo;  // unused
  }
  // only when the down projection has too few tiles to fill the machine (48 tiles on 74 CTA pairs at base B=8): with
  // many waves of tiles the split buys no balance and costs one extra write + read of the partial tensor
  if (partial && mlp_ksplit() == 2 && !epi1.out_hi && epi1.out_f32 && (W2.Cin / tc::BK) % 2 == 0 && p1.n_tiles < 2 * n_pairs) {
    p1.ksplit = 2;
    p1.epi2 = GemmEpi{};
    p1.epi2.out_f32 = partial;
    p1.epi2.ldo = epi1.ldo;
    p1.epi2.debug_skip = epi1.debug_skip;
    if (used_partial) *used_partial = true;
  }
  const int clusters = std::min(std::max(p0.n_tiles, p1.n_tiles * p1.ksplit), n_pairs);
  const double rows = (double)B * T;
  ProfScope prof(KC_MLP_FUSED, st, 2.0 * rows * ((double)W0.N * W0.K + (double)W2.N * W2.K),
                 rows * (W0.Cin * 4.0 + W0.N * 4.0 * 2 + W2.N * 8.0));
  const int smem = tc::Smem2<BN>::total(nprod > 1 ? 3 : 1);
  static int trace = -1;
  static unsigned long long* dbg = nullptr;
  if (trace < 0) {
    const char* e = debug_env("AFTER_DEBUG_TRACE_MLP");
    trace = e ? atoi(e) : 0;  // trace the n-th fused launch (1-based), run with AFTER_NO_GRAPH=1
    if (trace > 0) AFTER_CUDA_CHECK(cudaMalloc(&dbg, 128 * 16 * sizeof(unsigned long long)));
  }
  unsigned long long* dbg_now = nullptr;
  if (trace > 0 && --trace == 0) {
    dbg_now = dbg;
    AFTER_CUDA_CHECK(cudaMemsetAsync(dbg, 0, 128 * 16 * sizeof(unsigned long long), st));
  }
  // dependency counters per (row block, half of the hidden columns) when the down projection is K-split in two and the up
  // projection has an even number of N tiles: a K half of a down item then waits for its half of the up tiles only, and the
  // work list runs half-major (fused_item); else one counter per row block.  flags holds 2 ints per row block either way.
  const int groups = (p1.ksplit == 2 && p0.n_tiles_n % 2 == 0 && mlp_flag_groups() == 2) ? 2 : 1;
  launch_k(tc::mlp_fused_tc2_kernel<BN>, dim3(2 * clusters), dim3(tc::NUM_THREADS_MLP), (size_t)smem, st, p0, p1, T, nprod,
           m_tiles_per_b, flags, 2 * p0.n_tiles_n / groups, groups, dbg_now);
  AFTER_COUNT_LAUNCH();
  if (dbg_now) {
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<unsigned long long> hbuf(128 * 16);
    AFTER_CUDA_CHECK(cudaMemcpy(hbuf.data(), dbg, hbuf.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < clusters; ++c) if (hbuf[c * 16]) t0 = std::min(t0, hbuf[c * 16]);
    const char* names[16] = {"start", "p1_wait_begin", "p1_wait_end", "p0_loads_done", "p1_loads_done", "p0_epi_done", "p1_epi_done", "end",
                             "e0_acc_ready", "e0_first_ld", "e0_first_blk", "e0_done", "e1_acc_ready", "e1_first_ld", "e1_first_blk", "e1_done"};
    fprintf(stderr, "fused MLP trace (%d clusters; ns since first cluster start):\n", clusters);
    for (int c = 0; c < clusters; c += (clusters > 16 ? clusters / 12 : 1)) {
      fprintf(stderr, "  cluster %3d:", c);
      for (int k = 0; k < 16; ++k) fprintf(stderr, " %s=%lld", names[k], hbuf[c * 16 + k] ? (long long)(hbuf[c * 16 + k] - t0) : -1LL);
      fprintf(stderr, "\n");
    }
    trace = 0;
  }
  return true;
}

void gn_stats_launch(const float* x, double* stats, int B, int T, int C, int groups, cudaStream_t st);

// Dispatch.  `precision` is the handle's arithmetic mode.  A: the operand (B, T, P, Cin).  The output is
// (B, T, N) rows of epi.ldo floats.
// T_in: input rows per stream when they differ from the T output rows (a streaming conv reads its cached left context
// in front of the new frames; 0 = T); the operand's stream stride is A.bstride.
inline void tap_gemm(ActOperand& A, int B, int T, int P, const GemmWeight& W, GemmEpi epi, int precision,
                     cudaStream_t st, int T_in = 0) {
  if (T_in <= 0) T_in = T;
  const size_t a_bstride = A.bstride ? A.bstride : (size_t)T_in * P * W.Cin;
  const bool tc_mode = precision != AFTER_PRECISION_FP32_SIMT;
  epi.bias = W.bias;
  {
    static int skip = -1;
    if (skip < 0) { const char* e = debug_env("AFTER_DEBUG_SKIP_EPILOGUE"); skip = e ? atoi(e) : 0; }
    epi.debug_skip = skip;
  }
  // algorithmic work of this launch: 2*M*N*K flops; operand read once + result written once (+ residual read)
  const double rows = (double)B * T;
  const double g_flops = 2.0 * rows * W.N * W.K;
  const double g_bytes = rows * P * W.Cin * 4.0 + rows * W.N * 4.0 * (epi.res ? 2.0 : 1.0);
  ProfScope prof(tc_mode && W.tc_ok ? KC_TAP_GEMM_TC : KC_TAP_GEMM_SIMT, st, g_flops, g_bytes);
  if (tc_mode && W.tc_ok) {
    AFTER_REQUIRE(A.hi != nullptr, AFTER_ESTATE, "operand has no bf16 copy");
    const int nprod = precision == AFTER_PRECISION_BF16 ? 1 : 3;
    const ActOperand::Maps& am = A.maps(W.Cin, P, T_in, B);
    const bool flavour_ok = (!epi.rope || (!epi.bias && !epi.res && !epi.gelu && !epi.stats)) &&
                            (!epi.gelu || (!epi.res && !epi.stats)) && (!epi.stats || epi.stat_cpg % 8 == 0);
    if (W.tc2_ok && use_pair_kernel() && flavour_ok) {
      // epilogue flavour (compile-time specialisations; anything a flavour cannot express falls back to the 1-CTA kernel)
#define AFTER_LAUNCH_TC2(MODE)                                                                  \
  do {                                                                                          \
    if (W.bn2 == 256) launch_tap_gemm_tc2_bn<256, MODE>(am, W, epi, B, T, nprod, st);           \
    else if (W.bn2 == 128) launch_tap_gemm_tc2_bn<128, MODE>(am, W, epi, B, T, nprod, st);      \
    else launch_tap_gemm_tc2_bn<64, MODE>(am, W, epi, B, T, nprod, st);                         \
  } while (0)
      if (epi.rope && epi.out_hi && !epi.out_f32 && !epi.out_lo && W.bn2 == 256)
        launch_tap_gemm_tc2_bn<256, tc::EPI_ROPE_BF16>(am, W, epi, B, T, nprod, st);
      else if (epi.rope) AFTER_LAUNCH_TC2(tc::EPI_ROPE);
      else if (epi.gelu) AFTER_LAUNCH_TC2(tc::EPI_GELU);
      else AFTER_LAUNCH_TC2(tc::EPI_PLAIN);
#undef AFTER_LAUNCH_TC2
      AFTER_COUNT_LAUNCH();
      return;
    }
    if (W.bn == 128) launch_tap_gemm_tc_bn<128>(am, W, epi, B, T, nprod, st);
    else if (W.bn == 64) launch_tap_gemm_tc_bn<64>(am, W, epi, B, T, nprod, st);
    else launch_tap_gemm_tc_bn<32>(am, W, epi, B, T, nprod, st);
    AFTER_COUNT_LAUNCH();
    return;
  }
  AFTER_REQUIRE(A.f32 != nullptr, AFTER_ESTATE, "operand has no fp32 copy");
  const bool simt_ok = W.Cin % 16 == 0 && W.N % 4 == 0 && epi.ldo % 4 == 0 &&
                       (W.taps.n_per_phase == 0 || W.taps.n_per_phase % SG_BN == 0);
  if (simt_ok) {
    dim3 grid(ceil_div(W.N, SG_BN), ceil_div(T, SG_BM), B);
    tap_gemm_simt_kernel<<<grid, 256, 0, st>>>(A.f32, W.w, epi, W.taps, T, P, W.Cin, W.N, T_in, a_bstride);
    AFTER_CUDA_CHECK(cudaGetLastError());
    AFTER_COUNT_LAUNCH();
    return;
  }
  AFTER_REQUIRE(epi.out_f32 && !epi.out_hi && !epi.rope && epi.ldo == W.N, AFTER_EINVAL,
                "naive tap-GEMM supports fp32 output only");
  const size_t total = (size_t)B * T * W.N;
  tap_gemm_naive_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A.f32, W.w, W.bias, epi.res, epi.out_f32, W.taps,
                                                                         B, T, P, W.Cin, W.N, epi.gelu, T_in, a_bstride);
  AFTER_CUDA_CHECK(cudaGetLastError());
  AFTER_COUNT_LAUNCH();
  if (epi.stats) {
    const int C = epi.stat_cmod > 0 ? epi.stat_cmod : W.N;
    gn_stats_launch(epi.out_f32, epi.stats, B, T * (W.N / C), C, epi.stat_groups, st);
  }
}

// Unit-test entry: C[M,N] = A[M,K] W[N,K]^T (+bias), all device fp32, through the kernel `precision` selects.
inline void debug_gemm(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int precision,
                       cudaStream_t st) {
  AFTER_REQUIRE(precision >= 0 && precision <= 2, AFTER_EINVAL, "unknown precision");
  const bool tcm = precision != AFTER_PRECISION_FP32_SIMT;
  if (tcm) AFTER_REQUIRE(K % 64 == 0 && N % 32 == 0, AFTER_EINVAL, "tcgen05 GEMM needs K % 64 == 0 and N % 32 == 0");
  else AFTER_REQUIRE(K % 16 == 0 && N % 4 == 0, AFTER_EINVAL, "SIMT GEMM needs K % 16 == 0 and N % 4 == 0");
  Arena tmp;
  try {
    GemmWeight gw;
    gw.N = N; gw.Cin = K; gw.K = K; gw.bn = pick_bn(N);
    gw.w = const_cast<float*>(W);
    gw.bias = const_cast<float*>(bias);
    ActOperand a;
    a.capacity = (size_t)M * K;
    a.f32 = const_cast<float*>(A);
    if (tcm) {
      const size_t nw = (size_t)N * K, na = (size_t)M * K;
      gw.hi = tmp.alloc<__nv_bfloat16>(nw); gw.lo = tmp.alloc<__nv_bfloat16>(nw);
      a.hi = tmp.alloc<__nv_bfloat16>(na); a.lo = tmp.alloc<__nv_bfloat16>(na);
      split_bf16_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(W, gw.hi, gw.lo, nw);
      split_bf16_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(A, a.hi, a.lo, na);
      AFTER_CUDA_CHECK(cudaGetLastError());
      gw.map_hi = make_tmap_weight(gw.hi, N, K, gw.bn);
      gw.map_lo = make_tmap_weight(gw.lo, N, K, gw.bn);
      gw.tc_ok = true;
      gw.bn2 = pick_bn2(N, 0);
      gw.tc2_ok = gw.bn2 > 0;
      if (gw.tc2_ok) {
        gw.map2_hi = make_tmap_weight(gw.hi, N, K, gw.bn2 / 2);
        gw.map2_lo = make_tmap_weight(gw.lo, N, K, gw.bn2 / 2);
      }
    }
    GemmEpi e; e.out_f32 = C; e.ldo = N;
    const char* tr = debug_env("AFTER_DEBUG_TRACE");
    unsigned long long* ts = nullptr;
    if (tr && tr[0] == '1') {
      ts = tmp.alloc<unsigned long long>(32);
      AFTER_CUDA_CHECK(cudaMemsetAsync(ts, 0, 32 * 8, st));
      e.debug_ts = ts;
      // run twice so the traced launch is warm
      GemmEpi e0 = e; e0.debug_ts = nullptr;
      tap_gemm(a, 1, M, 1, gw, e0, precision, st);
    }
    tap_gemm(a, 1, M, 1, gw, e, precision, st);
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
    if (ts) {
      unsigned long long h[32];
      AFTER_CUDA_CHECK(cudaMemcpy(h, ts, sizeof(h), cudaMemcpyDeviceToHost));
      const char* names[22] = {"entry", "prologue_done", "full0", "full1", "full2", "full3", "", "", "tile0_issued", "tile1_issued",
                               "tile2_issued", "tile3_issued", "epi0_begin", "epi0_end", "epi1_begin", "epi1_end", "epi2_begin",
                               "epi2_end", "epi3_begin", "epi3_end", "teardown_sync", "dealloc_done"};
      fprintf(stderr, "trace M=%d N=%d K=%d prec=%d (ns since kernel entry of CTA 0):\n", M, N, K, precision);
      for (int i = 0; i < 22; ++i)
        if (h[i]) fprintf(stderr, "  %-14s %8lld\n", names[i], (long long)(h[i] - h[0]));
    }
  } catch (...) {
    cudaStreamSynchronize(st);
    tmp.release();
    throw;
  }
  tmp.release();
}

}  // namespace after
