// Shared device/host helpers for libafter_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <stdexcept>

namespace after {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define AFTER_CUDA_CHECK(expr)                                                                  \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::after::Error(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) +     \
                                   " (" __FILE__ ":" + std::to_string(__LINE__) + ")");        \
  } while (0)

#define AFTER_REQUIRE(cond, code, msg)                                                          \
  do {                                                                                          \
    if (!(cond)) throw ::after::Error((code), std::string(msg));                                \
  } while (0)

// A/B and trace knobs (AFTER_* environment variables, GemmEpi::debug_skip / debug_ts, the fused-MLP %globaltimer trace)
// exist only in builds compiled with -DAFTER_DEBUG (python -m after_b200.build --debug): the release library never reads
// the environment, and the kernels' debug branches are compiled out (kDebugBuild is a compile-time false).
#ifdef AFTER_DEBUG
constexpr bool kDebugBuild = true;
inline const char* debug_env(const char* name) { return getenv(name); }
#else
constexpr bool kDebugBuild = false;
inline const char* debug_env(const char*) { return nullptr; }
#endif

#define AFTER_STR_(x) #x
#define AFTER_STR(x) AFTER_STR_(x)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Kernels of the sampling loop start with pdl_wait() (blocks until the preceding kernel of the stream / graph has
// completed and its writes are visible; a no-op for a normal launch) followed by pdl_trigger() (lets the NEXT kernel be
// scheduled once every CTA of this one has got this far), and are launched through launch_k() with the
// programmatic-stream-serialization attribute: the next kernel's launch latency, block scheduling and on-chip set-up
// overlap the tail of this one.  Only kernels that call pdl_wait() before touching global memory may be launched this
// way.  AFTER_PDL=0 (debug builds only) turns the attribute off (plain stream order) for A/B measurements.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = debug_env("AFTER_PDL");
    v = (e && e[0] == '0') ? 0 : 1;  // on by default: +2.4 % steps/s at base B=8 (profiles/r01b_ab_knobs.jsonl)
  }
  return v == 1;
}
inline bool& pdl_scope() {  // set by the denoiser around its per-step kernels; the codec launches stay plain
  static thread_local bool on = false;
  return on;
}
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_scope()) { pdl_scope() = on; }
  ~PdlScope() { pdl_scope() = prev; }
};
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && pdl_scope()) ? 1 : 0;
  AFTER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// exact (erf) GELU, nn.GELU() default -- transformerv2.py:277,390
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// erf-form GELU with erf from Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32 rounding level), branch
// free: ~14 instructions instead of erff's ~35 with divergent ranges.  Used in the tensor-core GEMM epilogue, where
// the 9.4 M activations per MLP up-projection are evaluated by only 8 warps per SM.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfc_z = p * t * __expf(-z * z);
  const float erf_x = copysignf(1.0f - erfc_z, x);
  return 0.5f * x * (1.0f + erf_x);
}
// Two GELUs at once with Blackwell's packed fp32x2 FMA/MUL/ADD (half the issue slots of the scalar version).
// (u >= 1 and -z^2 log2 e <= 0 always, so the bare MUFU approximations need none of the range fix-ups that
// __fdividef / exp2f wrap around them: the epilogue is issue-bound, every instruction per element counts.)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 u = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_approx(u.x), rcp_approx(u.y));
  float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  const float2 a = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));  // -z^2 log2(e)
  const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
  const float2 erfc_z = __fmul2_rn(__fmul2_rn(p, t), e);
  const float2 erf_abs = __ffma2_rn(erfc_z, make_float2(-1.0f, -1.0f), make_float2(1.0f, 1.0f));
  const float2 erf_x = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
  return __fmul2_rn(__fmul2_rn(x, make_float2(0.5f, 0.5f)), __fadd2_rn(make_float2(1.0f, 1.0f), erf_x));
}
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
#endif

}  // namespace after
