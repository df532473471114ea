// C ABI of libafter_b200.so (see include/after_b200.h).  Everything here is argument checking, error
// translation and stream plumbing; the work lives in denoiser.cuh / codec.cuh / encoder1d.cuh.
#include <cstring>
#include <memory>
#include <mutex>

#include "codec.cuh"
#include "context.cuh"
#include "denoiser.cuh"
#include "ecapa.cuh"
#include "latent_map.cuh"
#include "unet.cuh"

namespace after {
std::atomic<int64_t> g_launches{0};
Profiler g_prof;
}

using namespace after;

struct after_ctx {
  after_config cfg{};
  int device = 0;
  int precision = -1;  // -1: weights not finalized
  TensorMap tensors[6];
  Arena arena;
  StreamBridge bridge;
  Denoiser denoiser;
  Codec codec;
  StructureEncoder structure;
  TimbreEncoder timbre;
  UNet unet;
  LatentMap latent_map;
  bool have_codec = false, have_structure = false, have_timbre = false, have_unet = false;
  bool net_is_unet() const { return have_unet && denoiser.D == 0; }
  // device staging of the *_host entry points (the copies go straight from / to the caller's host pointers)
  float* dev_io = nullptr;
  size_t dev_io_floats = 0;
  float* chain = nullptr;  // z_s | z_t | time_cond | cond | x scratch of after_generate
  size_t chain_floats = 0;
  std::string err;
};

static thread_local std::string g_err;

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename F>
int guarded(after_handle h, F&& f) {
  try {
    if (!h) throw Error(AFTER_EINVAL, "null handle");
    DeviceGuard g(h->device);
    f();
    return AFTER_OK;
  } catch (const Error& e) {
    (h ? h->err : g_err) = e.what();
    cudaGetLastError();  // clear a sticky-free error so the next call starts clean
    return e.code;
  } catch (const std::bad_alloc&) {
    (h ? h->err : g_err) = "out of host memory";
    return AFTER_ENOMEM;
  } catch (const std::exception& e) {
    (h ? h->err : g_err) = e.what();
    return AFTER_EINVAL;
  }
}

// Orders the library's work stream after the caller's stream on entry and the caller's stream after the work stream on
// EVERY exit path, exceptions included (work enqueued before a failure must still be ordered before whatever the
// caller enqueues next: outputs are allocated on the caller's stream).  All argument validation happens before it.
struct BridgeScope {
  StreamBridge& b;
  cudaStream_t user;
  BridgeScope(StreamBridge& bridge, void* stream) : b(bridge), user(reinterpret_cast<cudaStream_t>(stream)) { b.enter(user); }
  ~BridgeScope() {
    if (cudaEventRecord(b.e_out, b.work) == cudaSuccess) cudaStreamWaitEvent(user, b.e_out, 0);
  }
  BridgeScope(const BridgeScope&) = delete;
  BridgeScope& operator=(const BridgeScope&) = delete;
};

void require_ready(after_handle h) {
  AFTER_REQUIRE(h->precision >= 0, AFTER_ESTATE, "after_finalize_weights has not been called on this handle");
}

void ensure_staging(after_handle h, size_t floats) {
  if (floats > h->dev_io_floats) {
    if (h->dev_io) cudaFree(h->dev_io);
    h->dev_io = nullptr;
    AFTER_CUDA_CHECK(cudaMalloc(&h->dev_io, floats * sizeof(float)));
    h->dev_io_floats = floats;
  }
}

}  // namespace

extern "C" {

int after_abi_version(void) { return AFTER_B200_ABI_VERSION; }

const char* after_build_info(void) {
  return "libafter_b200 abi=3 arch=sm_100a (tcgen05/TMEM/TMA) nvcc=" AFTER_STR(__CUDACC_VER_MAJOR__) "." AFTER_STR(
      __CUDACC_VER_MINOR__) " built " __DATE__;
}

int after_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* after_last_error(after_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int after_create(const after_config* cfg, int device, after_handle* out) {
  try {
    AFTER_REQUIRE(cfg && out, AFTER_EINVAL, "null argument");
    AFTER_REQUIRE(cfg->abi_version == AFTER_B200_ABI_VERSION, AFTER_EINVAL, "after_config.abi_version mismatch");
    int n = after_device_count();
    AFTER_REQUIRE(n > 0, AFTER_ECUDA, "no CUDA device visible: libafter_b200 has no CPU fallback");
    AFTER_REQUIRE(device >= 0 && device < n, AFTER_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    AFTER_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    AFTER_REQUIRE(prop.major == 10, AFTER_ECUDA,
                  std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                      "; this library contains sm_100a code only");
    std::unique_ptr<after_ctx> h(new after_ctx());
    h->cfg = *cfg;
    h->device = device;
    DeviceGuard g(device);
    h->bridge.init();
    tc_init_kernels();
    *out = h.release();
    return AFTER_OK;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return AFTER_EINVAL;
  }
}

int after_destroy(after_handle h) {
  if (!h) return AFTER_OK;
  {
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    h->denoiser.destroy();
    h->codec.destroy();
    h->unet.destroy();
    h->bridge.destroy();
    h->arena.release();
    if (h->dev_io) cudaFree(h->dev_io);
    if (h->chain) cudaFree(h->chain);
  }
  delete h;
  return AFTER_OK;
}

int after_load_tensor(after_handle h, int module, const char* key, const void* data, const int64_t* shape, int ndim,
                      int dtype) {
  int ignored = 0;
  int rc = guarded(h, [&] {
    AFTER_REQUIRE(h->precision < 0, AFTER_ESTATE, "weights already finalized");
    AFTER_REQUIRE(module >= 0 && module < 6, AFTER_EINVAL, "unknown module id");
    AFTER_REQUIRE(key && data && (shape || ndim == 0) && ndim >= 0 && ndim <= 8, AFTER_EINVAL, "bad tensor arguments");
    const std::string k(key);
    // buffers the offline path never reads: streaming caches, unused position table, GroupNorm stream pads
    auto ends_with = [&](const char* s) {
      size_t n = strlen(s);
      return k.size() >= n && k.compare(k.size() - n, n, s) == 0;
    };
    if (k.find("cache") != std::string::npos || ends_with("precomputed_pos_enc") || ends_with(".pad") ||
        ends_with("num_batches_tracked") || ends_with("pqmf.hk") || ends_with("pqmf.h")) {
      ignored = 1;
      return;
    }
    HostTensor t;
    t.shape.assign(shape, shape + ndim);
    const int64_t n = t.numel();
    t.data.resize((size_t)n);
    if (dtype == AFTER_DTYPE_F32) memcpy(t.data.data(), data, (size_t)n * 4);
    else if (dtype == AFTER_DTYPE_F64)
      for (int64_t i = 0; i < n; ++i) t.data[i] = (float)reinterpret_cast<const double*>(data)[i];
    else if (dtype == AFTER_DTYPE_I64)
      for (int64_t i = 0; i < n; ++i) t.data[i] = (float)reinterpret_cast<const int64_t*>(data)[i];
    else throw Error(AFTER_EINVAL, "unsupported dtype");
    h->tensors[module][k] = std::move(t);
  });
  return rc == AFTER_OK && ignored ? AFTER_IGNORED : rc;
}

int after_finalize_weights(after_handle h, int precision) {
  return guarded(h, [&] {
    AFTER_REQUIRE(h->precision < 0, AFTER_ESTATE, "weights already finalized");
    AFTER_REQUIRE(precision >= 0 && precision <= 2, AFTER_EINVAL, "unknown precision mode");
    bool any = false;
    if (!h->tensors[AFTER_MODULE_DENOISER].empty()) {
      h->denoiser.finalize(h->cfg, h->tensors[AFTER_MODULE_DENOISER], precision, &h->arena);
      any = true;
    }
    if (!h->tensors[AFTER_MODULE_AUTOENCODER].empty()) {
      AFTER_REQUIRE(h->cfg.ae_channels > 0, AFTER_EINVAL, "autoencoder tensors loaded but after_config.ae_channels == 0");
      h->codec.finalize(h->cfg, h->tensors[AFTER_MODULE_AUTOENCODER], precision, &h->arena);
      h->have_codec = true;
      any = true;
    }
    if (!h->tensors[AFTER_MODULE_STRUCTURE_ENCODER].empty()) {
      AFTER_REQUIRE(h->cfg.se_n_blocks > 0, AFTER_EINVAL, "structure-encoder tensors loaded but after_config.se_n_blocks == 0");
      h->structure.finalize(h->cfg, h->tensors[AFTER_MODULE_STRUCTURE_ENCODER], precision, &h->arena);
      h->have_structure = true;
      any = true;
    }
    if (!h->tensors[AFTER_MODULE_TIMBRE_ENCODER].empty()) {
      AFTER_REQUIRE(h->cfg.te_n_blocks > 0, AFTER_EINVAL, "timbre-encoder tensors loaded but after_config.te_n_blocks == 0");
      h->timbre.finalize(h->cfg, h->tensors[AFTER_MODULE_TIMBRE_ENCODER], &h->arena);
      h->have_timbre = true;
      any = true;
    }
    if (!h->tensors[AFTER_MODULE_UNET].empty()) {
      AFTER_REQUIRE(h->cfg.un_n_levels > 0, AFTER_EINVAL, "UNET1D tensors loaded but after_config.un_n_levels == 0");
      h->unet.finalize(h->cfg, h->tensors[AFTER_MODULE_UNET], precision, &h->arena);
      h->have_unet = true;
      any = true;
    }
    if (!h->tensors[AFTER_MODULE_LATENT_MAP].empty()) h->latent_map.finalize(h->tensors[AFTER_MODULE_LATENT_MAP], &h->arena);
    AFTER_REQUIRE(any, AFTER_EMISSING, "no tensors were loaded");
    for (auto& m : h->tensors) m.clear();  // host copies are no longer needed
    h->precision = precision;
  });
}

int after_denoiser_forward(after_handle h, const float* x, const float* time, const float* cond, const float* time_cond,
                           float* out, int N, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x && time && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.forward(x, time, cond, time_cond, out, N, T, h->bridge.work);
  });
}

int after_unet_forward(after_handle h, const float* x, const float* time, const float* cond, const float* time_cond, float* out,
                       int N, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_unet, AFTER_ESTATE, "no UNET1D weights on this handle");
    AFTER_REQUIRE(x && time && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->unet.forward(x, time, cond, time_cond, out, N, T, h->bridge.work);
  });
}

int after_model_forward(after_handle h, const float* x, const float* time, const float* cond, const float* time_cond,
                        float* out, int B, int T, float guidance_timbre, float guidance_structure, int cfg_variant,
                        float clamp, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    if (h->net_is_unet()) {
      AFTER_REQUIRE(x && time && out, AFTER_EINVAL, "null tensor pointer");
      BridgeScope bridge(h->bridge, stream);
      h->unet.model_forward(x, time, cond, time_cond, out, B, T, guidance_timbre, guidance_structure, cfg_variant, clamp,
                            h->bridge.work);
      return;
    }
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x && time && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.model_forward(x, time, cond, time_cond, out, B, T, guidance_timbre, guidance_structure, cfg_variant, clamp,
                              h->bridge.work);
  });
}

int after_sample(after_handle h, const float* x0, const float* cond, const float* time_cond, float* out, int B, int T,
                 int nb_steps, float guidance_timbre, float guidance_structure, int cfg_variant, float clamp,
                 void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    if (h->net_is_unet()) {
      AFTER_REQUIRE(x0 && out, AFTER_EINVAL, "null tensor pointer");
      BridgeScope bridge(h->bridge, stream);
      h->unet.sample(x0, cond, time_cond, out, B, T, nb_steps, guidance_timbre, guidance_structure, cfg_variant, clamp,
                     h->bridge.work);
      return;
    }
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x0 && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.sample(x0, cond, time_cond, out, B, T, nb_steps, guidance_timbre, guidance_structure, cfg_variant, clamp,
                       h->bridge.work);
  });
}

int after_denoiser_forward_cached(after_handle h, const float* x, const float* time, const float* cond,
                                  const float* time_cond, float* out, int N, int T, int cache_index, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x && time && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    h->denoiser.check_cache_index(cache_index);
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.forward(x, time, cond, time_cond, out, N, T, h->bridge.work, cache_index);
  });
}

int after_model_forward_cached(after_handle h, const float* x, const float* time, const float* cond,
                               const float* time_cond, float* out, int B, int T, float guidance_timbre,
                               float guidance_structure, int cfg_variant, float clamp, int cache_index, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x && time && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    h->denoiser.check_cache_index(cache_index);
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.model_forward(x, time, cond, time_cond, out, B, T, guidance_timbre, guidance_structure, cfg_variant, clamp,
                              h->bridge.work, cache_index);
  });
}

int after_roll_cache(after_handle h, int roll_size, int cache_index, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.roll_cache(roll_size, cache_index, h->bridge.work);
  });
}

int after_reset_cache(after_handle h, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.reset_cache(h->bridge.work);
  });
}

int after_sample_stream(after_handle h, const float* x_last, const float* cond, const float* time_cond, float* out, int B,
                        int T, int nb_steps, float guidance_timbre, float guidance_structure, int cfg_variant, float clamp,
                        void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x_last && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->denoiser.sample(x_last, cond, time_cond, out, B, T, nb_steps, guidance_timbre, guidance_structure, cfg_variant, clamp,
                       h->bridge.work, /*stream=*/true);
  });
}

int after_sample_host(after_handle h, const float* x0, const float* cond, const float* time_cond, float* out, int B,
                      int T, int nb_steps, float guidance_timbre, float guidance_structure, int cfg_variant,
                      float clamp, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->denoiser.D > 0, AFTER_ESTATE, "no denoiser weights on this handle");
    AFTER_REQUIRE(x0 && cond && time_cond && out, AFTER_EINVAL, "null tensor pointer");
    AFTER_REQUIRE(B >= 1 && T >= 1, AFTER_EINVAL, "B and T must be >= 1");
    const Denoiser& d = h->denoiser;
    const size_t nx = (size_t)B * d.C * T, nc = (size_t)B * d.zt, nt = (size_t)B * d.zs * T;
    ensure_staging(h, 2 * nx + nc + nt);
    float* dx = h->dev_io;
    float* dc = dx + nx;
    float* dt = dc + nc;
    float* dout = dt + nt;
    cudaStream_t st = h->bridge.work;
    BridgeScope bridge(h->bridge, stream);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(dx, x0, nx * 4, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(dc, cond, nc * 4, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(dt, time_cond, nt * 4, cudaMemcpyHostToDevice, st));
    h->denoiser.sample(dx, dc, dt, dout, B, T, nb_steps, guidance_timbre, guidance_structure, cfg_variant, clamp, st);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(out, dout, nx * 4, cudaMemcpyDeviceToHost, st));
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int after_ae_encode(after_handle h, const float* audio, float* z, int B, int64_t samples, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_codec, AFTER_ESTATE, "no autoencoder weights on this handle");
    AFTER_REQUIRE(audio && z, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->codec.encode(audio, z, B, samples, h->bridge.work);
  });
}

int after_ae_decode(after_handle h, const float* z, float* audio, int B, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_codec, AFTER_ESTATE, "no autoencoder weights on this handle");
    AFTER_REQUIRE(audio && z, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->codec.decode(z, audio, B, T, h->bridge.work);
  });
}

int after_ae_encode_stream(after_handle h, int slot, const float* audio, float* z, int B, int64_t samples, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_codec, AFTER_ESTATE, "no autoencoder weights on this handle");
    AFTER_REQUIRE(audio && z, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->codec.encode_stream(slot, audio, z, B, samples, h->bridge.work);
  });
}

int after_ae_decode_stream(after_handle h, int slot, const float* z, float* audio, int B, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_codec, AFTER_ESTATE, "no autoencoder weights on this handle");
    AFTER_REQUIRE(audio && z, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->codec.decode_stream(slot, z, audio, B, T, h->bridge.work);
  });
}

int after_structure_encode_stream(after_handle h, int slot, const float* z, float* time_cond, int B, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_structure, AFTER_ESTATE, "no structure-encoder weights on this handle");
    AFTER_REQUIRE(z && time_cond, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->structure.forward_stream(slot, z, time_cond, B, T, h->bridge.work);
  });
}

int after_stream_reset(after_handle h, int slot, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_codec || h->have_structure, AFTER_ESTATE, "no streaming sub-model on this handle");
    BridgeScope bridge(h->bridge, stream);
    if (h->have_codec) h->codec.reset_stream_slot(slot, h->bridge.work);
    if (h->have_structure) h->structure.reset_stream_slot(slot, h->bridge.work);
  });
}

int after_structure_encode(after_handle h, const float* z, float* time_cond, int B, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_structure, AFTER_ESTATE, "no structure-encoder weights on this handle");
    AFTER_REQUIRE(z && time_cond, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->structure.forward(z, time_cond, B, T, h->bridge.work);
  });
}

int after_timbre_encode(after_handle h, const float* z, float* cond, int B, int T, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(h->have_timbre, AFTER_ESTATE, "no timbre-encoder weights on this handle");
    AFTER_REQUIRE(z && cond, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->timbre.forward(z, cond, B, T, h->bridge.work);
  });
}

int after_latent_map(after_handle h, int direction, const float* x, float* out, int B, int C_in, int T, int* C_out,
                     void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(x && out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    h->latent_map.run(direction, x, out, B, C_in, T, C_out, h->bridge.work);
  });
}

static void generate_device(after_handle h, const float* a_s, const float* a_t, const float* x0, float* out, int B,
                            int64_t samples, int nb_steps, float g_t, float g_s, cudaStream_t st) {
  AFTER_REQUIRE(h->denoiser.D > 0 && h->have_codec && h->have_structure && h->have_timbre, AFTER_ESTATE,
                "after_generate needs denoiser, autoencoder, structure-encoder and timbre-encoder weights");
  const int ratio = h->codec.ratio;
  AFTER_REQUIRE(samples > 0 && samples % ratio == 0, AFTER_EINVAL, "samples must be a positive multiple of the codec ratio");
  const int T = (int)(samples / ratio);
  const Denoiser& d = h->denoiser;
  AFTER_REQUIRE(d.C == h->cfg.ae_z_channels && d.C == h->cfg.se_in_size && d.C == h->cfg.te_in_size, AFTER_EINVAL,
                "latent channel counts of the sub-models disagree");
  const size_t nz = (size_t)B * d.C * T, ntc = (size_t)B * d.zs * T, nc = (size_t)B * d.zt;
  const size_t need = 3 * nz + ntc + nc;
  if (need > h->chain_floats) {
    if (h->chain) cudaFree(h->chain);
    h->chain = nullptr;
    AFTER_CUDA_CHECK(cudaMalloc(&h->chain, need * sizeof(float)));
    h->chain_floats = need;
  }
  float* z_s = h->chain;
  float* z_t = z_s + nz;
  float* x = z_t + nz;
  float* tcond = x + nz;
  float* cond = tcond + ntc;
  h->codec.encode_pair(a_s, a_t, z_s, z_t, B, samples, st);  // both inputs as one batch of 2 B chunks
  h->structure.forward(z_s, tcond, B, T, st);
  h->timbre.forward(z_t, cond, B, T, st);
  h->denoiser.sample(x0, cond, tcond, x, B, T, nb_steps, g_t, g_s, AFTER_CFG_AUDIO, 0.01f, st);
  h->codec.decode(x, out, B, T, st);
}

int after_generate(after_handle h, const float* audio_structure, const float* audio_timbre, const float* x0, float* audio_out,
                   int B, int64_t samples, int nb_steps, float guidance_timbre, float guidance_structure, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(audio_structure && audio_timbre && x0 && audio_out, AFTER_EINVAL, "null tensor pointer");
    BridgeScope bridge(h->bridge, stream);
    generate_device(h, audio_structure, audio_timbre, x0, audio_out, B, samples, nb_steps, guidance_timbre, guidance_structure,
                    h->bridge.work);
  });
}

int after_generate_host(after_handle h, const float* audio_structure, const float* audio_timbre, const float* x0,
                        float* audio_out, int B, int64_t samples, int nb_steps, float guidance_timbre,
                        float guidance_structure, void* stream) {
  return guarded(h, [&] {
    require_ready(h);
    AFTER_REQUIRE(audio_structure && audio_timbre && x0 && audio_out, AFTER_EINVAL, "null tensor pointer");
    AFTER_REQUIRE(B >= 1 && samples >= 1 && h->have_codec, AFTER_EINVAL, "bad batch / samples, or no codec on this handle");
    const int ratio = h->codec.ratio;
    AFTER_REQUIRE(samples % ratio == 0, AFTER_EINVAL, "samples must be a multiple of the codec ratio");
    const size_t na = (size_t)B * samples, nx = (size_t)B * h->denoiser.C * (samples / ratio);
    ensure_staging(h, 3 * na + nx);
    float* d_s = h->dev_io;
    float* d_t = d_s + na;
    float* d_o = d_t + na;
    float* d_x = d_o + na;
    cudaStream_t st = h->bridge.work;
    BridgeScope bridge(h->bridge, stream);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(d_s, audio_structure, na * 4, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(d_t, audio_timbre, na * 4, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(d_x, x0, nx * 4, cudaMemcpyHostToDevice, st));
    generate_device(h, d_s, d_t, d_x, d_o, B, samples, nb_steps, guidance_timbre, guidance_structure, st);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(audio_out, d_o, na * 4, cudaMemcpyDeviceToHost, st));
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int64_t after_launch_count(after_handle) { return g_launches.load(); }

int64_t after_device_bytes(after_handle h) { return h ? (int64_t)h->arena.bytes : 0; }

int after_ae_ratio(after_handle h) {
  if (!h || h->cfg.ae_channels <= 0) return 0;
  int r = h->cfg.ae_pqmf_bands > 0 ? h->cfg.ae_pqmf_bands : 1;
  for (int i = 0; i < h->cfg.ae_n_stages; ++i) r *= h->cfg.ae_factors[i];
  return r;
}

int after_profile_enable(after_handle h, int on) {
  return guarded(h, [&] {
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    g_prof.reset();
    g_prof.on = on != 0;
  });
}

int after_profile_read(after_handle h, int kernel_class, int64_t* launches, double* ms, double* flops, double* bytes) {
  return guarded(h, [&] {
    AFTER_REQUIRE(kernel_class >= 0 && kernel_class < KC_COUNT, AFTER_EINVAL, "unknown kernel class");
    AFTER_REQUIRE(launches && ms && flops && bytes, AFTER_EINVAL, "null output pointer");
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    g_prof.read(kernel_class, launches, ms, flops, bytes);
  });
}

int after_debug_gemm(after_handle h, const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                     int precision, void* stream) {
  return guarded(h, [&] {
    AFTER_REQUIRE(A && W && C, AFTER_EINVAL, "null tensor pointer");
    AFTER_REQUIRE(M >= 1 && N >= 1 && K >= 1, AFTER_EINVAL, "bad GEMM shape");
    BridgeScope bridge(h->bridge, stream);
    debug_gemm(A, W, bias, C, M, N, K, precision, h->bridge.work);
  });
}

}  // extern "C"
