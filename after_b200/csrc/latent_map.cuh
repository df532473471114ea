// The 2-D timbre map of the exported model (after_scripts/export.py:494-508): latent2map / map2latent average their
// input over time, run it through the encoder / decoder half of the export-time SmallAutoencoder
// (after/diffusion/latent_plot.py:20-37: Linear-GELU-Linear-GELU-Linear, widths 6-16-16-2 and 2-8-16-6) and repeat the
// result over the buffer.  Without a trained projection the reference uses an identity (export.py:143), which is what a
// handle without AFTER_MODULE_LATENT_MAP tensors does.
#pragma once
#include "context.cuh"

namespace after {

constexpr int LM_MAXW = 64;  // widest layer / channel count the kernel's shared-memory vectors hold

struct LatentMapNet {
  const float *w[3] = {nullptr, nullptr, nullptr}, *b[3] = {nullptr, nullptr, nullptr};
  int dim[4] = {0, 0, 0, 0};  // in, hidden 1, hidden 2, out
};

// one block per stream: mean over T of every input channel, three tiny dense layers, repeat over T
__global__ void __launch_bounds__(128)
latent_map_kernel(const float* __restrict__ x, float* __restrict__ out, LatentMapNet net, int C_in, int C_out, int T, int identity) {
  __shared__ float va[LM_MAXW], vb[LM_MAXW];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = warp; c < C_in; c += nw) {
    const float* xp = x + ((size_t)b * C_in + c) * T;
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += xp[t];
    s = warp_sum(s);
    if (lane == 0) va[c] = s / (float)T;
  }
  __syncthreads();
  float* src = va;
  float* dst = vb;
  if (!identity) {
    for (int l = 0; l < 3; ++l) {
      const int K = net.dim[l], N = net.dim[l + 1];
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = net.b[l][n];
        for (int k = 0; k < K; ++k) acc = fmaf(net.w[l][(size_t)n * K + k], src[k], acc);
        dst[n] = l < 2 ? gelu_erf(acc) : acc;
      }
      __syncthreads();
      float* tmp = src; src = dst; dst = tmp;
    }
  }
  for (int i = threadIdx.x; i < C_out * T; i += blockDim.x) out[(size_t)b * C_out * T + i] = src[i / T];
}

struct LatentMap {
  LatentMapNet enc, dec;  // latent -> map, map -> latent
  bool loaded = false;

  void finalize(const TensorMap& sd, Arena* arena) {
    auto load = [&](LatentMapNet& net, const std::string& prefix) {
      for (int l = 0; l < 3; ++l) {
        const std::string k = prefix + "." + std::to_string(2 * l);
        auto wi = sd.find(k + ".weight"), bi = sd.find(k + ".bias");
        AFTER_REQUIRE(wi != sd.end() && bi != sd.end(), AFTER_EMISSING, "missing tensor '" + k + ".weight' / '.bias'");
        const HostTensor& w = wi->second;
        AFTER_REQUIRE(w.shape.size() == 2 && bi->second.numel() == w.shape[0], AFTER_ESHAPE, "tensor '" + k + ".weight' has an unexpected shape");
        AFTER_REQUIRE(w.shape[0] <= LM_MAXW && w.shape[1] <= LM_MAXW, AFTER_EINVAL, "latent-map layer wider than 64");
        if (l == 0) net.dim[0] = (int)w.shape[1];
        AFTER_REQUIRE(net.dim[l] == (int)w.shape[1], AFTER_ESHAPE, "latent-map layer widths do not chain at '" + k + "'");
        net.dim[l + 1] = (int)w.shape[0];
        net.w[l] = arena->upload(w.data);
        net.b[l] = arena->upload(bi->second.data);
      }
    };
    load(enc, "encoder");
    load(dec, "decoder");
    AFTER_REQUIRE(enc.dim[3] == dec.dim[0] && dec.dim[3] == enc.dim[0], AFTER_ESHAPE, "latent-map encoder / decoder do not invert each other's shapes");
    loaded = true;
  }

  // direction 0: latent2map, 1: map2latent.  x dev (B, C_in, T) -> out dev (B, C_out, T); returns C_out through *c_out.
  void run(int direction, const float* x, float* out, int B, int C_in, int T, int* c_out, cudaStream_t st) const {
    AFTER_REQUIRE(direction == 0 || direction == 1, AFTER_EINVAL, "direction must be 0 (latent2map) or 1 (map2latent)");
    AFTER_REQUIRE(B >= 1 && T >= 1 && C_in >= 1 && C_in <= LM_MAXW, AFTER_EINVAL, "bad latent-map shape");
    const LatentMapNet& net = direction == 0 ? enc : dec;
    int C_out = C_in;
    if (loaded) {
      AFTER_REQUIRE(C_in == net.dim[0], AFTER_ESHAPE, "latent-map input has the wrong number of channels");
      C_out = net.dim[3];
    }
    if (c_out) *c_out = C_out;
    latent_map_kernel<<<B, 128, 0, st>>>(x, out, net, C_in, C_out, T, loaded ? 0 : 1);
    AFTER_CUDA_CHECK(cudaGetLastError());
    AFTER_COUNT_LAUNCH();
  }
};

}  // namespace after
