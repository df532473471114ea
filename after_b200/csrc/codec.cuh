// AutoEncoder.encode / decode (SimpleNetsStream.py:918-954) and Encoder1D.forward (encoder.py:273-298) on one B200.
//
// Data layout: every activation is frame-major fp32 (B, T, C) in HBM; a convolution is
//     act_operand_kernel  : x -> act(norm(x)) written as the GEMM A-operand (bf16 hi/lo for tcgen05, fp32 otherwise)
//     tap-GEMM            : operand x folded weights -> +bias (+residual) -> fp32 out, and the GroupNorm
//                           sum / sum-of-squares of the result for the next layer's norm (epilogue atomics)
// so each activation is written once and read twice (operand pass + residual), never re-read for statistics.
// Weight-norm (g * v / ||v||, SimpleNetsStream.py:84-92) and eval-mode BatchNorm are folded at load.
// Strided convs read the input as (B, T/f, f, C) phases; transposed convs write (B, T, f*C) = (B, T*f, C).
#pragma once
#include <cmath>
#include "codec_kernels.cuh"
#include "context.cuh"
#include "denoiser_kernels.cuh"
#include "gemm_host.cuh"

namespace after {

inline void gn_stats_launch(const float* x, double* stats, int B, int T, int C, int groups, cudaStream_t st) {
  dim3 grid(ceil_div(T, 256), B);
  launch_k(gn_stats_kernel, grid, dim3(256), 0, st, x, stats, T, C, groups);
  AFTER_COUNT_LAUNCH();
}

// Replayable CUDA graphs keyed by the call signature.
struct GraphCache {
  struct Entry {
    cudaGraphExec_t exec = nullptr;
    int64_t kernels = 0;
  };
  std::map<std::vector<int>, Entry> graphs;
  bool enabled = true;
  void init() {
    const char* ng = debug_env("AFTER_NO_GRAPH");
    enabled = !(ng && ng[0] == '1');
  }
  template <typename F>
  void run(const std::vector<int>& key, cudaStream_t st, F&& body) {
    if (!enabled || g_prof.on) {
      body();
      return;
    }
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      Entry ge;
      cudaGraph_t graph = nullptr;
      const int64_t before = g_launches.load();
      AFTER_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      try {
        body();
      } catch (...) {
        cudaStreamEndCapture(st, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      AFTER_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
      ge.kernels = g_launches.load() - before;
      g_launches.fetch_sub(ge.kernels);  // capture did not execute anything
      AFTER_CUDA_CHECK(cudaGraphInstantiate(&ge.exec, graph, 0));
      cudaGraphDestroy(graph);
      it = graphs.emplace(key, ge).first;
    }
    AFTER_CUDA_CHECK(cudaGraphLaunch(it->second.exec, st));
    g_launches.fetch_add(it->second.kernels);
  }
  void destroy() {
    for (auto& kv : graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }
};

// ------------------------------------------------------------------------------------------- tap tables
// cached_conv.get_padding semantics: centered -> left = (p-1)/2, causal -> left = p-1,
// with p = (k-1)*dilation + 1.
inline TapTable conv_taps(int k, int dilation, bool causal) {
  AFTER_REQUIRE(k >= 1 && k <= MAX_TAPS, AFTER_EINVAL, "conv kernel size not supported");
  TapTable t;
  t.ntaps = k;
  const int p = (k - 1) * dilation + 1;
  const int left = k == 1 ? 0 : (causal ? p / 2 + (p - 1) / 2 : (p - 1) / 2);
  for (int i = 0; i < k; ++i) {
    t.phase[0][i] = 0;
    t.shift[0][i] = (int16_t)(i * dilation - left);
  }
  return t;
}
// Downsample1d: conv k = 2f, stride f, padding get_padding(2f) = (f-1, f)   (SimpleNetsStream.py:32-48)
inline TapTable strided_taps(int f) {
  AFTER_REQUIRE(2 * f <= MAX_TAPS, AFTER_EINVAL, "downsampling factor too large");
  TapTable t;
  t.ntaps = 2 * f;
  for (int k = 0; k < 2 * f; ++k) {
    const int j = k - (f - 1);
    const int ph = ((j % f) + f) % f;
    t.phase[0][k] = (int8_t)ph;
    t.shift[0][k] = (int16_t)((j - ph) / f);
  }
  return t;
}
// Upsample1d: ConvTranspose1d k = 2f, stride f, padding f/2 (SimpleNetsStream.py:51-70).  Output phase r of frame
// t gets taps k0 = (r + pad) mod f from input frame t + (r + pad) / f and k0 + f from the frame before it.
inline TapTable transposed_taps(int f, int cout, int* kidx /*[f][2]*/) {
  AFTER_REQUIRE(f <= MAX_PHASES, AFTER_EINVAL, "upsampling factor too large");
  TapTable t;
  t.ntaps = 2;
  t.n_per_phase = cout;
  const int pad = f / 2;
  for (int r = 0; r < f; ++r) {
    const int k0 = (r + pad) % f, s = (r + pad) / f;
    t.phase[r][0] = 0; t.shift[r][0] = (int16_t)s;       kidx[r * 2 + 0] = k0;
    t.phase[r][1] = 0; t.shift[r][1] = (int16_t)(s - 1); kidx[r * 2 + 1] = k0 + f;
  }
  return t;
}

struct ConvLayer {
  GemmWeight w;
  int cin = 0, cout = 0;
  int in_phases = 1;   // input frames per GEMM row (stride of a strided conv)
  int out_phases = 1;  // output frames per GEMM row (stride of a transposed conv)
  int sid = -1;        // streaming: index of this conv's cached-context state (-1: stateless, e.g. kernel size 1)
};

struct NormAct {
  int C = 0, groups = 1;
  int norm = NORM_NONE, act = ACT_NONE;
  float *gamma = nullptr, *beta = nullptr, *alpha = nullptr, *inv_beta = nullptr;  // GroupNorm affine, Snake
  float *mu = nullptr, *rs = nullptr, *be = nullptr;                              // folded BatchNorm
  int gid = -1;  // streaming: index of this GroupNorm's statistics history (CachedGroupNorm stream branch)
};

struct ResBlock {
  NormAct a1, a2;
  ConvLayer c1, c2, skip;
  bool has_skip = false;
  int rid = -1, delay = 0;  // streaming: residual-branch delay state (cached_conv AlignBranches)
};

// Shared machinery of the two conv nets.
struct ConvNet {
  Arena* arena = nullptr;
  int precision = 0;
  const TensorMap* sd = nullptr;
  bool tc_mode() const { return precision != AFTER_PRECISION_FP32_SIMT; }
  int nprod() const { return precision == AFTER_PRECISION_BF16 ? 1 : 3; }

  // workspace
  float* buf[3] = {nullptr, nullptr, nullptr};
  ActOperand op;
  double* stats = nullptr;
  int stat_slots = 0, stat_next = 0;
  size_t slot_doubles = 0;
  size_t buf_elems = 0, op_elems = 0;
  int maxB = 0;
  int wsB = 0;  // chunks the offline workspace holds (Codec: 2 x max_batch for encode_pair; 0 = maxB)
  GraphCache graphs;

  const HostTensor& get(const std::string& key) const {
    auto it = sd->find(key);
    if (it == sd->end()) throw Error(AFTER_EMISSING, "missing tensor '" + key + "'");
    return it->second;
  }
  bool has(const std::string& key) const { return sd->find(key) != sd->end(); }

  std::vector<float> folded(const std::string& prefix) const {
    return fold_weight_norm(get(prefix + ".weight_v"), get(prefix + ".weight_g"));
  }

  // Conv1d weight (cout, cin, k) -> [cout][k*cin + ci]
  void make_conv(ConvLayer& L, const std::string& prefix, int cin, int cout, int k, const TapTable& taps, int in_phases) {
    const HostTensor& v = get(prefix + ".weight_v");
    AFTER_REQUIRE(v.shape.size() == 3 && v.shape[0] == cout && v.shape[1] == cin && v.shape[2] == k, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight_v' has an unexpected shape");
    const HostTensor& b = get(prefix + ".bias");
    AFTER_REQUIRE(b.numel() == cout, AFTER_ESHAPE, "tensor '" + prefix + ".bias' has an unexpected shape");
    make_conv_from(L, folded(prefix), b.data.data(), cin, cout, k, taps, in_phases);
  }
  // plain (not weight-normed) nn.Conv1d: prefix.weight (cout, cin, k), prefix.bias   (unet1d.py, blocks.py)
  void make_conv_plain(ConvLayer& L, const std::string& prefix, int cin, int cout, int k, const TapTable& taps, int in_phases) {
    const HostTensor& w = get(prefix + ".weight");
    AFTER_REQUIRE(w.shape.size() == 3 && w.shape[0] == cout && w.shape[1] == cin && w.shape[2] == k, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight' has an unexpected shape");
    const HostTensor& b = get(prefix + ".bias");
    AFTER_REQUIRE(b.numel() == cout, AFTER_ESHAPE, "tensor '" + prefix + ".bias' has an unexpected shape");
    make_conv_from(L, w.data, b.data.data(), cin, cout, k, taps, in_phases);
  }
  void make_conv_from(ConvLayer& L, const std::vector<float>& w, const float* bias, int cin, int cout, int k, const TapTable& taps,
                      int in_phases) {
    // 16/32-channel inputs (both ends of the codec) are zero-padded to the 64-channel K granule of the tcgen05 path
    const int cp = (tc_mode() && cin < 64 && cin % 4 == 0 && cout % 32 == 0 && in_phases == 1) ? 64 : cin;
    std::vector<float> m((size_t)cout * k * cp, 0.f);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int kk = 0; kk < k; ++kk) m[((size_t)o * k + kk) * cp + c] = w[((size_t)o * cin + c) * k + kk];
    build_gemm_weight(L.w, *arena, m, bias, cout, cp, taps, tc_mode());
    L.cin = cin; L.cout = cout; L.in_phases = in_phases; L.out_phases = 1;
  }

  // ConvTranspose1d weight (cin, cout, 2f) -> [r*cout + co][tap*cin + ci]
  void make_conv_transposed(ConvLayer& L, const std::string& prefix, int cin, int cout, int f) {
    const HostTensor& v = get(prefix + ".weight_v");
    AFTER_REQUIRE(v.shape.size() == 3 && v.shape[0] == cin && v.shape[1] == cout && v.shape[2] == 2 * f, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight_v' has an unexpected shape");
    const std::vector<float> w = folded(prefix);  // norm per input channel (dim 0)
    int kidx[MAX_PHASES * 2];
    TapTable taps = transposed_taps(f, cout, kidx);
    const int k = 2 * f;
    std::vector<float> m((size_t)f * cout * 2 * cin);
    for (int r = 0; r < f; ++r)
      for (int o = 0; o < cout; ++o)
        for (int tap = 0; tap < 2; ++tap)
          for (int c = 0; c < cin; ++c)
            m[(((size_t)r * cout + o) * 2 + tap) * cin + c] = w[((size_t)c * cout + o) * k + kidx[r * 2 + tap]];
    const HostTensor& b = get(prefix + ".bias");
    AFTER_REQUIRE(b.numel() == cout, AFTER_ESHAPE, "tensor '" + prefix + ".bias' has an unexpected shape");
    std::vector<float> bias((size_t)f * cout);
    for (int r = 0; r < f; ++r) std::copy(b.data.begin(), b.data.end(), bias.begin() + (size_t)r * cout);
    build_gemm_weight(L.w, *arena, m, bias.data(), f * cout, cin, taps, tc_mode());
    L.cin = cin; L.cout = cout; L.in_phases = 1; L.out_phases = f;
  }

  void make_snake(NormAct& a, const std::string& prefix, int C) {
    const HostTensor& al = get(prefix + ".alpha");
    const HostTensor& be = get(prefix + ".beta");
    AFTER_REQUIRE(al.numel() == C && be.numel() == C, AFTER_ESHAPE, "tensor '" + prefix + ".alpha' has an unexpected shape");
    std::vector<float> ib(C);
    for (int c = 0; c < C; ++c) ib[c] = 1.0f / (be.data[c] + 1e-9f);
    a.C = C;
    a.act = ACT_SNAKE;
    a.alpha = arena->upload(al.data);
    a.inv_beta = arena->upload(ib);
  }
  // ConvBlock1d prologue: CachedGroupNorm(min(C, groups)) -> SnakeBeta   (SimpleNetsStream.py:150-194)
  void make_gn_snake(NormAct& a, const std::string& prefix, int C, int groups) {
    make_snake(a, prefix + ".net.1", C);
    a.norm = NORM_GROUP;
    a.groups = std::min(C, groups);
    AFTER_REQUIRE(C % a.groups == 0, AFTER_EINVAL, "GroupNorm channels not divisible by groups");
    const HostTensor& g = get(prefix + ".net.0.gn.weight");
    const HostTensor& b = get(prefix + ".net.0.gn.bias");
    AFTER_REQUIRE(g.numel() == C && b.numel() == C, AFTER_ESHAPE, "tensor '" + prefix + ".net.0.gn.weight' has an unexpected shape");
    a.gamma = arena->upload(g.data);
    a.beta = arena->upload(b.data);
  }
  // eval-mode BatchNorm1d -> SiLU   (encoder.py:39-48)
  void make_bn_silu(NormAct& a, const std::string& prefix, int C) {
    const HostTensor& w = get(prefix + ".weight");
    const HostTensor& b = get(prefix + ".bias");
    const HostTensor& rm = get(prefix + ".running_mean");
    const HostTensor& rv = get(prefix + ".running_var");
    AFTER_REQUIRE(w.numel() == C && b.numel() == C && rm.numel() == C && rv.numel() == C, AFTER_ESHAPE,
                  "tensor '" + prefix + ".weight' has an unexpected shape");
    std::vector<float> rs(C);
    for (int c = 0; c < C; ++c) rs[c] = w.data[c] / std::sqrt(rv.data[c] + 1e-5f);
    a.C = C;
    a.norm = NORM_AFFINE;
    a.act = ACT_SILU;
    a.mu = arena->upload(rm.data);
    a.rs = arena->upload(rs);
    a.be = arena->upload(b.data);
  }

  // elems_per_stream: largest activation (frames x channels); op_elems_per_stream: largest conv operand, which may be
  // channel-padded to 64 (see make_conv)
  void alloc_workspace(size_t elems_per_stream, size_t op_elems_per_stream, int max_batch, int slots) {
    maxB = max_batch;
    buf_elems = elems_per_stream * (size_t)max_batch;
    op_elems = std::max(elems_per_stream, op_elems_per_stream) * (size_t)max_batch;
    for (auto& b : buf) b = arena->alloc<float>(buf_elems);
    alloc_operand(op, *arena, op_elems, tc_mode(), true);
    stat_slots = slots;
    slot_doubles = (size_t)max_batch * 8 * 2;
    stats = arena->alloc<double>(slot_doubles * slots);
    graphs.init();
  }

  // ---- streaming: cached_conv / CachedGroupNorm(stream=True) state ------------------------------------------------
  // The reference's streaming export builds its convs under cc.use_cached_conv(True): every conv with padding (l, r)
  // keeps the last l + r (+ stride delay) frames of its input and convolves [cache ; x] un-padded (CachedConv1d), the
  // residual branches are delayed to match (AlignBranches), GroupNorm normalises over [previous P frames ; x]
  // (SimpleNetsStream.py:134-144).  cached_conv (acids-ircam/cached_conv >= 2.5.0) is un-vendored: semantics restated in
  // oracle/after_oracle_stream.py.  Here the cache IS the front of the conv's persistent operand buffer: the operand
  // pass writes the new frames behind it, the tap-GEMM reads [cache ; new] through a TMA map whose taps are all
  // non-negative shifts, and one roll kernel at the end of the call moves every conv's tail to its front.
  struct StreamConv {
    GemmWeight w;                       // same device weights as the offline layer, streaming tap table
    int ns = 0, ns_al = 0, slab = 0;    // cached frames, rounded up to in_phases, frames per stream slab
    int Cp = 0, in_phases = 1, t_scale = 1;  // t_scale: input frames of this conv per latent frame
    bool tc = false;
  };
  struct StreamRes { int d = 0, C = 0, t_scale = 1; };
  struct StreamGn { int groups = 1, C = 0, t_scale = 1, dir = 0; };  // dir: which frame counter (0 encoder, 1 decoder)
  std::vector<StreamConv> sconv;
  std::vector<StreamRes> sres;
  std::vector<StreamGn> sgn;
  struct Slot {
    std::vector<ActOperand> op;   // per streaming conv
    std::vector<float*> res;      // per delayed residual: [B][d][C]
    std::vector<double*> hist;    // per GroupNorm: [B][cap][groups][2]
    RollDesc* descs = nullptr;
    long long* seen = nullptr;    // [2] latent frames consumed so far: encoder, decoder
  };
  std::vector<Slot> slots;
  int stream_max_lat = 0;  // latent frames per streaming call the state is sized for
  int gn_lat = 64;         // CachedGroupNorm padding_size in latent frames (the export's first call: 131072 samples)
  float* xd_buf = nullptr; // delayed residual of the current block
  struct StreamCtx { int slot = 0; int dir = 0; bool convs = false; };
  const StreamCtx* sctx = nullptr;  // non-null while a streaming body runs

  static int stride_delay(int r_pad, int cd, int stride) { return (stride - ((r_pad + cd) % stride)) % stride; }

  // Register a conv for streaming.  (l, r): its padding; cd: cumulative delay in front of it (CachedConv1d.__init__).
  // Returns the cumulative delay behind it.
  int reg_conv(ConvLayer& L, int k, int dil, int l, int r, int stride, int cd, int t_scale) {
    const int sd = stride_delay(r, cd, stride);
    const int ns = l + r + sd;
    const int out_cd = (r + sd + cd) / stride;
    if (ns == 0) return out_cd;
    AFTER_REQUIRE(L.in_phases == stride && L.out_phases == 1, AFTER_EINVAL, "streaming conv with an unexpected layout");
    StreamConv c;
    c.ns = ns; c.ns_al = ceil_div(ns, stride) * stride; c.in_phases = stride; c.t_scale = t_scale; c.Cp = L.w.Cin;
    c.slab = c.ns_al + stream_max_lat * t_scale;
    // (Tried: the fp32 SIMT kernel for live-sized buffers -- exact, but slower: 11.2 ms per Streamer.forward buffer with all
    // convs on it, 7.8 ms with only the narrow layers, 7.1 ms with the tcgen05 kernel everywhere.)
    c.tc = tc_mode() && L.w.tc_ok;
    c.w = L.w;
    TapTable t;
    t.ntaps = k;
    const int off = c.ns_al - ns;
    for (int i = 0; i < k; ++i) {
      const int p = i * dil + off;
      t.phase[0][i] = (int8_t)(p % stride);
      t.shift[0][i] = (int16_t)(p / stride);
    }
    c.w.taps = t;
    L.sid = (int)sconv.size();
    sconv.push_back(c);
    return out_cd;
  }
  int reg_res(ResBlock& r, int d, int C, int t_scale) {
    r.delay = d;
    if (d == 0) return -1;
    AFTER_REQUIRE((size_t)d * C <= (size_t)256 * RD_MAX, AFTER_EINVAL, "residual delay state too large");
    r.rid = (int)sres.size();
    sres.push_back(StreamRes{d, C, t_scale});
    return r.rid;
  }
  void reg_gn(NormAct& a, int t_scale, int dir) {
    if (a.norm != NORM_GROUP) return;
    AFTER_REQUIRE(256 % a.groups == 0, AFTER_EINVAL, "streaming GroupNorm needs a group count that divides 256");
    a.gid = (int)sgn.size();
    sgn.push_back(StreamGn{a.groups, a.C, t_scale, dir});
  }
  // ResnetBlock1d under cached_conv (SimpleNetsStream.py:197-254): block convs are built with cumulative_delay = 0, the
  // skip branch is delayed by block1's delay d, the block adds d to the running delay.
  int reg_res_block(ResBlock& r, int k, int dil, bool causal, int cd, int t_scale, int dir) {
    const int p = (k - 1) * dil + 1;
    const int l = k == 1 ? 0 : (causal ? p / 2 + (p - 1) / 2 : (p - 1) / 2), rr = k == 1 ? 0 : (causal ? 0 : p / 2);
    const int d = reg_conv(r.c1, k, dil, l, rr, 1, 0, t_scale);
    reg_res(r, d, r.c1.cin, t_scale);
    reg_gn(r.a1, t_scale, dir);
    reg_gn(r.a2, t_scale, dir);
    return cd + d;
  }
  // allocate the per-slot state once everything is registered
  void alloc_stream_state(int n_slots, int max_batch) {
    slots.resize(n_slots);
    for (Slot& s : slots) {
      std::vector<RollDesc> descs;
      for (const StreamConv& c : sconv) {
        ActOperand o;
        const size_t el = (size_t)c.slab * c.Cp * max_batch;
        o.capacity = el;
        o.bstride = (size_t)c.slab * c.Cp;
        RollDesc d{};
        if (c.tc) {
          o.hi = arena->alloc<__nv_bfloat16>(el);
          o.lo = arena->alloc<__nv_bfloat16>(el);
          AFTER_CUDA_CHECK(cudaMemset(o.hi, 0, el * 2));
          AFTER_CUDA_CHECK(cudaMemset(o.lo, 0, el * 2));
          d.a = o.hi; d.b = nprod() > 1 ? (void*)o.lo : nullptr; d.elem_bytes = 2;
        } else {
          o.f32 = arena->alloc<float>(el);
          AFTER_CUDA_CHECK(cudaMemset(o.f32, 0, el * 4));
          d.a = o.f32; d.b = nullptr; d.elem_bytes = 4;
        }
        AFTER_REQUIRE(((size_t)c.Cp * d.elem_bytes) % 16 == 0, AFTER_EINVAL, "streaming conv operand rows must be 16-byte multiples");
        d.ns = c.ns_al; d.Cp = c.Cp; d.slab = c.slab; d.t_scale = c.t_scale;
        descs.push_back(d);
        s.op.push_back(o);
      }
      for (const StreamRes& r : sres) {
        float* p = arena->alloc<float>((size_t)max_batch * r.d * r.C);
        AFTER_CUDA_CHECK(cudaMemset(p, 0, (size_t)max_batch * r.d * r.C * 4));
        s.res.push_back(p);
      }
      for (const StreamGn& g : sgn) {
        const size_t n = (size_t)max_batch * gn_cap(g) * g.groups * 2;
        double* p = arena->alloc<double>(n);
        AFTER_CUDA_CHECK(cudaMemset(p, 0, n * 8));
        s.hist.push_back(p);
      }
      if (descs.empty()) descs.push_back(RollDesc{});
      s.descs = arena->upload(descs);
      if (!xd_buf) xd_buf = arena->alloc<float>(buf_elems);
      s.seen = arena->alloc<long long>(2);
      AFTER_CUDA_CHECK(cudaMemset(s.seen, 0, 16));
    }
  }
  int gn_cap(const StreamGn& g) const { return (gn_lat + stream_max_lat + 8) * g.t_scale; }  // + 8: the decoder's z_buffer frames
  // back to the state of a freshly constructed model: zero caches, zero GroupNorm history
  void reset_stream(int slot, int max_batch, cudaStream_t st) {
    Slot& s = slots.at(slot);
    for (size_t i = 0; i < sconv.size(); ++i) {
      const size_t el = (size_t)sconv[i].slab * sconv[i].Cp * max_batch;
      if (s.op[i].hi) { AFTER_CUDA_CHECK(cudaMemsetAsync(s.op[i].hi, 0, el * 2, st)); AFTER_CUDA_CHECK(cudaMemsetAsync(s.op[i].lo, 0, el * 2, st)); }
      if (s.op[i].f32) AFTER_CUDA_CHECK(cudaMemsetAsync(s.op[i].f32, 0, el * 4, st));
    }
    for (size_t i = 0; i < sres.size(); ++i) AFTER_CUDA_CHECK(cudaMemsetAsync(s.res[i], 0, (size_t)max_batch * sres[i].d * sres[i].C * 4, st));
    for (size_t i = 0; i < sgn.size(); ++i)
      AFTER_CUDA_CHECK(cudaMemsetAsync(s.hist[i], 0, (size_t)max_batch * gn_cap(sgn[i]) * sgn[i].groups * 2 * 8, st));
    AFTER_CUDA_CHECK(cudaMemsetAsync(s.seen, 0, 16, st));
  }
  // end of a streaming call of `T_lat` latent frames: roll the conv caches, advance the frame counter
  void finish_stream_call(int B, int T_lat, cudaStream_t st) {
    const Slot& s = slots.at(sctx->slot);
    if (sctx->convs && !sconv.empty()) {
      launch_k(stream_roll_kernel, dim3((unsigned)sconv.size(), B), dim3(256), 0, st, (const RollDesc*)s.descs, T_lat, s.seen + sctx->dir);
    } else {  // no conv state on this path (offline decoder with streaming GroupNorm): only the counter moves
      launch_k(stream_advance_kernel, dim3(1), dim3(32), 0, st, T_lat, s.seen + sctx->dir);
    }
    AFTER_COUNT_LAUNCH();
  }

  // ---- run-time helpers -------------------------------------------------------------------
  double* new_slot() {
    AFTER_REQUIRE(stat_next < stat_slots, AFTER_ESTATE, "statistics arena exhausted");
    return stats + slot_doubles * (stat_next++);
  }
  void begin(cudaStream_t st) {
    stat_next = 0;
    AFTER_CUDA_CHECK(cudaMemsetAsync(stats, 0, slot_doubles * stat_slots * sizeof(double), st));
  }

  // operand <- act(norm(x)) in the format the consuming conv wants
  void produce(const float* x, const NormAct& a, const double* xstats, const ConvLayer& consumer, int B, int T, int C,
               cudaStream_t st) {
    const int Cp = consumer.w.Cin;  // >= C: operand channels the consumer's K loop walks (zero-padded)
    AFTER_REQUIRE(Cp >= C && (size_t)B * T * C <= buf_elems && (size_t)B * T * Cp <= op_elems, AFTER_EINVAL,
                  "activation exceeds the codec workspace");
    ActParams p;
    p.norm = a.norm; p.act = a.act;
    p.groups = a.groups; p.gamma = a.gamma; p.beta = a.beta;
    p.mu = a.mu; p.rs = a.rs; p.be = a.be; p.alpha = a.alpha; p.inv_beta = a.inv_beta;
    if (a.norm == NORM_GROUP && sctx) {
      // CachedGroupNorm stream branch: statistics over [previous P frames ; these T frames]
      AFTER_REQUIRE(a.gid >= 0, AFTER_ESTATE, "GroupNorm was not registered for streaming");
      const StreamGn& g = sgn[a.gid];
      const Slot& sl = slots.at(sctx->slot);
      const int P = gn_lat * g.t_scale;
      AFTER_REQUIRE(T <= gn_cap(g) - P, AFTER_EINVAL, "streaming buffer longer than the state was sized for");
      double* sw = new_slot();
      launch_k(gn_window_kernel, dim3(ceil_div(P + T, GN_WIN_ENTRIES), B), dim3(256), 0, st, x, sl.hist[a.gid], sw, sl.seen + g.dir,
               g.t_scale, P, gn_cap(g), T, C, g.groups);
      AFTER_COUNT_LAUNCH();
      xstats = sw;
      p.stat_frames = P + T;
    }
    p.stats = xstats;
    if (a.norm == NORM_GROUP) AFTER_REQUIRE(xstats != nullptr, AFTER_ESTATE, "GroupNorm input has no statistics");
    ActOperand* dst = &op;
    int out_T = T, out_t0 = 0;
    bool use_tc = tc_mode() && consumer.w.tc_ok;
    if (sctx && sctx->convs && consumer.sid >= 0) {  // streaming conv: the new frames go behind its cached context
      const StreamConv& sc = sconv[consumer.sid];
      AFTER_REQUIRE(T <= sc.slab - sc.ns_al, AFTER_EINVAL, "streaming buffer longer than the state was sized for");
      dst = &slots.at(sctx->slot).op[consumer.sid];
      out_T = sc.slab; out_t0 = sc.ns_al;
      use_tc = sc.tc;
    }
    OperandOut o;
    if (use_tc) { o.hi = dst->hi; o.lo = nprod() > 1 ? dst->lo : nullptr; }
    else o.f32 = dst->f32;
    const double el = (double)B * T;
    ProfScope prof(KC_ACT_OPERAND, st, 0.0, el * (4.0 * C + Cp * (o.f32 ? 4.0 : (o.lo ? 4.0 : 2.0))));
    const int q = Cp / 4;
    int R = 0;  // frame rows per block of the register-parameter kernel: q * R threads, a multiple of 32, <= 256
    if (C % 4 == 0 && Cp % 4 == 0 && q <= 256)
      for (int r = 256 / q; r >= 1; --r)
        if ((q * r) % 32 == 0) { R = r; break; }
    if (R > 0) {
      // ~32 K elements per block, but at least ~4 blocks per SM over the whole launch
      int fpb = std::max(R, 32768 / Cp);
      while (fpb > 4 * R && (long long)ceil_div(T, fpb) * B < 592) fpb /= 2;
      fpb = ceil_div(fpb, R) * R;
      launch_k(act_operand_rows_kernel, dim3(ceil_div(T, fpb), B), dim3(q * R), 0, st, x, o, p, T, C, Cp, fpb, out_T, out_t0, R);
    } else {
      const int fpb = std::max(1, 8192 / Cp);
      dim3 grid(ceil_div(T, fpb), B);
      const size_t smem = (size_t)5 * C * sizeof(float);
      if (C % 4 == 0 && Cp % 4 == 0) launch_k(act_operand_kernel<4>, grid, dim3(256), smem, st, x, o, p, T, C, Cp, fpb, out_T, out_t0);
      else launch_k(act_operand_kernel<1>, grid, dim3(256), smem, st, x, o, p, T, C, Cp, fpb, out_T, out_t0);
    }
    AFTER_COUNT_LAUNCH();
  }

  // out (B, T_rows * out_phases, cout) <- conv(operand) + bias (+ res); optional statistics of the result.
  // T_rows = GEMM rows per stream = input frames / in_phases.
  void conv(const ConvLayer& L, int B, int T_rows, float* out, const float* res, double* out_stats, int stat_groups,
            cudaStream_t st) {
    GemmEpi e;
    e.out_f32 = out;
    e.ldo = L.w.N;
    e.res = res;
    if (out_stats && !sctx) {  // streaming GroupNorms take their statistics from the window kernel, not from the producer
      e.stats = out_stats;
      e.stat_groups = stat_groups;
      e.stat_cpg = L.cout / stat_groups;
      e.stat_cmod = L.out_phases > 1 ? L.cout : 0;
    }
    if (sctx && sctx->convs && L.sid >= 0) {
      const StreamConv& sc = sconv[L.sid];
      tap_gemm(slots.at(sctx->slot).op[L.sid], B, T_rows, L.in_phases, sc.w, e, sc.tc ? precision : (int)AFTER_PRECISION_FP32_SIMT, st,
               sc.ns_al / L.in_phases + T_rows);
      return;
    }
    tap_gemm(op, B, T_rows, L.in_phases, L.w, e, precision, st);
  }

  // ResnetBlock1d: block2(block1(x)) + skip(x)   (SimpleNetsStream.py:197-254).  x -> out (distinct buffers), y1 scratch.
  void res_block(const ResBlock& r, const float* x, const double* xstats, float* y1, float* out, double* out_stats,
                 int out_groups, int B, int T, cudaStream_t st) {
    const float* res = x;
    if (sctx && sctx->convs && r.delay > 0) {  // AlignBranches: the skip branch sees x delayed by block1's conv delay
      launch_k(res_delay_kernel, dim3(B), dim3(256), 0, st, x, slots.at(sctx->slot).res[r.rid], xd_buf, r.delay, T, r.c1.cin);
      AFTER_COUNT_LAUNCH();
      res = xd_buf;
    }
    if (r.has_skip) {
      NormAct ident;
      produce(res, ident, nullptr, r.skip, B, T, r.skip.cin, st);
      conv(r.skip, B, T, out, nullptr, nullptr, 1, st);
      res = out;
    }
    double* s1 = new_slot();
    produce(x, r.a1, xstats, r.c1, B, T, r.c1.cin, st);
    conv(r.c1, B, T, y1, nullptr, s1, r.a2.groups, st);
    produce(y1, r.a2, s1, r.c2, B, T, r.c2.cin, st);
    conv(r.c2, B, T, out, res, out_stats, out_groups, st);
  }

  void make_res(ResBlock& r, const std::string& prefix, int cin, int cout, int k, int dilation, int groups, bool causal) {
    make_gn_snake(r.a1, prefix + ".net.branches.0.0", cin, groups);
    make_conv(r.c1, prefix + ".net.branches.0.0.net.2", cin, cout, k, conv_taps(k, dilation, causal), 1);
    make_gn_snake(r.a2, prefix + ".net.branches.0.1", cout, 8);
    make_conv(r.c2, prefix + ".net.branches.0.1.net.2", cout, cout, 1, conv_taps(1, 1, false), 1);
    r.has_skip = has(prefix + ".net.branches.1.weight_v");
    if (r.has_skip) make_conv(r.skip, prefix + ".net.branches.1", cin, cout, 1, conv_taps(1, 1, false), 1);
    else AFTER_REQUIRE(cin == cout, AFTER_EMISSING, "missing skip convolution '" + prefix + ".net.branches.1'");
  }
};

// ===========================================================================================
struct Codec : ConvNet {
  after_config cfg{};
  int M = 16, n_stages = 0, nb = 0, ratio = 1;
  std::vector<int> ech, dch;
  // encoder
  ResBlock to_in;
  struct Down { std::vector<ResBlock> res; NormAct snake; ConvLayer conv; int f; };
  std::vector<Down> downs;
  NormAct enc_out_snake; ConvLayer enc_out;
  // decoder
  ConvLayer dec_in;
  struct Up { NormAct snake; ConvLayer conv; std::vector<ResBlock> res; int f; };
  std::vector<Up> ups;
  NormAct out_a1, out_a2; ConvLayer out_c1, out_c2;
  // filter banks
  float *pq_fwd = nullptr, *pq_inv = nullptr;
  int pq_fwd_k = 0, pq_inv_k = 0;
  float *io_audio = nullptr, *io_z = nullptr;  // graph-stable staging of the public tensors

  void finalize(const after_config& c, const TensorMap& tensors, int prec, Arena* ar) {
    cfg = c; precision = prec; arena = ar; sd = &tensors;
    M = c.ae_pqmf_bands; n_stages = c.ae_n_stages; nb = c.ae_num_blocks;
    AFTER_REQUIRE(M == 16 && c.ae_in_channels == 16, AFTER_EINVAL, "codec requires the 16-band PQMF front end (pqmf_bands = in_channels = 16)");
    AFTER_REQUIRE(n_stages >= 1 && n_stages <= AFTER_MAX_STAGES && nb >= 1 && nb <= AFTER_MAX_STAGES, AFTER_EINVAL, "bad codec stage counts");
    AFTER_REQUIRE(c.ae_max_samples >= 1 && c.max_batch >= 1, AFTER_EINVAL, "ae_max_samples / max_batch must be >= 1");
    const int ks = c.ae_kernel_size, G = 8;
    ech.clear(); dch.clear();
    for (int i = 0; i <= n_stages; ++i) { ech.push_back(c.ae_channels * c.ae_multipliers[i]); dch.push_back(c.ae_channels * c.ae_dec_multipliers[i]); }
    ratio = M;
    for (int i = 0; i < n_stages; ++i) ratio *= c.ae_factors[i];

    // ---- encoder (SimpleNetsStream.py:400-459)
    make_res(to_in, "encoder.net.0", c.ae_in_channels, ech[0], ks, 1, G, false);
    downs.resize(n_stages);
    for (int i = 0; i < n_stages; ++i) {
      const std::string p = "encoder.net." + std::to_string(i + 1);
      Down& d = downs[i];
      d.f = c.ae_factors[i];
      d.res.resize(nb);
      for (int j = 0; j < nb; ++j) make_res(d.res[j], p + ".net." + std::to_string(j), ech[i], ech[i], ks, c.ae_dilations[j], G, false);
      make_snake(d.snake, p + ".net." + std::to_string(nb), ech[i]);
      make_conv(d.conv, p + ".net." + std::to_string(nb + 1), ech[i], ech[i + 1], 2 * d.f, strided_taps(d.f), d.f);
    }
    make_snake(enc_out_snake, "encoder.net." + std::to_string(n_stages + 1), ech[n_stages]);
    make_conv(enc_out, "encoder.net." + std::to_string(n_stages + 2), ech[n_stages], c.ae_z_channels, 3, conv_taps(3, 1, false), 1);

    // ---- decoder (SimpleNetsStream.py:552-651)
    make_conv(dec_in, "decoder.net.0", c.ae_z_channels, dch[0], ks, conv_taps(ks, 1, false), 1);
    ups.resize(n_stages);
    for (int i = 0; i < n_stages; ++i) {
      const std::string p = "decoder.net." + std::to_string(i + 1);
      Up& u = ups[i];
      u.f = c.ae_factors[n_stages - 1 - i];
      make_snake(u.snake, p + ".net.0", dch[i]);
      make_conv_transposed(u.conv, p + ".net.1", dch[i], dch[i + 1], u.f);
      u.res.resize(nb);
      for (int j = 0; j < nb; ++j) make_res(u.res[j], p + ".net." + std::to_string(2 + j), dch[i + 1], dch[i + 1], ks, c.ae_dilations[j], G, false);
    }
    const int out_c = c.ae_in_channels * (c.ae_use_loudness ? 2 : 1);
    make_gn_snake(out_a1, "decoder.synth.branches.0.net.0", dch[n_stages], G);
    make_conv(out_c1, "decoder.synth.branches.0.net.0.net.2", dch[n_stages], out_c, ks, conv_taps(ks, 1, false), 1);
    make_gn_snake(out_a2, "decoder.synth.branches.0.net.1", out_c, 8);
    make_conv(out_c2, "decoder.synth.branches.0.net.1.net.2", out_c, out_c, 1, conv_taps(1, 1, false), 1);

    // ---- PQMF banks (pqmf.py:186-279): forward (M, 1, K) -> [K][M]; inverse (M, M, K) -> [K][c][m]
    {
      const HostTensor& f = get("pqmf.forward_conv.weight");
      AFTER_REQUIRE(f.shape.size() == 3 && f.shape[0] == M && f.shape[1] == 1, AFTER_ESHAPE, "pqmf.forward_conv.weight has an unexpected shape");
      pq_fwd_k = (int)f.shape[2];
      std::vector<float> wt((size_t)pq_fwd_k * M);
      for (int m = 0; m < M; ++m)
        for (int k = 0; k < pq_fwd_k; ++k) wt[(size_t)k * M + m] = f.data[(size_t)m * pq_fwd_k + k];
      pq_fwd = arena->upload(wt);
      const HostTensor& v = get("pqmf.inverse_conv.weight");
      AFTER_REQUIRE(v.shape.size() == 3 && v.shape[0] == M && v.shape[1] == M, AFTER_ESHAPE, "pqmf.inverse_conv.weight has an unexpected shape");
      pq_inv_k = (int)v.shape[2];
      std::vector<float> it((size_t)pq_inv_k * M * M);
      for (int m = 0; m < M; ++m)
        for (int cc = 0; cc < M; ++cc)
          for (int k = 0; k < pq_inv_k; ++k) it[((size_t)k * M + cc) * M + m] = v.data[((size_t)m * M + cc) * pq_inv_k + k];
      pq_inv = arena->upload(it);
      const size_t smem_a = ((size_t)pq_fwd_k * M + PQ_FRAMES * M + pq_fwd_k) * sizeof(float);
      const size_t smem_s = ((size_t)pq_inv_k * M * M + (size_t)(PS_FRAMES + pq_inv_k - 1) * M) * sizeof(float);
      AFTER_REQUIRE(smem_a <= 200 * 1024 && smem_s <= 200 * 1024, AFTER_EINVAL, "PQMF filters too long for shared memory");
      AFTER_CUDA_CHECK(cudaFuncSetAttribute(pqmf_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
      AFTER_CUDA_CHECK(cudaFuncSetAttribute(pqmf_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    }

    // ---- workspace: the largest activation (frames x channels) of either net at the longest input
    AFTER_REQUIRE(c.ae_max_samples % ratio == 0, AFTER_EINVAL, "ae_max_samples must be a multiple of the codec ratio");
    size_t T = (size_t)(c.ae_max_samples / M), mx = T * std::max(c.ae_in_channels, out_c), mxp = T * 64;
    for (int i = 0; i <= n_stages; ++i) {
      mx = std::max(mx, T * (size_t)ech[i]);
      mxp = std::max(mxp, T * (size_t)std::max(ech[i], 64));
      if (i < n_stages) T /= c.ae_factors[i];
    }
    mx = std::max(mx, T * (size_t)c.ae_z_channels);
    for (int i = 0; i <= n_stages; ++i) {
      mx = std::max(mx, T * (size_t)dch[i]);
      mxp = std::max(mxp, T * (size_t)std::max(dch[i], 64));
      if (i < n_stages) T *= c.ae_factors[n_stages - 1 - i];
    }
    // the workspace holds 2 x max_batch chunks: the audio-to-audio chain encodes its structure and timbre inputs as ONE
    // batch (encode_pair) -- the encoder's kernels are launch- and latency-bound at 8 chunks (profiles/r01c_codec_sweep.jsonl)
    alloc_workspace(mx, mxp, 2 * c.max_batch, 4 * (n_stages * nb + 4) + 16);
    maxB = c.max_batch;  // the public limit (and the size of the streaming state) stays max_batch; wsB = what the workspace holds
    wsB = 2 * c.max_batch;
    io_audio = arena->alloc<float>((size_t)2 * c.max_batch * c.ae_max_samples);
    io_z = arena->alloc<float>((size_t)2 * c.max_batch * c.ae_z_channels * (c.ae_max_samples / ratio));

    // ---- streaming export (after_scripts/export_autoencoder.py:16-153, 305-319: AE_notcausal = cached encoder + offline
    // decoder over [z_buffer ; z] with overlap-add, CachedGroupNorm.stream = True in both)
    if (c.stream_slots > 0) {
      AFTER_REQUIRE(c.stream_slots <= 8, AFTER_EINVAL, "stream_slots must be <= 8");
      n_fade = 4;  // export_autoencoder.py:57
      gn_lat = c.stream_gn_frames > 0 ? c.stream_gn_frames : 64;
      const int64_t max_lat = c.ae_max_samples / ratio - n_fade;
      AFTER_REQUIRE(max_lat >= n_fade, AFTER_EINVAL, "ae_max_samples too small for the streaming decoder (needs n_fade + buffer frames)");
      stream_max_lat = (int)std::min<int64_t>(c.stream_max_frames > 0 ? c.stream_max_frames : 64, max_lat);
      int ts = ratio / M;  // frames at the PQMF rate per latent frame
      int cd = reg_res_block(to_in, ks, 1, false, 0, ts, 0);
      for (int i = 0; i < n_stages; ++i) {
        Down& d = downs[i];
        for (int j = 0; j < nb; ++j) cd = reg_res_block(d.res[j], ks, c.ae_dilations[j], false, cd, ts, 0);
        cd = reg_conv(d.conv, 2 * d.f, 1, d.f - 1, d.f, d.f, cd, ts);  // Downsample1d: padding get_padding(2f, f) = (f-1, f)
        AFTER_REQUIRE(ts % d.f == 0, AFTER_EINVAL, "codec ratio bookkeeping");
        ts /= d.f;
      }
      cd = reg_conv(enc_out, 3, 1, 1, 1, 1, cd, ts);
      enc_delay = cd;
      ts = 1;
      for (int i = 0; i < n_stages; ++i) {
        ts *= ups[i].f;
        for (int j = 0; j < nb; ++j) { reg_gn(ups[i].res[j].a1, ts, 1); reg_gn(ups[i].res[j].a2, ts, 1); }
      }
      reg_gn(out_a1, ts, 1);
      reg_gn(out_a2, ts, 1);
      alloc_stream_state(c.stream_slots, c.max_batch);
      for (int sl = 0; sl < c.stream_slots; ++sl) {
        const size_t nz = (size_t)c.max_batch * c.ae_z_channels * n_fade, no = (size_t)c.max_batch * ratio * n_fade;
        float* zb = arena->alloc<float>(nz);
        float* ob = arena->alloc<float>(no);
        AFTER_CUDA_CHECK(cudaMemset(zb, 0, nz * 4));
        AFTER_CUDA_CHECK(cudaMemset(ob, 0, no * 4));
        z_buffer.push_back(zb);
        out_buffer.push_back(ob);
      }
      zcat = arena->alloc<float>((size_t)c.max_batch * c.ae_z_channels * (stream_max_lat + n_fade));
      ycat = arena->alloc<float>((size_t)c.max_batch * ratio * (stream_max_lat + n_fade));
    }
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    sd = nullptr;
  }
  int n_fade = 4, enc_delay = 0;
  std::vector<float*> z_buffer, out_buffer;  // per slot: AE_notcausal's z_buffer (B, Z, n_fade), out_buffer (B, 1, ratio n_fade)
  float *zcat = nullptr, *ycat = nullptr;

  void destroy() { graphs.destroy(); }

  void check(int B, int64_t samples) {
    AFTER_REQUIRE(B >= 1 && B <= maxB, AFTER_EINVAL, "batch exceeds max_batch given at after_create");
    AFTER_REQUIRE(samples >= ratio && samples % ratio == 0, AFTER_EINVAL, "samples must be a positive multiple of the codec ratio");
    AFTER_REQUIRE(samples <= cfg.ae_max_samples, AFTER_EINVAL, "samples exceed ae_max_samples given at after_create");
  }

  // ------------------------------------------------------------------ AutoEncoder.encode
  void encode_body(const float* audio, float* z, int B, int64_t samples, cudaStream_t st) {
    PdlScope pdl(true);  // every kernel of the codec graphs starts with pdl_wait() (common.cuh)
    begin(st);
    int T = (int)(samples / M);
    float *x = buf[0], *y1 = buf[1], *o = buf[2];
    {
      dim3 grid(ceil_div(T, PQ_FRAMES), B);
      const size_t smem = ((size_t)pq_fwd_k * M + PQ_FRAMES * M + pq_fwd_k) * sizeof(float);
      const int p = (pq_fwd_k - 1) + 1;  // get_padding(k): left = (p - 1) / 2
      ProfScope prof(KC_PQMF, st, 2.0 * B * T * M * pq_fwd_k, (double)B * T * M * 8.0);
      launch_k(pqmf_analysis_kernel, grid, dim3(256), smem, st, audio, pq_fwd, x, T, pq_fwd_k, (p - 1) / 2);
      AFTER_COUNT_LAUNCH();
    }
    double* xs = new_slot();
    gn_stats_launch(x, xs, B, T, cfg.ae_in_channels, to_in.a1.groups, st);
    double* os = new_slot();
    res_block(to_in, x, xs, y1, o, os, 8, B, T, st);
    std::swap(x, o); xs = os;
    for (int i = 0; i < n_stages; ++i) {
      Down& d = downs[i];
      for (int j = 0; j < nb; ++j) {
        const bool need = j + 1 < nb;  // the stage's last block feeds a bare Snake, not a GroupNorm
        os = need ? new_slot() : nullptr;
        res_block(d.res[j], x, xs, y1, o, os, 8, B, T, st);
        std::swap(x, o); xs = os;
      }
      produce(x, d.snake, nullptr, d.conv, B, T, ech[i], st);
      T /= d.f;
      const bool need = i + 1 < n_stages;  // next stage starts with a GroupNorm; after the last one a Snake follows
      os = need ? new_slot() : nullptr;
      conv(d.conv, B, T, o, nullptr, os, 8, st);
      std::swap(x, o); xs = os;
    }
    produce(x, enc_out_snake, nullptr, enc_out, B, T, ech[n_stages], st);
    conv(enc_out, B, T, o, nullptr, nullptr, 1, st);
    // ReluBottleneck is the identity at inference (SimpleNetsStream.py:753-760); public layout is channel-first
    dim3 grid(ceil_div(T, 32), ceil_div(cfg.ae_z_channels, 32), B);
    launch_k(tokens_to_channels_kernel, grid, dim3(256), 0, st, o, z, (int)cfg.ae_z_channels, T);
    AFTER_COUNT_LAUNCH();
  }

  // The captured graphs work on library-owned staging buffers (io_audio / io_z), so they never bake caller pointers
  // in: one graph per (direction, B, length), replayed for any caller tensors.
  void encode(const float* audio, float* z, int B, int64_t samples, cudaStream_t st) {
    check(B, samples);
    const int T = (int)(samples / ratio);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_audio, audio, (size_t)B * samples * sizeof(float), cudaMemcpyDeviceToDevice, st));
    graphs.run({0, B, (int)(samples & 0x7fffffff)}, st, [&] { encode_body(io_audio, io_z, B, samples, st); });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(z, io_z, (size_t)B * cfg.ae_z_channels * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }

  // two encode calls with the same B and length as one batch of 2 B chunks: z1 = encode(audio1), z2 = encode(audio2)
  // (chunks are independent in every kernel of the encoder, GroupNorm statistics are per chunk)
  void encode_pair(const float* audio1, const float* audio2, float* z1, float* z2, int B, int64_t samples, cudaStream_t st) {
    check(B, samples);
    AFTER_REQUIRE(2 * B <= wsB, AFTER_ESTATE, "codec workspace does not hold 2 x B chunks");
    const int T = (int)(samples / ratio);
    const size_t na = (size_t)B * samples, nz = (size_t)B * cfg.ae_z_channels * T;
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_audio, audio1, na * sizeof(float), cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_audio + na, audio2, na * sizeof(float), cudaMemcpyDeviceToDevice, st));
    graphs.run({0, 2 * B, (int)(samples & 0x7fffffff)}, st, [&] { encode_body(io_audio, io_z, 2 * B, samples, st); });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(z1, io_z, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(z2, io_z + nz, nz * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }

  // ------------------------------------------------------------------ AutoEncoder.decode
  void decode_body(const float* z, float* audio, int B, int Tz, cudaStream_t st) {
    PdlScope pdl(true);
    begin(st);
    int T = Tz;
    float *x = buf[0], *y1 = buf[1], *o = buf[2];
    {
      dim3 grid(ceil_div(T, 32), ceil_div(cfg.ae_z_channels, 32), B);
      launch_k(channels_to_frames_kernel, grid, dim3(256), 0, st, z, o, (int)cfg.ae_z_channels, T);
      AFTER_COUNT_LAUNCH();
    }
    NormAct ident;
    produce(o, ident, nullptr, dec_in, B, T, cfg.ae_z_channels, st);
    conv(dec_in, B, T, x, nullptr, nullptr, 1, st);
    double *xs = nullptr, *os = nullptr;
    for (int i = 0; i < n_stages; ++i) {
      Up& u = ups[i];
      produce(x, u.snake, nullptr, u.conv, B, T, dch[i], st);
      os = new_slot();
      conv(u.conv, B, T, o, nullptr, os, 8, st);
      T *= u.f;
      std::swap(x, o); xs = os;
      for (int j = 0; j < nb; ++j) {
        // after a stage's last block comes a bare Snake (next stage) -- except the final stage, whose output
        // feeds the GroupNorm of the synthesis head
        const bool need = j + 1 < nb || i + 1 == n_stages;
        os = need ? new_slot() : nullptr;
        res_block(u.res[j], x, xs, y1, o, os, std::min(dch[i + 1], 8), B, T, st);
        std::swap(x, o); xs = os;
      }
    }
    // synthesis head: two ConvBlock1d without residual (SimpleNetsStream.py:612-633), loudness gate, inverse PQMF
    os = new_slot();
    produce(x, out_a1, xs, out_c1, B, T, dch[n_stages], st);
    conv(out_c1, B, T, y1, nullptr, os, out_a2.groups, st);
    produce(y1, out_a2, os, out_c2, B, T, out_c2.cin, st);
    conv(out_c2, B, T, o, nullptr, nullptr, 1, st);
    {
      dim3 grid(ceil_div(T, PS_FRAMES), B);
      const size_t smem = ((size_t)pq_inv_k * M * M + (size_t)(PS_FRAMES + pq_inv_k - 1) * M) * sizeof(float);
      ProfScope prof(KC_PQMF, st, 2.0 * B * T * M * M * pq_inv_k, (double)B * T * M * (cfg.ae_use_loudness ? 12.0 : 8.0));
      launch_k(pqmf_synthesis_kernel, grid, dim3(256), smem, st, o, pq_inv, audio, T, pq_inv_k, (pq_inv_k - 1) / 2,
               (int)cfg.ae_use_loudness);
      AFTER_COUNT_LAUNCH();
    }
  }

  void decode(const float* z, float* audio, int B, int T, cudaStream_t st) {
    AFTER_REQUIRE(T >= 1, AFTER_EINVAL, "T must be >= 1");
    check(B, (int64_t)T * ratio);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_z, z, (size_t)B * cfg.ae_z_channels * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    graphs.run({1, B, T}, st, [&] { decode_body(io_z, io_audio, B, T, st); });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(audio, io_audio, (size_t)B * T * ratio * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }

  // ------------------------------------------------------------------ streaming export (export_stream.ts)
  void check_stream(int slot, int B, int T_lat) {
    AFTER_REQUIRE(!slots.empty(), AFTER_ESTATE, "this handle was created with stream_slots == 0 (no streaming codec state)");
    AFTER_REQUIRE(slot >= 0 && slot < (int)slots.size(), AFTER_EINVAL, "stream slot out of range");
    AFTER_REQUIRE(B >= 1 && B <= maxB, AFTER_EINVAL, "batch exceeds max_batch given at after_create");
    AFTER_REQUIRE(T_lat >= 1 && T_lat <= stream_max_lat, AFTER_EINVAL, "buffer exceeds stream_max_frames given at after_create");
  }
  // AE_notcausal.encode on one buffer: offline PQMF of the buffer, cached-conv encoder, streaming GroupNorm
  void encode_stream(int slot, const float* audio, float* z, int B, int64_t samples, cudaStream_t st) {
    AFTER_REQUIRE(samples >= ratio && samples % ratio == 0, AFTER_EINVAL, "samples must be a positive multiple of the codec ratio");
    const int T = (int)(samples / ratio);
    check_stream(slot, B, T);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_audio, audio, (size_t)B * samples * sizeof(float), cudaMemcpyDeviceToDevice, st));
    graphs.run({10 + slot, B, T}, st, [&] {
      StreamCtx ctx{slot, 0, true};
      sctx = &ctx;
      struct Clear { const StreamCtx*& p; ~Clear() { p = nullptr; } } clear{sctx};
      encode_body(io_audio, io_z, B, samples, st);
      finish_stream_call(B, T, st);
    });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(z, io_z, (size_t)B * cfg.ae_z_channels * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // AE_notcausal.decode (export_autoencoder.py:128-153): decode [z_buffer ; z] offline (streaming GroupNorm), cross-fade the
  // first n_fade frames with the tail kept from the previous call, keep the new tail, return the first T frames of audio
  void decode_stream(int slot, const float* z, float* audio, int B, int T, cudaStream_t st) {
    check_stream(slot, B, T);
    AFTER_REQUIRE(T >= n_fade, AFTER_EINVAL, "the streaming decoder needs buffers of at least n_fade (4) latent frames");
    const int Z = cfg.ae_z_channels, Tc = T + n_fade;
    AFTER_CUDA_CHECK(cudaMemcpyAsync(io_z, z, (size_t)B * Z * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    graphs.run({20 + slot, B, T}, st, [&] {
      // zcat (B, Z, n_fade + T) = [z_buffer ; z]; z_buffer <- its last n_fade frames
      AFTER_CUDA_CHECK(cudaMemcpy2DAsync(zcat, (size_t)Tc * 4, z_buffer[slot], (size_t)n_fade * 4, (size_t)n_fade * 4, (size_t)B * Z,
                                         cudaMemcpyDeviceToDevice, st));
      AFTER_CUDA_CHECK(cudaMemcpy2DAsync(zcat + n_fade, (size_t)Tc * 4, io_z, (size_t)T * 4, (size_t)T * 4, (size_t)B * Z,
                                         cudaMemcpyDeviceToDevice, st));
      AFTER_CUDA_CHECK(cudaMemcpy2DAsync(z_buffer[slot], (size_t)n_fade * 4, zcat + T, (size_t)Tc * 4, (size_t)n_fade * 4, (size_t)B * Z,
                                         cudaMemcpyDeviceToDevice, st));
      StreamCtx ctx{slot, 1, false};
      sctx = &ctx;
      struct Clear { const StreamCtx*& p; ~Clear() { p = nullptr; } } clear{sctx};
      decode_body(zcat, ycat, B, Tc, st);
      finish_stream_call(B, Tc, st);
      const int n_out = T * ratio, nf = n_fade * ratio;
      launch_k(overlap_add_kernel, dim3(ceil_div(n_out, 256), B), dim3(256), 0, st, (const float*)ycat, out_buffer[slot], io_audio, n_out, nf);
      AFTER_COUNT_LAUNCH();
    });
    AFTER_CUDA_CHECK(cudaMemcpyAsync(audio, io_audio, (size_t)B * T * ratio * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  void reset_stream_slot(int slot, cudaStream_t st) {
    check_stream(slot, 1, 1);
    reset_stream(slot, maxB, st);
    AFTER_CUDA_CHECK(cudaMemsetAsync(z_buffer[slot], 0, (size_t)maxB * cfg.ae_z_channels * n_fade * 4, st));
    AFTER_CUDA_CHECK(cudaMemsetAsync(out_buffer[slot], 0, (size_t)maxB * ratio * n_fade * 4, st));
  }
};

// ===========================================================================================
// Encoder1D (structure encoder), all ratios == 1   (encoder.py:74-113, 116-237, 273-298)
struct StructureEncoder : ConvNet {
  after_config cfg{};
  struct Block { NormAct a1, a2; ConvLayer c1, c2; int rid = -1, delay = 0; };
  std::vector<Block> blocks;   // n + 1 V2ConvBlock1D
  std::vector<ConvLayer> pools;  // n 1x1 convs
  std::vector<int> cins, couts;
  int n = 0, maxT = 0;

  void make_block(Block& b, const std::string& prefix, int C, int k, bool causal) {
    make_bn_silu(b.a1, prefix + ".net.branches.0.0", C);
    make_conv(b.c1, prefix + ".net.branches.0.2", C, C, k, conv_taps(k, 1, causal), 1);
    make_bn_silu(b.a2, prefix + ".net.branches.0.3", C);
    make_conv(b.c2, prefix + ".net.branches.0.6", C, C, k, conv_taps(k, 1, causal), 1);
  }

  void finalize(const after_config& c, const TensorMap& tensors, int prec, Arena* ar) {
    cfg = c; precision = prec; arena = ar; sd = &tensors;
    n = c.se_n_blocks;
    AFTER_REQUIRE(n >= 1 && n <= AFTER_MAX_STAGES, AFTER_EINVAL, "bad structure-encoder block count");
    cins.clear(); couts.clear();
    for (int i = 0; i < n; ++i) { cins.push_back(i == 0 ? c.se_in_size : c.se_channels[i - 1]); couts.push_back(c.se_channels[i]); }
    blocks.resize(n + 1);
    pools.resize(n);
    int mxc = c.se_in_size;
    for (int i = 0; i < n; ++i) {
      const std::string p = "net." + std::to_string(i);
      make_block(blocks[i], p + ".net.0", cins[i], c.se_kernel_size, c.se_causal != 0);
      make_conv(pools[i], p + ".net.1", cins[i], couts[i], 1, conv_taps(1, 1, false), 1);
      mxc = std::max(mxc, std::max(cins[i], couts[i]));
    }
    make_block(blocks[n], "net." + std::to_string(n), couts[n - 1], c.se_kernel_size, c.se_causal != 0);
    maxT = c.seq_len;
    alloc_workspace((size_t)maxT * mxc, (size_t)maxT * std::max(mxc, 64), c.max_batch, 1);
    // ---- Encoder1D.forward_stream (encoder.py:300-322) under cc.use_cached_conv(True) (after_scripts/export.py:14-17):
    // V2ConvBlock1D chains conv1 -> conv2 delays and delays its identity branch by their sum (zero with causal padding)
    if (c.stream_slots > 0) {
      stream_max_lat = std::min(c.stream_max_frames > 0 ? c.stream_max_frames : 64, maxT);
      const int k = c.se_kernel_size;
      const bool causal = c.se_causal != 0;
      const int l = k == 1 ? 0 : (causal ? k - 1 : (k - 1) / 2), r = k == 1 ? 0 : (causal ? 0 : k / 2);
      for (Block& b : blocks) {
        int cd = reg_conv(b.c1, k, 1, l, r, 1, 0, 1);
        cd = reg_conv(b.c2, k, 1, l, r, 1, cd, 1);
        b.delay = cd;
        if (cd > 0) {
          AFTER_REQUIRE((size_t)cd * b.c1.cin <= (size_t)256 * RD_MAX, AFTER_EINVAL, "residual delay state too large");
          b.rid = (int)sres.size();
          sres.push_back(StreamRes{cd, b.c1.cin, 1});
        }
      }
      alloc_stream_state(c.stream_slots, c.max_batch);
    }
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
    sd = nullptr;
  }

  // x + conv(SiLU(BN(conv(SiLU(BN(x))))))   (encoder.py:25-71, dropout off)
  void run_block(const Block& b, const float* x, float* y1, float* out, int B, int T, cudaStream_t st) {
    const float* res = x;
    if (sctx && b.delay > 0) {
      launch_k(res_delay_kernel, dim3(B), dim3(256), 0, st, x, slots.at(sctx->slot).res[b.rid], xd_buf, b.delay, T, b.c1.cin);
      AFTER_COUNT_LAUNCH();
      res = xd_buf;
    }
    produce(x, b.a1, nullptr, b.c1, B, T, b.c1.cin, st);
    conv(b.c1, B, T, y1, nullptr, nullptr, 1, st);
    produce(y1, b.a2, nullptr, b.c2, B, T, b.c2.cin, st);
    conv(b.c2, B, T, out, res, nullptr, 1, st);
  }

  void forward_body(const float* z, float* out, int B, int T, cudaStream_t st) {
    float *x = buf[0], *y1 = buf[1], *o = buf[2];
    {
      dim3 grid(ceil_div(T, 32), ceil_div(cfg.se_in_size, 32), B);
      channels_to_frames_kernel<<<grid, 256, 0, st>>>(z, x, cfg.se_in_size, T);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    NormAct ident;
    for (int i = 0; i < n; ++i) {
      run_block(blocks[i], x, y1, o, B, T, st);
      produce(o, ident, nullptr, pools[i], B, T, cins[i], st);
      conv(pools[i], B, T, x, nullptr, nullptr, 1, st);
    }
    run_block(blocks[n], x, y1, o, B, T, st);
    const int C = couts[n - 1];
    if (cfg.se_use_tanh) {
      const size_t nel = (size_t)B * T * C;
      tanh_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(o, nel);
      AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    }
    dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B);
    tokens_to_channels_kernel<<<grid, 256, 0, st>>>(o, out, C, T);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  void forward(const float* z, float* out, int B, int T, cudaStream_t st) {
    AFTER_REQUIRE(B >= 1 && B <= maxB, AFTER_EINVAL, "batch exceeds max_batch given at after_create");
    AFTER_REQUIRE(T >= 1 && T <= maxT, AFTER_EINVAL, "T exceeds seq_len given at after_create");
    forward_body(z, out, B, T, st);
  }
  // Encoder1D.forward_stream: one buffer of T latent frames against the cached left context of every conv
  void forward_stream(int slot, const float* z, float* out, int B, int T, cudaStream_t st) {
    AFTER_REQUIRE(!slots.empty(), AFTER_ESTATE, "this handle was created with stream_slots == 0 (no streaming state)");
    AFTER_REQUIRE(slot >= 0 && slot < (int)slots.size(), AFTER_EINVAL, "stream slot out of range");
    AFTER_REQUIRE(B >= 1 && B <= maxB, AFTER_EINVAL, "batch exceeds max_batch given at after_create");
    AFTER_REQUIRE(T >= 1 && T <= stream_max_lat, AFTER_EINVAL, "buffer exceeds stream_max_frames given at after_create");
    StreamCtx ctx{slot, 0, true};
    sctx = &ctx;
    struct Clear { const StreamCtx*& p; ~Clear() { p = nullptr; } } clear{sctx};
    forward_body(z, out, B, T, st);
    finish_stream_call(B, T, st);
  }
  void reset_stream_slot(int slot, cudaStream_t st) {
    AFTER_REQUIRE(slot >= 0 && slot < (int)slots.size(), AFTER_EINVAL, "stream slot out of range");
    reset_stream(slot, maxB, st);
  }
  void destroy() { graphs.destroy(); }
};

}  // namespace after
