// DenoiserV2 forward + rectified-flow sampling loop on one B200.
//
// What is hoisted out of the per-step work (all exact rewrites):
//   * AdaLN-t (alpha_t, beta_t) of all layers depends only on time_cond  -> one table per call;
//     the three CFG rows share it (rows 0..2B-1) or use one constant row (dropped condition).
//   * AdaLN-c (alpha_c, beta_c) of all layers depends only on (t_i, cond) -> one table for all
//     steps and both condition classes, built before the loop (fourier -> MLP -> 6 stacked linears).
//   * the three CFG rows share x, so patchify_and_embed runs on B sequences, not 3B.
//   * the band mask is index arithmetic inside the attention kernel; RoPE cos/sin is one table.
// The per-step work is then 2 row-wise kernels + 3 GEMMs per layer, out_proj, and a fused
// CFG-combine + Euler update; the whole N-step loop is captured into one CUDA graph per
// (B, T, steps, variant) signature and replayed.
#pragma once
#include <cmath>
#include <cstring>
#include <tuple>
#include "context.cuh"
#include "denoiser_kernels.cuh"
#include "gemm_host.cuh"
#include "stream_step.cuh"

namespace after {

struct DenoiserLayer {
  GemmWeight qkv, mlp0, mlp2;
  float *n1_g, *n1_b, *n3_g, *n3_b;
};

struct Denoiser {
  after_config cfg{};
  int precision = 0;
  int D = 0, C = 0, L = 0, H = 0, zs = 0, zt = 0, HID = 0, NE = 0;
  int maxN = 0, maxT = 0, maxB = 0, maxRows = 0, maxR = 0;
  Arena* arena = nullptr;

  // weights
  float *emb0_w, *emb0_b, *emb2_w, *emb2_b, *pe_wt, *pe_b, *tc_w, *tc_b;
  float *adaC_w, *adaC_b, *adaT_w, *adaT_b;  // stacked over layers: [L*2D, D], [L*2D, zs]
  GemmWeight out_proj;
  std::vector<DenoiserLayer> layers;
  float *rope_inv, *fourier_freq;
  float2* rope_tab;

  // workspace
  float *x_state, *x_in, *cond_buf, *tc_buf, *time_buf;
  float *h0, *h, *qkv, *proj;
  __nv_bfloat16* qkv_bf = nullptr;  // bf16 mode: the QKV buffer the attention kernel reads (half the epilogue / attention bytes)
  float* h_part = nullptr;  // second K half of a K-split down projection (added by the next layer's AdaLN-t kernel)
  ActOperand a_op, hid_op;  // GEMM A operands: LN output [rows, D], MLP hidden [rows, HID]
  float *tcemb, *adaT, *E, *F1, *feat, *adaC;
  int *map_src, *map_trow, *map_tstride, *map_crow;
  float* guidance;
  int* mlp_flags = nullptr;  // fused-MLP dependency counters: [L][row blocks]
  int flag_stride = 0;

  // streaming state (transformerv2.py:143-188): un-rotated key/value history per (diffusion step, layer, sequence),
  // zero-initialised like the reference's registered buffers, plus the q|k|v of the last cached forward of every layer
  // (the reference's last_k / last_v) which roll_cache appends.
  int cacheW = 0;                               // after_config.max_cache_size; 0 = offline only
  float *kcache = nullptr, *vcache = nullptr;   // [max_steps][L][maxN][cacheW][D]
  float* qkv_stream = nullptr;                  // [L][maxRows][3D]
  int last_N = 0, last_T = 0;
  // skinny path of small streaming blocks (rows <= SKINNY_ROWS): fp32 operand rows for skinny_linear_kernel
  static constexpr int SKINNY_ROWS = 16;
  float *sk_a = nullptr, *sk_hid = nullptr;     // [SKINNY_ROWS][D], [SKINNY_ROWS][HID]
  bool skinny_now = false;
  int cfg_streams = 0;  // > 0 while run_network works on a CFG batch of 3 x cfg_streams sequences (row order of launch_adaln_t)
  unsigned* stream_barrier = nullptr;           // grid-barrier counter of the persistent streaming-block kernel
  int n_sms = 0;

  // graph cache
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t kernels = 0;
  };
  std::map<std::tuple<int, int, int, int, int>, GraphEntry> graphs;
  bool use_graph = true;

  bool tc_mode() const { return precision != AFTER_PRECISION_FP32_SIMT; }
  int nprod() const { return precision == AFTER_PRECISION_BF16 ? 1 : 3; }

  // ------------------------------------------------------------------ weights
  float* up(const HostTensor& t) { return arena->upload(t.data); }

  void make_linear(GemmWeight& lw, const std::vector<float>& w, const float* bias_host, int N, int K) {
    build_gemm_weight(lw, *arena, w, bias_host, N, K, TapTable{}, tc_mode());
  }

  void finalize(const after_config& c, const TensorMap& sd, int prec, Arena* ar) {
    cfg = c;
    precision = prec;
    arena = ar;
    D = c.embed_dim; C = c.n_channels; L = c.n_layers; zs = c.tcond_dim; zt = c.cond_dim;
    HID = c.mlp_multiplier * D; NE = c.noise_embed_dims; H = D / 64;
    AFTER_REQUIRE(D == 256 || D == 512, AFTER_EINVAL, "embed_dim must be 256 or 512 (head_dim 64, 4 or 8 heads)");
    AFTER_REQUIRE(C % 4 == 0 && C <= 256, AFTER_EINVAL, "n_channels must be a multiple of 4");
    AFTER_REQUIRE(c.attention_chunk_size >= 1 && c.local_attention_size >= 1 &&
                      c.attention_chunk_size + c.local_attention_size - 1 <= 32,
                  AFTER_EINVAL, "attention chunk + window - 1 must be <= 32");
    AFTER_REQUIRE(NE % 2 == 0 && NE >= 2, AFTER_EINVAL, "noise_embed_dims must be even");
    AFTER_REQUIRE(c.max_batch >= 1 && c.seq_len >= 1 && c.max_steps >= 1, AFTER_EINVAL, "max_batch/seq_len/max_steps must be >= 1");
    maxB = c.max_batch; maxN = 3 * maxB; maxT = c.seq_len; maxRows = maxN * maxT;
    maxR = std::max(c.max_steps * (maxB + 1), maxN);

    const std::string tb = "denoiser_trans_block.";
    const int EIN = NE + zt;
    emb0_w = up(need(sd, "embedding.0.weight", {D, EIN}));
    emb0_b = up(need(sd, "embedding.0.bias", {D}));
    emb2_w = up(need(sd, "embedding.2.weight", {D, D}));
    emb2_b = up(need(sd, "embedding.2.bias", {D}));
    {
      const HostTensor& w = need(sd, tb + "patchify_and_embed.1.weight", {D, C});
      std::vector<float> wt((size_t)C * D);
      for (int d = 0; d < D; ++d)
        for (int cc = 0; cc < C; ++cc) wt[(size_t)cc * D + d] = w.data[(size_t)d * C + cc];
      pe_wt = arena->upload(wt);
      pe_b = up(need(sd, tb + "patchify_and_embed.1.bias", {D}));
    }
    tc_w = up(need(sd, tb + "patchify_and_embed_tcond.1.weight", {zs, zs}));
    tc_b = up(need(sd, tb + "patchify_and_embed_tcond.1.bias", {zs}));

    std::vector<float> cw((size_t)L * 2 * D * D), cb((size_t)L * 2 * D), tw((size_t)L * 2 * D * zs), tbv((size_t)L * 2 * D);
    layers.resize(L);
    for (int i = 0; i < L; ++i) {
      const std::string p = tb + "decoder_blocks." + std::to_string(i) + ".";
      DenoiserLayer& ly = layers[i];
      make_linear(ly.qkv, need(sd, p + "self_attention.qkv_linear.weight", {3 * D, D}).data, nullptr, 3 * D, D);
      make_linear(ly.mlp0, need(sd, p + "mlp.mlp.0.weight", {HID, D}).data, need(sd, p + "mlp.mlp.0.bias", {HID}).data.data(), HID, D);
      make_linear(ly.mlp2, need(sd, p + "mlp.mlp.2.weight", {D, HID}).data, need(sd, p + "mlp.mlp.2.bias", {D}).data.data(), D, HID);
      ly.n1_g = up(need(sd, p + "norm1.weight", {D}));
      ly.n1_b = up(need(sd, p + "norm1.bias", {D}));
      ly.n3_g = up(need(sd, p + "norm3.weight", {D}));
      ly.n3_b = up(need(sd, p + "norm3.bias", {D}));
      const HostTensor& lw = need(sd, p + "linear.weight", {2 * D, D});
      const HostTensor& lb = need(sd, p + "linear.bias", {2 * D});
      const HostTensor& tlw = need(sd, p + "tcond_linear.weight", {2 * D, zs});
      const HostTensor& tlb = need(sd, p + "tcond_linear.bias", {2 * D});
      std::copy(lw.data.begin(), lw.data.end(), cw.begin() + (size_t)i * 2 * D * D);
      std::copy(lb.data.begin(), lb.data.end(), cb.begin() + (size_t)i * 2 * D);
      std::copy(tlw.data.begin(), tlw.data.end(), tw.begin() + (size_t)i * 2 * D * zs);
      std::copy(tlb.data.begin(), tlb.data.end(), tbv.begin() + (size_t)i * 2 * D);
    }
    adaC_w = arena->upload(cw); adaC_b = arena->upload(cb);
    adaT_w = arena->upload(tw); adaT_b = arena->upload(tbv);
    make_linear(out_proj, need(sd, tb + "out_proj.0.weight", {C, D}).data, need(sd, tb + "out_proj.0.bias", {C}).data.data(), C, D);

    // rotary inverse frequencies: the checkpoint's own parameter when present (rotary_embedding.py:69,88)
    {
      const int half = 16;
      std::vector<float> inv(half);
      auto it = sd.find(tb + "rotary_emb.freqs");
      if (it != sd.end() && it->second.numel() == half) inv = it->second.data;
      else
        for (int j = 0; j < half; ++j) inv[j] = 1.0f / powf(10000.0f, (float)(2 * j) / 32.0f);
      rope_inv = arena->upload(inv);
      const int npos = maxT + std::max(0, (int)c.max_cache_size);  // streaming: positions run over history + block
      rope_tab = arena->alloc<float2>((size_t)npos * half);
      rope_table_kernel<<<ceil_div(npos * half, 256), 256>>>(rope_tab, rope_inv, npos, half);
      AFTER_CUDA_CHECK(cudaGetLastError());
      // fourier frequencies (1/max_positions)^(k/half) as fp32 (transformerv2.py:34-40)
      const int fh = NE / 2;
      std::vector<float> fr(fh);
      for (int k = 0; k < fh; ++k) fr[k] = (float)std::pow((double)(float)(1.0 / 10000.0), (double)((float)k / (float)fh));
      fourier_freq = arena->upload(fr);
    }

    // workspace
    x_state = arena->alloc<float>((size_t)maxN * C * maxT);
    x_in = arena->alloc<float>((size_t)maxN * C * maxT);
    cond_buf = arena->alloc<float>((size_t)maxN * zt);
    tc_buf = arena->alloc<float>((size_t)maxN * zs * maxT);
    time_buf = arena->alloc<float>((size_t)std::max(maxN, c.max_steps));
    h0 = arena->alloc<float>((size_t)maxRows * D);
    h = arena->alloc<float>((size_t)maxRows * D);
    h_part = arena->alloc<float>((size_t)maxRows * D);
    // + 32 zeroed rows: attn_chunk_group_kernel reads up to MAXK - 1 rows past the last chunk (masked, must be finite)
    qkv = arena->alloc<float>((size_t)(maxRows + 32) * 3 * D);
    AFTER_CUDA_CHECK(cudaMemset(qkv, 0, (size_t)(maxRows + 32) * 3 * D * sizeof(float)));
    if (precision == AFTER_PRECISION_BF16) {
      qkv_bf = arena->alloc<__nv_bfloat16>((size_t)(maxRows + 32) * 3 * D);
      AFTER_CUDA_CHECK(cudaMemset(qkv_bf, 0, (size_t)(maxRows + 32) * 3 * D * sizeof(__nv_bfloat16)));
    }
    proj = arena->alloc<float>((size_t)maxRows * C);
    alloc_operand(a_op, *arena, (size_t)maxRows * D, tc_mode(), false);
    alloc_operand(hid_op, *arena, (size_t)maxRows * HID, tc_mode(), false);
    const size_t ada_ld = (size_t)L * 2 * D;
    tcemb = arena->alloc<float>((size_t)(maxRows + 1) * zs);
    adaT = arena->alloc<float>((size_t)(maxRows + 1) * ada_ld);
    E = arena->alloc<float>((size_t)maxR * EIN);
    F1 = arena->alloc<float>((size_t)maxR * D);
    feat = arena->alloc<float>((size_t)maxR * D);
    adaC = arena->alloc<float>((size_t)maxR * ada_ld);
    map_src = arena->alloc<int>(maxN); map_trow = arena->alloc<int>(maxN);
    map_tstride = arena->alloc<int>(maxN); map_crow = arena->alloc<int>(maxN);
    guidance = arena->alloc<float>(4);
    flag_stride = 2 * maxN * ceil_div(maxT, 2 * tc::BM);  // two counters per row block (launch_mlp_fused)
    mlp_flags = arena->alloc<int>((size_t)L * flag_stride);
#ifdef AFTER_DEBUG
    {  // opt-in dynamic shared memory of the staged A/B variant (must not first happen inside a stream capture)
      auto set = [](const void* f, size_t bytes) {
        AFTER_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      };
      auto wbytes = [](int nh, int maxk) { return 128 + (size_t)(16 + maxk) * 2 * nh * 64 * sizeof(float); };
      set((const void*)attn_warp_chunk_kernel<8, 12, true, 4>, wbytes(8, 12));
      set((const void*)attn_warp_chunk_kernel<8, 20, true, 4>, wbytes(8, 20));
      set((const void*)attn_warp_chunk_kernel<4, 12, true, 4>, wbytes(4, 12));
      set((const void*)attn_warp_chunk_kernel<4, 20, true, 4>, wbytes(4, 20));
    }
#endif
    cacheW = c.max_cache_size;
    AFTER_REQUIRE(cacheW >= 0 && cacheW <= 64, AFTER_EINVAL, "max_cache_size must be in [0, 64]");
    if (cacheW > 0) {
      const size_t cache_floats = (size_t)c.max_steps * L * maxN * cacheW * D;
      kcache = arena->alloc<float>(cache_floats);
      vcache = arena->alloc<float>(cache_floats);
      qkv_stream = arena->alloc<float>((size_t)L * maxRows * 3 * D);
      AFTER_CUDA_CHECK(cudaMemset(kcache, 0, cache_floats * sizeof(float)));
      AFTER_CUDA_CHECK(cudaMemset(vcache, 0, cache_floats * sizeof(float)));
      sk_a = arena->alloc<float>((size_t)SKINNY_ROWS * D);
      sk_hid = arena->alloc<float>((size_t)SKINNY_ROWS * HID);
      AFTER_CUDA_CHECK(cudaFuncSetAttribute(skinny_linear_kernel<SKINNY_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)((size_t)SKINNY_ROWS * std::max(D, HID) * sizeof(float))));
      stream_barrier = arena->alloc<unsigned>(4);
      // operand rows + the LayerNorm parameters of every layer (stream_block_kernel keeps them in shared memory)
      const int ss_bytes = (int)(((size_t)SKINNY_ROWS * std::max(D, HID) + (size_t)L * 4 * D) * sizeof(float));
      AFTER_REQUIRE(ss_bytes <= 200 * 1024, AFTER_EINVAL, "embed_dim * mlp_multiplier too large for the streaming kernels");
      auto set_ss = [&](const void* f) { AFTER_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, ss_bytes)); };
      set_ss((const void*)stream_block_kernel<8, 12, 512>); set_ss((const void*)stream_block_kernel<8, 20, 256>); set_ss((const void*)stream_block_kernel<8, 32, 256>);
      set_ss((const void*)stream_block_kernel<4, 12, 512>); set_ss((const void*)stream_block_kernel<4, 20, 256>); set_ss((const void*)stream_block_kernel<4, 32, 256>);
      int dev = 0;
      AFTER_CUDA_CHECK(cudaGetDevice(&dev));
      AFTER_CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const char* ng = debug_env("AFTER_NO_GRAPH");
    use_graph = !(ng && ng[0] == '1');
    AFTER_CUDA_CHECK(cudaDeviceSynchronize());
  }

  void destroy() {
    for (auto& kv : graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }

  // ------------------------------------------------------------------ building blocks
  SeqMap seqmap() const { return SeqMap{map_src, map_trow, map_tstride, map_crow}; }
  // bf16 mode, offline attention by attn_chunk_group_kernel (chunk 4, band <= 20) behind the 256-wide CTA-pair QKV GEMM:
  // q | k | v are written and read as bf16 (AFTER_QKV_BF16=0 in debug builds keeps the fp32 buffer for A/B runs)
  bool qkv_bf16_ok(int l) const {
    static int off = -1;
    if (off < 0) { const char* e = debug_env("AFTER_QKV_BF16"); off = (e && e[0] == '0') ? 1 : 0; }
    return !off && qkv_bf != nullptr && cfg.attention_chunk_size == 4 && cfg.attention_chunk_size + cfg.local_attention_size - 1 <= 20 &&
           layers[l].qkv.tc2_ok && layers[l].qkv.bn2 == 256 && use_pair_kernel();
  }

  void gemm(const GemmWeight& w, ActOperand& a, int nseq, int T, const GemmEpi& epi, cudaStream_t st) {
    tap_gemm(a, nseq, T, 1, w, epi, precision, st);
  }
  RowOperandOut operand_out(const ActOperand& o) const {
    RowOperandOut r;
    if (skinny_now) { r.f32 = sk_a; return r; }  // the skinny linears read fp32 rows
    r.f32 = o.f32; r.hi = o.hi; r.lo = nprod() > 1 ? o.lo : nullptr;
    return r;
  }
  // out[M, N] = act(A[M, K] W^T + bias) (+ res), M <= SKINNY_ROWS (skinny_linear_kernel)
  void skinny(const float* A, const GemmWeight& w, const float* res, float* out, int M, int gelu, cudaStream_t st) {
    ProfScope prof(KC_TAP_GEMM_SIMT, st, 2.0 * M * w.N * w.K, (double)w.N * w.K * 4.0);
    launch_k(skinny_linear_kernel<SKINNY_ROWS>, dim3(ceil_div(w.N, 8)), dim3(256), (size_t)M * w.K * sizeof(float), st, A, w.w,
             w.bias, res, out, w.N, M, w.N, w.K, gelu);
    AFTER_COUNT_LAUNCH();
  }
  // small fp32 linears that run once per call (tables): C[M,N] = A[M,K] W[N,K]^T + bias
  void small_linear(const float* A, const float* W, const float* bias, float* Cout, int M, int N, int K, int gelu,
                    cudaStream_t st) {
    TapTable tt;
    if (K % 16 == 0 && N % 4 == 0 && !gelu) {
      GemmEpi e; e.out_f32 = Cout; e.ldo = N; e.bias = bias;
      dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM), 1);
      tap_gemm_simt_kernel<<<grid, 256, 0, st>>>(A, W, e, tt, M, 1, K, N, M, (size_t)0);
      AFTER_CUDA_CHECK(cudaGetLastError());
      AFTER_COUNT_LAUNCH();
      return;
    }
    const size_t total = (size_t)M * N;
    tap_gemm_naive_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, W, bias, nullptr, Cout, tt, 1, M, 1, K, N, gelu, M, (size_t)0);
    AFTER_CUDA_CHECK(cudaGetLastError());
    AFTER_COUNT_LAUNCH();
  }

  // tables that depend on (time_cond) and on (times x cond classes)
  void build_tables(int n_tc_seq, int T, int n_time_rows, int times_per_row, int n_cond, int n_cls, cudaStream_t st) {
    const int ada_ld = L * 2 * D;
    const int tc_rows = n_tc_seq * T + 1;
    tcond_embed_kernel<<<ceil_div(tc_rows * zs, 256), 256, 0, st>>>(tc_buf, tc_w, tc_b, tcemb, n_tc_seq, zs, T, cfg.drop_value);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    small_linear(tcemb, adaT_w, adaT_b, adaT, tc_rows, ada_ld, zs, 0, st);
    const int R = n_time_rows;  // rows of the (time, class) table
    const int EIN = NE + zt;
    fourier_concat_kernel<<<ceil_div(R * EIN, 256), 256, 0, st>>>(time_buf, times_per_row, cond_buf, n_cond, zt, fourier_freq,
                                                                  NE / 2, 100.0f, cfg.drop_value, E, R, n_cls);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
    small_linear(E, emb0_w, emb0_b, F1, R, D, EIN, 1, st);
    small_linear(F1, emb2_w, emb2_b, feat, R, D, D, 0, st);
    small_linear(feat, adaC_w, adaC_b, adaC, R, ada_ld, D, 0, st);
  }

  template <int NV>
  void launch_adaln_t(const float* hin, const float* hadd, int use_src, int l, int rows, int T, cudaStream_t st) {
    RowOperandOut o = operand_out(a_op);
    ProfScope prof(KC_ROW_NORM, st, 0.0, (double)rows * D * (tc_mode() ? 4.0 * 2 + 2.0 * (nprod() > 1 ? 2 : 1) : 12.0));
    // CFG batches (rows = 3 B T): groups 0 and 1 of a (stream, frame) go to neighbouring warps (see the kernel)
    static int pair_off = -1;  // AFTER_ADALN_PAIR=0 (debug builds): plain row order, for A/B runs
    if (pair_off < 0) { const char* e = debug_env("AFTER_ADALN_PAIR"); pair_off = (e && e[0] == '0') ? 1 : 0; }
    const int Bs = cfg_streams;
    const int pair_frames = (!pair_off && Bs > 0 && rows == 3 * Bs * T) ? Bs * T : 0;
    const int blocks = pair_frames ? ceil_div(pair_frames, 4) + ceil_div(rows - 2 * pair_frames, 8) : ceil_div(rows, 8);
    launch_k(adaln_t_ln1_kernel<NV>, dim3(blocks), dim3(256), 0, st, hin, hadd, h, o, adaT, L * 2 * D, l * 2 * D, seqmap(),
             use_src, layers[l].n1_g, layers[l].n1_b, rows, T, pair_frames, Bs);
    AFTER_COUNT_LAUNCH();
  }
  template <int NH, int MAXK>
  void launch_attn(int l, const float* adaC_step, int rows, int T, cudaStream_t st) {
    RowOperandOut o = operand_out(a_op);
    // q,k,v read once + h read/write + operand write; ~2 * keys * 64 * 2 flops per (token, head)
    ProfScope prof(KC_ATTENTION, st, (double)rows * H * 64.0 * 4.0 * (cfg.attention_chunk_size + cfg.local_attention_size - 1),
                   (double)rows * D * (12.0 + 8.0 + (tc_mode() ? 2.0 * (nprod() > 1 ? 2 : 1) : 4.0)));
    const int n_seq = rows / T, chunks = n_seq * ((T + 3) / 4);
    bool done = false;
    if constexpr (MAXK <= 20) {  // wider bands: sc[4][MAXK] would spill -> one-warp-per-token kernel below
      if (cfg.attention_chunk_size == 4) {
#define AFTER_LAUNCH_ATTN_WARP(STAGED, Q, GRID, SMEM)                                                                     \
  launch_k(attn_warp_chunk_kernel<NH, MAXK, STAGED, Q>, dim3(GRID), dim3(128), SMEM, st, qkv, h, o, adaC_step, L * 2 * D,  \
           l * 2 * D, seqmap(), layers[l].n3_g, layers[l].n3_b, n_seq, T, cfg.local_attention_size,                       \
           mlp_flags + (size_t)l * flag_stride, flag_stride)
#ifdef AFTER_DEBUG
        // A/B variants (debug builds only; base B=8 fp32, profiles/r01b_ab_attn*.jsonl): AFTER_ATTN=staged stages the key
        // rows in shared memory (-3.4 %), AFTER_ATTN_QPW = 2 | 1 spreads a chunk's queries over 2 / 4 warps (-4 / -7 %)
        static int staged = -1, qpw = -1;
        if (staged < 0) { const char* e = debug_env("AFTER_ATTN"); staged = (e && !strcmp(e, "staged")) ? 1 : 0; }
        if (qpw < 0) { const char* e = debug_env("AFTER_ATTN_QPW"); qpw = e ? atoi(e) : 4; }
        // AFTER_ATTN: "warp" = the one-warp-per-chunk kernel, "g2" / "g4" = 2 / 4 chunks per block with all key rows in
        // one batch (profiles/r02m..o_ab_attn_*.jsonl); default = the release configuration below
        static int group = -1;
        if (group < 0) {
          const char* e = debug_env("AFTER_ATTN");
          group = !e ? 1 : !strcmp(e, "warp") ? 0 : !strcmp(e, "g4") ? 4 : !strcmp(e, "g2") ? 2 : 1;
        }
        constexpr int KB0 = MAXK <= 12 ? MAXK : MAXK / 2;
#define AFTER_LAUNCH_ATTN_GROUP(CPB, KB, MINB)                                                                            \
  launch_k(attn_chunk_group_kernel<NH, MAXK, CPB, KB, MINB, float>, dim3(ceil_div((T + 3) / 4, CPB), n_seq),              \
           dim3(CPB * (NH / 2) * 32), 0, st, qkv, h, o, adaC_step, L * 2 * D, l * 2 * D, seqmap(), layers[l].n3_g,        \
           layers[l].n3_b, T, cfg.local_attention_size, mlp_flags + (size_t)l * flag_stride, flag_stride)
        if (group == 1 && qkv_bf16_ok(l)) {
          launch_k(attn_chunk_group_kernel<NH, MAXK, 8 / NH, KB0 / 2, 6, __nv_bfloat16>, dim3(ceil_div((T + 3) / 4, 8 / NH), n_seq),
                   dim3(128), 0, st, qkv_bf, h, o, adaC_step, L * 2 * D, l * 2 * D, seqmap(), layers[l].n3_g, layers[l].n3_b, T,
                   cfg.local_attention_size, mlp_flags + (size_t)l * flag_stride, flag_stride);
        } else if (group == 1) {
          AFTER_LAUNCH_ATTN_GROUP(8 / NH, KB0 / 2, 6);
        } else if (group == 2) {
          AFTER_LAUNCH_ATTN_GROUP(2, KB0, 512 / (NH * 32));
        } else if (group == 4) {
          AFTER_LAUNCH_ATTN_GROUP(4, KB0, 256 / (NH * 32));
        } else
        if (staged && T % 16 == 0) {
          const size_t smem = 128 + (size_t)(16 + cfg.local_attention_size - 1) * 2 * D * sizeof(float);
          AFTER_LAUNCH_ATTN_WARP(true, 4, chunks / 4, smem);
        } else if (qpw == 2) {
          AFTER_LAUNCH_ATTN_WARP(false, 2, ceil_div(chunks * 2, 4), 0);
        } else if (qpw == 1) {
          AFTER_LAUNCH_ATTN_WARP(false, 1, chunks, 0);
        } else
          AFTER_LAUNCH_ATTN_WARP(false, 4, ceil_div(chunks, 4), 0);
#undef AFTER_LAUNCH_ATTN_GROUP
#else
        // NH / 2 warps per chunk, 128-thread blocks, key / value rows in two batches, 6 blocks (24 warps) per SM: the
        // best of the variants measured (profiles/r02m..o_ab_attn_*.jsonl); the one-warp-per-chunk kernel stays in debug
        // builds for A/B runs (AFTER_ATTN=warp)
        if (qkv_bf16_ok(l))
          launch_k(attn_chunk_group_kernel<NH, MAXK, 8 / NH, (MAXK <= 12 ? MAXK : MAXK / 2) / 2, 6, __nv_bfloat16>,
                   dim3(ceil_div((T + 3) / 4, 8 / NH), n_seq), dim3(128), 0, st, qkv_bf, h, o, adaC_step, L * 2 * D, l * 2 * D,
                   seqmap(), layers[l].n3_g, layers[l].n3_b, T, cfg.local_attention_size, mlp_flags + (size_t)l * flag_stride,
                   flag_stride);
        else
          launch_k(attn_chunk_group_kernel<NH, MAXK, 8 / NH, (MAXK <= 12 ? MAXK : MAXK / 2) / 2, 6, float>,
                   dim3(ceil_div((T + 3) / 4, 8 / NH), n_seq), dim3(128), 0, st, qkv, h, o, adaC_step, L * 2 * D, l * 2 * D,
                   seqmap(), layers[l].n3_g, layers[l].n3_b, T, cfg.local_attention_size, mlp_flags + (size_t)l * flag_stride,
                   flag_stride);
#endif
#undef AFTER_LAUNCH_ATTN_WARP
        done = true;
      }
    }
    if (!done) {
      launch_k(attn_adaln_c_ln3_kernel<NH, MAXK>, dim3(ceil_div(rows, 4)), dim3(128), 0, st, qkv, h, o, adaC_step, L * 2 * D,
               l * 2 * D, seqmap(), layers[l].n3_g, layers[l].n3_b, rows, T, cfg.attention_chunk_size,
               cfg.local_attention_size, mlp_flags + (size_t)l * flag_stride, flag_stride);
    }
    AFTER_COUNT_LAUNCH();
  }
  void attn(int l, const float* adaC_step, int rows, int T, cudaStream_t st) {
    const int mk = cfg.attention_chunk_size + cfg.local_attention_size - 1;
    if (H == 8) {
      if (mk <= 12) launch_attn<8, 12>(l, adaC_step, rows, T, st);
      else if (mk <= 20) launch_attn<8, 20>(l, adaC_step, rows, T, st);
      else launch_attn<8, 32>(l, adaC_step, rows, T, st);
    } else {
      if (mk <= 12) launch_attn<4, 12>(l, adaC_step, rows, T, st);
      else if (mk <= 20) launch_attn<4, 20>(l, adaC_step, rows, T, st);
      else launch_attn<4, 32>(l, adaC_step, rows, T, st);
    }
  }

  // ---- streaming (cached) attention: history rows come from the cache of (cache_index, layer), RoPE is applied on the fly
  size_t cache_slab() const { return (size_t)L * maxN * cacheW * D; }  // floats per diffusion step
  template <int NH, int MAXK>
  void launch_attn_stream(int l, int cache_index, const float* adaC_step, int rows, int T, cudaStream_t st) {
    RowOperandOut o = operand_out(a_op);
    ProfScope prof(KC_ATTENTION, st, (double)rows * H * 64.0 * 4.0 * (cfg.attention_chunk_size + cfg.local_attention_size - 1),
                   (double)rows * D * (12.0 + 8.0 + (tc_mode() ? 2.0 * (nprod() > 1 ? 2 : 1) : 4.0)));
    const size_t off = (size_t)cache_index * cache_slab() + (size_t)l * maxN * cacheW * D;
    launch_k(attn_stream_kernel<NH, MAXK>, dim3(rows), dim3(NH * 32), 0, st, qkv_stream + (size_t)l * maxRows * 3 * D,
             kcache + off, vcache + off, cacheW, rope_tab, h, o, adaC_step, L * 2 * D, l * 2 * D, seqmap(), layers[l].n3_g,
             layers[l].n3_b, rows, T, cfg.attention_chunk_size, cfg.local_attention_size,
             mlp_flags + (size_t)l * flag_stride, flag_stride);
    AFTER_COUNT_LAUNCH();
  }
  void attn_stream(int l, int cache_index, const float* adaC_step, int rows, int T, cudaStream_t st) {
    const int mk = cfg.attention_chunk_size + cfg.local_attention_size - 1;
    if (H == 8) {
      if (mk <= 12) launch_attn_stream<8, 12>(l, cache_index, adaC_step, rows, T, st);
      else if (mk <= 20) launch_attn_stream<8, 20>(l, cache_index, adaC_step, rows, T, st);
      else launch_attn_stream<8, 32>(l, cache_index, adaC_step, rows, T, st);
    } else {
      if (mk <= 12) launch_attn_stream<4, 12>(l, cache_index, adaC_step, rows, T, st);
      else if (mk <= 20) launch_attn_stream<4, 20>(l, cache_index, adaC_step, rows, T, st);
      else launch_attn_stream<4, 32>(l, cache_index, adaC_step, rows, T, st);
    }
  }

  // DenoiserV2.roll_cache (transformerv2.py:167-186, 433-435): every layer's history of `cache_index` takes the first
  // min(roll_size, T_last) frames of the last cached forward and keeps the newest cacheW.
  void roll_cache(int roll_size, int cache_index, cudaStream_t st) {
    AFTER_REQUIRE(cacheW > 0, AFTER_ESTATE, "this handle was created with max_cache_size == 0 (offline only)");
    AFTER_REQUIRE(cache_index >= 0 && cache_index < cfg.max_steps, AFTER_EINVAL, "cache_index must be in [0, max_steps)");
    AFTER_REQUIRE(roll_size >= 0, AFTER_EINVAL, "roll_size must be >= 0");
    AFTER_REQUIRE(last_N > 0, AFTER_ESTATE, "roll_cache before any cached forward");
    const int r = std::min(roll_size, last_T);
    if (r == 0) return;
    const size_t off = (size_t)cache_index * cache_slab();
    dim3 grid(L, last_N, 2);
    PdlScope pdl(true);
    launch_k(kv_roll_kernel, grid, dim3(256), 0, st, kcache + off, vcache + off, qkv_stream, maxN, maxRows, last_T, cacheW, r, D);
    AFTER_COUNT_LAUNCH();
  }

  void reset_cache(cudaStream_t st) {
    AFTER_REQUIRE(cacheW > 0, AFTER_ESTATE, "this handle was created with max_cache_size == 0 (offline only)");
    const size_t bytes = (size_t)cfg.max_steps * cache_slab() * sizeof(float);
    AFTER_CUDA_CHECK(cudaMemsetAsync(kcache, 0, bytes, st));
    AFTER_CUDA_CHECK(cudaMemsetAsync(vcache, 0, bytes, st));
    last_N = last_T = 0;
  }

  void check_cache_index(int cache_index) {
    AFTER_REQUIRE(cacheW > 0, AFTER_ESTATE, "this handle was created with max_cache_size == 0 (offline only)");
    AFTER_REQUIRE(cache_index >= 0 && cache_index < cfg.max_steps, AFTER_EINVAL, "cache_index must be in [0, max_steps)");
  }

  // One network evaluation: x_src (n_src, C, T) channel-first -> proj [N*T, C] token-major.
  // cache_index >= 0 selects the streaming attention (history of that diffusion step + this block).
  void run_network(const float* x_src, int n_src, int N, int T, const float* adaC_step, cudaStream_t st,
                   int cache_index = -1) {
    const int rows = N * T;
    cfg_streams = (N == 3 * n_src) ? n_src : 0;
    PdlScope pdl(true);  // every kernel below starts with pdl_wait(): programmatic dependent launches are safe
    // a live streaming block is a dozen rows: weight-streaming fp32 linears instead of 256-row tensor-core tiles
    skinny_now = cache_index >= 0 && rows <= SKINNY_ROWS && sk_a != nullptr;
    struct Reset { bool& f; ~Reset() { f = false; } } reset_skinny{skinny_now};
    {
      // 4 frames per block: 16-frame blocks (4x less W_in^T traffic from L2, 128 blocks) measured 1.4 % slower
      dim3 grid(ceil_div(T, 4), n_src);
      launch_k(patch_embed_kernel<4>, grid, dim3(256), C * 4 * sizeof(float), st, x_src, pe_wt, pe_b, h0, C, T, D);
      AFTER_COUNT_LAUNCH();
    }
    bool part = false;  // previous layer's down projection left its second K half in h_part
    for (int l = 0; l < L; ++l) {
      const float* hin = l == 0 ? h0 : h;
      const float* hadd = part ? h_part : nullptr;
      if (D == 512) launch_adaln_t<16>(hin, hadd, l == 0, l, rows, T, st);
      else launch_adaln_t<8>(hin, hadd, l == 0, l, rows, T, st);
      part = false;
      if (cache_index < 0) {
        GemmEpi e; e.ldo = 3 * D; e.rope = 1; e.D = D; e.rot_half = 16; e.rope_tab = rope_tab;
        if (qkv_bf16_ok(l)) e.out_hi = qkv_bf;
        else e.out_f32 = qkv;
        gemm(layers[l].qkv, a_op, N, T, e, st);
        attn(l, adaC_step, rows, T, st);
      } else {  // keys are cached un-rotated: plain epilogue into this layer's slot, rotation inside the attention kernel
        float* q_l = qkv_stream + (size_t)l * maxRows * 3 * D;
        if (skinny_now) {
          skinny(sk_a, layers[l].qkv, nullptr, q_l, rows, 0, st);
        } else {
          GemmEpi e; e.out_f32 = q_l; e.ldo = 3 * D;
          gemm(layers[l].qkv, a_op, N, T, e, st);
        }
        attn_stream(l, cache_index, adaC_step, rows, T, st);
      }
      if (skinny_now) {
        skinny(sk_a, layers[l].mlp0, nullptr, sk_hid, rows, 1, st);
        skinny(sk_hid, layers[l].mlp2, h, h, rows, 0, st);
      } else {
        GemmEpi e0; e0.ldo = HID; e0.gelu = 1;
        e0.out_f32 = hid_op.f32; e0.out_hi = hid_op.hi; e0.out_lo = nprod() > 1 ? hid_op.lo : nullptr;
        GemmEpi e1; e1.out_f32 = h; e1.ldo = D; e1.res = h;
        if (l == L - 1 && tc_mode()) { e1.out_hi = a_op.hi; e1.out_lo = nprod() > 1 ? a_op.lo : nullptr; }
        // one persistent launch for both MLP projections when the shapes allow it, else two launches
        int* fl = mlp_flags + (size_t)l * flag_stride;
        if (!(tc_mode() && launch_mlp_fused(a_op, layers[l].mlp0, e0, hid_op, layers[l].mlp2, e1, N, T, precision, fl, st,
                                            l < L - 1 ? h_part : nullptr, &part))) {
          gemm(layers[l].mlp0, a_op, N, T, e0, st);
          gemm(layers[l].mlp2, hid_op, N, T, e1, st);
        }
      }
    }
    if (skinny_now) {
      skinny(h, out_proj, nullptr, proj, rows, 0, st);
    } else {
      GemmEpi e; e.out_f32 = proj; e.ldo = C;
      if (!tc_mode()) {  // fp32 mode: the operand is h itself
        ActOperand hop; hop.f32 = h; hop.capacity = (size_t)maxRows * D;
        gemm(out_proj, hop, N, T, e, st);
      } else {
        gemm(out_proj, a_op, N, T, e, st);
      }
    }
    if (cache_index >= 0) { last_N = N; last_T = T; }
  }

  void upload_maps(const std::vector<int>& src, const std::vector<int>& trow, const std::vector<int>& tstr,
                   const std::vector<int>& crow, cudaStream_t st) {
    const size_t b = src.size() * sizeof(int);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(map_src, src.data(), b, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(map_trow, trow.data(), b, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(map_tstride, tstr.data(), b, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(map_crow, crow.data(), b, cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));  // host vectors are temporaries
  }

  void check_shape(int N, int T) {
    AFTER_REQUIRE(N >= 1 && N <= maxN, AFTER_EINVAL, "batch exceeds 3*max_batch given at after_create");
    AFTER_REQUIRE(T >= 1 && T <= maxT, AFTER_EINVAL, "T exceeds seq_len given at after_create");
  }

  // ------------------------------------------------------------------ DenoiserV2.forward
  void forward(const float* x, const float* time, const float* cond, const float* time_cond, float* out, int N, int T,
               cudaStream_t st, int cache_index = -1) {
    check_shape(N, T);
    if (cache_index >= 0) check_cache_index(cache_index);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_buf, cond, (size_t)N * zt * 4, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(tc_buf, time_cond, (size_t)N * zs * T * 4, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(time_buf, time, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
    std::vector<int> src(N), trow(N), tstr(N, 1), crow(N);
    for (int n = 0; n < N; ++n) { src[n] = n; trow[n] = n * T; crow[n] = n; }
    upload_maps(src, trow, tstr, crow, st);
    build_tables(N, T, N, 1, N, N, st);
    run_network(x, N, N, T, adaC, st, cache_index);
    dim3 grid(ceil_div(T, 32), ceil_div(C, 32), N);
    tokens_to_channels_kernel<<<grid, 256, 0, st>>>(proj, out, C, T);
    AFTER_CUDA_CHECK(cudaGetLastError()); AFTER_COUNT_LAUNCH();
  }

  // CFG row layout (model.py:730-743 / export_midi.py:322-345)
  void cfg_maps(int B, int T, int variant, cudaStream_t st) {
    const int N = 3 * B;
    std::vector<int> src(N), trow(N), tstr(N), crow(N);
    for (int n = 0; n < N; ++n) {
      const int b = n % B, grp = n / B;
      src[n] = b;
      const bool tc_on = variant == AFTER_CFG_AUDIO ? grp <= 1 : grp == 0;
      const bool c_on = variant == AFTER_CFG_AUDIO ? grp == 0 : grp <= 1;
      trow[n] = tc_on ? b * T : B * T;
      tstr[n] = tc_on ? 1 : 0;
      crow[n] = c_on ? b : B;
    }
    upload_maps(src, trow, tstr, crow, st);
  }

  void set_guidance(float g_timbre, float g_structure, int variant, float clamp, float dt, cudaStream_t st) {
    const float total = 0.5f * (g_structure + g_timbre);
    const float first = variant == AFTER_CFG_AUDIO ? g_timbre : g_structure;
    const float second = variant == AFTER_CFG_AUDIO ? g_structure : g_timbre;
    const float f = first / std::max(second, clamp);
    float hbuf[4] = {total, f, dt, 0.f};
    AFTER_CUDA_CHECK(cudaMemcpyAsync(guidance, hbuf, sizeof(hbuf), cudaMemcpyHostToDevice, st));
    AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
  }

  void combine(int B, int T, const float* xin, float* xout, int euler, cudaStream_t st) {
    dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B);
    PdlScope pdl(true);
    launch_k(cfg_combine_kernel, grid, dim3(256), 0, st, proj, guidance, xin, xout, B, C, T, euler);
    AFTER_COUNT_LAUNCH();
  }

  // ------------------------------------------------------------------ RectifiedFlow.model_forward
  void model_forward(const float* x, const float* time, const float* cond, const float* time_cond, float* out, int B,
                     int T, float g_t, float g_s, int variant, float clamp, cudaStream_t st, int cache_index = -1) {
    check_shape(3 * B, T);
    if (cache_index >= 0) check_cache_index(cache_index);
    AFTER_REQUIRE(variant == AFTER_CFG_AUDIO || variant == AFTER_CFG_MIDI, AFTER_EINVAL, "unknown cfg_variant");
    AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_buf, cond, (size_t)B * zt * 4, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(tc_buf, time_cond, (size_t)B * zs * T * 4, cudaMemcpyDeviceToDevice, st));
    // the reference repeats one time per stream over the 3 CFG rows; the (time, class) table is per stream:
    // rows r = b*(1+1)... use per-row times with n_cls = 2 classes per stream is not expressible with a shared
    // class table, so evaluate the table per (stream, {cond_b, drop}) explicitly: row 2b = (t_b, cond_b), 2b+1 = (t_b, drop).
    model_forward_tables(time, B, T, st);
    cfg_maps_per_stream(B, T, variant, st);
    set_guidance(g_t, g_s, variant, clamp, 1.0f, st);
    run_network(x, B, 3 * B, T, adaC, st, cache_index);
    combine(B, T, x, out, 0, st);
  }

  // (time_b, cond_b) / (time_b, drop) rows for model_forward with arbitrary per-stream times.
  void model_forward_tables(const float* time, int B, int T, cudaStream_t st) {
    // times_per_row table: row r -> time[r / 2]; build a 2B-long time vector on device via two strided copies
    AFTER_CUDA_CHECK(cudaMemcpy2DAsync(time_buf, 2 * sizeof(float), time, sizeof(float), sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpy2DAsync(time_buf + 1, 2 * sizeof(float), time, sizeof(float), sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    // condition classes: row 2b -> cond_b, row 2b+1 -> drop.  fourier_concat_kernel indexes cond by (r % n_cls);
    // with n_cls = 2B and a cond table whose odd entries are "drop" we get exactly that layout.
    build_tables_mf(B, T, st);
  }
  void build_tables_mf(int B, int T, cudaStream_t st);
  void cfg_maps_per_stream(int B, int T, int variant, cudaStream_t st) {
    const int N = 3 * B;
    std::vector<int> src(N), trow(N), tstr(N), crow(N);
    for (int n = 0; n < N; ++n) {
      const int b = n % B, grp = n / B;
      src[n] = b;
      const bool tc_on = variant == AFTER_CFG_AUDIO ? grp <= 1 : grp == 0;
      const bool c_on = variant == AFTER_CFG_AUDIO ? grp == 0 : grp <= 1;
      trow[n] = tc_on ? b * T : B * T;
      tstr[n] = tc_on ? 1 : 0;
      crow[n] = c_on ? 2 * b : 2 * b + 1;
    }
    upload_maps(src, trow, tstr, crow, st);
  }

  // ------------------------------------------------------------------ RectifiedFlow.sample
  // torch.linspace(0, 1, steps+1)[:-1] in fp32, the way ATen fills it (symmetric halves).
  static std::vector<float> time_grid(int nb_steps) {
    const int steps = nb_steps + 1;
    std::vector<float> t(nb_steps);
    const float step = (1.0f - 0.0f) / (float)(steps - 1);
    const int halfway = steps / 2;
    for (int i = 0; i < nb_steps; ++i) t[i] = i < halfway ? 0.0f + step * (float)i : 1.0f - step * (float)(steps - i - 1);
    return t;
  }

  // stream = true is one audio block of the exported Streamer.sample (after_scripts/export.py:398-416): Euler step s runs
  // against KV cache s, which is then rolled by the block length.
  // A live streaming block (rows = 3 B T <= 16) runs as ONE persistent kernel: all Euler steps, all layers, CFG combine and
  // roll_cache, separated by grid barriers instead of 28 launches per step (stream_step.cuh).
  bool persistent_stream_ok(int B, int T) const {
    static int off = -1;
    if (off < 0) { const char* e = debug_env("AFTER_STREAM_PERSISTENT"); off = (e && e[0] == '0') ? 1 : 0; }
    return !off && !g_prof.on && cacheW > 0 && sk_a != nullptr && 3 * B * T <= SKINNY_ROWS && L <= 8 && n_sms > 0 && D % 128 == 0 &&
           HID <= 1536;  // stream_block_kernel's K-quartered down projection holds 3 x 128 weights of a quarter per lane
  }
  void stream_block_persistent(int B, int T, int nb_steps, cudaStream_t st) {
    StreamNetDev net{};
    for (int l = 0; l < L; ++l) {
      const DenoiserLayer& ly = layers[l];
      net.layer[l] = StreamLayerDev{ly.qkv.w, ly.mlp0.w, ly.mlp0.bias, ly.mlp2.w, ly.mlp2.bias, ly.n1_g, ly.n1_b, ly.n3_g, ly.n3_b};
    }
    net.L = L; net.D = D; net.HID = HID; net.C = C; net.chunk = cfg.attention_chunk_size; net.window = cfg.local_attention_size;
    net.W = cacheW; net.maxN = maxN; net.maxRows = maxRows; net.ada_ld = L * 2 * D;
    net.pe_wt = pe_wt; net.pe_b = pe_b; net.out_w = out_proj.w; net.out_b = out_proj.bias;
    net.rope_tab = rope_tab; net.adaT = adaT; net.adaC = adaC; net.map = seqmap(); net.guidance = guidance;
    net.x_state = x_state; net.h0 = h0; net.hA = h_part; net.hB = h; net.qkv_stream = qkv_stream; net.sk_a = sk_a; net.sk_hid = sk_hid;
    net.proj = proj; net.kcache = kcache; net.vcache = vcache;
    net.cache_slab = cache_slab(); net.adaC_step_stride = (size_t)(B + 1) * L * 2 * D; net.barrier = stream_barrier;
    AFTER_CUDA_CHECK(cudaMemsetAsync(stream_barrier, 0, sizeof(unsigned), st));
    net.dbg = nullptr;
#ifdef AFTER_DEBUG
    static unsigned long long* dbg = nullptr;   // AFTER_DEBUG_TRACE_STREAM=1: per-barrier timeline of CTA 0 (run with AFTER_NO_GRAPH=1)
    static int trace = -1;
    if (trace < 0) { const char* e = debug_env("AFTER_DEBUG_TRACE_STREAM"); trace = e ? atoi(e) : 0; }
    if (trace > 0) {
      if (!dbg) AFTER_CUDA_CHECK(cudaMalloc(&dbg, 256 * sizeof(unsigned long long)));
      AFTER_CUDA_CHECK(cudaMemsetAsync(dbg, 0, 256 * sizeof(unsigned long long), st));
      net.dbg = dbg;
    }
#endif
    const size_t smem = ((size_t)SKINNY_ROWS * std::max(D, HID) + (size_t)L * 4 * D) * sizeof(float);
    const int mk = cfg.attention_chunk_size + cfg.local_attention_size - 1;
    PdlScope pdl(false);  // plain launch: every CTA must become resident for the grid barrier, nothing to overlap with
#define AFTER_LAUNCH_SS(NH, MK, NT) launch_k(stream_block_kernel<NH, MK, NT>, dim3(n_sms), dim3(NT), smem, st, net, B, T, nb_steps)
    if (H == 8) { if (mk <= 12) AFTER_LAUNCH_SS(8, 12, 512); else if (mk <= 20) AFTER_LAUNCH_SS(8, 20, 256); else AFTER_LAUNCH_SS(8, 32, 256); }
    else        { if (mk <= 12) AFTER_LAUNCH_SS(4, 12, 512); else if (mk <= 20) AFTER_LAUNCH_SS(4, 20, 256); else AFTER_LAUNCH_SS(4, 32, 256); }
#undef AFTER_LAUNCH_SS
    AFTER_COUNT_LAUNCH();
    last_N = 3 * B; last_T = T;
#ifdef AFTER_DEBUG
    if (net.dbg && --trace == 0) {
      AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
      std::vector<unsigned long long> hb(256);
      AFTER_CUDA_CHECK(cudaMemcpy(hb.data(), net.dbg, 256 * 8, cudaMemcpyDeviceToHost));
      fprintf(stderr, "stream block trace of CTA 0 (ns): barrier k: [work before it] [wait in it]\n");
      for (int k = 0; k < 60 && hb[2 * k]; ++k)
        fprintf(stderr, "  %2d: work %6lld wait %6lld\n", k, k ? (long long)(hb[2 * k] - hb[2 * k - 1]) : 0LL, (long long)(hb[2 * k + 1] - hb[2 * k]));
    }
#endif
  }

  void sample_body(int B, int T, int nb_steps, cudaStream_t st, bool stream = false) {
    build_tables(B, T, nb_steps * (B + 1), 0, B, B + 1, st);
    if (stream && persistent_stream_ok(B, T)) {
      stream_block_persistent(B, T, nb_steps, st);
      return;
    }
    const size_t step_stride = (size_t)(B + 1) * L * 2 * D;
    for (int s = 0; s < nb_steps; ++s) {
      run_network(x_state, B, 3 * B, T, adaC + (size_t)s * step_stride, st, stream ? s : -1);
      combine(B, T, x_state, x_state, 1, st);
      if (stream) roll_cache(T, s, st);
    }
  }

  void sample(const float* x0, const float* cond, const float* time_cond, float* out, int B, int T, int nb_steps,
              float g_t, float g_s, int variant, float clamp, cudaStream_t st, bool stream = false) {
    check_shape(3 * B, T);
    if (stream) check_cache_index(0);
    AFTER_REQUIRE(nb_steps >= 1 && nb_steps <= cfg.max_steps, AFTER_EINVAL, "nb_steps exceeds max_steps given at after_create");
    AFTER_REQUIRE(variant == AFTER_CFG_AUDIO || variant == AFTER_CFG_MIDI, AFTER_EINVAL, "unknown cfg_variant");
    const size_t xb = (size_t)B * C * T * 4;
    AFTER_CUDA_CHECK(cudaMemcpyAsync(x_state, x0, xb, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_buf, cond, (size_t)B * zt * 4, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(tc_buf, time_cond, (size_t)B * zs * T * 4, cudaMemcpyDeviceToDevice, st));
    std::vector<float> tg = time_grid(nb_steps);
    AFTER_CUDA_CHECK(cudaMemcpyAsync(time_buf, tg.data(), tg.size() * 4, cudaMemcpyHostToDevice, st));
    cfg_maps(B, T, variant, st);  // synchronises: tg stays alive until here
    set_guidance(g_t, g_s, variant, clamp, 1.0f / (float)nb_steps, st);

    if (!use_graph || g_prof.on) {
      sample_body(B, T, nb_steps, st, stream);
    } else {
      auto key = std::make_tuple(B, T, nb_steps, variant, stream ? 1 : 0);
      auto it = graphs.find(key);
      if (it == graphs.end()) {
        GraphEntry ge;
        cudaGraph_t graph = nullptr;
        const int64_t before = g_launches.load();
        AFTER_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
          sample_body(B, T, nb_steps, st, stream);
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
      if (stream) { last_N = 3 * B; last_T = T; }
    }
    AFTER_CUDA_CHECK(cudaMemcpyAsync(out, x_state, xb, cudaMemcpyDeviceToDevice, st));
  }
};

// model_forward tables: rows (2b) = (t_b, cond_b), (2b+1) = (t_b, drop)
inline void Denoiser::build_tables_mf(int B, int T, cudaStream_t st) {
  // interleave cond with drop rows into x_in scratch (reused as a [2B, zt] table), then run the generic builder
  // with per-row times and n_cond = n_cls = 2B.
  float* ctab = x_in;
  std::vector<float> dropv((size_t)zt, cfg.drop_value);
  for (int b = 0; b < B; ++b) {
    AFTER_CUDA_CHECK(cudaMemcpyAsync(ctab + (size_t)(2 * b) * zt, cond_buf + (size_t)b * zt, zt * 4, cudaMemcpyDeviceToDevice, st));
    AFTER_CUDA_CHECK(cudaMemcpyAsync(ctab + (size_t)(2 * b + 1) * zt, dropv.data(), zt * 4, cudaMemcpyHostToDevice, st));
  }
  AFTER_CUDA_CHECK(cudaStreamSynchronize(st));
  AFTER_CUDA_CHECK(cudaMemcpyAsync(cond_buf, ctab, (size_t)2 * B * zt * 4, cudaMemcpyDeviceToDevice, st));
  build_tables(B, T, 2 * B, 1, 2 * B, 2 * B, st);
}

}  // namespace after
