/* Minimal C host of libafter_b200.so: what a non-Python caller (e.g. an nn_tilde-style C++ backend) does.
 *
 *   gcc -I include examples/c_host.c -o c_host -L after_b200/lib -lafter_b200 -Wl,-rpath,$PWD/after_b200/lib
 *
 * It creates a handle for the tiny denoiser configuration, feeds a state dict of zeros (real callers pass the tensors of a
 * checkpoint, key by key, exactly as torch's state_dict() names them -- SURVEY.md appendix A.4), finalizes and runs one
 * sample() through host buffers.  Without a CUDA device after_create fails loudly and the program reports it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "after_b200.h"

static int load(after_handle h, const char* key, int ndim, int64_t d0, int64_t d1) {
  int64_t shape[2] = {d0, d1};
  size_t n = (size_t)d0 * (ndim > 1 ? (size_t)d1 : 1);
  float* zeros = (float*)calloc(n, sizeof(float));
  int rc = after_load_tensor(h, AFTER_MODULE_DENOISER, key, zeros, shape, ndim, AFTER_DTYPE_F32);
  free(zeros);
  return rc < 0 ? rc : 0;
}

int main(void) {
  after_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = AFTER_B200_ABI_VERSION;
  cfg.n_channels = 64; cfg.seq_len = 32; cfg.embed_dim = 256; cfg.cond_dim = 6; cfg.noise_embed_dims = 64;
  cfg.n_layers = 1; cfg.mlp_multiplier = 3; cfg.tcond_dim = 12; cfg.local_attention_size = 8; cfg.attention_chunk_size = 4;
  cfg.drop_value = -4.0f; cfg.max_batch = 1; cfg.max_steps = 2;

  printf("%s (abi %d), %d CUDA device(s)\n", after_build_info(), after_abi_version(), after_device_count());
  after_handle h = NULL;
  int rc = after_create(&cfg, 0, &h);
  if (rc != AFTER_OK) {
    printf("after_create failed (%d): %s\n", rc, after_last_error(NULL));
    return 2;
  }
  const int D = 256;
  char key[128];
  rc |= load(h, "embedding.0.weight", 2, D, 70);  rc |= load(h, "embedding.0.bias", 1, D, 0);
  rc |= load(h, "embedding.2.weight", 2, D, D);   rc |= load(h, "embedding.2.bias", 1, D, 0);
  const char* tb = "denoiser_trans_block.";
  snprintf(key, sizeof key, "%spatchify_and_embed.1.weight", tb);       rc |= load(h, key, 2, D, 64);
  snprintf(key, sizeof key, "%spatchify_and_embed.1.bias", tb);         rc |= load(h, key, 1, D, 0);
  snprintf(key, sizeof key, "%spatchify_and_embed_tcond.1.weight", tb); rc |= load(h, key, 2, 12, 12);
  snprintf(key, sizeof key, "%spatchify_and_embed_tcond.1.bias", tb);   rc |= load(h, key, 1, 12, 0);
  const char* blk = "denoiser_trans_block.decoder_blocks.0.";
  struct { const char* name; int nd; int64_t a, b; } t[] = {
      {"self_attention.qkv_linear.weight", 2, 3 * D, D}, {"mlp.mlp.0.weight", 2, 3 * D, D}, {"mlp.mlp.0.bias", 1, 3 * D, 0},
      {"mlp.mlp.2.weight", 2, D, 3 * D}, {"mlp.mlp.2.bias", 1, D, 0}, {"norm1.weight", 1, D, 0}, {"norm1.bias", 1, D, 0},
      {"norm3.weight", 1, D, 0}, {"norm3.bias", 1, D, 0}, {"linear.weight", 2, 2 * D, D}, {"linear.bias", 1, 2 * D, 0},
      {"tcond_linear.weight", 2, 2 * D, 12}, {"tcond_linear.bias", 1, 2 * D, 0}};
  for (size_t i = 0; i < sizeof t / sizeof t[0]; ++i) {
    snprintf(key, sizeof key, "%s%s", blk, t[i].name);
    rc |= load(h, key, t[i].nd, t[i].a, t[i].b);
  }
  snprintf(key, sizeof key, "%sout_proj.0.weight", tb); rc |= load(h, key, 2, 64, D);
  snprintf(key, sizeof key, "%sout_proj.0.bias", tb);   rc |= load(h, key, 1, 64, 0);
  if (rc != 0 || (rc = after_finalize_weights(h, AFTER_PRECISION_FP32)) != AFTER_OK) {
    printf("loading weights failed (%d): %s\n", rc, after_last_error(h));
    after_destroy(h);
    return 3;
  }
  enum { T = 32 };
  static float x0[64 * T], cond[6], tcond[12 * T], out[64 * T];
  for (int i = 0; i < 64 * T; ++i) x0[i] = (float)(i % 7) - 3.0f;
  rc = after_sample_host(h, x0, cond, tcond, out, 1, T, 2, 2.0f, 1.0f, AFTER_CFG_AUDIO, 0.01f, NULL);
  if (rc != AFTER_OK) printf("after_sample_host failed (%d): %s\n", rc, after_last_error(h));
  else printf("sample ok: out[0] = %g (all-zero weights: the velocity is 0, so out == x0: %s)\n", out[0], out[0] == x0[0] ? "yes" : "no");
  after_destroy(h);
  return rc == AFTER_OK ? 0 : 4;
}
