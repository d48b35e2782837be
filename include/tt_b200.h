/*
 * tt_b200.h — C ABI of the B200-native two-tower hot path (libtt_b200.so).
 *
 * The reference (DiegoPaniagua23/music-recommendation-multimodal) has no FFI of its
 * own: its hot path is PyTorch library calls inside three nn.Modules and two loops.
 * Each entry point below names the reference call site (file:line under the
 * reference root) whose device work it replaces.
 *
 * Conventions (all entry points):
 *   - return 0 on success, non-zero on error; tt_last_error() returns the message
 *     for the calling thread;
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *     _host; sizes are explicit; matrices are row-major with an explicit leading
 *     dimension in ELEMENTS;
 *   - no allocation, no host synchronisation, no global mutable state: launches go
 *     to the caller's stream (a cudaStream_t passed as void*) and are CUDA-graph
 *     capturable;
 *   - "bf16" buffers are raw uint16 bfloat16 bit patterns.
 */
#ifndef TT_B200_H_
#define TT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_B200_VERSION 1

/* ---- library ------------------------------------------------------------------ */
const char* tt_last_error(void);
int tt_version(void);
int tt_num_sms(void);

/* ---- dense contraction on tcgen05 ------------------------------------------------
 * C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 *
 * Replaces every cuBLAS GEMM the reference reaches through nn.Linear /
 * nn.MultiheadAttention in_proj/out_proj (src/models/user_tower.py:37-57,
 * src/models/item_tower.py:121-129), torch.matmul for the InfoNCE logits
 * (src/models/two_tower.py:106) and their autograd dgrad/wgrad counterparts
 * (src/train.py:63).
 *
 * Operand layouts (bf16):
 *   a_mn = 0 : A is [M, lda]  (K contiguous;  "K-major")
 *   a_mn = 1 : A is [K, lda]  (M contiguous; "MN-major", i.e. A^T stored) — M % 64 == 0
 *   b_mn = 0 : B is [N, ldb]  (K contiguous; nn.Linear weight layout)
 *   b_mn = 1 : B is [K, ldb]  (N contiguous) — N % 64 == 0
 * Epilogue, applied in this order to v = alpha * acc:
 *   v += bias[n]; relu; dropout(seed, site, m*N+n); gate (v = gate[m,n] > 0 ? v*gate_scale : 0);
 *   v += residual[m,n]; store fp32 (optionally atomically accumulated) and/or bf16.
 */
typedef struct tt_gemm_args {
  const void* A;
  const void* B;
  int32_t lda, ldb;
  int32_t a_mn, b_mn;
  int32_t M, N, K;
  float alpha;
  const float* bias;       /* [N] or NULL */
  int32_t relu;
  float drop_p;            /* 0 => no dropout */
  uint64_t drop_seed;
  uint32_t drop_site;
  const void* gate;        /* bf16 [M, ld_gate] or NULL */
  int32_t ld_gate;
  float gate_scale;
  const float* residual;   /* fp32 [M, ld_res] or NULL */
  int32_t ld_res;
  float* out_f32;          /* fp32 [M, ld_f32] or NULL */
  int32_t ld_f32;
  void* out_bf16;          /* bf16 [M, ld_bf16] or NULL */
  int32_t ld_bf16;
  int32_t accumulate;      /* 1: out_f32 += (red.add), required when k_splits > 1 */
  int32_t k_splits;        /* 0 => choose automatically (only >1 when accumulate) */
  int32_t block_n;         /* 0 => choose automatically (64/128/256) */
} tt_gemm_args;

int tt_gemm_bf16(const tt_gemm_args* args, void* stream);

/* ---- causal self-attention, d_head = 64 -------------------------------------------
 * Replaces F.scaled_dot_product_attention inside nn.TransformerEncoderLayer
 * (src/models/user_tower.py:37-45, 111-116) for right-padded histories.
 *   qkv  : bf16 [B*L, 3*H*64], columns [Q | K | V], head h at h*64 inside each third
 *          (the packed in_proj output, rows [Q;K;V] of in_proj_weight)
 *   ctx  : bf16 [B*L, H*64]     attention output before out_proj
 *   lse  : fp32 [B, H, L]       natural-log sum-exp of the scaled scores (for backward)
 * Dropout (p > 0) is applied to the attention probabilities with the counter-based hash
 * (seed, site, ((b*H+h)*L+i)*L+j). L <= 512 forward, L <= 256 backward.
 */
int tt_attn_causal_fwd(const void* qkv, void* ctx, float* lse, int B, int L, int H, float drop_p,
                       uint64_t drop_seed, uint32_t drop_site, void* stream);
int tt_attn_causal_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse,
                       void* dqkv, int B, int L, int H, float drop_p, uint64_t drop_seed,
                       uint32_t drop_site, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* TT_B200_H_ */
