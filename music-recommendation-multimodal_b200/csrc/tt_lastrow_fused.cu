// Last encoder layer, single-row form taken one step further: K and V are never materialised.
//
// Only out[b, len_b - 1] of the last layer is read (src/models/user_tower.py:122-132), so its attention has ONE
// query per (sequence, head) (tt_lastrow.cu). With k_j = Wk h1_j + bk and v_j = Wv h1_j + bv (h1 = norm1 output,
// in_proj of nn.MultiheadAttention, src/models/user_tower.py:37-45) the scores and the context are
//     s_j   = scale * q_h . k_j  = (scale * Wk_h^T q_h) . h1_j + const      (the constant drops out of the softmax)
//     ctx_h = sum_j pd_j v_j     = Wv_h (sum_j pd_j h1_j) + bv_h * sum_j pd_j
// i.e. a 256-vector a_h = scale * Wk_h^T q_h per head, L dot products against the h1 rows, a weighted sum r_h of
// those rows, and one 64 x 256 product. The [T, 512] K|V GEMM of the layer, its dgrad and wgrad GEMMs over all
// T tokens and the bias column sums disappear (three encoder-sized launches of the c2 step); the layer's
// gradient w.r.t. h1 becomes dh1_j = sum_h (ds_jh a_h + pd_jh dr_h) with dr_h = Wv_h^T dctx_h, and the K / V
// weight gradients are sums over the BATCH of outer products (q_h x da_h, dctx_h x r_h): B-row GEMMs.
// Exact algebra; K and V are no longer rounded to bf16 on the way.
//
// One block (256 threads) per sequence; weights (Wq, Wk, Wv: 3 x 128 KB bf16) come from L2, the sequence's
// h1 rows (L x 512 B) are read twice (second time from L2).
//   tt_lastrow_attn_fwd : gather row len-1 -> q -> a -> scores -> softmax / dropout -> r -> ctx
//   tt_lastrow_attn_bwd : dctx -> dr -> (scores again) ds, pd -> da, dh1 (every position) -> dq -> + Wq^T dq
#include "../../include/tt_b200.h"
#include "tt_common.cuh"

namespace tt {

static constexpr int kLrD = 256;       // model width
static constexpr int kLrH = 4;         // heads
static constexpr int kLrDh = 64;
static constexpr int kLrMaxL = 512;
static constexpr int kLrThreads = 256;
static constexpr float kLrLog2e = 1.4426950408889634f;

struct LastRowParams {
  const __nv_bfloat16* h1;      // [B*L, 256] norm1 output of the layer
  const float* x_in;            // [B*L, 256] layer input (residual stream); forward only
  const int32_t* last_idx;      // [B]
  const __nv_bfloat16* Wqkv;    // [768, 256] in_proj_weight (bf16 shadow)
  const float* bqkv;            // [768]
  int B, L;
  float scale;
  uint32_t drop_thresh; float drop_scale; uint64_t seed; const uint64_t* seed_dev; uint32_t site;
  // forward outputs (all but ctx are kept for backward / the weight-gradient GEMMs)
  __nv_bfloat16* hq_bf; float* xq_in; __nv_bfloat16* q_bf; float* a_f32; float* r_f32;
  __nv_bfloat16* r_bf; float* lse; float* sumpd; __nv_bfloat16* ctx_bf;
  // backward
  const __nv_bfloat16* dctx;    // [B, 256]
  float* dh;                    // [B*L, 256] written for every position
  __nv_bfloat16* dq_bf;         // [B, 256]
  __nv_bfloat16* da_bf;         // [B, 4*256]  (scale folded in: dWk_h = q_h^T da_h)
  float* dbv;                   // [256] accumulated: d(bias of V)
};

// 8 consecutive bf16 of a row (16 B, lane-contiguous) -> fp32
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 f = unpack_bf16(w[h]);
    v[2 * h] = f.x;
    v[2 * h + 1] = f.y;
  }
}

// out[n] = sum_k W[n][k] x[k] for the block's 256 rows n = 0..255 of a [256, 256] bf16 row-major matrix:
// warp w takes rows 32 w .. 32 w + 31, a lane holds x[8 lane .. 8 lane + 7] (x_sel(row) picks the vector, warp-
// uniform), one 16-byte load per lane and row, five shuffles. Lane 0 returns the sums through `emit`.
template <typename XSel, typename Emit>
__device__ __forceinline__ void rowdot_256(const __nv_bfloat16* W, int warp, int lane, XSel x_sel, Emit emit) {
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    const int n = warp * 32 + i;
    const float* x = x_sel(n);
    float w[8];
    ld8(W + static_cast<size_t>(n) * kLrD + lane * 8, w);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += w[e] * x[lane * 8 + e];
    acc = warp_sum(acc);
    if (lane == 0) emit(n, acc);
  }
}

// out[c] = sum_{d < rows} W[d][c] y[d] for column c = threadIdx.x of a [rows, 256] bf16 row-major matrix
// (a warp reads 64 contiguous bytes per row)
__device__ __forceinline__ float coldot(const __nv_bfloat16* W, int rows, const float* y, int c) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
  for (int d = 0; d < rows; d += 2) {
    acc0 += __bfloat162float(W[static_cast<size_t>(d) * kLrD + c]) * y[d];
    acc1 += __bfloat162float(W[static_cast<size_t>(d + 1) * kLrD + c]) * y[d + 1];
  }
  return acc0 + acc1;
}

__global__ void __launch_bounds__(kLrThreads) lastrow_attn_fwd_kernel(const LastRowParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_x[kLrD];                 // h1 row of the query position
  __shared__ float s_q[kLrD];
  __shared__ float s_a[kLrH * kLrD];          // a_h, later reused for r_h
  __shared__ float s_r[kLrH * kLrD];
  __shared__ float s_s[kLrH * kLrMaxL];       // scores -> normalised, dropped probabilities
  __shared__ float s_stat[kLrH * 2];          // 1 / l, sum of pd per head
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int last = p.last_idx[b];
  const int len = last + 1;
  const size_t seq0 = static_cast<size_t>(b) * p.L;
  const __nv_bfloat16* Wq = p.Wqkv;
  const __nv_bfloat16* Wk = p.Wqkv + static_cast<size_t>(kLrD) * kLrD;
  const __nv_bfloat16* Wv = p.Wqkv + static_cast<size_t>(2 * kLrD) * kLrD;

  // ---- the query position's rows: residual stream (kept for out_proj's residual add) and h1 -------------
  {
    const size_t row = seq0 + last;
    const __nv_bfloat16 hv = p.h1[row * kLrD + tid];
    s_x[tid] = __bfloat162float(hv);
    p.hq_bf[static_cast<size_t>(b) * kLrD + tid] = hv;
    p.xq_in[static_cast<size_t>(b) * kLrD + tid] = p.x_in[row * kLrD + tid];
  }
  for (int i = tid; i < kLrH * kLrD; i += kLrThreads) s_r[i] = 0.f;
  __syncthreads();
  // ---- q = Wq h + bq ----------------------------------------------------------------------------------
  rowdot_256(Wq, warp, lane, [&](int) { return s_x; },
             [&](int n, float v) {
               v += p.bqkv[n];
               s_q[n] = v;
               p.q_bf[static_cast<size_t>(b) * kLrD + n] = __float2bfloat16(v);
             });
  __syncthreads();
  // ---- a_h = scale * Wk_h^T q_h -------------------------------------------------------------------------
#pragma unroll
  for (int h = 0; h < kLrH; ++h) {
    const float v = p.scale * coldot(Wk + static_cast<size_t>(h * kLrDh) * kLrD, kLrDh, s_q + h * kLrDh, tid);
    s_a[h * kLrD + tid] = v;
    p.a_f32[(static_cast<size_t>(b) * kLrH + h) * kLrD + tid] = v;
  }
  __syncthreads();
  // ---- scores s[h][j] = a_h . h1_j : a warp per position, a lane per 8 columns ----------------------------
  {
    float a[kLrH][8];
#pragma unroll
    for (int h = 0; h < kLrH; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) a[h][e] = s_a[h * kLrD + lane * 8 + e];
    for (int j = warp; j < len; j += kLrThreads / 32) {
      float x[8];
      ld8(p.h1 + (seq0 + j) * kLrD + lane * 8, x);
      float d[kLrH];
#pragma unroll
      for (int h = 0; h < kLrH; ++h) {
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += a[h][e] * x[e];
        d[h] = warp_sum(acc);
      }
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < kLrH; ++h) s_s[h * kLrMaxL + j] = d[h];
      }
    }
  }
  __syncthreads();
  // ---- softmax + dropout: warp h < 4 owns head h --------------------------------------------------------
  if (warp < kLrH) {
    const int h = warp;
    float* s = s_s + h * kLrMaxL;
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    const uint64_t seed = p.seed + ((p.drop_thresh && p.seed_dev) ? *p.seed_dev : 0ull);
    const uint32_t dkey = drop_key(seed, p.site);
    const uint64_t didx0 = (static_cast<uint64_t>(b * kLrH + h) * p.L + last) * p.L;
    float l = 0.f, spd = 0.f;
    for (int j = lane; j < len; j += 32) {
      float e = exp2f((s[j] - m) * kLrLog2e);
      l += e;
      if (p.drop_thresh) e = drop_keep_k(dkey, didx0 + j, p.drop_thresh) ? e * p.drop_scale : 0.f;
      spd += e;
      s[j] = e;
    }
    l = warp_sum(l);
    spd = warp_sum(spd);
    const float inv = 1.f / l;
    for (int j = lane; j < len; j += 32) s[j] *= inv;
    if (lane == 0) {
      s_stat[h * 2] = inv;
      s_stat[h * 2 + 1] = spd * inv;
      p.lse[b * kLrH + h] = m + logf(l);          // natural-log sum-exp of the scaled scores (up to the dropped constant)
      p.sumpd[b * kLrH + h] = spd * inv;
    }
  }
  __syncthreads();
  // ---- r_h = sum_j pd[h][j] h1_j ------------------------------------------------------------------------
  {
    float r[kLrH][8];
#pragma unroll
    for (int h = 0; h < kLrH; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) r[h][e] = 0.f;
    for (int j = warp; j < len; j += kLrThreads / 32) {
      float x[8];
      ld8(p.h1 + (seq0 + j) * kLrD + lane * 8, x);
#pragma unroll
      for (int h = 0; h < kLrH; ++h) {
        const float w = s_s[h * kLrMaxL + j];
#pragma unroll
        for (int e = 0; e < 8; ++e) r[h][e] += w * x[e];
      }
    }
#pragma unroll
    for (int h = 0; h < kLrH; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_r[h * kLrD + lane * 8 + e], r[h][e]);
  }
  __syncthreads();
  for (int i = tid; i < kLrH * kLrD; i += kLrThreads) {
    const float v = s_r[i];
    p.r_f32[static_cast<size_t>(b) * kLrH * kLrD + i] = v;
    p.r_bf[static_cast<size_t>(b) * kLrH * kLrD + i] = __float2bfloat16(v);
  }
  // ---- ctx[n] = Wv[n] . r_head(n) + bv[n] * sum_j pd_j ----------------------------------------------------
  rowdot_256(Wv, warp, lane, [&](int n) { return s_r + (n / kLrDh) * kLrD; },
             [&](int n, float v) {
               v += p.bqkv[2 * kLrD + n] * s_stat[(n / kLrDh) * 2 + 1];
               p.ctx_bf[static_cast<size_t>(b) * kLrD + n] = __float2bfloat16(v);
             });
}

__global__ void __launch_bounds__(kLrThreads) lastrow_attn_bwd_kernel(const LastRowParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_g[kLrD];                 // dctx
  __shared__ float s_a[kLrH * kLrD];          // a_h
  __shared__ float s_dr[kLrH * kLrD];         // dr_h = Wv_h^T dctx_h
  __shared__ float s_da[kLrH * kLrD];         // da_h (accumulated)
  __shared__ float s_ds[kLrH * kLrMaxL];      // scores -> ds
  __shared__ float s_pd[kLrH * kLrMaxL];      // dpd -> pd
  __shared__ float s_dq[kLrD];
  __shared__ float s_beta[kLrH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int last = p.last_idx[b];
  const int len = last + 1;
  const size_t seq0 = static_cast<size_t>(b) * p.L;
  const __nv_bfloat16* Wq = p.Wqkv;
  const __nv_bfloat16* Wk = p.Wqkv + static_cast<size_t>(kLrD) * kLrD;
  const __nv_bfloat16* Wv = p.Wqkv + static_cast<size_t>(2 * kLrD) * kLrD;

  s_g[tid] = __bfloat162float(p.dctx[static_cast<size_t>(b) * kLrD + tid]);
  for (int i = tid; i < kLrH * kLrD; i += kLrThreads) {
    s_a[i] = p.a_f32[static_cast<size_t>(b) * kLrH * kLrD + i];
    s_da[i] = 0.f;
  }
  __syncthreads();
  // ---- beta_h = dctx_h . bv_h (the bias part of d pd_j), d(bv) += dctx * sum_j pd_j ----------------------
  if (warp < kLrH) {
    const int h = warp;
    float acc = s_g[h * kLrDh + lane] * p.bqkv[2 * kLrD + h * kLrDh + lane] +
                s_g[h * kLrDh + 32 + lane] * p.bqkv[2 * kLrD + h * kLrDh + 32 + lane];
    acc = warp_sum(acc);
    if (lane == 0) s_beta[h] = acc;
  }
  if (p.dbv) atomicAdd(p.dbv + tid, s_g[tid] * p.sumpd[b * kLrH + tid / kLrDh]);
  // ---- dr_h = Wv_h^T dctx_h -------------------------------------------------------------------------------
#pragma unroll
  for (int h = 0; h < kLrH; ++h)
    s_dr[h * kLrD + tid] = coldot(Wv + static_cast<size_t>(h * kLrDh) * kLrD, kLrDh, s_g + h * kLrDh, tid);
  __syncthreads();
  // ---- per position: score (again) and d pd_j = dr_h . h1_j + beta_h ------------------------------------------
  float a[kLrH][8], dr[kLrH][8];
#pragma unroll
  for (int h = 0; h < kLrH; ++h)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a[h][e] = s_a[h * kLrD + lane * 8 + e];
      dr[h][e] = s_dr[h * kLrD + lane * 8 + e];
    }
  for (int j = warp; j < len; j += kLrThreads / 32) {
    float x[8];
    ld8(p.h1 + (seq0 + j) * kLrD + lane * 8, x);
    float d0[kLrH], d1[kLrH];
#pragma unroll
    for (int h = 0; h < kLrH; ++h) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc0 += a[h][e] * x[e];
        acc1 += dr[h][e] * x[e];
      }
      d0[h] = warp_sum(acc0);
      d1[h] = warp_sum(acc1);
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < kLrH; ++h) {
        s_ds[h * kLrMaxL + j] = d0[h];
        s_pd[h * kLrMaxL + j] = d1[h] + s_beta[h];
      }
    }
  }
  __syncthreads();
  // ---- per head: p_j, pd_j, ds_j = p_j (dp_j - delta) -----------------------------------------------------------
  if (warp < kLrH) {
    const int h = warp;
    float* sc = s_ds + h * kLrMaxL;
    float* dp = s_pd + h * kLrMaxL;
    const float lse = p.lse[b * kLrH + h];
    const uint64_t seed = p.seed + ((p.drop_thresh && p.seed_dev) ? *p.seed_dev : 0ull);
    const uint32_t dkey = drop_key(seed, p.site);
    const uint64_t didx0 = (static_cast<uint64_t>(b * kLrH + h) * p.L + last) * p.L;
    float delta = 0.f;
    for (int j = lane; j < len; j += 32) {
      const float pr = exp2f((sc[j] - lse) * kLrLog2e);
      const bool keep = !p.drop_thresh || drop_keep_k(dkey, didx0 + j, p.drop_thresh);
      const float dpj = keep ? dp[j] * p.drop_scale : 0.f;
      delta += pr * dpj;
      sc[j] = pr;          // parked: ds needs delta first
      dp[j] = dpj;
    }
    delta = warp_sum(delta);
    for (int j = lane; j < len; j += 32) {
      const float pr = sc[j];
      const float dpj = dp[j];
      sc[j] = pr * (dpj - delta);                                                    // ds_j
      const bool keep = !p.drop_thresh || drop_keep_k(dkey, didx0 + j, p.drop_thresh);
      dp[j] = keep ? pr * p.drop_scale : 0.f;                                        // pd_j
    }
  }
  __syncthreads();
  // ---- da_h = sum_j ds_j h1_j ; dh1_j = sum_h ds_jh a_h + pd_jh dr_h (zero beyond the sequence) ----------------
  {
    float da[kLrH][8];
#pragma unroll
    for (int h = 0; h < kLrH; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) da[h][e] = 0.f;
    for (int j = warp; j < p.L; j += kLrThreads / 32) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
      if (j < len) {
        float x[8];
        ld8(p.h1 + (seq0 + j) * kLrD + lane * 8, x);
#pragma unroll
        for (int h = 0; h < kLrH; ++h) {
          const float ds = s_ds[h * kLrMaxL + j], pd = s_pd[h * kLrMaxL + j];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            da[h][e] += ds * x[e];
            o[e] += ds * a[h][e] + pd * dr[h][e];
          }
        }
      }
      float4* dst = reinterpret_cast<float4*>(p.dh + (seq0 + j) * kLrD + lane * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
#pragma unroll
    for (int h = 0; h < kLrH; ++h)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_da[h * kLrD + lane * 8 + e], da[h][e]);
  }
  __syncthreads();
  for (int i = tid; i < kLrH * kLrD; i += kLrThreads)
    p.da_bf[static_cast<size_t>(b) * kLrH * kLrD + i] = __float2bfloat16(s_da[i] * p.scale);
  // ---- dq[n] = scale * Wk[n] . da_head(n) --------------------------------------------------------------------
  rowdot_256(Wk, warp, lane, [&](int n) { return s_da + (n / kLrDh) * kLrD; },
             [&](int n, float v) {
               v *= p.scale;
               s_dq[n] = v;
               p.dq_bf[static_cast<size_t>(b) * kLrD + n] = __float2bfloat16(v);
             });
  __syncthreads();     // (also: this block's dh rows are written)
  // ---- the query position's h1 row also fed q: dh1_last += Wq^T dq ---------------------------------------------
  {
    const float v = coldot(Wq, kLrD, s_dq, tid);
    p.dh[(seq0 + last) * kLrD + tid] += v;
  }
}

static uint32_t lr_drop_threshold(float p) {
  if (p <= 0.f) return 0;
  double t = static_cast<double>(p) * 4294967296.0;
  uint32_t v = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
  return v == 0 ? 1 : v;
}

static int fill_lastrow(LastRowParams& p, const void* h1, const int32_t* last_idx, const void* Wqkv, const float* bqkv,
                        int B, int L, int H, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                        const char* who) {
  TT_REQUIRE(h1 && last_idx && Wqkv && bqkv && B > 0 && L > 0, "%s: bad arguments", who);
  TT_REQUIRE(H == kLrH, "%s: %d heads unsupported (4 heads of 64)", who, H);
  TT_REQUIRE(L <= kLrMaxL, "%s: L=%d > %d unsupported", who, L, kLrMaxL);
  TT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "%s: drop_p out of range", who);
  p = LastRowParams{};
  p.h1 = static_cast<const __nv_bfloat16*>(h1);
  p.last_idx = last_idx;
  p.Wqkv = static_cast<const __nv_bfloat16*>(Wqkv);
  p.bqkv = bqkv;
  p.B = B; p.L = L;
  p.scale = 0.125f;
  p.drop_thresh = lr_drop_threshold(drop_p);
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  p.seed = seed; p.seed_dev = seed_dev; p.site = site;
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_lastrow_attn_fwd(const void* h1, const float* x_in, const int32_t* last_idx, const void* Wqkv,
                                   const float* bqkv, int B, int L, int H, float drop_p, uint64_t seed,
                                   const uint64_t* seed_dev, uint32_t site, void* hq_bf16, float* xq_in,
                                   void* q_bf16, float* a_f32, float* r_f32, void* r_bf16, float* lse, float* sumpd,
                                   void* ctx_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LastRowParams p;
  int rc = fill_lastrow(p, h1, last_idx, Wqkv, bqkv, B, L, H, drop_p, seed, seed_dev, site, "tt_lastrow_attn_fwd");
  if (rc) return rc;
  TT_REQUIRE(x_in && hq_bf16 && xq_in && q_bf16 && a_f32 && r_f32 && r_bf16 && lse && sumpd && ctx_bf16,
             "tt_lastrow_attn_fwd: null output");
  p.x_in = x_in;
  p.hq_bf = static_cast<__nv_bfloat16*>(hq_bf16); p.xq_in = xq_in;
  p.q_bf = static_cast<__nv_bfloat16*>(q_bf16); p.a_f32 = a_f32; p.r_f32 = r_f32;
  p.r_bf = static_cast<__nv_bfloat16*>(r_bf16); p.lse = lse; p.sumpd = sumpd;
  p.ctx_bf = static_cast<__nv_bfloat16*>(ctx_bf16);
  TT_CHECK_CUDA(launch_k(lastrow_attn_fwd_kernel, dim3(B), dim3(kLrThreads), 0, stream, p));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_lastrow_attn_bwd(const void* h1, const int32_t* last_idx, const void* Wqkv, const float* bqkv,
                                   const void* dctx_bf16, const float* a_f32, const float* lse, const float* sumpd,
                                   int B, int L, int H, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                                   uint32_t site, float* dh, void* dq_bf16, void* da_bf16, float* dbv, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LastRowParams p;
  int rc = fill_lastrow(p, h1, last_idx, Wqkv, bqkv, B, L, H, drop_p, seed, seed_dev, site, "tt_lastrow_attn_bwd");
  if (rc) return rc;
  TT_REQUIRE(dctx_bf16 && a_f32 && lse && sumpd && dh && dq_bf16 && da_bf16, "tt_lastrow_attn_bwd: null pointer");
  p.dctx = static_cast<const __nv_bfloat16*>(dctx_bf16);
  p.a_f32 = const_cast<float*>(a_f32); p.lse = const_cast<float*>(lse); p.sumpd = const_cast<float*>(sumpd);
  p.dh = dh; p.dq_bf = static_cast<__nv_bfloat16*>(dq_bf16); p.da_bf = static_cast<__nv_bfloat16*>(da_bf16);
  p.dbv = dbv;
  TT_CHECK_CUDA(launch_k(lastrow_attn_bwd_kernel, dim3(B), dim3(kLrThreads), 0, stream, p));
  TT_LAUNCH_CHECK();
  return TT_OK;
}
