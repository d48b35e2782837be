// Latency-optimised GEMM for the SMALL problems of the step (the "B-row" GEMMs: user head, item tower, the
// single-row last encoder layer, the InfoNCE logits and their dgrad / wgrad counterparts — M, N, K of a few
// hundred each, 33-270 MFLOP).
//
// The persistent tcgen05 kernel of tt_gemm.cu spends ~10 us on any of them whatever their size: 640-thread
// CTAs, a 200 KB shared-memory carve-out, TMEM allocation, tensor-map fetches, a TMA ring that never fills and a
// staged epilogue are fixed costs that only pay off on the 51,200-row launches. 27 of the step's 41 GEMM launches
// are of this kind and ~19 of them sit on the critical dependency chain. Here a 128-thread CTA owns a 32 x 32 output
// tile and takes the reduction 256 deep per iteration — each of its four warps multiplies the whole tile over its own
// 64-deep slice, so a K = 256 problem is ONE load round trip (32 KB in flight per CTA, next iteration prefetched in
// registers, any of the four operand layouts transposed on the way into shared memory) and one burst of warp-level
// mma.sync m16n8k16 (bf16 in, fp32 accumulate); the four partial tiles are summed through shared memory.
// Nothing to set up, 64-256 CTAs, a few microseconds. Tensor-core peak is irrelevant at this size;
// tcgen05 stays the path for everything large. Same epilogue contract as tt_gemm_bf16 (alpha, bias, ReLU,
// dropout with the same counter hash, gate, residual, fp32 / bf16 outputs, atomically accumulated split-K).
#include "../../include/tt_b200.h"
#include "tt_common.cuh"
#include <stdlib.h>

namespace tt {

static constexpr int kSBM = 32, kSBN = 32;
static constexpr int kSBK = 256;              // reduction depth per iteration: four 64-deep slices, one per warp
static constexpr int kSPitch = kSBK + 8;      // bf16 elements per smem row (528 B): fragment reads are conflict-free
static constexpr int kVec = kSBM * kSBK / 8 / 128;   // 16-byte vectors per thread and operand tile (8)

struct SmallGemmParams {
  const __nv_bfloat16* A; const __nv_bfloat16* B;
  int lda, ldb, M, N, K;
  int iters_per_split;
  float alpha;
  const float* bias;
  int relu;
  uint32_t drop_thresh; float drop_scale; uint64_t drop_seed; const uint64_t* drop_seed_dev; uint32_t drop_site;
  const __nv_bfloat16* gate; int ld_gate; float gate_scale;
  const float* residual; int ld_res;
  float* out_f32; int ld_f32;
  __nv_bfloat16* out_bf16; int ld_bf16;
  int accumulate;
};

// One operand tile (32 rows of the output dimension x 256 of the reduction) as eight 16-byte vectors per thread.
// MN = false: source is [rows, ld] with the reduction contiguous; vector v covers row v / 32, k (v % 32) * 8 .. + 8.
// MN = true : source is [K, ld] with the output dimension contiguous; vector v covers k v / 4, rows (v % 4) * 8 .. + 8.
// Out-of-range parts are zero (the launcher guarantees the dimensions are multiples of 8 where vectors run).
template <bool MN>
__device__ __forceinline__ void tile_load(const __nv_bfloat16* __restrict__ src, int ld, int row0, int rows, int k0,
                                          int K, int tid, uint4 (&r)[kVec]) {
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int v = tid + i * 128;
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (!MN) {
      const int rr = row0 + (v >> 5), kk = k0 + (v & 31) * 8;
      if (rr < rows && kk < K) x = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(rr) * ld + kk));
    } else {
      const int kk = k0 + (v >> 2), rr = row0 + (v & 3) * 8;
      if (rr < rows && kk < K) x = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(kk) * ld + rr));
    }
    r[i] = x;
  }
}
template <bool MN>
__device__ __forceinline__ void tile_store(__nv_bfloat16* __restrict__ s, int tid, const uint4 (&r)[kVec]) {
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int v = tid + i * 128;
    if constexpr (!MN) {
      *reinterpret_cast<uint4*>(s + (v >> 5) * kSPitch + (v & 31) * 8) = r[i];
    } else {
      const int kk = v >> 2, rr = (v & 3) * 8;
      const uint32_t w[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        reinterpret_cast<uint16_t*>(s)[(rr + 2 * j) * kSPitch + kk] = static_cast<uint16_t>(w[j] & 0xFFFFu);
        reinterpret_cast<uint16_t*>(s)[(rr + 2 * j + 1) * kSPitch + kk] = static_cast<uint16_t>(w[j] >> 16);
      }
    }
  }
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 128 threads = 4 warps. Every warp multiplies the WHOLE 32 x 32 output tile over its own 64-deep slice of the
// 256-deep iteration (so a K = 256 problem is one load round trip + one multiply for the CTA, 32 KB in flight), the
// four partial tiles are summed through shared memory at the end and thread `tid` finishes 8 consecutive columns of
// row tid / 4.
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) rowgemm_kernel(const SmallGemmParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) __nv_bfloat16 smem[(kSBM + kSBN) * kSPitch];     // 33 KB; reused for the partial sums
  __nv_bfloat16* sA = smem;
  __nv_bfloat16* sB = smem + kSBM * kSPitch;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
  const int iters = (p.K + kSBK - 1) / kSBK;
  const int it0 = blockIdx.z * p.iters_per_split;
  const int it1 = min(iters, it0 + p.iters_per_split);

  float acc[2][4][4];          // [16-row block][8-column block][fragment]
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  uint4 ra[kVec], rb[kVec];
  tile_load<A_MN>(p.A, p.lda, m0, p.M, it0 * kSBK, p.K, tid, ra);
  tile_load<B_MN>(p.B, p.ldb, n0, p.N, it0 * kSBK, p.K, tid, rb);
  for (int it = it0; it < it1; ++it) {
    tile_store<A_MN>(sA, tid, ra);
    tile_store<B_MN>(sB, tid, rb);
    __syncthreads();
    if (it + 1 < it1) {          // next iteration's global loads fly while this one is multiplied
      tile_load<A_MN>(p.A, p.lda, m0, p.M, (it + 1) * kSBK, p.K, tid, ra);
      tile_load<B_MN>(p.B, p.ldb, n0, p.N, (it + 1) * kSBK, p.K, tid, rb);
    }
    const __nv_bfloat16* a_base = sA + g * kSPitch + warp * 64 + 2 * t;
    const __nv_bfloat16* b_base = sB + g * kSPitch + warp * 64 + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[2][4], b[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const __nv_bfloat16* q = a_base + i * 16 * kSPitch + ks * 16;
        a[i][0] = *reinterpret_cast<const uint32_t*>(q);
        a[i][1] = *reinterpret_cast<const uint32_t*>(q + 8 * kSPitch);
        a[i][2] = *reinterpret_cast<const uint32_t*>(q + 8);
        a[i][3] = *reinterpret_cast<const uint32_t*>(q + 8 * kSPitch + 8);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16* q = b_base + j * 8 * kSPitch + ks * 16;
        b[j][0] = *reinterpret_cast<const uint32_t*>(q);
        b[j][1] = *reinterpret_cast<const uint32_t*>(q + 8);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }

  // ---- combine the four warps' partial tiles: part[warp][row][col], fp32, pitch 36 (float2 / float4 friendly)
  float* part = reinterpret_cast<float*>(smem);
  constexpr int kPP = 36;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* q = part + (warp * 32 + i * 16 + g) * kPP + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(q) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(q + 8 * kPP) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
  __syncthreads();
  const int lrow = tid >> 2, lcol = (tid & 3) * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float4 x = *reinterpret_cast<const float4*>(part + (w * 32 + lrow) * kPP + lcol);
    const float4 y = *reinterpret_cast<const float4*>(part + (w * 32 + lrow) * kPP + lcol + 4);
    v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
    v[4] += y.x; v[5] += y.y; v[6] += y.z; v[7] += y.w;
  }

  // ---- epilogue (same order as tt_gemm.cu): alpha / bias, ReLU, dropout, gate, residual, stores
  const int row = m0 + lrow, col0 = n0 + lcol;
  if (row >= p.M || col0 >= p.N) return;
  const int ncol = min(8, p.N - col0);            // even (launcher)
  const uint64_t seed = p.drop_seed + ((p.drop_thresh && p.drop_seed_dev) ? *p.drop_seed_dev : 0ull);
  const uint32_t dkey = drop_key(seed, p.drop_site);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (e >= ncol) break;
    float x = v[e];
    if (p.bias) x = fmaf(x, p.alpha, p.bias[col0 + e]);
    else if (p.alpha != 1.f) x *= p.alpha;
    if (p.relu) x = x < 0.f ? 0.f : x;
    if (p.drop_thresh) {
      const uint32_t idx = static_cast<uint32_t>(row) * static_cast<uint32_t>(p.N) + static_cast<uint32_t>(col0 + e);
      x = drop_keep_k(dkey, idx, p.drop_thresh) ? x * p.drop_scale : 0.f;
    }
    if (p.gate) {
      const float gv = __bfloat162float(p.gate[static_cast<size_t>(row) * p.ld_gate + col0 + e]);
      x = gv > 0.f ? x * p.gate_scale : 0.f;
    }
    if (p.residual) x += p.residual[static_cast<size_t>(row) * p.ld_res + col0 + e];
    v[e] = x;
  }
  if (p.out_f32) {
    float* o = p.out_f32 + static_cast<size_t>(row) * p.ld_f32 + col0;
    if (p.accumulate) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e < ncol) atomicAdd(o + e, v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; e += 2)
        if (e < ncol) *reinterpret_cast<float2*>(o + e) = make_float2(v[e], v[e + 1]);
    }
  }
  if (p.out_bf16) {
    __nv_bfloat16* o = p.out_bf16 + static_cast<size_t>(row) * p.ld_bf16 + col0;
#pragma unroll
    for (int e = 0; e < 8; e += 2)
      if (e < ncol) *reinterpret_cast<uint32_t*>(o + e) = pack_bf16(v[e], v[e + 1]);
  }
}

static int rowgemm_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TT_ROWGEMM");     // 0 = every GEMM on the tcgen05 kernel (A/B timing)
    v = e ? atoi(e) : 1;
  }
  return v;
}

// The small-problem path takes a launch when it is small in every dimension and the vector loads line up.
bool rowgemm_eligible(const tt_gemm_args* a) {
  if (rowgemm_mode() == 0 || a->block_n != 0) return false;     // an explicit tile width asks for the tcgen05 kernel
  if (a->M > 1024 || a->N > 1024 || a->K > 2048) return false;
  if (static_cast<long long>(a->M) * a->N * a->K > 300ll * 1000 * 1000) return false;
  if (a->K % 8 != 0 || a->N % 2 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return false;
  if (a->lda % 8 != 0 || a->ldb % 8 != 0) return false;
  if (a->a_mn && a->M % 8 != 0) return false;
  if (a->b_mn && a->N % 8 != 0) return false;
  if (a->out_f32 && ((reinterpret_cast<uintptr_t>(a->out_f32) & 7) || a->ld_f32 % 2 != 0)) return false;
  if (a->out_bf16 && ((reinterpret_cast<uintptr_t>(a->out_bf16) & 3) || a->ld_bf16 % 2 != 0)) return false;
  if (a->residual && ((reinterpret_cast<uintptr_t>(a->residual) & 7) || a->ld_res % 2 != 0)) return false;
  if (a->gate && ((reinterpret_cast<uintptr_t>(a->gate) & 3) || a->ld_gate % 2 != 0)) return false;
  return true;
}

static uint32_t small_drop_threshold(float p) {
  if (p <= 0.f) return 0;
  double t = static_cast<double>(p) * 4294967296.0;
  uint32_t v = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
  return v == 0 ? 1 : v;
}

int rowgemm_launch(const tt_gemm_args* a, cudaStream_t stream) {
  SmallGemmParams p;
  p.A = static_cast<const __nv_bfloat16*>(a->A); p.B = static_cast<const __nv_bfloat16*>(a->B);
  p.lda = a->lda; p.ldb = a->ldb; p.M = a->M; p.N = a->N; p.K = a->K;
  p.alpha = a->alpha; p.bias = a->bias; p.relu = a->relu;
  p.drop_thresh = small_drop_threshold(a->drop_p);
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = a->drop_seed; p.drop_seed_dev = a->drop_seed_dev; p.drop_site = a->drop_site;
  p.gate = static_cast<const __nv_bfloat16*>(a->gate); p.ld_gate = a->ld_gate; p.gate_scale = a->gate_scale;
  p.residual = a->residual; p.ld_res = a->ld_res;
  p.out_f32 = a->out_f32; p.ld_f32 = a->ld_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); p.ld_bf16 = a->ld_bf16;
  p.accumulate = a->accumulate;
  const int iters = (a->K + kSBK - 1) / kSBK;
  const int tiles = ((a->M + kSBM - 1) / kSBM) * ((a->N + kSBN - 1) / kSBN);
  int ks = 1;
  if (a->accumulate) {          // split the reduction until the grid covers the machine (atomics combine)
    ks = a->k_splits > 0 ? a->k_splits : (2 * num_sms() + tiles - 1) / tiles;
    if (ks > iters) ks = iters;
    if (ks < 1) ks = 1;
  }
  p.iters_per_split = (iters + ks - 1) / ks;
  ks = (iters + p.iters_per_split - 1) / p.iters_per_split;
  dim3 grid((a->N + kSBN - 1) / kSBN, (a->M + kSBM - 1) / kSBM, ks);
  cudaError_t e;
  if (a->a_mn && a->b_mn) e = launch_k(rowgemm_kernel<true, true>, grid, dim3(128), 0, stream, p);
  else if (a->a_mn) e = launch_k(rowgemm_kernel<true, false>, grid, dim3(128), 0, stream, p);
  else if (a->b_mn) e = launch_k(rowgemm_kernel<false, true>, grid, dim3(128), 0, stream, p);
  else e = launch_k(rowgemm_kernel<false, false>, grid, dim3(128), 0, stream, p);
  if (e != cudaSuccess) return cuda_fail(e, "rowgemm launch");
  TT_LAUNCH_CHECK();
  return TT_OK;
}

}  // namespace tt
