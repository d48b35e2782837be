// Duplicate-free traffic for a row-sharded ID table.
//
// Reference: nn.Embedding(vocab_size, 256, padding_idx=0) lookups and their dense gradient
// (src/models/user_tower.py:26, 86; autograd's embedding_dense_backward). Histories are Zipfian in the item id
// (SURVEY.md §8d): at batch 256 x 200 only about a third of the 51,200 tokens carry a distinct id, and the ten most
// popular items make up a quarter of all tokens. With the table row-sharded over NVLink that matters twice — every
// duplicate would cross the link again, and every duplicate's gradient would be one more remote atomic on the same
// hot row. So a step first reduces its ids to the distinct ones:
//
//   tt_ids_dedup        ids[T] -> uniq[0..n) (uniq[0] = 0, the padding id), inverse[t] = slot of ids[t]
//                       (direct-address table over the vocabulary: one atomicExch per token decides who registers an id)
//   tt_rows_gather      cache[s] = table row uniq[s], read from the OWNER's memory over NVLink: each distinct row
//                       crosses the link once; the embedding kernels then run on the compact cache with the slot
//                       numbers as ids (it stays in L2: 18 MB at c2)
//   tt_rows_scatter_add owner's grad row uniq[s] += gacc[s] (red.global.add.v4.f32 over NVLink), gacc[s] = 0:
//                       the per-token gradients were combined locally in gacc (by the embedding backward's
//                       atomics on the L2-resident compact buffer), one remote reduction per distinct row remains.
#include "../../include/tt_b200.h"
#include "tt_common.cuh"

namespace tt {

int make_sharded_table(TableRef& t, const ::tt_symm_team* team, int64_t offset, const char* who) {
  if (team == nullptr || team->world < 1 || team->world > TT_SYMM_MAX_RANKS || offset < 0 || offset % 16 != 0) {
    set_last_error("%s: bad team / shard offset", who);
    return TT_ERR_INVALID;
  }
  for (int r = 0; r < TT_SYMM_MAX_RANKS; ++r)
    t.base[r] = r < team->world ? reinterpret_cast<float*>(static_cast<uint8_t*>(team->bufs[r]) + offset) : nullptr;
  for (int r = 0; r < team->world; ++r)
    if (team->bufs[r] == nullptr) {
      set_last_error("%s: rank %d has no mapping", who, r);
      return TT_ERR_INVALID;
    }
  t.world = team->world;
  return TT_OK;
}

// state[0] = working counter (next free slot; slot 0 is always the padding id), state[1] = published number of
// distinct ids of the last completed call (what the gather / scatter kernels and the next call's reset read).
// flag / state are zero-filled by the caller once; every call clears exactly the flags the previous one set.
__global__ void __launch_bounds__(256) dedup_reset_kernel(int* __restrict__ flag, int64_t* __restrict__ uniq,
                                                          int* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  const int n_prev = state[1];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x + 1; s < n_prev; s += gridDim.x * blockDim.x) flag[uniq[s]] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state[0] = 1;
    uniq[0] = 0;
  }
}

__global__ void __launch_bounds__(256) dedup_assign_kernel(const int64_t* __restrict__ ids, int T, int64_t V,
                                                           int* __restrict__ flag, int* __restrict__ slot,
                                                           int64_t* __restrict__ uniq, int* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int64_t id = ids[t];
    if (id <= 0 || id >= V) continue;          // padding (and out-of-table ids, treated as padding)
    if (atomicExch(&flag[id], 1) == 0) {
      const int s = atomicAdd(&state[0], 1);
      slot[id] = s;
      uniq[s] = id;
    }
  }
}

__global__ void __launch_bounds__(256) dedup_inverse_kernel(const int64_t* __restrict__ ids, int T, int64_t V,
                                                            const int* __restrict__ slot, int64_t* __restrict__ inverse,
                                                            int* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int64_t id = ids[t];
    inverse[t] = (id <= 0 || id >= V) ? 0 : slot[id];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = state[0];     // published count (read by later kernels / host)
}

// one warp per distinct id: 1 KB row from the owner -> cache[s]
__global__ void __launch_bounds__(256) rows_gather_kernel(const TableRef table, const int64_t* __restrict__ uniq,
                                                          const int* __restrict__ n_uniq, float* __restrict__ cache) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = *n_uniq, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const float4* src = reinterpret_cast<const float4*>(table.row(uniq[s]));
    const float4 a = src[lane], b = src[32 + lane];
    float4* dst = reinterpret_cast<float4*>(cache + static_cast<size_t>(s) * 256);
    dst[lane] = a;
    dst[32 + lane] = b;
  }
}

// one warp per distinct id (slot 0 = padding: no gradient): owner's grad row += gacc[s]; gacc[s] = 0
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(const TableRef grad, const int64_t* __restrict__ uniq,
                                                               const int* __restrict__ n_uniq, float* __restrict__ gacc) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = *n_uniq, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1; s < n; s += warps) {
    float4* src = reinterpret_cast<float4*>(gacc + static_cast<size_t>(s) * 256);
    const float4 a = src[lane], b = src[32 + lane];
    float* dst = grad.row(uniq[s]);
    red_add_f32x4(dst + lane * 4, a.x, a.y, a.z, a.w);
    red_add_f32x4(dst + (32 + lane) * 4, b.x, b.y, b.z, b.w);
    src[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    src[32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __threadfence_system();      // remote reductions performed before the grid retires (owners' AdamW follows a barrier)
}

// Deterministic mode: one warp per distinct id; the 64-bit fixed-point sum of its tokens' gradient rows (2^-40
// units, accumulated by tt_embed_ln_bwd_det) is rounded ONCE to fp32 and added to the table's gradient row by
// the only warp that touches that row; the accumulator row is cleared for the next step.
__global__ void __launch_bounds__(256) rows_scatter_add_i64_kernel(float* __restrict__ grad, const int64_t* __restrict__ uniq,
                                                                   const int* __restrict__ n_uniq,
                                                                   long long* __restrict__ acc) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = *n_uniq, lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  constexpr double kInv = 1.0 / 1099511627776.0;   // 2^-40
  for (int s = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1; s < n; s += warps) {
    longlong2* src = reinterpret_cast<longlong2*>(acc + static_cast<size_t>(s) * 256);
    float2* dst = reinterpret_cast<float2*>(grad + static_cast<size_t>(uniq[s]) * 256);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const longlong2 a = src[k * 32 + lane];
      float2 d = dst[k * 32 + lane];
      d.x += static_cast<float>(static_cast<double>(a.x) * kInv);
      d.y += static_cast<float>(static_cast<double>(a.y) * kInv);
      dst[k * 32 + lane] = d;
      src[k * 32 + lane] = make_longlong2(0, 0);
    }
  }
}

}  // namespace tt

using namespace tt;

static int small_grid(int n_threads_wanted) {
  int g = (n_threads_wanted + 255) / 256;
  const int cap = num_sms() * 8;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

extern "C" int tt_ids_dedup(const int64_t* ids, int T, int64_t V, int32_t* flag, int32_t* slot, int64_t* uniq,
                            int32_t* state, int64_t* inverse, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TT_REQUIRE(ids && flag && slot && uniq && state && inverse && T > 0 && V > 1, "tt_ids_dedup: bad arguments");
  // clear the flags the previous call set (its list is still in uniq / state[2]), then start a new list at slot 1
  TT_CHECK_CUDA(launch_k(dedup_reset_kernel, dim3(small_grid(T)), dim3(256), 0, stream, flag, uniq, state));
  TT_LAUNCH_CHECK();
  TT_CHECK_CUDA(launch_k(dedup_assign_kernel, dim3(small_grid(T)), dim3(256), 0, stream, ids, T, V, flag, slot, uniq, state));
  TT_LAUNCH_CHECK();
  TT_CHECK_CUDA(launch_k(dedup_inverse_kernel, dim3(small_grid(T)), dim3(256), 0, stream, ids, T, V, static_cast<const int*>(slot), inverse, state));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_rows_gather(const tt_symm_team* team, int64_t weight_offset, const float* table_local,
                              const int64_t* uniq, const int32_t* n_uniq, int max_rows, float* cache, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TT_REQUIRE(uniq && n_uniq && cache && max_rows > 0 && (team != nullptr) != (table_local != nullptr),
             "tt_rows_gather: bad arguments (exactly one of team / table_local)");
  TableRef t;
  if (team) {
    int rc = make_sharded_table(t, team, weight_offset, "tt_rows_gather");
    if (rc) return rc;
  } else {
    for (int r = 0; r < 16; ++r) t.base[r] = nullptr;
    t.base[0] = const_cast<float*>(table_local);
    t.world = 0;
  }
  TT_CHECK_CUDA(launch_k(rows_gather_kernel, dim3(small_grid(max_rows * 32)), dim3(256), 0, stream, t, uniq, n_uniq, cache));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_rows_scatter_add(const tt_symm_team* team, int64_t grad_offset, float* grad_local, const int64_t* uniq,
                                   const int32_t* n_uniq, int max_rows, float* gacc, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TT_REQUIRE(uniq && n_uniq && gacc && max_rows > 0 && (team != nullptr) != (grad_local != nullptr),
             "tt_rows_scatter_add: bad arguments (exactly one of team / grad_local)");
  TableRef t;
  if (team) {
    int rc = make_sharded_table(t, team, grad_offset, "tt_rows_scatter_add");
    if (rc) return rc;
  } else {
    for (int r = 0; r < 16; ++r) t.base[r] = nullptr;
    t.base[0] = grad_local;
    t.world = 0;
  }
  TT_CHECK_CUDA(launch_k(rows_scatter_add_kernel, dim3(small_grid(max_rows * 32)), dim3(256), 0, stream, t, uniq, n_uniq, gacc));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_rows_scatter_add_i64(float* grad_local, const int64_t* uniq, const int32_t* n_uniq, int max_rows,
                                       int64_t* acc64, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TT_REQUIRE(grad_local && uniq && n_uniq && acc64 && max_rows > 0, "tt_rows_scatter_add_i64: bad arguments");
  TT_CHECK_CUDA(launch_k(rows_scatter_add_i64_kernel, dim3(small_grid(max_rows * 32)), dim3(256), 0, stream, grad_local,
                         uniq, n_uniq, reinterpret_cast<long long*>(acc64)));
  TT_LAUNCH_CHECK();
  return TT_OK;
}
