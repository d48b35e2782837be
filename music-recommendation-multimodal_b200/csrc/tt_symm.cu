// Cross-GPU steps of the data-parallel training step written as kernels over NVLink peer memory.
//
// The reference's multi-GPU wiring is DistributedDataParallel over NCCL (src/train.py:29-35, 300): one
// gradient all-reduce per step, then every rank runs the same torch.optim.AdamW (:302) on the full
// parameter set. Here every rank owns 1/G of the flat parameter buffer and ONE kernel per step does
//     gradient reduce-scatter  ->  AdamW on the owned slice  ->  parameter (+ bf16 shadow) all-gather
// over a symmetric arena (the same allocation mapped into every rank, plus its NVLS multicast mapping):
// `multimem.ld_reduce` pulls the G-way gradient sum of the owned slice through the switch, the update happens
// in registers, `multimem.st` pushes the new fp32 values and their bf16 copies to all ranks. The receive
// direction carries the reduced gradients while the send direction carries the parameters, there is no
// intermediate buffer and no launch gap between the three stages. Without a multicast mapping the same kernel
// reads / writes the peers' unicast mappings in rank order (deterministic sum).
// The small exchanges of the step (normalised embeddings + user ids, row log-sum-exps) use the same arena:
// every rank stores its rows into all ranks' gathered buffers and the kernel ends with a cross-rank barrier,
// so the whole step is one CUDA graph with no library collective inside.
//
// Cross-rank synchronisation: every arena has a 4 KB control block. A collective is numbered by a local epoch
// counter (all ranks issue the same sequence of collectives on an arena). "Arrive" flags are monotonic epoch
// stamps written by the peers into MY control block (slot = writer's rank) with release semantics at system
// scope; waiting is an acquire spin on local memory. A wait that lasts longer than ~10 s sets the block's
// error word and gives up (a peer died): the host reads the word with the loss and raises.
#include "../../include/tt_b200.h"
#include "tt_common.cuh"

namespace tt {

struct SymmCtrl {
  uint32_t epoch;          // collectives completed on this arena (local)
  uint32_t done_ctas;      // last-block detection
  uint32_t error;          // != 0: a wait timed out
  uint32_t pad0[13];
  uint32_t ready[16];      // phase A: "my inputs are complete / you may overwrite my gathered buffers"
  uint32_t done[16];       // phase B: "I have finished reading your memory and writing into it"
};

struct Team {
  int rank, world;
  uint8_t* bufs[TT_SYMM_MAX_RANKS];
  uint8_t* mc;             // nullptr: no multicast mapping
  long long ctrl_off;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ SymmCtrl* ctrl_of(const Team& t, int r) {
  return reinterpret_cast<SymmCtrl*>(t.bufs[r] + t.ctrl_off);
}

// spin until every rank's stamp in `flags` (local memory) has reached epoch e
__device__ __forceinline__ void wait_all(const Team& t, const uint32_t* flags, uint32_t e, uint32_t* err) {
  const unsigned long long t0 = global_ns();
  for (int q = 0; q < t.world; ++q) {
    while (static_cast<int32_t>(ld_acquire_sys(flags + q) - e) < 0) {
      if (global_ns() - t0 > 10000000000ull) {
        atomicExch(err, 1u);
        return;
      }
      __nanosleep(64);
    }
  }
}
// stamp epoch e into slot `rank` of the given flag array of every rank's control block
__device__ __forceinline__ void signal_all(const Team& t, size_t flag_off, uint32_t e) {
  __threadfence_system();
  for (int p = 0; p < t.world; ++p) {
    uint32_t* f = reinterpret_cast<uint32_t*>(t.bufs[p] + t.ctrl_off + flag_off) + t.rank;
    st_release_sys(f, e);
  }
}

// Called by every thread of every CTA when the CTA's remote reads / writes are issued. The last CTA of the grid
// tells the peers this rank is done and waits for all of them, so when the KERNEL completes every rank has
// finished touching this rank's memory and every write into it is visible.
__device__ __forceinline__ void finish_collective(const Team& t, uint32_t e) {
  __syncthreads();
  if (threadIdx.x == 0) {
    SymmCtrl* c = ctrl_of(t, t.rank);
    __threadfence_system();
    const uint32_t prev = atomicAdd(&c->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence();
      c->done_ctas = 0;
      signal_all(t, offsetof(SymmCtrl, done), e);
      wait_all(t, c->done, e, &c->error);
      c->epoch = e;
      __threadfence();
    }
  }
}

__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const void* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f32x4(void* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st_b32x2(void* mc_addr, uint32_t a, uint32_t b) {
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(mc_addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 ld_peer_f32x4(const void* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// --------------------------------------------------------------------------------------------
// All-gather of up to four row blocks into every rank's arena (rank r's block lands at dst_off + r * nbytes).
// --------------------------------------------------------------------------------------------
struct GatherSeg {
  const uint8_t* src;
  long long dst_off, nbytes;      // nbytes per rank, multiple of 16
};
struct GatherParams {
  Team team;
  GatherSeg seg[4];
  int n_seg;
  int pre_barrier;                // 1: peers may still be reading the destination (no barrier since their last use)
};

__global__ void __launch_bounds__(256) symm_allgather_kernel(const GatherParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const Team& t = p.team;
  SymmCtrl* c = ctrl_of(t, t.rank);
  const uint32_t e = c->epoch + 1u;
  if (p.pre_barrier) {
    if (blockIdx.x == 0 && threadIdx.x == 0) signal_all(t, offsetof(SymmCtrl, ready), e);
    if (threadIdx.x == 0) wait_all(t, c->ready, e, &c->error);
    __syncthreads();
  }
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int s = 0; s < p.n_seg; ++s) {
    const GatherSeg g = p.seg[s];
    const long long n16 = g.nbytes >> 4;
    const long long base = g.dst_off + static_cast<long long>(t.rank) * g.nbytes;
    for (long long i = tid; i < n16; i += nthreads) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g.src) + i);
      if (t.mc != nullptr) {
        mc_st_f32x4(t.mc + base + (i << 4), v);
      } else {
        for (int q = 0; q < t.world; ++q) *reinterpret_cast<float4*>(t.bufs[q] + base + (i << 4)) = v;
      }
    }
  }
  finish_collective(t, e);
}

// --------------------------------------------------------------------------------------------
// reduce-scatter (mean over ranks) -> AdamW -> all-gather, one pass over this rank's 1/G of the flat buffer.
// Arithmetic per element is that of adamw_kernel (tt_rowwise.cu): torch.optim.AdamW, decoupled decay.
// --------------------------------------------------------------------------------------------
struct DpAdamwParams {
  Team team;
  long long flat_off, grad_off, shadow_off;   // byte offsets inside the arena
  long long n4;                               // float4 elements of the WHOLE flat buffer (multiple of world)
  long long shadow_begin4;                    // first float4 element that has a bf16 shadow
  float* m; float* v;                         // local moments of the owned slice [n4 / world * 4]
  float lr, beta1, beta2, eps, wd;
  const int64_t* step_dev;
  // optional: this rank's shard of a row-sharded ID table (plain local memory). It is updated in the same kernel,
  // between the two barriers: phase A guarantees every rank's gradient rows have landed in loc_g, phase B that
  // every owner's rows are final before anybody's next step gathers them.
  float* loc_p; float* loc_g; float* loc_m; float* loc_v;
  long long loc_n4;
  float loc_scale;                            // gradient used = loc_scale * loc_g (1 / world: loc_g holds the SUM)
};

template <bool kMulticast>
__global__ void __launch_bounds__(256) dp_adamw_kernel(const DpAdamwParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const Team& t = p.team;
  SymmCtrl* c = ctrl_of(t, t.rank);
  const uint32_t e = c->epoch + 1u;
  // phase A: every rank's backward pass is complete (its gradient buffer is final)
  if (blockIdx.x == 0 && threadIdx.x == 0) signal_all(t, offsetof(SymmCtrl, ready), e);
  if (threadIdx.x == 0) wait_all(t, c->ready, e, &c->error);
  __syncthreads();

  const float step = static_cast<float>(*p.step_dev);
  const float bc1 = 1.f - powf(p.beta1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(p.beta2, step));
  const float step_size = p.lr / bc1;
  const float decay = 1.f - p.lr * p.wd;
  const float inv_world = 1.f / static_cast<float>(t.world);
  const long long per = p.n4 / t.world;
  const long long lo = per * t.rank;
  constexpr int kUnroll = 4;      // gradient loads in flight per thread: NVLink round trips are microseconds
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; j0 < per; j0 += nthreads * kUnroll) {
    float4 g[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long j = j0 + u * nthreads;
      if (j < per) {
        const long long off = p.grad_off + ((lo + j) << 4);
        if constexpr (kMulticast) {
          g[u] = mc_ld_reduce_f32x4(t.mc + off);
        } else {
          g[u] = ld_peer_f32x4(t.bufs[0] + off);
          for (int q = 1; q < t.world; ++q) {
            const float4 x = ld_peer_f32x4(t.bufs[q] + off);
            g[u].x += x.x; g[u].y += x.y; g[u].z += x.z; g[u].w += x.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long j = j0 + u * nthreads;
      if (j >= per) continue;
      const long long i = lo + j;
      float4 pp = *reinterpret_cast<const float4*>(t.bufs[t.rank] + p.flat_off + (i << 4));
      float4 mm = reinterpret_cast<float4*>(p.m)[j];
      float4 vv = reinterpret_cast<float4*>(p.v)[j];
      float* pa = reinterpret_cast<float*>(&pp);
      const float* ga = reinterpret_cast<const float*>(&g[u]);
      float* ma = reinterpret_cast<float*>(&mm);
      float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = ga[k] * inv_world;
        const float x = pa[k] * decay;
        ma[k] = p.beta1 * ma[k] + (1.f - p.beta1) * gk;
        va[k] = p.beta2 * va[k] + (1.f - p.beta2) * gk * gk;
        const float denom = sqrtf(va[k]) / bc2_sqrt + p.eps;
        pa[k] = x - step_size * ma[k] / denom;
      }
      reinterpret_cast<float4*>(p.m)[j] = mm;
      reinterpret_cast<float4*>(p.v)[j] = vv;
      const bool sh = i >= p.shadow_begin4;
      const uint32_t s0 = pack_bf16(pp.x, pp.y), s1 = pack_bf16(pp.z, pp.w);
      const long long foff = p.flat_off + (i << 4);
      const long long soff = p.shadow_off + ((i - p.shadow_begin4) << 3);
      if constexpr (kMulticast) {
        mc_st_f32x4(t.mc + foff, pp);
        if (sh) mc_st_b32x2(t.mc + soff, s0, s1);
      } else {
        for (int q = 0; q < t.world; ++q) {
          *reinterpret_cast<float4*>(t.bufs[q] + foff) = pp;
          if (sh) *reinterpret_cast<uint2*>(t.bufs[q] + soff) = make_uint2(s0, s1);
        }
      }
    }
  }
  // the owned rows of the row-sharded ID table: dense AdamW, gradient cleared in the same pass
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.loc_n4; i += nthreads) {
    float4 pp = reinterpret_cast<float4*>(p.loc_p)[i];
    const float4 gg = reinterpret_cast<const float4*>(p.loc_g)[i];
    float4 mm = reinterpret_cast<float4*>(p.loc_m)[i];
    float4 vv = reinterpret_cast<float4*>(p.loc_v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * p.loc_scale;
      const float x = pa[k] * decay;
      ma[k] = p.beta1 * ma[k] + (1.f - p.beta1) * gk;
      va[k] = p.beta2 * va[k] + (1.f - p.beta2) * gk * gk;
      const float denom = sqrtf(va[k]) / bc2_sqrt + p.eps;
      pa[k] = x - step_size * ma[k] / denom;
    }
    reinterpret_cast<float4*>(p.loc_p)[i] = pp;
    reinterpret_cast<float4*>(p.loc_m)[i] = mm;
    reinterpret_cast<float4*>(p.loc_v)[i] = vv;
    if ((__float_as_uint(gg.x) | __float_as_uint(gg.y) | __float_as_uint(gg.z) | __float_as_uint(gg.w)) != 0u)
      reinterpret_cast<float4*>(p.loc_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // phase B: all ranks have read my gradients and written my parameters when the kernel completes
  finish_collective(t, e);
}

__global__ void symm_barrier_kernel(const Team t) {
  pdl_launch_dependents();
  pdl_wait();      // the epoch must be read after the previous collective on this arena has published it
  SymmCtrl* c = ctrl_of(t, t.rank);
  const uint32_t e = c->epoch + 1u;
  if (threadIdx.x == 0) {
    signal_all(t, offsetof(SymmCtrl, ready), e);
    wait_all(t, c->ready, e, &c->error);
  }
  finish_collective(t, e);
}

static int fill_team(Team& t, const tt_symm_team* a, const char* who) {
  if (a == nullptr || a->world < 1 || a->world > TT_SYMM_MAX_RANKS || a->rank < 0 || a->rank >= a->world) {
    set_last_error("%s: bad team (world must be in [1, %d])", who, TT_SYMM_MAX_RANKS);
    return TT_ERR_INVALID;
  }
  t.rank = a->rank; t.world = a->world; t.mc = static_cast<uint8_t*>(a->multicast); t.ctrl_off = a->ctrl_offset;
  for (int r = 0; r < TT_SYMM_MAX_RANKS; ++r) t.bufs[r] = r < a->world ? static_cast<uint8_t*>(a->bufs[r]) : nullptr;
  for (int r = 0; r < a->world; ++r)
    if (t.bufs[r] == nullptr || (reinterpret_cast<uintptr_t>(t.bufs[r]) & 15) != 0) {
      set_last_error("%s: arena mapping of rank %d is null or not 16-byte aligned", who, r);
      return TT_ERR_INVALID;
    }
  if (a->ctrl_offset < 0 || (a->ctrl_offset & 127) != 0) {
    set_last_error("%s: control block offset must be a non-negative multiple of 128", who);
    return TT_ERR_INVALID;
  }
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" int tt_symm_ctrl_bytes(void) { return 4096; }

extern "C" int tt_symm_allgather(const tt_symm_team* team, const tt_symm_segment* segs, int n_seg, int pre_barrier,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GatherParams p;
  int rc = fill_team(p.team, team, "tt_symm_allgather");
  if (rc) return rc;
  TT_REQUIRE(segs && n_seg >= 1 && n_seg <= 4, "tt_symm_allgather: 1..4 segments");
  long long total16 = 0;
  for (int s = 0; s < 4; ++s) {
    if (s < n_seg) {
      TT_REQUIRE(segs[s].src && segs[s].nbytes > 0 && segs[s].nbytes % 16 == 0 && segs[s].dst_offset % 16 == 0 &&
                     (reinterpret_cast<uintptr_t>(segs[s].src) & 15) == 0,
                 "tt_symm_allgather: segment %d must be 16-byte aligned and a multiple of 16 bytes", s);
      p.seg[s].src = static_cast<const uint8_t*>(segs[s].src);
      p.seg[s].dst_off = segs[s].dst_offset;
      p.seg[s].nbytes = segs[s].nbytes;
      if (segs[s].nbytes / 16 > total16) total16 = segs[s].nbytes / 16;
    } else {
      p.seg[s].src = nullptr; p.seg[s].dst_off = 0; p.seg[s].nbytes = 0;
    }
  }
  p.n_seg = n_seg;
  p.pre_barrier = pre_barrier;
  long long grid = (total16 + 255) / 256;
  if (grid > num_sms()) grid = num_sms();
  if (grid < 1) grid = 1;
  TT_CHECK_CUDA(launch_k(symm_allgather_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream, p));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_dp_adamw_step(const tt_symm_team* team, int64_t flat_offset, int64_t grad_offset,
                                int64_t shadow_offset, int64_t n, int64_t shadow_begin, float* m, float* v, float lr,
                                float beta1, float beta2, float eps, float weight_decay, const int64_t* step_dev,
                                float* shard_p, float* shard_g, float* shard_m, float* shard_v, int64_t shard_n,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DpAdamwParams p;
  int rc = fill_team(p.team, team, "tt_dp_adamw_step");
  if (rc) return rc;
  TT_REQUIRE(m && v && step_dev && n > 0, "tt_dp_adamw_step: null pointer");
  TT_REQUIRE(n % (4 * team->world) == 0 && shadow_begin % 4 == 0 && shadow_begin >= 0 && shadow_begin <= n,
             "tt_dp_adamw_step: n (%lld) must split into %d shards of whole float4s, shadow_begin into float4s",
             static_cast<long long>(n), team->world);
  TT_REQUIRE(flat_offset % 16 == 0 && grad_offset % 16 == 0 && shadow_offset % 16 == 0,
             "tt_dp_adamw_step: arena offsets must be multiples of 16");
  p.flat_off = flat_offset; p.grad_off = grad_offset; p.shadow_off = shadow_offset;
  p.n4 = n / 4; p.shadow_begin4 = shadow_begin / 4;
  p.m = m; p.v = v; p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.wd = weight_decay; p.step_dev = step_dev;
  TT_REQUIRE(shard_n >= 0 && shard_n % 4 == 0 && (shard_n == 0 || (shard_p && shard_g && shard_m && shard_v)),
             "tt_dp_adamw_step: the local table shard needs p, g, m, v and a multiple of 4 elements");
  p.loc_p = shard_p; p.loc_g = shard_g; p.loc_m = shard_m; p.loc_v = shard_v; p.loc_n4 = shard_n / 4;
  p.loc_scale = 1.f / static_cast<float>(team->world);
  const long long per = p.n4 / team->world;
  long long grid = (per + 256 * 4 - 1) / (256 * 4);
  const long long grid_loc = (p.loc_n4 + 256 * 4 - 1) / (256 * 4);
  if (grid_loc > grid) grid = grid_loc;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (team->multicast != nullptr)
    TT_CHECK_CUDA(launch_k(dp_adamw_kernel<true>, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream, p));
  else
    TT_CHECK_CUDA(launch_k(dp_adamw_kernel<false>, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream, p));
  TT_LAUNCH_CHECK();
  return TT_OK;
}

extern "C" int tt_symm_barrier(const tt_symm_team* team, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Team t;
  int rc = fill_team(t, team, "tt_symm_barrier");
  if (rc) return rc;
  TT_CHECK_CUDA(launch_k(symm_barrier_kernel, dim3(1), dim3(32), 0, stream, t));
  TT_LAUNCH_CHECK();
  return TT_OK;
}
