"""Host-side partitioning rules of the multi-GPU paths (pure Python / torch CPU, no device code)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_bounds(num_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row-sharding of the item table (V rows, row 0 = padding): rank r owns
    [r * ceil(V / world), ...). Returns (first_row, num_rows_local)."""
    per = (num_rows + world - 1) // world
    first = min(rank * per, num_rows)
    return first, max(0, min(num_rows, first + per) - first)


def merge_canonical(scores: torch.Tensor, idx: torch.Tensor):
    """Reference semantics of the cross-shard merge (the CUDA kernel tt_topk_merge implements the
    same rule): top-K of the union of [G, U, K] lists under (score desc, global index asc)."""
    G, U, K = scores.shape
    s = scores.permute(1, 0, 2).reshape(U, G * K)
    i = idx.permute(1, 0, 2).reshape(U, G * K).long()
    order = torch.argsort(i, dim=1, stable=True)                 # index ascending ...
    s, i = s.gather(1, order), i.gather(1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)  # ... then score descending (stable)
    return i.gather(1, order)[:, :K].to(idx.dtype), s.gather(1, order)[:, :K]


def merge_bounded_reference(scores: torch.Tensor, idx: torch.Tensor, bounds: torch.Tensor, K: int):
    """Reference semantics of the bounded shard protocol (retrieval.sharded_topk / merge_bounded on the device):
    lists [G, U, K'] hold a shard's best K' items with exact scores (-1 / -inf padded), bounds [G, U] say "no
    item of shard g outside its list scores above bounds[g, u]". Returns (idx (U, K), score (U, K), certified
    bool (U,)): the top K of the union is the exact global top K whenever its K-th score beats every bound —
    then every item that could enter the top K is in some list."""
    G, U, Kin = scores.shape
    s = scores.permute(1, 0, 2).reshape(U, G * Kin).clone()
    i = idx.permute(1, 0, 2).reshape(U, G * Kin).long()
    s[i < 0] = float("-inf")
    big = torch.iinfo(torch.int64).max
    order = torch.argsort(torch.where(i < 0, torch.full_like(i, big), i), dim=1, stable=True)
    s, i = s.gather(1, order), i.gather(1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)
    s, i = s.gather(1, order)[:, :K], i.gather(1, order)[:, :K]
    bmax = bounds.max(dim=0).values
    certified = torch.where(i[:, K - 1] >= 0, s[:, K - 1] > bmax, torch.isneginf(bmax))
    return i.to(idx.dtype), s, certified


class RowShardedTable:
    """Row-sharded item-ID embedding table for catalogs that do not fit replicated.

    Reference semantics: ``nn.Embedding(vocab_size, 256, padding_idx=0)`` and its dense gradient
    (src/models/user_tower.py:26,86; autograd `embedding_dense_backward`), AdamW over every row
    (src/train.py:302). Rank r owns the contiguous rows `shard_bounds(V, world, r)`. A lookup is the exchange
    step the path really has: token ids go to their owners (all-to-all), owners gather the rows, rows come back
    (all-to-all); the backward sends gradient rows the same way and the owner scatter-adds them. The optimizer
    touches local rows only. Written with torch ops + torch.distributed so the exchange logic is testable with
    gloo on CPU; on GPUs the same calls run over NCCL/NVLink.
    """

    def __init__(self, vocab_size: int, dim: int, rank: int, world: int, device, group=None):
        self.vocab_size, self.dim, self.rank, self.world, self.group = vocab_size, dim, rank, world, group
        self.per = (vocab_size + world - 1) // world
        self.first, self.rows = shard_bounds(vocab_size, world, rank)
        self.weight = torch.zeros(self.rows, dim, device=device)
        self.grad = torch.zeros(self.rows, dim, device=device)
        self.exp_avg = None
        self.exp_avg_sq = None
        self._saved = None

    def describe(self) -> str:
        return (f"contiguous row shards ({self.per} rows/rank); per step: all_to_all of counts and token ids, "
                f"index_select on the owner, all_to_all of {self.dim * 4} B rows back; backward: all_to_all of "
                f"gradient rows + index_add_ on the owner (torch ops over NCCL)")

    def load_full(self, full_table: torch.Tensor) -> None:
        self.weight.copy_(full_table[self.first:self.first + self.rows])

    def _a2a(self, out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits) -> None:
        if self.world == 1:
            out.copy_(inp)
            return
        import torch.distributed as dist
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        """rows[t] = table[ids[t]] for this rank's tokens (ids int64 [T], global item ids). Collective."""
        ids = ids.reshape(-1)
        owner = torch.div(ids, self.per, rounding_mode="floor")
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=self.world)
        recv_counts = torch.empty_like(send_counts)
        self._a2a(recv_counts, send_counts, None, None)
        send_splits, recv_splits = send_counts.tolist(), recv_counts.tolist()      # host sync: variable message sizes
        ids_sorted = ids[order]
        ids_recv = torch.empty(sum(recv_splits), dtype=ids.dtype, device=ids.device)
        self._a2a(ids_recv, ids_sorted, recv_splits, send_splits)
        local = ids_recv - self.first
        rows_out = self.weight.index_select(0, local)
        rows_back = torch.empty(ids.numel(), self.dim, device=ids.device, dtype=self.weight.dtype)
        self._a2a(rows_back, rows_out, send_splits, recv_splits)
        rows = torch.empty_like(rows_back)
        rows[order] = rows_back
        self._saved = (order, send_splits, recv_splits, ids_recv, local)
        return rows

    def backward(self, drows: torch.Tensor, scale: float = 1.0) -> None:
        """grad[id] += scale * sum over all ranks' tokens with that id of drows (rows of id 0 excluded: padding_idx).
        `drows` [T, dim] is the gradient w.r.t. what `lookup` returned. Collective."""
        order, send_splits, recv_splits, ids_recv, local = self._saved
        d_sorted = drows.reshape(-1, self.dim)[order]
        d_recv = torch.empty(ids_recv.numel(), self.dim, device=drows.device, dtype=drows.dtype)
        self._a2a(d_recv, d_sorted, recv_splits, send_splits)
        d_recv = d_recv * scale
        d_recv[ids_recv == 0] = 0          # padding_idx=0 never receives a gradient
        self.grad.index_add_(0, local, d_recv.to(self.grad.dtype))

    def adamw_step(self, step_dev: torch.Tensor, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                   weight_decay: float = 0.01) -> None:
        """Dense AdamW over the local rows (every row decays every step, like the replicated table); `step_dev`
        is the engine's already-advanced device step counter. CUDA only (tt_adamw_step)."""
        from . import ops
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.weight)
            self.exp_avg_sq = torch.zeros_like(self.weight)
        n = self.weight.numel()
        ops.adamw_step(self.weight.view(n), self.grad.view(n), self.exp_avg.view(n), self.exp_avg_sq.view(n), step_dev,
                       lr, betas[0], betas[1], eps, weight_decay, shadow=None, zero_grad=True)

    def gather_full(self) -> torch.Tensor:
        """The whole (V, dim) table on every rank (checkpointing / tests)."""
        if self.world == 1:
            return self.weight.clone()
        import torch.distributed as dist
        pad = torch.zeros(self.per, self.dim, device=self.weight.device, dtype=self.weight.dtype)
        pad[:self.rows] = self.weight
        out = torch.empty(self.world * self.per, self.dim, device=self.weight.device, dtype=self.weight.dtype)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        return out[:self.vocab_size]


class SymmShardedTable:
    """Row-sharded ID table in a symmetric NVLink arena (BASELINE.json configs[4]): the rows of
    ``nn.Embedding(vocab_size, 256, padding_idx=0)`` (src/models/user_tower.py:26) and of its dense gradient are
    dealt round-robin — row id lives at local row id // world of rank id % world — in one allocation that every rank
    maps. Round-robin because item popularity is Zipfian in the id: contiguous ranges would put ~90 % of every
    rank's lookups on rank 0. There is no lookup exchange: a step's distinct rows are read from, and its combined
    gradient rows added into, the owners' memory over NVLink (tt_rows_gather / tt_rows_scatter_add). The owner runs
    dense AdamW (src/train.py:302) over its rows with the gradient SUM scaled by 1 / world (data-parallel mean)."""

    def __init__(self, vocab_size: int, dim: int, group=None, device=None):
        from .symm import SymmArena
        import torch.distributed as dist
        self.vocab_size, self.dim = vocab_size, dim
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.per = (vocab_size + self.world - 1) // self.world
        self.rows = max(0, (vocab_size - self.rank + self.world - 1) // self.world)     # ids rank, rank + world, ...
        self.arena = SymmArena({"weight": self.per * dim * 4, "grad": self.per * dim * 4}, group, device)
        self.weight = self.arena.view("weight", torch.float32, (self.per, dim))
        self.grad = self.arena.view("grad", torch.float32, (self.per, dim))
        self.exp_avg = None
        self.exp_avg_sq = None

    def describe(self) -> str:
        how = "NVLS-capable arena" if self.arena.multicast else "peer-mapped arena"
        return (f"rows dealt round-robin over {self.world} ranks ({self.per} rows/rank) in a symmetric {how}; forward: "
                f"the step's DISTINCT rows are gathered once from the owners' memory over NVLink into a compact cache, "
                f"backward: per-token gradients combined locally, one red.global.add.v4.f32 per distinct row into the "
                f"owner's gradient shard; no id/row exchange between ranks; owners run AdamW on V/world rows")

    def load_full(self, full_table: torch.Tensor) -> None:
        self.weight[:self.rows].copy_(full_table[self.rank::self.world])
        self.barrier()           # peers gather from this shard: nobody reads it before every owner has loaded

    def barrier(self) -> None:
        self.arena.barrier()

    def adamw_step(self, step_dev: torch.Tensor, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                   weight_decay: float = 0.01) -> None:
        from . import ops
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(self.per * self.dim, device=self.weight.device)
            self.exp_avg_sq = torch.zeros(self.per * self.dim, device=self.weight.device)
        n = self.per * self.dim
        ops.adamw_step(self.weight.view(n), self.grad.view(n), self.exp_avg, self.exp_avg_sq, step_dev, lr, betas[0],
                       betas[1], eps, weight_decay, shadow=None, zero_grad=True, grad_scale=1.0 / self.world)

    def shard_tensors(self):
        """(p, g, m, v) of the owned rows as flat tensors — what tt_dp_adamw_step updates inside its barriers."""
        n = self.per * self.dim
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(n, device=self.weight.device)
            self.exp_avg_sq = torch.zeros(n, device=self.weight.device)
        return self.weight.view(n), self.grad.view(n), self.exp_avg, self.exp_avg_sq

    def gather_full(self) -> torch.Tensor:
        import torch.distributed as dist
        parts = torch.empty(self.world, self.per, self.dim, device=self.weight.device)
        dist.all_gather_into_tensor(parts.view(self.world * self.per, self.dim), self.weight.contiguous(),
                                    group=self.arena.group)
        out = torch.empty(self.vocab_size, self.dim, device=self.weight.device)
        for r in range(self.world):
            n = max(0, (self.vocab_size - r + self.world - 1) // self.world)
            out[r::self.world] = parts[r, :n]
        return out
