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
