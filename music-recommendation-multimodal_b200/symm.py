"""Symmetric arenas: one allocation per rank, every rank's copy mapped into every process (NVLink peer
memory) plus, where the fabric offers it, the NVLS multicast mapping — the memory the cross-GPU kernels of
``csrc/tt_symm.cu`` work on (gradient reduce-scatter -> AdamW -> parameter all-gather in one kernel, the small
all-gathers of the gathered-negatives exchange, the row-sharded ID table's peer gathers).

torch supplies the plumbing only: ``torch.distributed._symmetric_memory`` allocates the memory and exchanges the
handles (one rendezvous per arena, over the process group's store); every byte that moves afterwards is moved by
this repo's kernels. Reference counterpart: DistributedDataParallel over NCCL (src/train.py:29-35, 300).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import SymmSegment, SymmTeam, check, lib

_ALIGN = 256


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class SymmArena:
    """``layout``: {name: nbytes}; every block is 256-byte aligned inside one symmetric allocation that ends with
    the kernels' control block. Collective: every rank of ``group`` must construct the same arena."""

    def __init__(self, layout: Dict[str, int], group=None, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 16:
            raise ValueError("symmetric arenas serve one NVLink domain: at most 16 ranks")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, nbytes in layout.items():
            self.offsets[name] = (off, nbytes)
            off = _round_up(off + nbytes, _ALIGN)
        self.ctrl_offset = off
        self.nbytes = off + lib().tt_symm_ctrl_bytes()
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        ptrs = list(self.hdl.buffer_ptrs)
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr(), "unexpected symmetric-memory handle"
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.team = SymmTeam()
        self.team.rank, self.team.world = self.rank, self.world
        for r in range(self.world):
            self.team.bufs[r] = ptrs[r]
        self.team.multicast = mc if mc else None
        self.team.ctrl_offset = self.ctrl_offset
        self.multicast = bool(mc)
        # nobody may signal into a control block that its owner has not zeroed yet
        dist.barrier(self.group)
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()

    def view(self, name: str, dtype: torch.dtype, shape: Sequence[int]) -> torch.Tensor:
        off, nbytes = self.offsets[name]
        t = self.buf[off:off + nbytes].view(dtype)
        return t.view(*shape)

    def offset(self, name: str) -> int:
        return self.offsets[name][0]

    def error_word(self) -> torch.Tensor:
        """int32 device scalar: non-zero after a cross-rank wait timed out (a peer never arrived)."""
        return self.buf[self.ctrl_offset + 8:self.ctrl_offset + 12].view(torch.int32)

    def check(self) -> None:
        self._err_host.copy_(self.error_word())
        torch.cuda.current_stream().synchronize()
        if int(self._err_host[0]) != 0:
            raise RuntimeError("a cross-GPU wait timed out (peer rank missing or stalled)")

    # ---- collectives -------------------------------------------------------------------------
    def allgather(self, blocks: Sequence[Tuple[torch.Tensor, str]], pre_barrier: bool = False) -> None:
        """blocks: up to four (source tensor, destination block name); rank r's source lands at byte
        offset(name) + r * source bytes of every rank's arena."""
        segs = (SymmSegment * 4)()
        for i, (src, name) in enumerate(blocks):
            assert src.is_cuda and src.is_contiguous()
            nb = src.numel() * src.element_size()
            assert nb % 16 == 0 and nb * self.world <= self.offsets[name][1], (name, nb)
            segs[i].src, segs[i].dst_offset, segs[i].nbytes = src.data_ptr(), self.offsets[name][0], nb
        check(lib().tt_symm_allgather(ctypes.byref(self.team), segs, len(blocks), int(pre_barrier),
                                      torch.cuda.current_stream().cuda_stream), "tt_symm_allgather")

    def barrier(self) -> None:
        check(lib().tt_symm_barrier(ctypes.byref(self.team), torch.cuda.current_stream().cuda_stream), "tt_symm_barrier")

    def dp_adamw_step(self, flat: str, grad: str, shadow: str, n: int, shadow_begin: int, m: torch.Tensor,
                      v: torch.Tensor, step_dev: torch.Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                      weight_decay: float = 0.01, table_shard=None) -> None:
        """``table_shard`` = (p, g, m, v) flat fp32 tensors of this rank's rows of a row-sharded ID table: updated in
        the same kernel, between its two barriers (gradient = g / world, g cleared)."""
        assert m.dtype == torch.float32 and v.dtype == torch.float32 and m.numel() == n // self.world == v.numel()
        sp = [None] * 4 if table_shard is None else [t.data_ptr() for t in table_shard]
        sn = 0 if table_shard is None else table_shard[0].numel()
        if table_shard is not None:
            assert all(t.dtype == torch.float32 and t.is_contiguous() and t.numel() == sn for t in table_shard)
        check(lib().tt_dp_adamw_step(ctypes.byref(self.team), self.offsets[flat][0], self.offsets[grad][0],
                                     self.offsets[shadow][0], n, shadow_begin, m.data_ptr(), v.data_ptr(), lr, betas[0],
                                     betas[1], eps, weight_decay, step_dev.data_ptr(), sp[0], sp[1], sp[2], sp[3], sn,
                                     torch.cuda.current_stream().cuda_stream), "tt_dp_adamw_step")


def available(group=None) -> bool:
    """True when this process group can build symmetric arenas (CUDA + NCCL group on one NVLink domain)."""
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:
        return False
    if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
        return False
    return dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 16
