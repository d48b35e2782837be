"""Thin tensor-level wrappers over the C ABI (pointer + size marshalling only)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import GemmArgs, check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("tt ops take CUDA tensors only (no CPU fallback)")


def gemm(A: torch.Tensor, B: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, relu: bool = False,
         drop_p: float = 0.0, drop_seed: int = 0, drop_site: int = 0,
         gate: Optional[torch.Tensor] = None, gate_scale: float = 1.0,
         residual: Optional[torch.Tensor] = None,
         out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
         accumulate: bool = False, k_splits: int = 0, block_n: int = 0,
         M: Optional[int] = None, N: Optional[int] = None, K: Optional[int] = None) -> None:
    """C[M,N] = epi(alpha * A @ B^T). See ``tt_gemm_bf16`` in include/tt_b200.h.

    a_mn=False: A is [M,K]; a_mn=True: A is stored [K,M]. Same for B with N.
    """
    _require_cuda(A, B, bias, gate, residual, out_f32, out_bf16)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1
    if a_mn:
        k_a, m_a = A.shape
    else:
        m_a, k_a = A.shape
    if b_mn:
        k_b, n_b = B.shape
    else:
        n_b, k_b = B.shape
    M = m_a if M is None else M
    N = n_b if N is None else N
    K = min(k_a, k_b) if K is None else K
    args = GemmArgs()
    args.A, args.B = A.data_ptr(), B.data_ptr()
    args.lda, args.ldb = A.stride(0), B.stride(0)
    args.a_mn, args.b_mn = int(a_mn), int(b_mn)
    args.M, args.N, args.K = M, N, K
    args.alpha = alpha
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() >= N
    args.bias = _ptr(bias)
    args.relu = int(relu)
    args.drop_p, args.drop_seed, args.drop_site = drop_p, drop_seed, drop_site
    if gate is not None:
        assert gate.dtype == torch.bfloat16 and gate.stride(1) == 1
        args.ld_gate = gate.stride(0)
    args.gate = _ptr(gate)
    args.gate_scale = gate_scale
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
        args.ld_res = residual.stride(0)
    args.residual = _ptr(residual)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.stride(1) == 1
        args.ld_f32 = out_f32.stride(0)
    args.out_f32 = _ptr(out_f32)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.stride(1) == 1
        args.ld_bf16 = out_bf16.stride(0)
    args.out_bf16 = _ptr(out_bf16)
    args.accumulate = int(accumulate)
    args.k_splits = k_splits
    args.block_n = block_n
    check(lib().tt_gemm_bf16(ctypes.byref(args), _stream()), "tt_gemm_bf16")


def attn_fwd(qkv: torch.Tensor, ctx: torch.Tensor, lse: Optional[torch.Tensor], B: int, L: int, H: int,
             drop_p: float = 0.0, drop_seed: int = 0, drop_site: int = 0) -> None:
    """Causal self-attention forward (tt_attn_causal_fwd). qkv bf16 [B*L, 3*H*64] -> ctx bf16 [B*L, H*64]."""
    _require_cuda(qkv, ctx, lse)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * L, 3 * H * 64)
    assert ctx.dtype == torch.bfloat16 and ctx.is_contiguous() and ctx.shape == (B * L, H * 64)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * L
    check(lib().tt_attn_causal_fwd(qkv.data_ptr(), ctx.data_ptr(), _ptr(lse), B, L, H, drop_p,
                                   drop_seed, drop_site, _stream()), "tt_attn_causal_fwd")


def attn_bwd(qkv: torch.Tensor, ctx: torch.Tensor, dctx: torch.Tensor, lse: torch.Tensor,
             dqkv: torch.Tensor, B: int, L: int, H: int, drop_p: float = 0.0, drop_seed: int = 0,
             drop_site: int = 0) -> None:
    """Causal self-attention backward (tt_attn_causal_bwd): dqkv bf16 [B*L, 3*H*64]."""
    _require_cuda(qkv, ctx, dctx, lse, dqkv)
    for t in (qkv, ctx, dctx, dqkv):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
    assert lse.dtype == torch.float32 and lse.is_contiguous()
    check(lib().tt_attn_causal_bwd(qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(),
                                   dqkv.data_ptr(), B, L, H, drop_p, drop_seed, drop_site, _stream()),
          "tt_attn_causal_bwd")
