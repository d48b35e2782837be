#!/usr/bin/env python
"""Tile / split-K sweep of tt_gemm_bf16 on the encoder's weight-gradient and forward shapes.
Operands rotate over 3 buffer sets (> L2 in total) so a launch does not find its inputs cached."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mrm_b200 import ops  # noqa: E402

T, D, FF = 51200, 256, 1024
dev = "cuda"
bf = dict(device=dev, dtype=torch.bfloat16)
NSET = 3
xs = [torch.randn(T, FF, **bf) for _ in range(NSET)]
ys = [torch.randn(T, FF, **bf) for _ in range(NSET)]
w = torch.randn(FF, FF, **bf) * 0.05
bias = torch.randn(FF, device=dev)
g32 = torch.zeros(FF, FF, device=dev)
o16 = [torch.empty(T, FF, **bf) for _ in range(NSET)]
o32 = [torch.empty(T, D, device=dev) for _ in range(NSET)]


def t(fn, iters=9):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % NSET)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


print("== weight gradients  dW[M,N] += A[K,M]^T B[K,N], K = 51200")
for M, N in [(512, 256), (256, 1024), (1024, 256), (256, 256), (768, 256)]:
    for bn in (64, 128, 256):
        if bn > N:
            continue
        row = []
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
        for ks in sorted({max(1, 148 // tiles), max(1, 296 // tiles), max(1, 74 // tiles)}):
            us = t(lambda i: ops.gemm(xs[i][:, :M], ys[i][:, :N], a_mn=True, b_mn=True, out_f32=g32[:M, :N],
                                      accumulate=True, block_n=bn, k_splits=ks))
            row.append(f"ks={ks:3d}: {us:6.1f} us")
        print(f"M={M:5d} N={N:5d} bn={bn:3d} tiles={tiles:3d}  " + "  ".join(row))

print("== forward / dgrad shapes, M = 51200")
cases = [
    ("qkv bias bf16 N=768 K=256", 768, 256, lambda i, bn: ops.gemm(xs[i][:, :256], w[:768, :256], bias=bias, out_bf16=o16[i][:, :768], block_n=bn)),
    ("kv  bias bf16 N=512 K=256", 512, 256, lambda i, bn: ops.gemm(xs[i][:, :256], w[:512, :256], bias=bias, out_bf16=o16[i][:, :512], block_n=bn)),
    ("ffn1 bias relu drop N=1024", 1024, 256, lambda i, bn: ops.gemm(xs[i][:, :256], w[:, :256], bias=bias, relu=True, drop_p=0.1, drop_site=2, out_bf16=o16[i], block_n=bn)),
    ("dgrad f32 N=256 K=1024   ", 256, 1024, lambda i, bn: ops.gemm(xs[i], w[:, :256], b_mn=True, out_f32=o32[i], block_n=bn)),
    ("dgrad f32 N=256 K=768    ", 256, 768, lambda i, bn: ops.gemm(xs[i][:, :768], w[:768, :256], b_mn=True, out_f32=o32[i], block_n=bn)),
    ("dgrad bf16 N=256 K=256   ", 256, 256, lambda i, bn: ops.gemm(xs[i][:, :256], w[:256, :256], b_mn=True, out_bf16=o16[i][:, :256], block_n=bn)),
]
for name, N, K, fn in cases:
    row = []
    for bn in (64, 128, 256):
        us = t(lambda i: fn(i, bn))
        row.append(f"bn={bn:3d}: {us:6.1f} us")
    print(f"{name}  " + "  ".join(row))

print("== B-row (M = 256) shapes: latency-bound, more CTAs vs wider tiles")
xb = torch.randn(256, FF, **bf)
ob16 = torch.empty(256, FF, **bf)
ob32 = torch.empty(256, FF, device=dev)
tiny = [
    ("256x256x256  bias f32      ", lambda i, bn: ops.gemm(xb[:, :256], w[:256, :256], bias=bias, out_f32=ob32[:, :256], block_n=bn)),
    ("256x1024x256 bias relu bf16", lambda i, bn: ops.gemm(xb[:, :256], w[:, :256], bias=bias, relu=True, out_bf16=ob16, block_n=bn)),
    ("256x256x1024 bias f32      ", lambda i, bn: ops.gemm(xb, w[:256], bias=bias, out_f32=ob32[:, :256], block_n=bn)),
    ("256x512x512  bias f32      ", lambda i, bn: ops.gemm(xb[:, :512], w[:512, :512], bias=bias, out_f32=ob32[:, :512], block_n=bn)),
    ("256x256x256  B^T f32       ", lambda i, bn: ops.gemm(xb[:, :256], w[:256, :256], b_mn=True, out_f32=ob32[:, :256], block_n=bn)),
    ("256x256x1024 B^T f32       ", lambda i, bn: ops.gemm(xb, w[:, :256], b_mn=True, out_f32=ob32[:, :256], block_n=bn)),
    ("256x1024x256 B^T bf16      ", lambda i, bn: ops.gemm(xb[:, :256], w[:256], b_mn=True, out_bf16=ob16, block_n=bn)),
]
for name, fn in tiny:
    row = []
    for bn in (64, 128, 256):
        us = t(lambda i: fn(i, bn), iters=20)
        row.append(f"bn={bn:3d}: {us:6.2f} us")
    print(f"{name}  " + "  ".join(row))
