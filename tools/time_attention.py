"""Times tt_attn_causal_fwd / tt_attn_causal_bwd alone at the c2 shape (B=256, L=200, H=4) with CUDA events.
TT_ATTN_BWD=legacy / TT_ATTN_FWD=legacy select the one-CTA-per-item kernels for an A/B on the same box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mrm_b200 import ops  # noqa: E402


def main():
    B, L, H = int(os.environ.get("B", 256)), int(os.environ.get("L", 200)), 4
    drop = float(os.environ.get("DROP", 0.1))
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B * L, 3 * H * 64, generator=g) * 1.0).cuda().bfloat16()
    dctx = torch.randn(B * L, H * 64, generator=g).cuda().bfloat16()
    ctx = torch.empty((B * L, H * 64), device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, L, device="cuda")
    dqkv = torch.empty((B * L, 3 * H * 64), device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    for mode_f in ("legacy", "persistent"):
        os.environ["TT_ATTN_FWD"] = mode_f
        t = timed(lambda: ops.attn_fwd(qkv, ctx, lse, B, L, H, drop_p=drop, drop_seed=5, drop_site=1))
        print(f"attn_fwd  {mode_f:10s} B={B} L={L}: {t:7.1f} us")
    for mode_b in ("legacy", "persistent"):
        os.environ["TT_ATTN_BWD"] = mode_b
        t = timed(lambda: ops.attn_bwd(qkv, ctx, dctx, lse, dqkv, B, L, H, drop_p=drop, drop_seed=5, drop_site=1))
        print(f"attn_bwd  {mode_b:10s} B={B} L={L}: {t:7.1f} us")


if __name__ == "__main__":
    main()
