#!/usr/bin/env python
"""Per-launch CUDA-event timing of every tt_gemm_bf16 call of one eager training step (c2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mrm_b200 import ops, synthetic  # noqa: E402
from mrm_b200.engine import TwoTowerEngine  # noqa: E402

cfg = synthetic.TwoTowerConfig(vocab_size=100_001, max_seq_len=200, dropout=0.1)
eng = TwoTowerEngine(cfg)
eng.load_state_dict(synthetic.make_state_dict(cfg, seed=0))
batch = {k: v.cuda() for k, v in synthetic.make_batch(cfg, 256, seed=1, full_length=True, num_users=10**6).items()}
for _ in range(2):
    eng.train_step(batch)
descr = []
orig = ops.gemm


def spy(A, B, **kw):
    a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    flags = [k for k in ("bias", "relu", "gate", "residual", "accumulate") if kw.get(k) is not None and kw.get(k) is not False]
    if kw.get("drop_p", 0) > 0:
        flags.append("drop")
    flags.append("f32" if kw.get("out_f32") is not None else "bf16")
    descr.append((M, N, K, ("A^T " if a_mn else "") + ("B^T " if b_mn else "") + "+".join(flags)))
    return orig(A, B, **kw)


ops.gemm = spy
eng.serialize = True            # one stream: per-launch times without overlap
for it in range(4):
    descr.clear()
    eng.gemm_log = []
    torch.cuda._sleep(30_000_000)   # ~15 ms head start: the whole step is queued before the GPU starts it
    eng.forward(batch, training=True)
    eng.backward()
    torch.cuda.synchronize()
    if it == 3:
        total = 0.0
        for i, ((e0, e1, f, nb, _), d) in enumerate(zip(eng.gemm_log, descr)):
            us = e0.elapsed_time(e1) * 1e3
            total += us
            print(f"{i:2d} M={d[0]:6d} N={d[1]:5d} K={d[2]:6d} {us:8.1f} us {f / us / 1e6:8.1f} TF/s {nb / us / 1e3:8.1f} GB/s  {d[3]}")
        print(f"total {total:.1f} us")
    eng.gemm_log = None
    eng.grad.zero_()
