#!/usr/bin/env python
"""Run-to-run repeatability of the train step on one GPU (race hunting aid).

1. forward+backward from the SAME state, N times: gradients may differ only by fp32 atomic-add order
   (relative to the largest gradient element ~1e-6); anything larger points at a stream-ordering race.
2. three optimizer steps from the same state, several times, eager and graph-replayed: loss trajectories.

usage: python tools/determinism_check.py [--big] [--trials N]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mrm_b200 import synthetic  # noqa: E402
from mrm_b200.engine import TwoTowerEngine  # noqa: E402
from mrm_b200.train import TrainStepRunner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="c2 shape (B=256, L=200, V=100k) instead of the small one")
    ap.add_argument("--trials", type=int, default=20)
    ap.add_argument("--dropout", type=float, default=0.0)
    a = ap.parse_args()
    if a.big:
        cfg, B, L = synthetic.TwoTowerConfig(vocab_size=100001, max_seq_len=200, dropout=a.dropout), 256, 200
    else:
        cfg, B, L = synthetic.TwoTowerConfig(vocab_size=3001, max_seq_len=50, dropout=a.dropout), 32, 50
    sd = synthetic.make_state_dict(cfg, seed=7)
    batches = [{k: v.cuda() for k, v in synthetic.make_batch(cfg, B, seed=100 + s, num_users=20).items()}
               for s in range(3)]
    eng = TwoTowerEngine(cfg)
    eng.load_state_dict(sd)

    # ---- 1. gradient repeatability
    ref = None
    worst = 0.0
    for t in range(a.trials):
        eng.grad.zero_()
        eng.seed_dev.zero_()                # same dropout masks every trial
        eng.forward(batches[0], training=True)
        eng.backward()
        torch.cuda.synchronize()
        g = eng.grad.clone()
        if ref is None:
            ref = g
            continue
        d = (g - ref).abs()
        rel = (d.max() / ref.abs().max()).item()
        worst = max(worst, rel)
        if rel > 1e-4:
            i = int(d.argmax())
            where = "?"
            for n, v in eng.g.items():
                off = v.data_ptr() - eng.grad.data_ptr()
                if 0 <= (i * 4 - off) < v.numel() * 4:
                    where = n
            print(f"trial {t}: max |dgrad| / max|grad| = {rel:.3e} at flat index {i} ({where})")
    print(f"gradient repeatability over {a.trials} trials: worst relative-to-max difference {worst:.3e}")

    # ---- 2. loss trajectories
    for use_graph in (False, True):
        traj = []
        for t in range(6):
            e = TwoTowerEngine(cfg)
            e.load_state_dict(sd)
            r = TrainStepRunner(e, B, L, lr=1e-3, use_graph=use_graph)
            ls = []
            for s in range(4):
                r.load_batch(batches[s % 3])
                ls.append(r.step_resident().item())
            traj.append(ls)
        tt = torch.tensor(traj, dtype=torch.float64)
        print(f"graph={use_graph}: per-step loss spread over 6 runs (max-min): "
              f"{[f'{x:.2e}' for x in (tt.max(0).values - tt.min(0).values).tolist()]}  mean {tt.mean(0).tolist()}")


if __name__ == "__main__":
    main()
