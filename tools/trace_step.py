#!/usr/bin/env python
"""Kernel timeline of the graph-replayed c2 training step (CUPTI activity records via torch.profiler):
start offset, duration and stream of every kernel of one replay, plus the idle gaps on the main stream.
Not a bench: tracing perturbs timing slightly; use it for ordering / overlap / gap analysis."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from mrm_b200 import synthetic  # noqa: E402
from mrm_b200.engine import TwoTowerEngine  # noqa: E402
from mrm_b200.train import TrainStepRunner, make_dp_engine  # noqa: E402


def main():
    """Single process, or under torchrun (rank 0 prints the timeline of ITS GPU): the data-parallel step."""
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = int(os.environ.get("TT_TRACE_BATCH", 256))
    L, V = 200, 100_001
    cfg = synthetic.TwoTowerConfig(vocab_size=V, max_seq_len=L, dropout=0.1)
    eng, table = make_dp_engine(cfg, world)
    eng.load_state_dict(synthetic.make_state_dict(cfg, seed=0))
    runner = TrainStepRunner(eng, B, L, world_size=world, sharded_table=table)
    batch = synthetic.make_batch(cfg, B, seed=1 + rank, full_length=True, num_users=1_000_000)
    runner.load_batch({k: v.cuda() for k, v in batch.items()})
    for _ in range(5):
        runner.step_resident()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            runner.step_resident()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    path = os.path.join(tempfile.gettempdir(), "trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") == "kernel"]
    ks.sort(key=lambda e: e["ts"])
    n = len(ks) // 3
    ks = ks[n:2 * n]          # the middle replay
    t0 = ks[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in ks)
    print(f"# {n} kernels, span {end - t0:.1f} us")
    streams = sorted({e["args"].get("stream") for e in ks})
    print(f"# streams: {streams}")
    busy = 0.0
    last_end = t0
    for e in ks:
        s = e["args"].get("stream")
        gap = e["ts"] - last_end
        print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} s{streams.index(s)} gap={gap:6.1f} {e['name'][:60]}")
        last_end = max(last_end, e["ts"] + e["dur"])
    # union busy time
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ks)
    cur_s, cur_e = iv[0]
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f"# union busy {busy:.1f} us of {end - t0:.1f} us span; sum of durations {sum(e['dur'] for e in ks):.1f} us")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
