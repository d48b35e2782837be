#!/usr/bin/env python
"""Real-NCCL check of the sharded retrieval (run under torchrun on >= 2 GPUs): the bounded protocol
(per-shard candidate lists + completeness bounds, merge + certificate) must return exactly what the
per-shard exact top-K protocol returns; both are timed (CUDA events, max over ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import mrm_b200  # noqa: F401
    from mrm_b200 import retrieval
    U, N, K = int(os.environ.get("USERS", 10_000)), int(os.environ.get("ITEMS", 1_000_000)), 100
    rows = (N + 1 + world - 1) // world
    first = rank * rows
    n_local = max(0, min(N + 1, first + rows) - first)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.nn.functional.normalize(torch.randn(n_local, 256, device=dev, generator=g), dim=1)
    if rank == 0:
        shard[0] = 0
    index = retrieval.CatalogIndex.from_shard(shard, first, N + 1, device=dev)
    gu = torch.Generator(device=dev).manual_seed(99)
    t = torch.randint(1, n_local, (U,), device=dev, generator=gu)
    users = torch.nn.functional.normalize(shard[t] + 3.3 / 16.0 * torch.randn(U, 256, device=dev, generator=gu), dim=1)
    dist.broadcast(users, 0)
    out = {}
    for name, bounded in (("bounded", True), ("per-shard exact", False)):
        for _ in range(2):
            i, s = retrieval.sharded_topk(users, index, K, bounded=bounded)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            i, s = retrieval.sharded_topk(users, index, K, bounded=bounded)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / 5], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = (i.clone(), s.clone(), ms.item())
    same = torch.equal(out["bounded"][0], out["per-shard exact"][0]) and torch.equal(out["bounded"][1], out["per-shard exact"][1])
    flag = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world={world} users={U} items={N}: bounded {out['bounded'][2]:.3f} ms/pass = {U / out['bounded'][2] * 1e3:.0f} users/s, "
              f"per-shard exact {out['per-shard exact'][2]:.3f} ms/pass; identical results on every rank: {bool(flag.item())}")
        print("DIST_RETRIEVAL_OK" if flag.item() else "DIST_RETRIEVAL_MISMATCH")
    # host-side anatomy of one metrics_from_embeddings call (what bench.py's e2e_users_per_s times): every stage
    # followed by a device synchronisation so wall-clock differences are attributable
    import time
    targets = (t + first).clone()
    dist.broadcast(targets, 0)
    hu, ht = users.cpu().pin_memory(), targets.cpu().pin_memory()
    kl = [10, 20, 50, 100]
    for _ in range(2):
        retrieval.metrics_from_embeddings(users, targets, index, kl)
    torch.cuda.synchronize()
    marks = []

    def mark(name):
        torch.cuda.synchronize()
        marks.append((name, time.perf_counter()))

    dist.barrier()
    mark("start")
    du, dt = hu.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
    mark("h2d")
    i, s, bad = retrieval.sharded_topk(du, index, K, defer_check=True)
    mark("sharded_topk")
    rec, nd = retrieval.rank_metrics(i, dt, kl)
    mark("rank_metrics")
    packed = torch.cat([rec.flatten(), nd.flatten(), bad.max().float().view(1)]).cpu()
    mark("readback")
    t0 = time.perf_counter()
    m = retrieval.metrics_from_embeddings(hu.to(dev, non_blocking=True), ht.to(dev, non_blocking=True), index, kl)
    torch.cuda.synchronize()
    whole = time.perf_counter() - t0
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):       # bench.py's e2e loop, verbatim
        m = retrieval.metrics_from_embeddings(hu.to(dev, non_blocking=True), ht.to(dev, non_blocking=True), index, kl)
    torch.cuda.synchronize()
    loop = (time.perf_counter() - t0) / 5
    if rank == 0:
        print(f"5-call loop: {1e3 * loop:.2f} ms per call")
        print("e2e anatomy (ms):", ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks[:-1], marks[1:])),
              f"| whole call {1e3 * whole:.2f}", f"recall@10 {m['Recall@10']:.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
