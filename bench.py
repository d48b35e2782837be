#!/usr/bin/env python
"""Benchmark of the two-tower hot path on B200 (one JSON line on rank 0).

Headline metric (BASELINE.json configs[1]): two-tower train samples/sec — one step = forward +
backward + InfoNCE + dense AdamW on batch 256/GPU, seq len 200, 100k-item ID table, dropout 0.1,
all histories full length, synthetic precomputed modality embeddings.
  value : device-resident inputs, the whole step replayed as one CUDA graph, CUDA-event timing.
  e2e   : through the public train_one_epoch-style call with HOST (pinned) batches: H2D of every
          step's batch and D2H of its loss inside the timed region.
  roofline     : the tcgen05 GEMM (dominant kernel): algorithmic FLOPs / event-timed duration of
                 the step's own GEMM launches, against MEASURED_PEAKS.json bf16 sustained.
  cpu_baseline : the oracle port of the reference step on this box's host cores (bounded sample).
  retrieval    : evaluate_metrics workload (BASELINE.json configs[2]): 10k users x 1M items,
                 top-100 + Recall/NDCG, catalog sharded over the N GPUs.
`--impl reference` times the reference's CPU implementation (oracle port) for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

C2 = dict(batch=256, seq_len=200, vocab=100_001)      # BASELINE.json configs[1]
C3 = dict(users=10_000, items=1_000_000, k=100, hist_len=50)  # BASELINE.json configs[2]
CPU_SAMPLE_BATCH = 64


def flops_per_sample_fwd(L, B_neg, D=256, FF=1024, NL=2):
    """SURVEY.md §8d algorithmic work per sample (all positions valid)."""
    return (L * NL * (2 * D * 3 * D + 2 * D * D + 4 * D * FF + 4 * ((L + 1) / 2) * D)
            + (2 * 304 * 256 + 2 * 256 * 256) + (2 * 512 * 512 + 2 * 512 * 256) + 2 * B_neg * D)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(s[2 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


def pin(batch):
    return {k: v.pin_memory() for k, v in batch.items()}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference step / retrieval on the host cores
# ------------------------------------------------------------------------------------------
def cpu_train_samples_per_s(steps, warmup, batch_size):
    from mrm_b200 import synthetic
    from oracle import two_tower_oracle as oracle
    cfg = synthetic.TwoTowerConfig(vocab_size=C2["vocab"], max_seq_len=C2["seq_len"], dropout=0.0)
    sd = synthetic.make_state_dict(cfg, seed=0)
    batch = synthetic.make_batch(cfg, batch_size, seed=1, full_length=True, num_users=1_000_000)
    p = {k: v.clone() for k, v in sd.items()}
    m = {k: torch.zeros_like(v) for k, v in p.items() if v.is_floating_point()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items() if v.is_floating_point()}
    times = []
    for t in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        _, _, _, _, grads, _ = oracle.loss_and_grads(p, batch, cfg.temperature, cfg.num_heads)
        for k, g in grads.items():
            p[k], m[k], v2[k] = oracle.adamw_step(p[k], g, m[k], v2[k], t)
        dt = time.perf_counter() - t0
        if t > warmup:
            times.append(dt)
    total = sum(times)
    return batch_size * len(times) / total, total / len(times) * 1e3


def cpu_retrieval_users_per_s(num_users=128):
    from mrm_b200 import synthetic
    from oracle import two_tower_oracle as oracle
    table = synthetic.make_catalog(C3["items"], 256, seed=2)
    users, targets = synthetic.make_queries(table, num_users, seed=3)
    oracle.calculate_metrics_global(users[:64], table, targets[:64], [10, 20, 50, 100])
    t0 = time.perf_counter()
    oracle.calculate_metrics_global(users, table, targets, [10, 20, 50, 100])
    return num_users / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sps, ms = cpu_train_samples_per_s(args.steps, args.warmup, CPU_SAMPLE_BATCH)
    sample = (f"oracle port of the reference step (fp32 torch CPU: fwd+bwd+InfoNCE+AdamW), batch {CPU_SAMPLE_BATCH} "
              f"of the c2 workload (L={C2['seq_len']}, V={C2['vocab']}) per step")
    line = {
        "impl": "reference", "metric": "two-tower train samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: two-tower train step, batch 256/GPU, seq_len 200, 100k items",
                   "reference_step_batch": CPU_SAMPLE_BATCH},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def time_gemm_roofline(eng, static_batch, peaks, iters=5):
    """Event-time every tcgen05 GEMM launch of the training step (the step's own operands, fused epilogues
    and launch arguments, on the launching stream). The step is issued eagerly on ONE stream behind a
    ~15 ms device-side sleep, so the whole launch sequence is already queued when the GPU starts it and
    the events bracket kernel execution, not host launch latency. Per launch the roofline time is
    max(flops / bf16 peak, algorithmic bytes / HBM peak); d=256 makes the encoder GEMMs HBM-bound.
    Returns (all launches, the encoder-sized launches (>= 4096 rows or a >= 4096-deep reduction), n)."""
    peak_f = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")) * 1e12
    peak_b = peaks["hbm_gbs"] * 1e9
    zero = dict(flops=0.0, bytes=0.0, secs=0.0, roof_secs=0.0, flop_secs=0.0, byte_secs=0.0, n=0.0)
    acc, big = dict(zero), dict(zero)
    n_launch = 0
    eng.serialize = True
    for it in range(iters + 1):
        eng.gemm_log = []
        torch.cuda._sleep(30_000_000)
        eng.forward(static_batch, training=True)
        eng.backward()
        torch.cuda.synchronize()
        if it > 0:       # first pass warms the eager path
            for e0, e1, f, b, (M, N, K) in eng.gemm_log:
                for d in ((acc, big) if max(M, K) >= 4096 else (acc,)):
                    d["flops"] += f
                    d["bytes"] += b
                    d["secs"] += e0.elapsed_time(e1) * 1e-3
                    d["roof_secs"] += max(f / peak_f, b / peak_b)
                    d["flop_secs"] += f / peak_f
                    d["byte_secs"] += b / peak_b
                    d["n"] += 1
        n_launch = len(eng.gemm_log)
        eng.gemm_log = None
        eng.grad.zero_()
    eng.serialize = False
    return {k: v / iters for k, v in acc.items()}, {k: v / iters for k, v in big.items()}, n_launch


def run_ours(args, rank, world, local_rank):
    import mrm_b200
    from mrm_b200 import _lib, retrieval, synthetic
    from mrm_b200.engine import TwoTowerEngine
    from mrm_b200.train import TrainStepRunner

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks, peak_src = load_peaks()
    B, L, V = C2["batch"], C2["seq_len"], C2["vocab"]
    cfg = synthetic.TwoTowerConfig(vocab_size=V, max_seq_len=L, dropout=0.1)
    eng = TwoTowerEngine(cfg, dev)
    eng.load_state_dict(synthetic.make_state_dict(cfg, seed=0))
    runner = TrainStepRunner(eng, B, L, world_size=world)
    host_batches = [pin(synthetic.make_batch(cfg, B, seed=100 + rank * 17 + i, full_length=True,
                                             num_users=1_000_000)) for i in range(4)]
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host_batches[0].items() if k in runner.static)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (includes graph capture) -------------------------------------------------
    runner.load_batch(host_batches[0])
    for _ in range(max(args.warmup, 3)):
        runner.step_resident()
    torch.cuda.synchronize()

    # ---- value: device-resident inputs ----------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = _lib.launch_count
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        runner.step_resident()
        marks[i].record()
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    per_step = sorted([e0.elapsed_time(marks[0])] + [marks[i - 1].elapsed_time(marks[i]) for i in range(1, args.steps)])
    pct = {f"p{q}": per_step[min(len(per_step) - 1, int(q / 100.0 * len(per_step)))] for q in (10, 50, 90)}
    value = world * B * 1e3 / ms_step
    launches = runner.kernels_per_step * args.steps + (_lib.launch_count - l0)

    # ---- e2e: host batches, H2D + step + D2H of the loss every step -----------------------
    # Every step's batch crosses PCIe from pinned memory inside the timed region; the copy of batch i+1 is
    # issued while step i computes (double-buffered staging, TrainStepRunner.stage_batch), the loss of every
    # step is read back to the host before the next one is launched.
    for i in range(3):
        runner.step_from_host(host_batches[i % 4])
    packed = [runner.pack_host(b) for b in host_batches]      # one pinned buffer per batch: one H2D copy
    runner.stage_batch(packed[0])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        runner.step_from_host(None, prefetch=packed[(i + 1) % 4])   # K copies for K steps
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / e2e_s.item()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- forward + backward only (no optimizer step) -------------------------
    fb = None
    if world == 1:
        eng2 = TwoTowerEngine(cfg, dev)
        eng2.load_state_dict(synthetic.make_state_dict(cfg, seed=0))
        r2 = TrainStepRunner(eng2, B, L, with_optimizer=False)
        r2.load_batch(host_batches[0])
        for _ in range(4):
            r2.step_resident()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(20):
            r2.step_resident()
        f1.record()
        torch.cuda.synchronize()
        fb_ms = f0.elapsed_time(f1) / 20
        fb = {"ms_per_step": fb_ms, "samples_per_s": B * 1e3 / fb_ms,
              "what": "forward + backward + InfoNCE, gradient buffer cleared instead of the AdamW step"}
        del r2, eng2

    # ---- roofline of the dominant kernel (tcgen05 GEMM) -----------------------------------
    g_all, g, gemm_launches = time_gemm_roofline(eng, runner.static, peaks)
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    hbm_bound = g["byte_secs"] >= g["flop_secs"]
    if hbm_bound:
        roof = {"bound": "hbm", "achieved": g["bytes"] / g["secs"] / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": g["flops"] / g["secs"] / 1e12, "peak": peak_tf, "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gemm_dram_traffic.json")
    if os.path.exists(tpath):      # written by tools/ncu_summary.py from the committed `ncu --set full` capture
        with open(tpath) as f:
            traffic = json.load(f).get("encoder_gemm_dram_bytes_per_launch")
    roof.update({
        "kernel": "tt::gemm_bf16_kernel (tcgen05), the encoder-sized launches (51200-row operands)",
        "traffic": traffic, "peak_source": f"{peak_src}",
        "launches_per_step": g["n"], "avg_launch_us": g["secs"] * 1e6 / max(g["n"], 1),
        "algorithmic_bytes_per_launch": g["bytes"] / max(g["n"], 1),
        "frac_of_per_launch_roofline": g["roof_secs"] / g["secs"],
        "tensor_tflops": g["flops"] / g["secs"] / 1e12, "tensor_frac": g["flops"] / g["secs"] / 1e12 / peak_tf,
        "us_per_step": g["secs"] * 1e6, "share_of_step": g["secs"] * 1e3 / ms_step,
        "flops_per_step": g["flops"], "bytes_per_step": g["bytes"],
        "all_gemm_launches": {"launches_per_step": gemm_launches, "us_per_step": g_all["secs"] * 1e6,
                              "flops_per_step": g_all["flops"], "bytes_per_step": g_all["bytes"],
                              "hbm_gbs": g_all["bytes"] / g_all["secs"] / 1e9},
        "how": "CUDA events around each tt_gemm_bf16 launch of the step issued on one stream behind a device-side "
               "sleep (launch queue full, so events bracket execution only; same operands and fused epilogues as "
               "the timed step); algorithmic bytes = operands + outputs + residual/gate per launch"})
    step_flops = 3.0 * flops_per_sample_fwd(L, B) * B

    # ---- retrieval (configs[2]) -----------------------------------------------------------
    retr = None if args.skip_retrieval else bench_retrieval(eng, rank, world, dev, peaks)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        sps, _ = cpu_train_samples_per_s(2, 1, CPU_SAMPLE_BATCH)
        cpu = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"oracle port of the reference step on batch {CPU_SAMPLE_BATCH} of the c2 workload, "
                         f"1 warm-up + 2 timed steps",
               "retrieval_users_per_s": cpu_retrieval_users_per_s(128),
               "retrieval_sample": "oracle calculate_metrics_global, 128 users x 1M items, top-100"}

    if rank == 0:
        line = {
            "metric": "two-tower train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "c2: two-tower train step (fwd+bwd+InfoNCE+dense AdamW), batch 256/GPU, "
                                   "seq_len 200 (all positions valid), 100k-item ID table, dropout 0.1",
                       "global_batch": world * B, "seq_len": L, "vocab_size": V,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "negatives": ("all-gathered across ranks (NCCL all-gather of embeddings + row log-sum-exps)"
                                     if world > 1 else "in-batch"),
                       "last_layer": "exact single-row form (only out[b, len-1] of the last encoder layer is ever "
                                     "read: K/V for all positions, query/out_proj/FFN for one row per sequence)",
                       "l2": "per-step working set (~1.2 GB activations + 410 MB optimizer state) exceeds the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "ms_per_step_percentiles": pct,
            "fwd_bwd_only": fb,
            "gpu_launches": launches,
            "kernels_per_step": runner.kernels_per_step,
            "step_tflops": step_flops / (ms_step * 1e-3) / 1e12,
            "step_tflops_note": "reference-algorithm FLOPs of the step (SURVEY.md §8d formula, every layer on every "
                                "position) per second; the executed GEMM FLOPs are roofline.flops_per_step",
            "roofline": roof,
            "clocks": sampler.summary(),
            "retrieval": retr,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)


def bench_retrieval(eng, rank, world, dev, peaks):
    """10k users x 1M items, top-100 + Recall/NDCG; the catalog is sharded over the ranks."""
    from mrm_b200 import retrieval
    U, N, K = C3["users"], C3["items"], C3["k"]
    rows = (N + 1 + world - 1) // world
    first = rank * rows
    n_local = max(0, min(N + 1, first + rows) - first)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = torch.nn.functional.normalize(torch.randn(n_local, 256, device=dev, generator=g), dim=1)
    if rank == 0:
        shard[0] = 0
    index = retrieval.CatalogIndex(shard, device=dev)
    index.item_base, index.vocab_size = first, N + 1
    gu = torch.Generator(device=dev).manual_seed(99)
    targets = torch.randint(1, n_local, (U,), device=dev, generator=gu)
    users = torch.nn.functional.normalize(shard[targets] + 3.3 / 16.0 * torch.randn(U, 256, device=dev, generator=gu),
                                          dim=1)
    targets = targets + first
    if world > 1:
        dist.broadcast(users, 0)
        dist.broadcast(targets, 0)
    host_users = users.cpu().pin_memory()
    host_targets = targets.cpu().pin_memory()
    kl = [10, 20, 50, 100]
    for _ in range(2):
        m = retrieval.metrics_from_embeddings(users, targets, index, kl)
    torch.cuda.synchronize()
    iters = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    nfb = 0
    for _ in range(iters):
        if world > 1:    # the whole sharded pass: per-shard candidates, all-gather, merge + certificate
            idx, score = retrieval.sharded_topk(users, index, K)
        else:
            idx, score, nfb = retrieval.retrieve_topk(users, index, K, exact_fallback=False)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    # e2e: host user embeddings -> device, retrieval, merge, metrics -> host
    t0 = time.perf_counter()
    for _ in range(iters):
        m = retrieval.metrics_from_embeddings(host_users.to(dev, non_blocking=True),
                                              host_targets.to(dev, non_blocking=True), index, kl)
    torch.cuda.synchronize()
    e2e = torch.tensor([(time.perf_counter() - t0) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    # ---- the same pass with the user embeddings produced by the CUDA user tower (eval mode) from HOST
    # histories (L = 50, the reference default): H2D ids -> SASRec forward -> retrieval -> merge -> metrics
    from mrm_b200 import synthetic
    from mrm_b200.engine import TwoTowerEngine
    Lh, UB = C3["hist_len"], 2000
    tcfg = synthetic.TwoTowerConfig(vocab_size=4096, max_seq_len=Lh, dropout=0.0)   # small ID table: ids are synthetic
    teng = TwoTowerEngine(tcfg, dev)
    teng.load_state_dict(synthetic.make_state_dict(tcfg, seed=0))
    hb = synthetic.make_batch(tcfg, U, seed=7, full_length=True)
    h_ids, h_mask = hb["history_ids"].pin_memory(), hb["history_mask"].pin_memory()
    h_g, h_c = hb["user_gender"].pin_memory(), hb["user_country"].pin_memory()
    uemb = torch.empty(U, 256, device=dev)

    def tower_pass():
        ws = teng.workspace(UB, Lh)
        if not teng.shadow_valid:
            teng.refresh_shadow()
        for s0 in range(0, U, UB):
            sl = slice(s0, s0 + UB)
            u = teng.user_forward(ws, h_ids[sl].to(dev, non_blocking=True), h_mask[sl].to(dev, non_blocking=True),
                                  h_g[sl].to(dev, non_blocking=True), h_c[sl].to(dev, non_blocking=True), training=False)
            uemb[sl].copy_(u)
        return retrieval.metrics_from_embeddings(uemb, host_targets.to(dev, non_blocking=True), index, kl)

    for _ in range(2):
        tower_pass()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        tower_pass()
    torch.cuda.synchronize()
    e2e_tower = torch.tensor([(time.perf_counter() - t0) / iters], device=dev)
    if world > 1:
        dist.all_reduce(e2e_tower, op=dist.ReduceOp.MAX)

    # single-user recommendation latency (src/inference.py:283-306): one user, 50 history items excluded, top-10
    rec_ms = None
    if world == 1:
        from mrm_b200 import inference
        g1 = torch.Generator(device=dev).manual_seed(9)
        hist = torch.randint(1, N, (1, 50), device=dev, generator=g1)
        one = users[:1].contiguous()
        for _ in range(3):
            inference.recommend_topk(one, index, hist, 10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            ridx, _ = inference.recommend_topk(one, index, hist, 10)
            ridx.cpu()                      # the caller reads the ids: one sync per query
        rec_ms = (time.perf_counter() - t0) / 20 * 1e3

    flops = 2.0 * U * (N + 1) * 256 / world
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    return {"metric": "top-100 retrieval users/sec @1M items", "recommend_1user_ms": rec_ms, "users_per_s": U / (ms.item() * 1e-3),
            "ms_per_pass": ms.item(), "e2e_users_per_s": U / e2e.item(),
            "e2e_users_per_s_incl_user_tower": U / e2e_tower.item(),
            "scoring_tflops_per_gpu": flops / (ms.item() * 1e-3) / 1e12,
            "roofline_frac_tensor": flops / (ms.item() * 1e-3) / 1e12 / peak_tf,
            "config": {"workload": f"c3: {U} users x {N} items, top-{K}, Recall/NDCG@10/20/50/100, catalog sharded "
                                   f"over {world} GPU(s)",
                       "users_per_s": "device-resident user embeddings: scoring + top-K + exact re-score (events)"
                                      + ("; sharded: per-shard pass, all-gather, merge (retrieval.sharded_topk: candidate "
                                         "lists + bounds + certificate from 4 shards, per-shard exact top-K below)"
                                         if world > 1 else ""),
                       "e2e_users_per_s": "host user embeddings -> device, retrieval, cross-shard merge, metrics -> host",
                       "e2e_users_per_s_incl_user_tower": f"host histories (L={Lh}) -> CUDA user tower (eval) -> same",
                       "kprime": 256},
            "recall_at_10": m["Recall@10"], "fallback_users": int(nfb)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-retrieval", action="store_true", help="skip the c3 retrieval section (A/B timing runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
