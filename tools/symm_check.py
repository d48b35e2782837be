#!/usr/bin/env python
"""Real-NVLink check of the symmetric-arena kernels (run under torchrun on >= 2 GPUs):
  1. tt_symm_allgather of three ragged blocks == torch.distributed.all_gather;
  2. tt_dp_adamw_step == mean-all-reduce of the gradients + the single-GPU tt_adamw_step on every rank,
     parameters and bf16 shadow identical on all ranks afterwards, over several steps;
  3. timing of tt_dp_adamw_step at the c2 flat-buffer size (27.8 M fp32) against the NCCL sequence it replaces
     (reduce_scatter AVG + sharded AdamW + all_gather + shadow cast)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from mrm_b200 import ops  # noqa: E402
from mrm_b200.symm import SymmArena  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # ---- 1. all-gather
    g = torch.Generator(device=dev).manual_seed(10 + rank)
    a = torch.randn(96, 256, device=dev, generator=g).to(torch.bfloat16)
    b = torch.randint(0, 1 << 40, (96,), device=dev, generator=g)
    c = torch.randn(100, device=dev, generator=g)
    arena = SymmArena({"a": world * a.numel() * 2, "b": world * b.numel() * 8, "c": world * c.numel() * 4}, device=dev)
    if rank == 0:
        print(f"arena: {arena.nbytes} B, multicast {arena.multicast}")
    for it in range(3):
        arena.allgather([(a, "a"), (b, "b"), (c, "c")], pre_barrier=(it > 0))
        torch.cuda.synchronize()
        for t, name, dt in ((a, "a", torch.bfloat16), (b, "b", torch.long), (c, "c", torch.float32)):
            ref = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(ref, t)
            got = arena.view(name, dt, (world,) + tuple(t.shape))
            ok &= bool(torch.equal(got, torch.stack(ref)))
        a += 1
        c += 1
    arena.check()
    if rank == 0:
        print("allgather", "ok" if ok else "MISMATCH")
    # ---- 2. fused reduce-scatter + AdamW + all-gather
    for n, dense_begin in ((4096 * world, 1024), (27_793_408 // (2048) * 2048, 25_658_368 // 4 * 4)):
        n = n // (4 * world) * (4 * world)
        layout = {"flat": n * 4, "grad": n * 4, "shadow": (n - dense_begin) * 2}
        ar = SymmArena(layout, device=dev)
        flat, grad = ar.view("flat", torch.float32, (n,)), ar.view("grad", torch.float32, (n,))
        shadow = ar.view("shadow", torch.bfloat16, (n - dense_begin,))
        g0 = torch.Generator(device=dev).manual_seed(5)
        flat.copy_(torch.randn(n, device=dev, generator=g0))
        ref_p = flat.clone()
        ref_m, ref_v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        ref_shadow = torch.zeros(n - dense_begin, device=dev, dtype=torch.bfloat16)
        m, v = torch.zeros(n // world, device=dev), torch.zeros(n // world, device=dev)
        step = torch.zeros((), device=dev, dtype=torch.long)
        # a local table shard updated inside the same kernel: gradient SUM over ranks sits in sg, scaled by 1 / world
        sn = 8192
        gs = torch.Generator(device=dev).manual_seed(77 + rank)
        shard = [torch.randn(sn, device=dev, generator=gs), torch.zeros(sn, device=dev), torch.zeros(sn, device=dev),
                 torch.zeros(sn, device=dev)]
        ref_shard = [t.clone() for t in shard]
        for it in range(3):
            gr = torch.Generator(device=dev).manual_seed(100 * it + rank)
            grad.copy_(torch.randn(n, device=dev, generator=gr) * (0.1 if it else 1.0))
            gsum = grad.clone()
            dist.all_reduce(gsum, op=dist.ReduceOp.SUM)
            step += 1
            sg = torch.randn(sn, device=dev, generator=gs)
            sg[::3] = 0
            shard[1].copy_(sg)
            ar.dp_adamw_step("flat", "grad", "shadow", n, dense_begin, m, v, step, 1e-3, table_shard=shard)
            ops.adamw_step(ref_shard[0], sg.clone(), ref_shard[2], ref_shard[3], step, 1e-3, shadow=None, zero_grad=False,
                           grad_scale=1.0 / world)
            gmean = gsum / world
            ops.adamw_step(ref_p, gmean, ref_m, ref_v, step, 1e-3, shadow=ref_shadow, shadow_begin=dense_begin,
                           shadow_end=n, zero_grad=False)
            torch.cuda.synchronize()
            # the switch's reduction order is its own: compare with one ulp of slack on the gradient sum
            d = (flat - ref_p).abs().max().item()
            ds = (shadow.float() - ref_shadow.float()).abs().max().item()
            allp = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(allp, flat)
            same = all(torch.equal(allp[0], x) for x in allp[1:])
            alls = [torch.empty_like(shadow) for _ in range(world)]
            dist.all_gather(alls, shadow)
            same &= all(torch.equal(alls[0], x) for x in alls[1:])
            same &= bool(torch.equal(shadow, flat[dense_begin:].to(torch.bfloat16)))
            if rank == 0:
                print(f"n={n} step {it}: max |p - ref| {d:.3e}, shadow {ds:.3e}, replicas identical {same}")
            dsh = (shard[0] - ref_shard[0]).abs().max().item()
            ok &= d < 2e-6 and same and dsh == 0.0 and shard[1].abs().max().item() == 0.0
        ar.check()
        if n > 1_000_000:
            def timed(fn, iters=20):
                for _ in range(3):
                    fn()
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return t.item()

            ms_fused = timed(lambda: ar.dp_adamw_step("flat", "grad", "shadow", n, dense_begin, m, v, step, 1e-3))
            gshard = torch.empty(n // world, device=dev)
            lo = rank * (n // world)

            def nccl_seq():
                dist.reduce_scatter_tensor(gshard, grad, op=dist.ReduceOp.AVG)
                ops.adamw_step(flat[lo:lo + n // world], gshard, m, v, step, 1e-3, shadow=None, zero_grad=False)
                dist.all_gather_into_tensor(flat, flat[lo:lo + n // world])
                ops.cast_bf16(flat[dense_begin:], shadow)

            ms_nccl = timed(nccl_seq)
            if rank == 0:
                nv = n * 4 * (world - 1) / world
                print(f"dp_adamw_step n={n} world={world}: fused {ms_fused * 1e3:.1f} us "
                      f"({2 * nv / (ms_fused * 1e-3) / 1e9:.0f} GB/s NVLink in+out per GPU), "
                      f"NCCL reduce_scatter + AdamW + all_gather + cast {ms_nccl * 1e3:.1f} us")
        del ar
    if rank == 0:
        print("SYMM_CHECK_OK" if ok else "SYMM_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
