#!/usr/bin/env python
"""Real-NCCL check of the data-parallel step (run under torchrun on >= 2 GPUs):
every rank trains 3 steps with all-gathered negatives + reduce-scatter / sharded AdamW / all-gather;
rank 0 then replays the same 3 steps in ONE process with `world` virtual ranks (exchanges done by
torch.cat, full AdamW on the averaged gradient) and compares the resulting parameters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from mrm_b200 import synthetic  # noqa: E402
from mrm_b200.engine import TwoTowerEngine  # noqa: E402
from mrm_b200.train import TrainStepRunner  # noqa: E402


def main():
    sharded = "--sharded-table" in sys.argv     # config-5 layout: row-sharded ID table (all-to-all lookups)
    comm = "nccl" if "--nccl" in sys.argv else "auto"   # auto = symmetric-arena kernels (csrc/tt_symm.cu)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = synthetic.TwoTowerConfig(vocab_size=3001, max_seq_len=50, dropout=0.0)
    sd = synthetic.make_state_dict(cfg, seed=7)
    B, L, steps, lr = 32, 50, 3, 1e-3
    batches = [[synthetic.make_batch(cfg, B, seed=100 + 10 * s + r, num_users=20) for r in range(world)] for s in range(steps)]
    tname = "user_tower.item_embedding.weight"
    if sharded:
        import dataclasses
        from mrm_b200.sharding import RowShardedTable, SymmShardedTable
        eng = TwoTowerEngine(dataclasses.replace(cfg, vocab_size=2))
        if comm == "nccl":      # torch-op exchange (all_to_all of ids / rows), also what the gloo tests cover
            table = RowShardedTable(cfg.vocab_size, 256, rank, world, eng.device)
        else:                   # symmetric arena: rows read from / gradients added into the owner over NVLink
            table = SymmShardedTable(cfg.vocab_size, 256, device=eng.device)
        runner = TrainStepRunner(eng, B, L, world_size=world, lr=lr, use_graph=True, sharded_table=table, comm=comm)
        eng.load_state_dict(sd)
        table.load_full(sd[tname].cuda())
        if rank == 0:
            print("exchanges:", runner.comm_description())
    else:
        eng = TwoTowerEngine(cfg)
        eng.load_state_dict(sd)
        runner = TrainStepRunner(eng, B, L, world_size=world, lr=lr, use_graph=True, comm=comm)
        if rank == 0:
            print("exchanges:", runner.comm_description())
    losses = [runner.step_from_host(batches[s][rank]) for s in range(steps)]
    torch.cuda.synchronize()
    if sharded:     # same flat layout as the replicated engine: [full table | everything else]
        ref_layout = TwoTowerEngine(cfg)
        ref_layout.load_state_dict({**eng.state_dict(), tname: table.gather_full()})
        mine = ref_layout.flat.clone()
        del ref_layout
    else:
        mine = eng.flat.clone()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = True
    if rank == 0:
        for r in range(1, world):
            d = (gathered[r] - gathered[0]).abs().max().item()
            print(f"replica divergence rank {r}: {d:.3e}")
            ok &= d == 0.0
        # single-process replay with virtual ranks
        engs = []
        for r in range(world):
            e = TwoTowerEngine(cfg)
            e.load_state_dict(sd)
            engs.append(e)
        ref_losses = []
        for s in range(steps):
            wss = [e.forward_towers({k: v.cuda() for k, v in batches[s][r].items()}, training=True) for r, e in enumerate(engs)]
            U_all = torch.cat([ws["un_bf"] for ws in wss]); I_all = torch.cat([ws["in_bf"] for ws in wss])
            uid_all = torch.cat([batches[s][r]["user_idx"] for r in range(world)]).cuda()
            gs = []
            for r, (e, ws) in enumerate(zip(engs, wss)):
                g = e.gathered_workspace(ws, world)
                g["U_all"].copy_(U_all); g["I_all"].copy_(I_all); g["uid_all"].copy_(uid_all)
                e.loss_forward(ws, batches[s][r]["user_idx"].cuda(), gathered=g, rank=r)
                gs.append(g)
            lr_all = torch.cat([ws["lse_r"] for ws in wss]); lc_all = torch.cat([ws["lse_c"] for ws in wss])
            tot = 0.0
            for r, (e, ws, g) in enumerate(zip(engs, wss, gs)):
                g["lse_r_all"].copy_(lr_all); g["lse_c_all"].copy_(lc_all)
                tot += e.loss_value(ws, world * B).item()
                e.backward()
            ref_losses.append(tot)
            avg = sum(e.grad for e in engs) / world
            for e in engs:
                e.grad.copy_(avg)
                e.adamw_step(lr=lr)
        # Adam normalises every element's update to ~lr, so elements whose gradient is pure rounding noise
        # (atomics order) may move differently; compare the UPDATE vectors in L2 and the loss trajectories.
        # Bounds: tools/determinism_check.py shows ONE GPU re-running these steps from the same state spreads
        # by 1e-4 / 4e-4 in the step-2 / step-3 loss at lr 1e-3 (gradients themselves repeat to 1e-7), so the
        # cross-implementation bound sits just above that run-to-run spread.
        p0 = TwoTowerEngine(cfg)
        p0.load_state_dict(sd)
        da, db = gathered[0] - p0.flat, engs[0].flat - p0.flat
        d = ((da - db).norm() / db.norm()).item()
        dl = max(abs(a - b) for a, b in zip(losses, ref_losses))
        print(f"NCCL run vs single-process virtual ranks: relative update diff {d:.3e}, max |loss diff| {dl:.3e}, "
              f"losses {losses}")
        ok &= d < 1e-1 and dl < 2e-3
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
