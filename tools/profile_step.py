#!/usr/bin/env python
"""One eager training step (c2) or one retrieval pass (c3) inside a cudaProfilerStart/Stop window,
for `ncu --profile-from-start off`. Prints nothing that could be mistaken for a bench value."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="train", choices=["train", "retrieval", "shard"])
    ap.add_argument("--shards", type=int, default=8, help="--what shard: the per-rank pass of a catalog sharded this many ways")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seq-len", type=int, default=200)
    ap.add_argument("--vocab", type=int, default=100_001)
    ap.add_argument("--users", type=int, default=10_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    args = ap.parse_args()
    import mrm_b200
    from mrm_b200 import retrieval, synthetic
    from mrm_b200.engine import TwoTowerEngine
    dev = torch.device("cuda", 0)
    if args.what == "train":
        cfg = synthetic.TwoTowerConfig(vocab_size=args.vocab, max_seq_len=args.seq_len, dropout=0.1)
        eng = TwoTowerEngine(cfg, dev)
        eng.load_state_dict(synthetic.make_state_dict(cfg, seed=0))
        batch = {k: v.cuda() for k, v in synthetic.make_batch(cfg, args.batch, seed=1, full_length=True,
                                                              num_users=1_000_000).items()}
        for _ in range(3):
            eng.train_step(batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eng.train_step(batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    elif args.what == "shard":
        # what ONE rank of an N-way sharded catalog runs per pass (retrieval.sharded_topk, bounded protocol): the
        # candidate pass over items/N rows writing the exchange layout, then the merge of N gathered buffers
        G = args.shards
        n = (args.items + 1 + G - 1) // G
        g = torch.Generator(device=dev).manual_seed(1)
        table = torch.nn.functional.normalize(torch.randn(n, 256, device=dev, generator=g), dim=1)
        index = retrieval.CatalogIndex.from_shard(table, n, args.items + 1, device=dev)     # a middle shard
        t = torch.randint(1, n, (args.users,), device=dev, generator=g)
        users = torch.nn.functional.normalize(table[t] + 3.3 / 16 * torch.randn(args.users, 256, device=dev, generator=g), dim=1)
        kps = retrieval.shard_kprime(256, G)
        buf = retrieval.exchange_buffers(index, args.users, retrieval._pitch(kps), G)

        def one():
            retrieval.retrieve_candidates(users, index, kps, pack=buf["pack"])
            for r in range(G):                       # stands in for the all-gather
                buf["all"][r * args.users:(r + 1) * args.users].copy_(buf["pack"])
            return retrieval.merge_packed(buf["all"], G, args.users, kps, 100, buf["bad"])

        for _ in range(2):
            one()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        one()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        g = torch.Generator(device=dev).manual_seed(1)
        table = torch.nn.functional.normalize(torch.randn(args.items + 1, 256, device=dev, generator=g), dim=1)
        table[0] = 0
        index = retrieval.CatalogIndex(table, device=dev)
        t = torch.randint(1, args.items, (args.users,), device=dev, generator=g)
        users = torch.nn.functional.normalize(table[t] + 3.3 / 16 * torch.randn(args.users, 256, device=dev, generator=g), dim=1)
        for _ in range(2):
            retrieval.retrieve_topk(users, index, 100, exact_fallback=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        retrieval.retrieve_topk(users, index, 100, exact_fallback=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profile window done")


if __name__ == "__main__":
    main()
