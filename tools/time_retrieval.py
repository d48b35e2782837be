#!/usr/bin/env python
"""A/B timing of one c3 retrieval pass (10k users x 1M items, top-100) with CUDA events; environment
knobs (e.g. TT_TOPK_STAGES) are read by the library at first launch, so one process per setting."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    import mrm_b200  # noqa: F401
    from mrm_b200 import retrieval
    users_n, items_n = int(os.environ.get("USERS", 10_000)), int(os.environ.get("ITEMS", 1_000_000))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    table = torch.nn.functional.normalize(torch.randn(items_n + 1, 256, device=dev, generator=g), dim=1)
    table[0] = 0
    index = retrieval.CatalogIndex(table, device=dev)
    t = torch.randint(1, items_n, (users_n,), device=dev, generator=g)
    users = torch.nn.functional.normalize(table[t] + 3.3 / 16 * torch.randn(users_n, 256, device=dev, generator=g), dim=1)
    kps = int(os.environ.get("SHARD_KPRIME", 0))     # > 0: time the per-shard candidate pass of the sharded protocol

    def one_pass():
        if kps:
            return retrieval.retrieve_candidates(users, index, kps)
        return retrieval.retrieve_topk(users, index, 100, exact_fallback=False)

    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one_pass()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    _, _, n_fb = retrieval.retrieve_topk(users, index, 100, exact_fallback=True)   # certificate failures -> exact path
    hit = (retrieval.retrieve_topk(users, index, 100, exact_fallback=False)[0][:, :10] == t[:, None].int()).any(1).float().mean().item()
    print(f"users={users_n} items={items_n} fallback_users={n_fb} recall@10={hit:.3f}")
    print(f"knobs={ {k: v for k, v in os.environ.items() if k.startswith('TT_')} } ms_per_pass median={ts[len(ts)//2]:.3f} min={ts[0]:.3f}")


if __name__ == "__main__":
    main()
