"""world_size-2 gloo tests (CPU) of the multi-GPU protocols' host-side logic:
 - retrieval: contiguous catalog shards, padding row masked only where it lives, per-shard canonical
   top-K, all-gather, merge == unsharded result (the CUDA merge kernel implements the same rule);
 - data-parallel training: averaging per-rank gradients of per-rank losses (reference DDP semantics,
   src/train.py:300) == gradient of the mean of the per-rank losses.
The per-rank compute here is the CPU oracle; the CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mrm_b200 import synthetic
from mrm_b200.sharding import merge_bounded_reference, merge_canonical, shard_bounds
from oracle import two_tower_oracle as oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _retrieval_worker(rank, world, K):
    torch.set_num_threads(1)
    table = synthetic.make_catalog(2999, 256, seed=5, grid=2.0 ** -3)      # heavy ties
    users, _ = synthetic.make_queries(table, 40, seed=6, noise=5.0, grid=2.0 ** -3)
    first, rows = shard_bounds(table.shape[0], world, rank)
    local = table[first:first + rows]
    s = users @ local.t()
    if first == 0:
        s[:, 0] = float("-inf")
    v, i = oracle.canonical_topk(s, K)
    i = (i + first).to(torch.int32)
    all_v = [torch.empty_like(v) for _ in range(world)]
    all_i = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(all_v, v)
    dist.all_gather(all_i, i)
    mi, mv = merge_canonical(torch.stack(all_v), torch.stack(all_i))
    rv, ri = oracle.canonical_topk(oracle.retrieval_scores(users, table), K)
    assert torch.equal(mi.long(), ri), f"rank {rank}: merged shards differ from the unsharded top-K"
    assert torch.equal(mv, rv)


def test_sharded_retrieval_protocol_world2():
    _run(_retrieval_worker, 2, 50)


def _bounded_retrieval_worker(rank, world, K, kps, noise, eps):
    """The bounded shard protocol (retrieval.sharded_topk from 4 shards on): a shard ships its kps best items
    with exact scores plus a completeness bound; the merged top K is exact wherever it is certified. The shard's
    selection is done on PERTURBED scores (standing in for the bf16 scoring pass, |error| <= eps), the shipped
    scores are exact — as on the device."""
    torch.set_num_threads(1)
    table = synthetic.make_catalog(2999, 256, seed=5)
    users, _ = synthetic.make_queries(table, 64, seed=6, noise=noise)
    first, rows = shard_bounds(table.shape[0], world, rank)
    exact = (users.double() @ table[first:first + rows].double().t()).float()
    if first == 0:
        exact[:, 0] = float("-inf")
    g = torch.Generator().manual_seed(100 + rank)
    approx = exact + (torch.rand(exact.shape, generator=g) * 2 - 1) * eps      # the selection sees these
    n = min(kps, rows)
    sel_v, sel_i = torch.topk(approx, n, dim=1)
    # completeness bound: everything NOT selected has approx score <= the smallest selected one, hence
    # exact score <= that + eps; a shard that lists all of its items has nothing outside the list
    bound = sel_v[:, -1] + eps if n < rows else torch.full((users.shape[0],), float("-inf"))
    lv = torch.full((users.shape[0], kps), float("-inf"))
    li = torch.full((users.shape[0], kps), -1, dtype=torch.int32)
    ev = exact.gather(1, sel_i)
    order = torch.argsort(ev, dim=1, descending=True, stable=True)
    lv[:, :n] = ev.gather(1, order)
    li[:, :n] = (sel_i.gather(1, order) + first).to(torch.int32)
    all_v = [torch.empty_like(lv) for _ in range(world)]
    all_i = [torch.empty_like(li) for _ in range(world)]
    all_b = [torch.empty_like(bound) for _ in range(world)]
    dist.all_gather(all_v, lv)
    dist.all_gather(all_i, li)
    dist.all_gather(all_b, bound)
    mi, mv, ok = merge_bounded_reference(torch.stack(all_v), torch.stack(all_i), torch.stack(all_b), K)
    rv, ri = oracle.canonical_topk(oracle.retrieval_scores(users.double(), table.double()).float(), K)
    assert ok.float().mean() > 0.5, f"rank {rank}: only {ok.float().mean():.2f} certified — the test would prove nothing"
    assert torch.equal(mi.long()[ok], ri[ok]), f"rank {rank}: a certified list differs from the unsharded top-K"
    assert torch.equal(mv[ok], rv[ok])
    # every rank holds the same merged result and the same certificate (the fallback decision is collective)
    flat = torch.cat([mi.float().flatten(), ok.float()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref)


def test_bounded_shard_protocol_world2():
    _run(_bounded_retrieval_worker, 2, 50, 96, 0.35, 1e-3)


def test_bounded_shard_protocol_world3_short_lists():
    # lists of 24 for a top-50: certificates fail for part of the users — whatever certifies must still be exact
    _run(_bounded_retrieval_worker, 3, 50, 24, 0.35, 1e-3)


class _CpuShard:
    """Stand-in for retrieval.CatalogIndex holding a CPU shard (the device kernels are replaced below)."""

    def __init__(self, table, first, rows):
        from mrm_b200 import retrieval
        self.table, self.item_base, self.vocab_size = table[first:first + rows], first, table.shape[0]
        self.is_sharded, self._group_checked, self._scratch = True, set(), {}
        self.check_group = lambda U, group=None: retrieval.CatalogIndex.check_group(self, U, group)

    @property
    def num_rows(self):
        return self.table.shape[0]


def _sharded_topk_plumbing_worker(rank, world, K, kps):
    """retrieval.sharded_topk itself (packing, the all-gathers, certificate, collective fallback on the padded
    subset) with the four device entry points replaced by their CPU reference semantics."""
    torch.set_num_threads(1)
    from mrm_b200 import retrieval, sharding
    table = synthetic.make_catalog(2999, 256, seed=5)
    users, _ = synthetic.make_queries(table, 70, seed=6, noise=0.35)
    first, rows = shard_bounds(table.shape[0], world, rank)
    index = _CpuShard(table, first, rows)

    def shard_scores(u, ix):
        sc = (u.double() @ ix.table.double().t()).float()
        if ix.item_base == 0:
            sc[:, 0] = float("-inf")
        return sc

    def retrieve_topk(u, ix, k, kprime=256, mask_item0=True, exact_fallback=True, flags_out=None):
        v, i = oracle.canonical_topk(shard_scores(u, ix), k)
        return (i + ix.item_base).to(torch.int32), v, 0

    def retrieve_candidates(u, ix, kprime, mask_item0=True, pack=None):
        """writes the exchange layout [scores | ids | bound | flag] like tt_topk_finalize_bounded"""
        sc = shard_scores(u, ix)
        n = min(kprime, sc.shape[1])
        v, i = oracle.canonical_topk(sc, n)
        out_v = torch.full((u.shape[0], kprime), float("-inf"))
        out_i = torch.full((u.shape[0], kprime), -1, dtype=torch.int32)
        out_v[:, :n], out_i[:, :n] = v, (i + ix.item_base).to(torch.int32)
        bound = v[:, -1].clone() if n < sc.shape[1] else torch.full((u.shape[0],), float("-inf"))
        pack[:, :kprime] = out_v.view(torch.int32)
        pack[:, kprime:2 * kprime] = out_i
        pack[:, 2 * kprime] = bound.view(torch.int32)
        pack[:, 2 * kprime + 1] = 0
        return out_i, out_v, bound, torch.zeros(u.shape[0], dtype=torch.int32)

    def merge_packed(allp, G, U, kp, k, bad):
        """reference semantics of tt_topk_merge_packed on the gathered exchange buffer"""
        a = allp.view(G, U, allp.shape[1])
        i, v, ok = sharding.merge_bounded_reference(a[:, :, :kp].contiguous().view(torch.float32),
                                                    a[:, :, kp:2 * kp].contiguous(),
                                                    a[:, :, 2 * kp].contiguous().view(torch.float32), k)
        bad.copy_((~ok | a[:, :, 2 * kp + 1].any(dim=0)).to(torch.int32))
        return i, v

    calls = {"fallback_users": 0}

    def merge_topk(sc, ix):
        calls["fallback_users"] = sc.shape[1]
        return merge_canonical(sc, ix)

    saved = {n: getattr(retrieval, n) for n in ("retrieve_topk", "retrieve_candidates", "merge_packed", "merge_topk",
                                                "shard_kprime")}
    try:
        retrieval.retrieve_topk, retrieval.retrieve_candidates = retrieve_topk, retrieve_candidates
        retrieval.merge_packed, retrieval.merge_topk = merge_packed, merge_topk
        retrieval.shard_kprime = lambda kprime, shards: kps
        mi, mv = retrieval.sharded_topk(users, index, K, bounded=True)
        fallback_users = calls["fallback_users"]
        ei, ev = retrieval.sharded_topk(users, index, K, bounded=False)
    finally:
        for n, f in saved.items():
            setattr(retrieval, n, f)
    rv, ri = oracle.canonical_topk(oracle.retrieval_scores(users.double(), table.double()).float(), K)
    assert torch.equal(mi.long(), ri) and torch.equal(mv, rv), f"rank {rank}: bounded protocol"
    assert torch.equal(ei.long(), ri) and torch.equal(ev, rv), f"rank {rank}: per-shard exact protocol"
    return fallback_users


def _bad_tiling_worker(rank, world):
    """CatalogIndex.check_group: shards that overlap / leave a gap, or ranks that bring different user counts, are
    refused on every rank alike (the merge kernels assume distinct items and paired user rows)."""
    torch.set_num_threads(1)
    table = synthetic.make_catalog(299, 256, seed=5)
    first, rows = shard_bounds(table.shape[0], world, rank)
    ok = _CpuShard(table, first, rows)
    ok.check_group(10)
    overlap = _CpuShard(table, 0, table.shape[0])            # every rank claims the whole catalog (unsharded index)
    with pytest.raises(ValueError, match="tile"):
        overlap.check_group(10)
    ragged = _CpuShard(table, first, rows)
    with pytest.raises(ValueError, match="disagree"):
        ragged.check_group(10 + rank)


def test_shard_tiling_check_world2():
    _run(_bad_tiling_worker, 2)


def _plumbing_short_lists(rank, world):
    # lists of 26 for a top-50 over 2 shards: 52 candidates, the 50th of the union almost never beats both
    # shards' bounds -> the uncertified users go through the collective fallback, padded to 256 users
    n = _sharded_topk_plumbing_worker(rank, world, 50, 26)
    assert n == 256, n


def _plumbing_long_lists(rank, world):
    _sharded_topk_plumbing_worker(rank, world, 50, 64)


def test_sharded_topk_plumbing_world2_all_fallback():
    _run(_plumbing_short_lists, 2)


def test_sharded_topk_plumbing_world3():
    _run(_plumbing_long_lists, 3)


def test_shard_kprime_budget():
    """Per-shard candidate budget of the bounded protocol: the whole budget on one shard, about 1.5x the even
    share (+16) on several, a multiple of 8 inside [48, K'], and enough lists to hold a top-100 from 2 shards on."""
    from mrm_b200.retrieval import shard_kprime
    assert shard_kprime(256, 1) == 256
    got = {g: shard_kprime(256, g) for g in (2, 3, 4, 8, 16, 64)}
    assert got[2] == 208 and got[4] == 112 and got[8] == 64, got
    for g, k in got.items():
        assert k % 8 == 0 and 48 <= k <= 256
        assert k >= min(256, 1.5 * 256 / g)
        if g >= 2:
            assert g * k >= 100
    assert all(got[a] >= got[b] for a, b in zip((2, 3, 4, 8, 16), (3, 4, 8, 16, 64)))


def test_shard_bounds_cover_catalog():
    for V in (1, 7, 1000, 1_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(V, world, r) for r in range(world)]
            assert spans[0][0] == 0
            pos = 0
            for first, rows in spans:
                assert first == min(pos, V) or rows == 0
                pos = first + rows
            assert pos == V


def _dp_worker(rank, world):
    torch.set_num_threads(1)
    cfg = synthetic.TwoTowerConfig(vocab_size=301, num_countries=7, max_seq_len=10, embedding_dim=64,
                                   modality_dim=16, dropout=0.0)
    sd = synthetic.make_state_dict(cfg, seed=1)
    batch = synthetic.make_batch(cfg, 8, seed=10 + rank)
    _, _, _, _, grads, _ = oracle.loss_and_grads(sd, batch, cfg.temperature, cfg.num_heads, dtype=torch.float64)
    flat = torch.cat([g.flatten() for g in grads.values()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    # every rank now holds the same averaged gradient = gradient of mean_r(loss_r)
    ref = None
    for r in range(world):
        b = synthetic.make_batch(cfg, 8, seed=10 + r)
        _, _, _, _, g, _ = oracle.loss_and_grads(sd, b, cfg.temperature, cfg.num_heads, dtype=torch.float64)
        f = torch.cat([x.flatten() for x in g.values()])
        ref = f if ref is None else ref + f
    ref /= world
    assert (flat - ref).abs().max().item() < 1e-12


def test_dp_gradient_average_world2():
    _run(_dp_worker, 2)


def _sharded_table_worker(rank, world):
    """Row-sharded ID table: lookups and gradient scatter through the all-to-all
    exchange == the replicated table's index_select / dense embedding backward (padding row excluded)."""
    from mrm_b200.sharding import RowShardedTable
    torch.set_num_threads(1)
    V, D, T = 1003, 16, 700                      # V not divisible by the world size: ragged last shard
    full = torch.randn(V, D, generator=torch.Generator().manual_seed(1))
    t = RowShardedTable(V, D, rank, world, "cpu")
    t.load_full(full)
    g = torch.Generator().manual_seed(10 + rank)
    ids = torch.randint(0, V, (T,), generator=g)
    ids[:50] = 0                                  # padding tokens
    ids[50:90] = V - 1                            # hot row on the last shard
    rows = t.lookup(ids)
    assert torch.equal(rows, full[ids]), f"rank {rank}: lookup differs"
    drows = torch.randn(T, D, generator=g)
    t.backward(drows, scale=1.0 / world)
    # reference: every rank's tokens contribute, averaged over ranks, row 0 gets nothing
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_d = [torch.empty_like(drows) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    dist.all_gather(all_d, drows)
    ref = torch.zeros(V, D)
    for i_r, d_r in zip(all_ids, all_d):
        d_r = d_r.clone() / world
        d_r[i_r == 0] = 0
        ref.index_add_(0, i_r, d_r)
    assert torch.allclose(t.grad, ref[t.first:t.first + t.rows], atol=1e-5), f"rank {rank}: gradient differs"
    assert t.first != 0 or t.grad[0].abs().max().item() == 0.0
    assert torch.equal(t.gather_full(), full)


def test_row_sharded_table_exchange_world2():
    _run(_sharded_table_worker, 2)


def test_row_sharded_table_exchange_world3():
    _run(_sharded_table_worker, 3)
