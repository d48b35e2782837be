"""Duplicate-free ID-table traffic (csrc/tt_sparse.cu) on one GPU, against torch.unique / index_add_:
tt_ids_dedup, tt_rows_gather, tt_rows_scatter_add in their local-table form (the sharded form uses the same kernels
with the owner's shard as the table; tools/dist_check.py --sharded-table covers it over real NVLink), and the
embedding kernels running on the compact cache == running on the table itself (bit for bit in the forward)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scratch(V, T):
    dev = "cuda"
    return dict(flag=torch.zeros(V, device=dev, dtype=torch.int32), slot=torch.zeros(V, device=dev, dtype=torch.int32),
                uniq=torch.zeros(T + 1, device=dev, dtype=torch.int64), state=torch.zeros(2, device=dev, dtype=torch.int32),
                inverse=torch.zeros(T, device=dev, dtype=torch.int64))


@pytest.mark.parametrize("V,T", [(1000, 5000), (100_001, 51_200)])
def test_ids_dedup_gather_scatter_match_torch(V, T):
    from mrm_b200 import ops, synthetic
    cfg = synthetic.TwoTowerConfig(vocab_size=V, max_seq_len=50)
    g = torch.Generator(device="cuda").manual_seed(V + T)
    table = torch.randn(V, 256, device="cuda", generator=g)
    sc = _scratch(V, T)
    cache = torch.zeros(T + 1, 256, device="cuda")
    gacc = torch.zeros(T + 1, 256, device="cuda")
    grad = torch.zeros(V, 256, device="cuda")
    ref_grad = torch.zeros(V, 256, device="cuda", dtype=torch.float64)
    ref_abs = torch.zeros(V, 256, device="cuda", dtype=torch.float64)     # sum of |terms|: scale of the fp32 summation error
    for step in range(3):              # the flag table is cleaned by the NEXT call: several steps, different ids
        ids = synthetic.make_batch(cfg, T // 50, seed=step, zipf=True)["history_ids"].cuda().view(-1)[:T].contiguous()
        ops.ids_dedup(ids, V, sc["flag"], sc["slot"], sc["uniq"], sc["state"], sc["inverse"])
        n = int(sc["state"][1].item())
        uniq = sc["uniq"][:n]
        want = torch.unique(ids[ids != 0])
        assert uniq[0].item() == 0 and torch.equal(torch.sort(uniq[1:]).values, want)
        assert torch.equal(uniq[sc["inverse"]], ids)                      # every token finds its id again
        assert int(sc["flag"].sum().item()) == n - 1                      # exactly the registered ids are flagged
        ops.rows_gather(sc["uniq"], sc["state"], cache, table_local=table)
        assert torch.equal(cache[:n], table[uniq])
        # backward side: per-token rows combined in the compact buffer, then one add per distinct row
        d = torch.randn(T, 256, device="cuda", generator=g)
        d[ids == 0] = 0
        gacc.index_add_(0, sc["inverse"], d)
        gacc[0] = 0
        ref_grad.index_add_(0, ids, d.double())
        ref_abs.index_add_(0, ids, d.double().abs())
        ops.rows_scatter_add(sc["uniq"], sc["state"], gacc, grad_local=grad)
        torch.cuda.synchronize()
        assert gacc.abs().max().item() == 0.0                             # cleared for the next step
        # hot rows sum thousands of terms in an unspecified order: fp32 error scales with the sum of |terms|
        assert bool(((grad.double() - ref_grad).abs() <= 1e-6 * ref_abs + 1e-6).all())
    assert grad[0].abs().max().item() == 0.0


def test_embedding_kernels_on_the_compact_cache_equal_the_table_path():
    """tt_embed_ln_fwd on (cache, slots) == on (table, ids), bit for bit; the backward's per-id gradient sums agree
    (atomic order differs)."""
    from mrm_b200 import ops, synthetic
    V, B, L = 5001, 32, 50
    cfg = synthetic.TwoTowerConfig(vocab_size=V, max_seq_len=L)
    sd = synthetic.make_state_dict(cfg, seed=1)
    ids = synthetic.make_batch(cfg, B, seed=2)["history_ids"].cuda()
    T = B * L
    table = sd["user_tower.item_embedding.weight"].cuda()
    P = sd["user_tower.position_embedding.weight"].cuda()
    w, b = sd["user_tower.layer_norm.weight"].cuda(), sd["user_tower.layer_norm.bias"].cuda()
    nw = sd["user_tower.transformer_encoder.layers.0.norm1.weight"].cuda()
    nb = sd["user_tower.transformer_encoder.layers.0.norm1.bias"].cuda()
    outs = []
    sc = _scratch(V, T)
    cache = torch.zeros(T + 1, 256, device="cuda")
    ops.ids_dedup(ids.view(-1), V, sc["flag"], sc["slot"], sc["uniq"], sc["state"], sc["inverse"])
    ops.rows_gather(sc["uniq"], sc["state"], cache, table_local=table)
    for e_ids, e_table in ((ids.view(-1), table), (sc["inverse"], cache)):
        x0 = torch.empty(T, 256, device="cuda")
        h = torch.empty(T, 256, device="cuda", dtype=torch.bfloat16)
        ops.embed_ln_fwd(e_ids, e_table, P, w, b, nw, nb, B, L, x0, h)
        outs.append((x0, h))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    dx = torch.randn(T, 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    dE = torch.zeros(V, 256, device="cuda")
    gacc = torch.zeros(T + 1, 256, device="cuda")
    aux = [torch.zeros(L, 256, device="cuda"), torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")]
    ops.embed_ln_bwd(ids.view(-1), table, P, w, b, dx, B, L, dE, *aux)
    aux2 = [torch.zeros_like(t) for t in aux]
    ops.embed_ln_bwd(sc["inverse"], cache, P, w, b, dx, B, L, gacc, *aux2)
    dE2 = torch.zeros(V, 256, device="cuda")
    ops.rows_scatter_add(sc["uniq"], sc["state"], gacc, grad_local=dE2)
    torch.cuda.synchronize()
    assert (dE - dE2).abs().max().item() <= 1e-4 * max(1.0, dE.abs().max().item())
    assert dE2[0].abs().max().item() == 0.0
    for a, c in zip(aux, aux2):
        assert (a - c).abs().max().item() <= 1e-3 * max(1.0, a.abs().max().item())


def test_deterministic_table_gradient():
    """Deterministic mode of the embedding backward (tt_embed_ln_bwd_det + tt_rows_scatter_add_i64): per distinct id
    the token gradient rows are summed in 64-bit fixed point, so heavy duplicates (Zipfian ids: one id on a fifth of
    all tokens) give the SAME bits on every run; the result equals the fp64 sum rounded to fp32 up to the 2^-40
    quantisation of each addend, and dP / d(ln) equal the default kernel's."""
    import torch.nn.functional as F
    from mrm_b200 import ops
    B, L, V = 48, 200, 5001
    g = torch.Generator().manual_seed(21)
    ids = (torch.rand(B, L, generator=g) ** 6 * (V - 1)).long() + 1     # Zipf-like: small ids dominate
    ids[torch.rand(B, L, generator=g) < 0.2] = 7                        # one very hot id
    ids[:, -5:] = 0                                                     # padding
    ids = ids.cuda()
    T = B * L
    E = (torch.randn(V, 256, generator=g) * 0.1).cuda()
    P = (torch.randn(L, 256, generator=g) * 0.1).cuda()
    w = (1 + 0.1 * torch.randn(256, generator=g)).cuda()
    b = (0.1 * torch.randn(256, generator=g)).cuda()
    dx0 = torch.randn(T, 256, generator=g).cuda() * 1e-3
    flag = torch.zeros(V, device="cuda", dtype=torch.int32)
    slot = torch.zeros(V, device="cuda", dtype=torch.int32)
    uniq = torch.zeros(T + 1, device="cuda", dtype=torch.int64)
    state = torch.zeros(2, device="cuda", dtype=torch.int32)
    inverse = torch.zeros(T, device="cuda", dtype=torch.int64)
    acc = torch.zeros(T + 1, 256, device="cuda", dtype=torch.int64)
    runs = []
    for _ in range(4):
        dE, dP = torch.zeros_like(E), torch.zeros_like(P)
        dg, db = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
        ops.ids_dedup(ids.view(-1), V, flag, slot, uniq, state, inverse)
        ops.embed_ln_bwd_det(ids.view(-1), E, P, w, b, dx0, B, L, inverse, acc, dP, dg, db, drop_p=0.1, seed=9, site=2)
        ops.rows_scatter_add_i64(uniq, state, acc, dE)
        torch.cuda.synchronize()
        assert int(acc.abs().max()) == 0                     # accumulator cleared for the next step
        runs.append(dE)
    for r in runs[1:]:
        assert torch.equal(r, runs[0])                       # bit-identical from run to run
    assert float(runs[0][0].abs().max()) == 0.0              # padding row untouched
    # the default kernel: same values up to the order of its floating-point atomics
    dE2, dP2 = torch.zeros_like(E), torch.zeros_like(P)
    dg2, db2 = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    ops.embed_ln_bwd(ids.view(-1), E, P, w, b, dx0, B, L, dE2, dP2, dg2, db2, drop_p=0.1, seed=9, site=2)
    torch.cuda.synchronize()
    scale = float(dE2.abs().max())
    assert float((runs[0] - dE2).abs().max()) <= 2e-6 * scale + 1e-10
    # without dropout: against autograd in fp64 (exact sum, rounded once)
    dE3, dP3 = torch.zeros_like(E), torch.zeros_like(P)
    dg3, db3 = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    ops.ids_dedup(ids.view(-1), V, flag, slot, uniq, state, inverse)
    ops.embed_ln_bwd_det(ids.view(-1), E, P, w, b, dx0, B, L, inverse, acc, dP3, dg3, db3)
    ops.rows_scatter_add_i64(uniq, state, acc, dE3)
    Er = E.double().requires_grad_(True)
    x = F.layer_norm(Er[ids] + P.double()[:L].unsqueeze(0), (256,), w.double(), b.double()).view(T, 256)
    x.backward(dx0.double())
    ref = Er.grad.clone()
    ref[0] = 0
    err = float((dE3.double() - ref).abs().max())
    assert err <= 3e-6 * float(ref.abs().max()), err       # fp32 LayerNorm-backward arithmetic per token, exact sum
