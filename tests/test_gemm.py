"""tcgen05 GEMM (tt_gemm_bf16) against torch fp32 matmul of the same bf16 operands.

The operands are bf16 on both sides and both accumulate in fp32, so the only
difference is summation order: tolerance 2e-3 * sqrt(K)-scaled magnitude is generous;
in practice errors are ~1e-5 relative.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(device=device, dtype=torch.bfloat16)


def _operands(M, N, K, a_mn, b_mn, seed=0):
    A = _mk((M, K), seed)
    B = _mk((N, K), seed + 1)
    A_store = A.t().contiguous() if a_mn else A
    B_store = B.t().contiguous() if b_mn else B
    return A, B, A_store, B_store


def _check(out, ref, K, what=""):
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= 2e-3 * scale + 1e-4 * math.sqrt(K), f"{what}: max err {err} (scale {scale})"


LAYOUTS = [(False, False), (False, True), (True, False), (True, True)]
LAYOUT_IDS = ["kk", "kmn", "mnk", "mnmn"]


@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS, ids=LAYOUT_IDS)
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 256), (384, 768, 256), (1024, 256, 1024),
                                   (640, 1024, 320)])
def test_gemm_plain(a_mn, b_mn, M, N, K):
    from mrm_b200 import ops
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(As, Bs, a_mn=a_mn, b_mn=b_mn, out_f32=out)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    _check(out, ref, K, f"plain {M}x{N}x{K}")


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_block_n(block_n):
    from mrm_b200 import ops
    M, N, K = 512, 512, 192
    A, B, As, Bs = _operands(M, N, K, False, False, seed=3)
    out = torch.empty((M, N), device="cuda", dtype=torch.float32)
    ops.gemm(As, Bs, out_f32=out, block_n=block_n)
    torch.cuda.synchronize()
    _check(out, A.float() @ B.float().t(), K, f"block_n={block_n}")


def test_gemm_ragged_edges():
    """M not a multiple of 128, N not a multiple of 32, K not a multiple of 64 (K-major)."""
    from mrm_b200 import ops
    M, N, K = 200, 104, 304
    A = _mk((M, 320), 5)[:, :K]
    B = _mk((N, 320), 6)[:, :K]
    out = torch.full((M, 112), -7.0, device="cuda", dtype=torch.float32)
    ops.gemm(A, B, out_f32=out, M=M, N=N, K=K)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    _check(out[:, :N], ref, K, "ragged")
    assert (out[:, N:] == -7.0).all(), "wrote outside N"


def test_gemm_many_tiles_persistent():
    """More tiles than SMs: exercises the persistent loop, smem ring wrap and both TMEM buffers."""
    from mrm_b200 import ops
    M, N, K = 128 * 40, 1024, 256
    A, B, As, Bs = _operands(M, N, K, False, False, seed=9)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(As, Bs, out_bf16=out)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item(), err


def test_gemm_epilogue_bias_relu_residual_bf16():
    from mrm_b200 import ops
    M, N, K = 384, 512, 256
    A, B, As, Bs = _operands(M, N, K, False, False, seed=11)
    g = torch.Generator().manual_seed(1)
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    o32 = torch.empty((M, N), device="cuda")
    o16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(As, Bs, alpha=0.5, bias=bias, relu=True, residual=res, out_f32=o32, out_bf16=o16)
    torch.cuda.synchronize()
    ref = torch.relu(0.5 * (A.float() @ B.float().t()) + bias) + res
    _check(o32, ref, K, "epilogue fp32")
    assert (o16.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


def test_gemm_gate():
    from mrm_b200 import ops
    M, N, K = 256, 256, 128
    A, B, As, Bs = _operands(M, N, K, False, False, seed=13)
    gate = _mk((M, N), 14)
    out = torch.empty((M, N), device="cuda")
    ops.gemm(As, Bs, gate=gate, gate_scale=1.25, out_f32=out)
    torch.cuda.synchronize()
    ref = (A.float() @ B.float().t()) * 1.25 * (gate.float() > 0)
    _check(out, ref, K, "gate")


@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (False, False)], ids=["mnmn", "kk"])
def test_gemm_splitk_accumulate(a_mn, b_mn):
    """wgrad shape: small output, long contraction, split-K with red.add into fp32."""
    from mrm_b200 import ops
    M, N, K = 256, 256, 64 * 100
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, seed=17)
    out = torch.ones((M, N), device="cuda")
    ops.gemm(As, Bs, a_mn=a_mn, b_mn=b_mn, out_f32=out, accumulate=True)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + 1.0
    _check(out, ref, K, "split-K")


def test_gemm_dropout_is_deterministic_and_scaled():
    from mrm_b200 import ops
    M, N, K = 256, 256, 64
    A, B, As, Bs = _operands(M, N, K, False, False, seed=19)
    o1 = torch.empty((M, N), device="cuda")
    o2 = torch.empty((M, N), device="cuda")
    ops.gemm(As, Bs, drop_p=0.25, drop_seed=123, drop_site=2, out_f32=o1)
    ops.gemm(As, Bs, drop_p=0.25, drop_seed=123, drop_site=2, out_f32=o2)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
    ref = A.float() @ B.float().t()
    kept = o1 != 0
    frac = kept.float().mean().item()
    assert 0.72 < frac < 0.78, frac
    assert torch.allclose(o1[kept], ref[kept] / 0.75, rtol=2e-3, atol=1e-3)


def test_gemm_rejects_bad_args():
    from mrm_b200 import TTError, ops
    A = _mk((128, 64), 0)
    B = _mk((100, 64), 1)
    out = torch.empty((128, 100), device="cuda")
    with pytest.raises(TTError):
        ops.gemm(A, B.t().contiguous(), b_mn=True, out_f32=out)  # ldb = 100 is not a multiple of 8


def test_gemm_mn_major_ragged_n():
    """MN-major B with N = 304 (the user-fusion input width): ragged last 64-chunk."""
    from mrm_b200 import ops
    M, N, K = 256, 304, 256
    A = _mk((M, K), 21)
    Bfull = _mk((K + 1, N), 22)            # one spare row keeps the ragged chunk readable
    B = Bfull[:K]
    out = torch.empty((M, N), device="cuda")
    ops.gemm(A, B, b_mn=True, out_f32=out, N=N, K=K)
    torch.cuda.synchronize()
    _check(out, A.float() @ B.float(), K, "ragged MN-major N")


# ---- encoder-sized launches: many waves of the persistent loop, odd row-block counts, ragged edges ----
@pytest.mark.parametrize("a_mn,b_mn", LAYOUTS, ids=LAYOUT_IDS)
@pytest.mark.parametrize("K", [320, 712], ids=["k320", "k712"])
def test_gemm_large_layouts(a_mn, b_mn, K):
    """75 row blocks with a ragged last one, ragged N (520) and K, every operand layout. K = 712 (12 k-blocks)
    takes the CTA-pair path (cta_group::2, the odd last pair has an empty partner), K = 320 the single-CTA one."""
    from mrm_b200 import ops
    M, N = 128 * 75 - 56, 520
    A, B, As, Bs = _operands(M, N, K, a_mn, b_mn, seed=31)
    if b_mn:                                  # keep the ragged 64-chunk of an MN-major B readable
        Bfull = torch.zeros(K + 1, N, device="cuda", dtype=torch.bfloat16)
        Bfull[:K] = Bs
        Bs = Bfull[:K]
    if a_mn:
        Afull = torch.zeros(K + 1, M, device="cuda", dtype=torch.bfloat16)
        Afull[:K] = As
        As = Afull[:K]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(As, Bs, a_mn=a_mn, b_mn=b_mn, out_f32=out, M=M, N=N, K=K)
    torch.cuda.synchronize()
    _check(out, A.float() @ B.float().t(), K, "large")


@pytest.mark.parametrize("K", [256, 768], ids=["k256", "k768-pairs"])
def test_gemm_large_epilogues(K):
    """Encoder-sized launch (M = 51 200 / 4): bias + ReLU + dropout + residual, fp32 and bf16 outputs, gate;
    K = 768 runs on CTA pairs, K = 256 on single CTAs."""
    from mrm_b200 import ops
    M, N = 12800, 768
    A, B, As, Bs = _operands(M, N, K, False, False, seed=41)
    g = torch.Generator().manual_seed(2)
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    o32 = torch.empty((M, N), device="cuda")
    o16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(As, Bs, alpha=0.5, bias=bias, relu=True, residual=res, out_f32=o32, out_bf16=o16)
    torch.cuda.synchronize()
    ref = torch.relu(0.5 * (A.float() @ B.float().t()) + bias) + res
    _check(o32, ref, K, "large epilogue fp32")
    assert (o16.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    gate = _mk((M, N), 43)
    og = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(As, Bs, gate=gate, gate_scale=1.25, out_bf16=og)
    torch.cuda.synchronize()
    refg = (A.float() @ B.float().t()) * 1.25 * (gate.float() > 0)
    assert (og.float() - refg).abs().max().item() <= 2 ** -7 * refg.abs().max().item()
    o1 = torch.empty((M, N), device="cuda")
    o2 = torch.empty((M, N), device="cuda")
    ops.gemm(As, Bs, drop_p=0.25, drop_seed=5, drop_site=1, out_f32=o1)
    ops.gemm(As, Bs, drop_p=0.25, drop_seed=5, drop_site=1, out_f32=o2, block_n=64)   # other tile width, same mask
    torch.cuda.synchronize()
    assert torch.equal(o1 != 0, o2 != 0)
    assert torch.allclose(o1, o2, rtol=1e-5, atol=1e-5)


def test_gemm_large_splitk_wgrad():
    """Weight-gradient shape (1024 x 256 over 19 200 tokens, both operands MN-major): 8 row blocks x 18 splits."""
    from mrm_b200 import ops
    M, N, K = 1024, 256, 64 * 300
    A, B, As, Bs = _operands(M, N, K, True, True, seed=51)
    out = torch.ones((M, N), device="cuda")
    ops.gemm(As, Bs, a_mn=True, b_mn=True, out_f32=out, accumulate=True)
    torch.cuda.synchronize()
    _check(out, A.float() @ B.float().t() + 1.0, K, "large split-K")


@pytest.mark.parametrize("M,N,K,b_mn", [
    (256, 256, 256, True),        # B-row dgrad of the single-row last layer: several column blocks, only block 0 sums
    (200, 104, 304, False),       # ragged rows / columns / reduction (zero-filled tails must not count)
    (128 * 13 + 40, 256, 1024, True),   # CTA pairs (deep reduction), odd number of row blocks
    (51200, 256, 768, True),      # the c2 dgrad of the packed in_proj: dY = dqkv
    (51200, 256, 1024, True),     # the c2 dgrad of linear1: dY = dpre
    (2048, 512, 512, True),       # two column blocks per row block with a 256-wide tile
])
def test_gemm_a_colsum(M, N, K, b_mn):
    """a_colsum[k] += sum_m A[m, k] next to the product (the bias gradient that belongs to a dgrad GEMM's dY
    operand), against an fp64 column sum of the same bf16 values; the product itself must not change."""
    from mrm_b200 import ops
    lda = ((K + 63) // 64) * 64 + 64          # a view into a wider buffer, like dqkv[:, D:]
    A = _mk((M, lda), 21)[:, 64:64 + K]
    B = _mk((N, K), 22)
    Bs = B.t().contiguous() if b_mn else B
    out = torch.empty((M, N), device="cuda", dtype=torch.float32)
    ref_out = torch.empty_like(out)
    cs = torch.full((K,), 0.25, device="cuda", dtype=torch.float32)      # accumulated, not overwritten
    ops.gemm(A, Bs, b_mn=b_mn, out_f32=ref_out)
    ops.gemm(A, Bs, b_mn=b_mn, out_f32=out, a_colsum=cs)
    torch.cuda.synchronize()
    assert torch.equal(out, ref_out)
    ref = A.double().sum(dim=0) + 0.25
    err = (cs.double() - ref).abs().max().item()
    assert err <= 1e-6 * A.float().abs().sum(dim=0).max().item() + 1e-6, err


def test_gemm_a_colsum_rejects_mn_major_and_splitk():
    from mrm_b200 import ops
    from mrm_b200._lib import TTError
    A, B = _mk((256, 128), 1), _mk((128, 128), 2)
    out = torch.zeros((128, 128), device="cuda", dtype=torch.float32)
    cs = torch.zeros(512, device="cuda")
    with pytest.raises((TTError, AssertionError)):
        ops.gemm(A, B.t().contiguous(), a_mn=True, b_mn=True, out_f32=out, a_colsum=cs)
    with pytest.raises(TTError):
        ops.gemm(_mk((128, 512), 3), _mk((128, 512), 4), out_f32=out, accumulate=True, k_splits=2, a_colsum=cs)
