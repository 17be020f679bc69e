"""Grouped persistent weight-gradient GEMM (tulip_gemm_tn_group): the launch the executor uses for the weight gradients of a
Swin half-block (autograd of tulip.py:194-200 fc1 / fc2 and :282-324 qkv / proj).  fp32 outputs: rel-L2 <= 1e-4 against the
float64 product of the same bf16-exact operands; every problem of a launch is checked, so are accumulation into a non-zero
gradient and the permuted-row store of the PixelShuffle-feeding weights."""
import pytest
import torch

from tests.util import bf16r, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from tulip_b200 import ops
    return ops


def rnd(*shape, seed=0):
    gen = torch.Generator().manual_seed(seed)
    return bf16r(torch.randn(*shape, generator=gen))


def run_group(ops, shapes, bias=True, prefill=0.0, perm=None):
    probs, keep = [], []
    for i, (M, N, K) in enumerate(shapes):
        dy, x = rnd(M, N, seed=10 + i).cuda().bfloat16(), rnd(M, K, seed=20 + i).cuda().bfloat16()
        dW = torch.full((N, K), prefill, dtype=torch.float32, device="cuda")
        db = torch.full((N,), prefill, dtype=torch.float32, device="cuda") if bias else None
        d = dict(dY=dy, ldy=N, X=x, ldx=K, K1=K, M=M, N=N, K=K, dW=dW, lddw=K)
        if db is not None:
            d["db"] = db
        if perm and perm[i]:
            d["perm_R2"], d["perm_Cc"] = perm[i]
        probs.append(d)
        keep.append((dy, x, dW, db))
    ops.gemm_tn_group(probs)
    torch.cuda.synchronize()
    return keep


GROUPS = [
    [(4096, 96, 384), (4096, 384, 96)],                       # fc2 + fc1, stage 0 (C = 96)
    [(4096, 96, 96), (4096, 288, 96)],                        # proj + qkv, stage 0: 96 / 32 valid rows in the last row tile
    [(2048, 192, 768), (2048, 768, 192)],                     # stage 1
    [(1000, 384, 1536), (1000, 1536, 384)],                   # stage 2, token count not a multiple of the 64-token block
    [(256, 768, 3072), (256, 3072, 768)],                     # stage 3
    [(130, 768, 768), (130, 2304, 768)],
    [(8192, 1536, 96)],                                       # single problem (head expand)
    [(512, 96, 96), (512, 288, 96), (512, 96, 384), (512, 384, 96)],   # four problems
    [(64, 96, 96), (40, 288, 96)],                            # fewer tokens than one ring
]


@pytest.mark.parametrize("shapes", GROUPS, ids=lambda s: "+".join("x".join(map(str, t)) for t in s))
def test_tn_group_matches_float64(ops, shapes):
    for dy, x, dW, db in run_group(ops, shapes):
        assert rel_l2(dW.cpu(), dy.double().cpu().T @ x.double().cpu()) <= 1e-4
        assert rel_l2(db.cpu(), dy.double().cpu().sum(0)) <= 1e-4


def test_tn_group_accumulates_and_skips_bias(ops):
    shapes = [(2048, 96, 384), (2048, 384, 96)]
    for dy, x, dW, db in run_group(ops, shapes, bias=False, prefill=0.5):
        assert db is None
        assert rel_l2(dW.cpu(), 0.5 + dy.double().cpu().T @ x.double().cpu()) <= 1e-4


def test_tn_group_permuted_rows(ops):
    # rows n' = ij * Cc + c of the product land in row c * R2 + ij (weights stored shuffle-slot-major, tulip.py:117-123, 174-178)
    R2, Cc = 16, 96
    shapes = [(1024, R2 * Cc, 96), (1024, 96, 96)]
    out = run_group(ops, shapes, perm=[(R2, Cc), None])
    dy, x, dW, db = out[0]
    want = (dy.double().cpu().T @ x.double().cpu()).reshape(R2, Cc, 96).transpose(0, 1).reshape(R2 * Cc, 96)
    assert rel_l2(dW.cpu(), want) <= 1e-4
    assert rel_l2(db.cpu(), dy.double().cpu().sum(0).reshape(R2, Cc).T.reshape(-1)) <= 1e-4
    dy, x, dW, db = out[1]
    assert rel_l2(dW.cpu(), dy.double().cpu().T @ x.double().cpu()) <= 1e-4


def test_tn_group_agrees_with_single_launches(ops):
    # the executor's grouped launches against the stand-alone kernel (same operands, same fp32 atomics: not bitwise)
    shapes = [(32768, 192, 768), (32768, 768, 192)]
    for dy, x, dW, db in run_group(ops, shapes):
        dW1, db1 = ops.linear_wgrad(dy, x, impl=2)
        assert rel_l2(dW.cpu(), dW1.cpu()) <= 1e-5 and rel_l2(db.cpu(), db1.cpu()) <= 1e-5


def test_tn_group_rejects_gathered_operands(ops):
    dy, x = rnd(256, 96, seed=1).cuda().bfloat16(), rnd(256, 96, seed=2).cuda().bfloat16()
    dW = torch.zeros(96, 96, device="cuda")
    with pytest.raises(RuntimeError):
        ops.gemm_tn_group([dict(dY=dy, ldy=96, X=x, ldx=96, K1=96, M=256, N=96, K=96, dW=dW, lddw=96, y_mode=ops.A_UNSHUFFLE)])
