"""Programmatic dependent launch (ma_set_pdl): overlapping a kernel's prologue with its predecessor's tail must not change
a single bit of the results -- every chained kernel waits (griddepcontrol.wait) before its first global memory access."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pdl_restore():
    from mapanything_b200 import ops

    before = ops.set_pdl(None)
    sk_before = ops.set_stream_k(False)   # bitwise comparisons: the stream-K reduce-add order is not bit reproducible
    yield ops
    ops.set_pdl(before)
    ops.set_stream_k(sk_before)


def _chain(ops, x0, w1, w2, g, b, n_iter):
    """LayerNorm -> GEMM(GELU) -> GEMM(residual reduce-add) repeated on one stream: each kernel consumes what the previous
    one wrote, and the residual stream is updated in place (the hazard pattern of a transformer block)."""
    x = x0.clone()
    h = torch.empty(x.shape[0], x.shape[1], device="cuda", dtype=torch.bfloat16)
    f = torch.empty(x.shape[0], w1.shape[0], device="cuda", dtype=torch.bfloat16)
    for _ in range(n_iter):
        ops.layernorm(x, h, g, b)
        ops.gemm(h, w1, f, act=ops.MA_ACT_GELU)
        ops.gemm(f, w2, x, residual=x)
    torch.cuda.synchronize()
    return x


@pytest.mark.parametrize("rows", [300, 10953])
def test_pdl_chain_bitwise(pdl_restore, rows):
    ops = pdl_restore
    g = torch.Generator(device="cuda").manual_seed(3)
    C, Hd = 768, 3072
    x0 = torch.randn(rows, C, device="cuda", generator=g)
    w1 = (torch.randn(Hd, C, device="cuda", generator=g) * C ** -0.5).bfloat16()
    w2 = (torch.randn(C, Hd, device="cuda", generator=g) * (0.1 * Hd ** -0.5)).bfloat16()
    gam = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    bet = 0.1 * torch.randn(C, device="cuda", generator=g)
    ops.set_pdl(False)
    ref = _chain(ops, x0, w1, w2, gam, bet, 12)
    ops.set_pdl(True)
    for _ in range(3):  # repeated: a race would not show on every run
        got = _chain(ops, x0, w1, w2, gam, bet, 12)
        assert torch.equal(got, ref)
    assert torch.isfinite(ref).all()


def test_pdl_attention_chain_bitwise(pdl_restore):
    """qkv GEMM -> attention -> proj GEMM (in-place residual), the other hazard chain of a block."""
    ops = pdl_restore
    g = torch.Generator(device="cuda").manual_seed(5)
    H, L, nseq = 12, 1369, 3
    D = H * 64
    x0 = torch.randn(nseq * L, D, device="cuda", generator=g)
    wqkv = (torch.randn(3 * D, D, device="cuda", generator=g) * D ** -0.5).bfloat16()
    wproj = (torch.randn(D, D, device="cuda", generator=g) * (0.2 * D ** -0.5)).bfloat16()
    gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")

    def run():
        x = x0.clone()
        h = torch.empty(nseq * L, D, device="cuda", dtype=torch.bfloat16)
        qkv = torch.empty(nseq * L, 3 * D, device="cuda", dtype=torch.bfloat16)
        a = torch.empty(nseq * L, D, device="cuda", dtype=torch.bfloat16)
        for _ in range(6):
            ops.layernorm(x, h, gam, bet)
            ops.gemm(h, wqkv, qkv)
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], a, num_heads=H, num_seqs=nseq, q_len=L, kv_len=L,
                          q_seq_stride=L, kv_seq_stride=L)
            ops.gemm(a, wproj, x, residual=x)
        torch.cuda.synchronize()
        return x

    ops.set_pdl(False)
    ref = run()
    ops.set_pdl(True)
    for _ in range(3):
        assert torch.equal(run(), ref)


def test_pdl_model_bitwise(pdl_restore):
    """Whole tiny model through infer(): identical outputs with and without programmatic dependent launch."""
    ops = pdl_restore
    from mapanything_b200 import MapAnything, tiny_config
    from oracle.config import tiny_config as oracle_tiny
    from oracle.model import MapAnythingOracle
    from oracle.weights import init_reference_style

    oracle = init_reference_style(MapAnythingOracle(**oracle_tiny()).eval(), 0)
    model = MapAnything(**tiny_config())
    model.load_state_dict(oracle.state_dict(), strict=True)
    model = model.to("cuda").eval()
    g = torch.Generator().manual_seed(11)
    views = [{"img": torch.randn(1, 3, 70, 98, generator=g), "data_norm_type": ["dinov2"]} for _ in range(4)]
    ops.set_pdl(False)
    ref = model.infer([dict(v) for v in views])
    ops.set_pdl(True)
    got = model.infer([dict(v) for v in views])
    torch.cuda.synchronize()
    for a, b in zip(got, ref):
        for k in b:
            if torch.is_tensor(b[k]):
                assert torch.equal(a[k], b[k]), k
