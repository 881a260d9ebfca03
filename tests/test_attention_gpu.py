"""GPU parity tests for the tcgen05 flash attention (ma_attention_fwd) against fp32 PyTorch softmax attention."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, nseq, L, H, q_stride=None):
    D = H * 64
    q_stride = q_stride or L
    outs = []
    for s in range(nseq):
        blk = qkv[s * q_stride : s * q_stride + L].float()
        q, k, v = blk[:, :D], blk[:, D : 2 * D], blk[:, 2 * D :]
        q = q.view(L, H, 64).transpose(0, 1)
        k = k.view(L, H, 64).transpose(0, 1)
        v = v.view(L, H, 64).transpose(0, 1)
        a = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1)
        outs.append((a @ v).transpose(0, 1).reshape(L, D))
    return outs


@pytest.mark.parametrize("nseq,L,H", [(1, 128, 1), (1, 256, 2), (2, 1370, 16), (3, 1369, 12), (1, 2739, 12), (1, 100, 4)])
def test_attention_self(nseq, L, H):
    from mapanything_b200 import ops

    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(nseq * 1000 + L + H)
    qkv = (torch.randn(nseq * L, 3 * D, device="cuda", generator=g) * 1.5).bfloat16()
    out = torch.full((nseq * L, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv[:, :D], qkv[:, D : 2 * D], qkv[:, 2 * D :], out, num_heads=H, num_seqs=nseq, q_len=L, kv_len=L)
    torch.cuda.synchronize()
    refs = _ref(qkv, nseq, L, H)
    for s in range(nseq):
        got = out[s * L : (s + 1) * L].float()
        assert torch.isfinite(got).all(), f"non-finite output in seq {s}"
        err = (got - refs[s]).abs().max().item()
        assert err < 2e-2, f"seq {s}: max abs err {err}"


def test_attention_frame_inside_global_buffer():
    """Frame attention over V views stored inside a (V*N+1)-row buffer: strided sequences, trailing extra token untouched."""
    from mapanything_b200 import ops

    V, N, H = 3, 1369, 12
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(V * N + 1, 3 * D, device="cuda", generator=g).bfloat16()
    out = torch.zeros(V * N + 1, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv[:, :D], qkv[:, D : 2 * D], qkv[:, 2 * D :], out, num_heads=H, num_seqs=V, q_len=N, kv_len=N)
    refs = _ref(qkv, V, N, H)
    for s in range(V):
        err = (out[s * N : (s + 1) * N].float() - refs[s]).abs().max().item()
        assert err < 2e-2, f"view {s}: {err}"
    assert out[-1].abs().max().item() == 0.0


def test_attention_cross_lengths():
    """q_len != kv_len with separate q / kv buffers (the sequence-parallel global-attention shape)."""
    from mapanything_b200 import ops

    H, Lq, Lk = 12, 700, 2739
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(6)
    q = torch.randn(Lq, D, device="cuda", generator=g).bfloat16()
    kv = torch.randn(Lk, 2 * D, device="cuda", generator=g).bfloat16()
    out = torch.empty(Lq, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, kv[:, :D], kv[:, D:], out, num_heads=H, num_seqs=1, q_len=Lq, kv_len=Lk)
    qf = q.float().view(Lq, H, 64).transpose(0, 1)
    kf = kv[:, :D].float().view(Lk, H, 64).transpose(0, 1)
    vf = kv[:, D:].float().view(Lk, H, 64).transpose(0, 1)
    ref = (torch.softmax(qf @ kf.transpose(1, 2) / 8.0, -1) @ vf).transpose(0, 1).reshape(Lq, D)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
