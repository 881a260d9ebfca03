"""Short single-purpose runs for `ncu --set full` captures (run under gpurun):  python tools/prof_kernels.py attn|gemm"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200 import ops  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "attn"
    V = 8
    if what == "attn":
        H, L = 12, 1369 * V + 1
        D = H * 64
        qkv = torch.randn(L, 3 * D, device="cuda").bfloat16()
        o = torch.empty(L, D, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=1, q_len=L, kv_len=L)
        # encoder shape too
        H, L = 16, 1370
        D = H * 64
        qkv = torch.randn(V * L, 3 * D, device="cuda").bfloat16()
        o = torch.empty(V * L, D, device="cuda", dtype=torch.bfloat16)
        for _ in range(2):
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=V, q_len=L, kv_len=L)
    else:
        M = 1370 * V
        for (n, k, f32res) in ((4096, 1024, False), (1024, 4096, True), (3072, 1024, False), (1024, 1024, True)):
            x = torch.randn(M, k, device="cuda").bfloat16()
            w = torch.randn(n, k, device="cuda").bfloat16()
            if f32res:
                out = torch.randn(M, n, device="cuda")
                b = torch.randn(n, device="cuda")
                for _ in range(2):
                    ops.gemm(x, w, out, bias=b, colscale=b, residual=out)
            else:
                out = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
                for _ in range(2):
                    ops.gemm(x, w, out)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
