"""Encoder-shape attention (V sequences of 1370 tokens, 16 heads x 64) through ma_attention_fwd, a few launches: the target of an
ncu capture.  python tools/prof_attn_enc.py [views]"""
import sys

import torch

sys.path.insert(0, "map-anything_b200")
from mapanything_b200 import ops  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, L = 16, 1370
D = H * 64
qkv = torch.randn(V * L, 3 * D, device="cuda").bfloat16()
o = torch.empty(V * L, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=V, q_len=L, kv_len=L)
torch.cuda.synchronize()
print("ok")
