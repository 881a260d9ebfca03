import sys, torch
sys.path.insert(0, "map-anything_b200")
from mapanything_b200 import ops
V = int(sys.argv[1]) if len(sys.argv) > 1 else 24
H, L = 12, 1369 * V + 1
D = H * 64
qkv = torch.randn(L, 3 * D, device="cuda").bfloat16()
o = torch.empty(L, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=1, q_len=L, kv_len=L)
torch.cuda.synchronize()
print("ok")
