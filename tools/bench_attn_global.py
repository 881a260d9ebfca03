"""Global-attention shape (1 sequence, 12 heads x 64, 1369 * V + 1 tokens) through ma_attention_fwd: TFLOP/s per view count and the
largest deviation from F.scaled_dot_product_attention at the first one.  python tools/bench_attn_global.py 8 16 24
(MAPANYTHING_B200_LIB selects a library variant, tools/build_variants.sh.)"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, "map-anything_b200")
from mapanything_b200 import ops  # noqa: E402

H, D = 12, 768
out = {"lib": os.environ.get("MAPANYTHING_B200_LIB", "default")}
for n, V in enumerate(int(a) for a in sys.argv[1:]):
    L = 1369 * V + 1
    torch.manual_seed(V)
    qkv = torch.randn(L, 3 * D, device="cuda").bfloat16()
    o = torch.empty(L, D, device="cuda", dtype=torch.bfloat16)

    def run():
        ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=1, q_len=L, kv_len=L)

    for _ in range(3):
        run()
    if n == 0:
        q, k, v = (qkv[:, i * D:(i + 1) * D].view(L, H, 64).transpose(0, 1)[None] for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v)[0].transpose(0, 1).reshape(L, D)
        out["max_abs_diff_vs_sdpa"] = (o.float() - ref.float()).abs().max().item()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 10)
    ms = sorted(ts)[len(ts) // 2]
    out[f"v{V}_tflops"] = round(4.0 * L * L * 64 * H / ms / 1e9, 1)
print(json.dumps(out), flush=True)
