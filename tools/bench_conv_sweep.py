"""Implicit-GEMM 3x3 conv: time against the input-channel count at fixed output shape.  The intercept of the line is the per-tile
floor that does not depend on the main loop (epilogue, tile scheduling), the slope the main-loop cost per 64-channel block.
python tools/bench_conv_sweep.py"""
import json
import sys

import torch

sys.path.insert(0, "map-anything_b200")
from mapanything_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / n)
    return sorted(ts)[len(ts) // 2]


for (n, H, W, Co), cins in {(2, 518, 518, 128): (64, 128, 256), (4, 296, 296, 128): (64, 128, 256, 512),
                            (8, 148, 148, 256): (64, 96, 128, 256, 512), (8, 148, 148, 128): (64, 128, 256)}.items():
    for C in cins:
        x = torch.randn(n, H, W, C, device="cuda").bfloat16()
        w = torch.randn(Co, 9 * C, device="cuda").bfloat16()
        o = torch.empty(n * H * W, Co, device="cuda", dtype=torch.bfloat16)
        o2 = torch.empty_like(o)
        t = timeit(lambda: ops.conv3x3(x, w, o))
        t2 = timeit(lambda: ops.conv3x3(x, w, o, out_relu=o2))
        fl = 2.0 * n * H * W * Co * 9 * C
        print(json.dumps({"shape": f"{n}x{H}x{W}x{C}->{Co}", "us": round(t * 1e3, 1), "tflops": round(fl / t / 1e9, 1),
                          "us_two_outputs": round(t2 * 1e3, 1)}), flush=True)
