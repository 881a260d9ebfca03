"""Experimental ping-pong attention (MA_ATTN_PINGPONG flag) vs the default kernel: same results? faster?

    python tools/pingpong_check.py          # run under gpurun, with your own `timeout`: the variant is unmeasured

One JSON line per shape: max |diff| between the two kernels' outputs and TFLOP/s of each."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200 import ops  # noqa: E402


def timeit(fn, iters=8):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, (nseq, L, H) in {"small": (1, 300, 2), "global_8v": (1, 1369 * 8 + 1, 12), "encoder_8v": (8, 1370, 16)}.items():
        D = H * 64
        qkv = torch.randn(nseq * L, 3 * D, device="cuda", generator=g).bfloat16()
        outs = []
        res = {"shape": name}
        for pp in (False, True):
            o = torch.empty(nseq * L, D, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, num_heads=H, num_seqs=nseq, q_len=L,  # noqa: E731
                                       kv_len=L, pingpong=pp)
            ms = timeit(fn)
            outs.append(o.float())
            res["pingpong_tflops" if pp else "default_tflops"] = round(4 * nseq * H * L * L * 64 / ms / 1e9, 1)
        res["max_abs_diff"] = (outs[0] - outs[1]).abs().max().item()
        res["finite"] = bool(torch.isfinite(outs[1]).all())
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
