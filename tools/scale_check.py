"""Single-GPU scale check of forward() at large view counts (BASELINE config[4] is 1000 views, memory-efficient mode):

    python tools/scale_check.py 1000 [more counts ...]

Random-init full-size model, synthetic 518-px views created on the device, one forward per count (timed with CUDA events,
first-call effects included), peak memory, finiteness of the outputs.  One JSON line per count."""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200 import MapAnything, mapanything_config  # noqa: E402


def main():
    counts = [int(a) for a in sys.argv[1:]] or [250]
    torch.manual_seed(0)
    model = MapAnything(**mapanything_config()).to("cuda").eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    for V in counts:
        views = [{"img": torch.randn(1, 3, 518, 518, device="cuda", generator=g), "data_norm_type": ["dinov2"]} for _ in range(V)]
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        s.record()
        out = model.forward(views, memory_efficient_inference=True)
        e.record()
        torch.cuda.synchronize()
        sec = s.elapsed_time(e) / 1e3
        finite = all(bool(torch.isfinite(o["pts3d"]).all()) for o in out[:: max(1, V // 16)])
        tflop = V * (1013.6 + 467.3 + 69.1 + 69.1 * V + 177.7 + 131.2 + 35.5) / 1e3
        print(json.dumps({"views": V, "seconds": round(sec, 3), "wall_seconds": round(time.time() - t0, 3),
                          "views_per_s": round(V / sec, 2), "tflop": round(tflop, 1), "achieved_tflops": round(tflop / sec, 1),
                          "frac_of_sustained_peak": round(tflop / sec / 1401.9, 3),
                          "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "finite": finite}), flush=True)
        del out, views
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
