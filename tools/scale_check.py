"""Scale check of forward() at large view counts (BASELINE config[4]: 1000 views, memory-efficient mode).

    python tools/scale_check.py 1000 [more counts ...]                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/scale_check.py 1000                                                       # one scene sharded by view

Random-init full-size model, synthetic 518-px views created on the device, ONE timed forward per count (CUDA events,
barrier on both sides, max over ranks) after a small warm-up scene (kernel / NCCL initialisation), peak memory per rank,
finiteness of the outputs.  Rank 0 prints one JSON line per count."""
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
from mapanything_b200 import MapAnything, mapanything_config  # noqa: E402
from mapanything_b200.sharding import partition_views  # noqa: E402


def main():
    counts = [int(a) for a in sys.argv[1:]] or [250]
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    model = MapAnything(**mapanything_config()).to(f"cuda:{local}").eval()
    g = torch.Generator(device="cuda").manual_seed(1 + rank)

    def make(n):
        return [{"img": torch.randn(1, 3, 518, 518, device="cuda", generator=g), "data_norm_type": ["dinov2"]} for _ in range(n)]

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(V, timed):
        mine = partition_views(V, world)[rank] if world > 1 else V
        if world > 1:
            model.enable_view_sharding(views_per_rank=partition_views(V, world))
        views = make(mine)
        sync()
        torch.cuda.reset_peak_memory_stats()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        s.record()
        out = model.forward(views, memory_efficient_inference=True)
        e.record()
        sync()
        sec = s.elapsed_time(e) / 1e3
        stats = torch.tensor([sec, torch.cuda.max_memory_allocated() / 2 ** 30,
                              float(all(bool(torch.isfinite(o["pts3d"]).all()) for o in out[:: max(1, mine // 16)]))], device="cuda")
        if dist is not None:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(stats, op=dist.ReduceOp.MIN)
            sec, mem, finite = mx[0].item(), mx[1].item(), stats[2].item() > 0
        else:
            sec, mem, finite = stats[0].item(), stats[1].item(), stats[2].item() > 0
        if timed and rank == 0:
            tflop = V * (1013.6 + 467.3 + 69.1 + 69.1 * V + 177.7 + 131.2 + 35.5) / 1e3
            print(json.dumps({"views": V, "n_gpus": world, "views_per_gpu": partition_views(V, world) if world > 1 else [V],
                              "seconds": round(sec, 3), "wall_seconds": round(time.time() - t0, 3),
                              "views_per_s": round(V / sec, 2), "tflop": round(tflop, 1),
                              "achieved_tflops_per_gpu": round(tflop / sec / world, 1),
                              "frac_of_sustained_peak": round(tflop / sec / world / 1401.9, 3),
                              "peak_mem_gib_per_gpu": round(mem, 2), "finite": finite,
                              "mode": "memory_efficient_inference=True" + (", view-sharded + K/V all-gather" if world > 1 else "")}),
                  flush=True)
        del out, views
        torch.cuda.empty_cache()

    run(2 * world, timed=False)  # warm-up scene
    for V in counts:
        run(V, timed=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
