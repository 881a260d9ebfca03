"""torchrun worker for tests/test_sharded_gpu.py: one scene sharded by view over WORLD_SIZE GPUs vs the same scene on one GPU."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "map-anything_b200"))
sys.path.insert(0, str(ROOT))


def main():
    cfg_name, V, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import mapanything_b200 as mb
    from mapanything_b200.sharding import partition_views

    cfg = getattr(mb, cfg_name)()
    torch.manual_seed(0)  # same random-init weights on every rank
    model = mb.MapAnything(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(99)
    views = [{"img": torch.randn(1, 3, size, size, generator=g).to(dev), "data_norm_type": ["dinov2"]} for _ in range(V)]
    counts = partition_views(V, world)
    lo = sum(counts[:rank])
    mine = views[lo:lo + counts[rank]]

    full = model([dict(v) for v in views])  # the whole scene on this GPU alone
    model.enable_view_sharding()
    part = model([dict(v) for v in mine])
    part2 = model.infer([dict(v) for v in mine])  # infer() over the shard, count exchange included
    model.disable_view_sharding()
    torch.cuda.synchronize()

    worst = {}
    for i, p in enumerate(part):
        f = full[lo + i]
        for k in ("pts3d", "depth_along_ray", "ray_directions", "conf", "cam_trans", "cam_quats", "metric_scaling_factor"):
            a, b = p[k].float(), f[k].float()
            rel = ((a - b).norm() / b.norm().clamp(min=1e-12)).item()
            worst[k] = max(worst.get(k, 0.0), rel)
        assert torch.isfinite(p["pts3d"]).all()
    assert len(part2) == len(mine) and "mask" in part2[0]
    t = torch.tensor([max(worst.values())], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("SHARDED_RESULT " + json.dumps({"worst_rel": t.item(), "per_key_rank0": worst, "counts": counts}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
