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
    mode = sys.argv[4] if len(sys.argv) > 4 else "img"
    multimodal = mode in ("mm", "mm0")  # mm0: only the views of rank 0's shard carry geometric inputs
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
    if multimodal:  # internal keys of forward(): unit rays, depth along ray, poses (view 0 = identity), metric scale
        from mapanything_b200.preprocess import preprocess_input_views_for_inference

        from mapanything_b200.sharding import partition_views as _pv

        n_geo = _pv(V, world)[0] if mode == "mm0" else V
        for i, v in enumerate(views):
            if i >= n_geo:
                continue
            f = 0.8 * size + 0.4 * size * float(torch.rand(1, generator=g))
            v["intrinsics"] = torch.tensor([[[f, 0, size / 2], [0, f, size / 2], [0, 0, 1.0]]], device=dev)
            if i != 2:
                v["depth_z"] = (1.0 + 3.0 * torch.rand(1, size, size, 1, generator=g)).to(dev)
            if i != 3:
                q = torch.randn(4, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 0.0, 1.0])
                t = torch.randn(3, generator=g)
                if i == 0:
                    q, t = torch.tensor([0.0, 0.0, 0.0, 1.0]), torch.zeros(3)
                v["camera_poses"] = ((q / q.norm())[None].to(dev), t[None].to(dev))
            v["is_metric_scale"] = torch.tensor([True], device=dev)
        views = preprocess_input_views_for_inference(views)
        model.geometric_input_config.update({"overall_prob": 1.0, "dropout_prob": 0.0, "ray_dirs_prob": 1.0, "depth_prob": 1.0,
                                             "cam_prob": 1.0})
    counts = partition_views(V, world)
    lo = sum(counts[:rank])
    mine = views[lo:lo + counts[rank]]

    full = model([dict(v) for v in views])  # the whole scene on this GPU alone
    model.enable_view_sharding()
    part = model([dict(v) for v in mine])
    part2 = model.infer([dict(v) for v in mine]) if not multimodal else part  # infer() over the shard, count exchange included
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
    assert len(part2) == len(mine) and (multimodal or "mask" in part2[0])
    t = torch.tensor([max(worst.values())], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("SHARDED_RESULT " + json.dumps({"worst_rel": t.item(), "per_key_rank0": worst, "counts": counts}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
