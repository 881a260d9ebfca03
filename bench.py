#!/usr/bin/env python
"""Benchmark of the MapAnything feed-forward inference hot path (BASELINE.json metric: views/sec @518 px).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--views V] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward pass over one synthetic V-view 518x518 image-only scene (random-init ViT-L + 24-layer
alternating attention + DPT + pose/scale heads), inputs resident in HBM -> per-view output dicts resident in HBM.
N = 1 runs BASELINE config[1] (8 views, bf16, 1xB200).  N > 1 (one process per GPU under torchrun): ONE scene of 8*N
views sharded by view (8 views per GPU); the 12 global-attention blocks all-gather K/V over NCCL, overlapped with the
attention over the local keys.  Per-GPU view count is fixed ("weak"), but the algorithmic work per view
grows with the scene (global attention is quadratic): config.tflop_per_view and roofline.step_frac carry the
FLOP-normalised picture.  `--multi replicas` runs N independent 8-view scenes instead.  Rank 0 prints ONE JSON line.

  value     views/sec of `model.forward` (device-resident inputs), CUDA events, max over ranks
  e2e       views/sec of `model.infer` from pinned HOST images (H2D inside the timed region, GPU post-processing,
            D2H of pts3d / conf / mask / poses / intrinsics / scale)
  roofline  dominant kernel family = the tcgen05 GEMM (every Linear + im2col'ed conv): algorithmic FLOPs (2MNK summed
            over its launches) / its summed launch durations, both taken live with CUDA events around each launch on
            the launching stream inside an instrumented pass of the same step; peak = MEASURED_PEAKS.json
  cpu_baseline  the fp32 CPU oracle (a "port": the reference's uniception dependency is not installable) on a bounded
            sample of the workload (2 of the views), all host threads
`--impl reference` times that CPU oracle alone, same metric/config, on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "map-anything_b200"))
sys.path.insert(0, str(ROOT))

IMG = 518
METRIC = "views_per_sec_518px"

# Algorithmic GFLOP per view (SURVEY.md section 8d / BASELINE.md section 3), D = 768, regressor hidden 128
F_ENC, F_IS_LIN, F_FRAME, F_GLOBAL_PER_VIEW, F_DPT, F_POSE = 1013.6, 467.3, 69.1, 69.1, 308.9, 35.5


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM (10960 x 4096 x 1024, bf16 out), from the
# `ncu --set full` capture profiles/r1_gemm_2cta_tma_epilogue.ncu-rep (summary: ..._ncu.csv, rows 1-2): 31.8 MB read +
# 45.2 MB written.  Algorithmic bytes of that launch: 22.4 MB (X) + 8.4 MB (W) + 89.8 MB (Y) = 120.6 MB; DRAM traffic is
# BELOW it because part of Y is still resident in the 126 MB L2 when the kernel ends -- no re-reads.
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 77.0e6
NCU_GEMM_TRAFFIC_NOTE = ("bytes per launch of gemm 10960x4096x1024 (ncu, profiles/r1_gemm_2cta_tma_epilogue_ncu.csv); "
                         "algorithmic 120.6 MB; the kernel is tensor-bound")


def gflop_per_view(v: int) -> float:
    return F_ENC + F_IS_LIN + F_FRAME + F_GLOBAL_PER_VIEW * v + F_DPT + F_POSE


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1401.9), d.get("hbm_gbs", 6542.7), "measured"
    return 1400.0, 6650.0, "fallback"  # B200_PROFILING.md fallback (sustained figure)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, pw, mx, reasons = [], [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                pw.append(float(r[2]))
                mx = max(mx, float(r[1]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_host_views(v: int, seed: int):
    import torch

    g = torch.Generator().manual_seed(seed)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return [((torch.rand(1, 3, IMG, IMG, generator=g) - mean) / std).pin_memory() for _ in range(v)]


def make_host_geometry(v: int, seed: int, first_view_identity: bool = True):
    """SURVEY 8d config C3 inputs per view: pinhole intrinsics (f ~ U[400,600], c = 259), depth_z ~ U[1,4] m,
    cam2world pose (view 0 identity), metric scale.  Pinned host tensors."""
    import torch

    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(v):
        f = float(torch.empty(1).uniform_(400, 600, generator=g))
        k = torch.tensor([[[f, 0, 259.0], [0, f, 259.0], [0, 0, 1.0]]])
        d = torch.empty(1, IMG, IMG, 1).uniform_(1.0, 4.0, generator=g)
        q = torch.randn(4, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 0.0, 1.0])
        t = torch.randn(3, generator=g)
        if i == 0 and first_view_identity:
            q, t = torch.tensor([0.0, 0.0, 0.0, 1.0]), torch.zeros(3)
        x, y, z, w = (q / q.norm()).tolist()
        pose = torch.eye(4)
        pose[:3, :3] = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        pose[:3, 3] = t
        out.append({"intrinsics": k.pin_memory(), "depth_z": d.pin_memory(), "camera_poses": pose[None].pin_memory(),
                    "is_metric_scale": torch.tensor([True]).pin_memory()})
    return out


def run_reference(args):
    """The reference arm: the fp32 CPU oracle of the path, all host threads, bounded sample of the workload."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.config import mapanything_config
    from oracle.model import MapAnythingOracle
    from oracle.weights import init_reference_style

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_views = 2
    model = init_reference_style(MapAnythingOracle(**mapanything_config()).eval(), 0)
    g = torch.Generator().manual_seed(1234)
    views = [{"img": torch.randn(1, 3, IMG, IMG, generator=g), "data_norm_type": ["dinov2"]} for _ in range(sample_views)]

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            model(views)
        return time.perf_counter() - t0

    t_first = step()  # warm-up (also sizes the run: the whole arm must end within a few minutes)
    budget = 200.0
    warm = max(0, min(args.warmup - 1, int(0.25 * budget / max(t_first, 1e-3))))
    for _ in range(warm):
        step()
    steps = max(1, min(args.steps, int(0.75 * budget / max(t_first, 1e-3))))
    times = [step() for _ in range(steps)]
    t = sum(times) / len(times)
    vps = sample_views / t
    sample = (f"{sample_views} of the {args.views} views (518x518, full ViT-L + 24-layer info sharing + DPT), fp32, "
              f"{steps} timed step(s) of {t:.2f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": "views/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm + 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.views, args.gpus, args.multi), "views": args.views * args.gpus,
                   "views_per_gpu": args.views, "image": IMG, "l2": "inputs >> L2"},
        "cpu_baseline": {"value": vps, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": vps, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(v: int, gpus: int, multi: str = "shard", multimodal: bool = False) -> str:
    if multimodal:
        return (f"MapAnything multi-modal (images+intrinsics+poses+depth), {v * gpus} views 518x518 bf16 on {gpus}xB200"
                + (" (BASELINE config[2])" if v * gpus == 24 else ""))
    if gpus == 1:
        return f"MapAnything image-only, {v} views 518x518 bf16 on 1xB200 (BASELINE config[1])" if v == 8 else \
            f"MapAnything image-only, {v} views 518x518 bf16 on 1xB200"
    if multi == "replicas":
        return f"MapAnything image-only, {v} views 518x518 bf16 per GPU, {gpus} independent scenes on {gpus}xB200"
    return (f"MapAnything image-only, ONE scene of {v * gpus} views 518x518 bf16 sharded by view over {gpus}xB200 "
            f"({v} views per GPU; global attention = local queries x all-gathered K/V over NCCL); algorithmic work per view "
            f"grows with the scene: {gflop_per_view(v * gpus) / 1e3:.2f} TFLOP/view vs {gflop_per_view(v) / 1e3:.2f} at {v} views")


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mapanything_b200 import MapAnything, mapanything_config, ops
    from mapanything_b200 import _lib

    _lib.load()  # fail loudly if the CUDA library is missing
    torch.manual_seed(0)
    model = MapAnything(**mapanything_config()).to(dev).eval()  # random-init weights of the architecture
    V = args.views  # views per GPU
    shard = world > 1 and args.multi == "shard"
    v_scene = V * world if shard else V  # views of the scene each forward pass works on
    host_imgs = make_host_views(V, 1234 + rank)
    host_extra = make_host_geometry(V, 4321 + rank, first_view_identity=(rank == 0)) if args.multimodal else [{} for _ in range(V)]
    dev_views = [{"img": im.to(dev), "data_norm_type": ["dinov2"]} for im in host_imgs]
    if args.multimodal:
        # forward() takes the model's internal keys: preprocess once (device resident), enable the geometric inputs
        from mapanything_b200.preprocess import preprocess_input_views_for_inference

        dev_views = preprocess_input_views_for_inference(
            [{**v, **{k: x.to(dev) for k, x in e.items()}} for v, e in zip(dev_views, host_extra)])
        model.geometric_input_config.update({"overall_prob": 1.0, "dropout_prob": 0.0, "ray_dirs_prob": 1.0, "depth_prob": 1.0,
                                             "cam_prob": 1.0})
    model.engine()
    if shard:
        model.enable_view_sharding(views_per_rank=[V] * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fwd_step():
        return model(dev_views)

    d2h_keys = ("pts3d", "conf", "mask", "camera_poses", "intrinsics", "metric_scaling_factor")

    def e2e_step():
        views = [{"img": im, "data_norm_type": ["dinov2"], **e} for im, e in zip(host_imgs, host_extra)]  # pinned HOST tensors
        preds = model.infer(views)
        outs = [p[k].to("cpu", non_blocking=True) for p in preds for k in d2h_keys]
        torch.cuda.current_stream().synchronize()
        return outs

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    if args.profile_mode:  # short, un-reported run for ncu: one warm-up step, one step, nothing else
        fwd_step()
        torch.cuda.synchronize()
        fwd_step()
        torch.cuda.synchronize()
        if rank == 0:
            print(json.dumps({"profile_mode": True, "views": V, "launches_per_step": ops.LAUNCHES // 2}))
        return
    for _ in range(max(args.warmup, 3)):
        fwd_step()
    with ClockSampler(local_rank) as clocks:
        launches0 = ops.LAUNCHES
        ms_fwd = timed(fwd_step, args.steps)
        launches = ops.LAUNCHES - launches0
        # instrumented pass of the same step for the per-kernel roofline (events around every GEMM launch)
        # (single stream for this pass: with the encoder's two view groups on two streams the kernels overlap and the
        # per-launch event times would count the time spent waiting for SMs)
        eng = model.engine()
        n_streams, eng.encoder_streams = eng.encoder_streams, 1
        ops.PROFILE = []
        barrier()
        fwd_step()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        eng.encoder_streams = n_streams
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
    ms_step = ms_fwd / args.steps
    vps = world * V / (ms_step * 1e-3)
    e2e_ms_step = ms_e2e / args.steps
    e2e_vps = world * V / (e2e_ms_step * 1e-3)

    peak_tf, peak_hbm, peak_src = measured_peaks()
    if os.environ.get("MA_BENCH_DUMP") and rank == 0:  # per-launch list (family, algorithmic flops, ms) for offline analysis
        Path(os.environ["MA_BENCH_DUMP"]).write_text(json.dumps([[n, f, s.elapsed_time(e), t] for n, f, s, e, t in prof]))
    fam = {}
    for name, flops, s, e, _tag in prof:
        d = fam.setdefault(name, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += s.elapsed_time(e)
        d[2] += 1
    gemm = [a + b for a, b in zip(fam.get("gemm", [0.0, 1e-9, 0]), fam.get("conv3x3", [0.0, 0.0, 0]))]  # one kernel
    achieved = gemm[0] / (gemm[1] * 1e-3) / 1e12
    step_tf = V * gflop_per_view(v_scene) * 1e9 / (ms_step * 1e-3) / 1e12  # per GPU

    outs = e2e_step()  # every rank: the sharded step is collective
    d2h = sum(o.numel() * o.element_size() for o in outs) * world
    h2d = (sum(im.numel() * 4 for im in host_imgs)
           + sum(x.numel() * x.element_size() for e in host_extra for x in e.values())) * world
    if rank == 0:
        cpu = cpu_baseline(V) if world == 1 else None  # the host baseline is reported at N = 1 only
        line = {
            "metric": METRIC, "value": vps, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(V, world, args.multi, args.multimodal), "views": V * world, "views_per_gpu": V,
                       "scene_views": v_scene, "tflop_per_view": gflop_per_view(v_scene) / 1e3, "image": IMG,
                       "weights": "random-init", "parallelism": "single" if world == 1 else
                       (f"view-shard x{world} + K/V all-gather" if shard else f"replicas x{world}"),
                       "l2": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush"},
            "roofline": {
                "bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH,
                "traffic_note": NCU_GEMM_TRAFFIC_NOTE, "peak_source": peak_src,
                "launches": gemm[2], "kernel_ms_per_step": gemm[1],
                "gemm_share_of_step": gemm[1] / ms_step,
                "step_achieved_tflops": step_tf, "step_frac": step_tf / peak_tf,
                "families_ms": {k: round(v[1], 3) for k, v in fam.items()},
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_vps, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(v: int):
    """fp32 CPU oracle on a bounded sample (2 views of the workload), all host threads."""
    import torch

    from oracle.config import mapanything_config
    from oracle.model import MapAnythingOracle
    from oracle.weights import init_reference_style

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = init_reference_style(MapAnythingOracle(**mapanything_config()).eval(), 0)
    g = torch.Generator().manual_seed(1234)
    views = [{"img": torch.randn(1, 3, IMG, IMG, generator=g), "data_norm_type": ["dinov2"]} for _ in range(2)]
    t0 = time.perf_counter()
    with torch.no_grad():
        model(views)
    t = time.perf_counter() - t0
    return {"value": 2 / t, "unit": "views/s", "cores": cores, "kind": "port",
            "sample": f"2 of the {v} views, one fp32 forward of the full model on the host ({t:.1f} s)"}


def main():
    # rank 0 prints exactly ONE line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--views", type=int, default=8, help="views per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--multimodal", action="store_true",
                    help="BASELINE config[2]: every view also carries intrinsics, depth and a camera pose (use with --views 24)")
    ap.add_argument("--multi", default="shard", choices=["shard", "replicas"],
                    help="N > 1: one scene of views*N views sharded by view (default) or N independent scenes")
    ap.add_argument("--profile-mode", action="store_true", help="1 warm-up + 1 step only, for ncu captures (prints no bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every host core (set before torch is imported)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
