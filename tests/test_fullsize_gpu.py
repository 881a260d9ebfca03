"""Full-size GPU parity at the sizes BASELINE.json benchmarks (round-1 verdict: parity existed at V = 2 only):
config[1] (8 views, image-only) and config[2] (24 views, images + intrinsics + poses + depth), 518 px, ViT-L + 24-layer
alternating attention + DPT, reference-style random-init weights.

The fp32 CPU oracle needs minutes at these sizes, so the yardstick here is the SAME oracle modules moved to the GPU and run in
fp32 with TF32 off (SURVEY 8c tier 1) -- after checking, at V = 2, that this GPU-fp32 oracle reproduces the CPU oracle
(test_gpu_fp32_oracle_matches_cpu_oracle).  Attention of the GPU oracle goes through F.scaled_dot_product_attention (what the
reference calls; the explicit softmax of the CPU form would materialise a 32857^2 matrix per head at 24 views).

Bound: the north_star tolerances as stated, or 2x the reference's OWN bf16-autocast distance from fp32 on the same inputs
(the AMP oracle) where that floor is already above the stated figure.  The actual numbers are printed (pytest -s) and
recorded in DESIGN.md section 4.
"""
import contextlib

import pytest
import torch

from test_geometric_gpu import _cuda, _multimodal_views
from test_model_gpu import _assert_within, _build, _fmt, _metrics, _views

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def _gpu_fp32_oracle(oracle):
    """The oracle on cuda:0 in strict fp32 (no TF32 anywhere), attention through SDPA."""
    import oracle.vit as ov

    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, ov.USE_SDPA,
             torch.get_float32_matmul_precision())
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    ov.USE_SDPA = True
    try:
        yield oracle.cuda()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, ov.USE_SDPA = saved[:3]
        torch.set_float32_matmul_precision(saved[3])
        oracle.cpu()
        torch.cuda.empty_cache()


def _cpu(preds):
    return [{k: v.cpu() for k, v in p.items()} for p in preds]


def test_gpu_fp32_oracle_matches_cpu_oracle():
    """V = 2, full size: the GPU-fp32 form of the oracle (cuBLAS fp32, cuDNN fp32, SDPA) against the CPU oracle the golden
    vectors pin.  Two fp32 evaluations of a 70-layer network differ by summation order only."""
    from oracle.config import mapanything_config

    oracle, _ = _build(mapanything_config, seed=0, init="reference")
    views = _views(2, 518, seed=1234 + 1)
    with torch.no_grad():
        ref = oracle([dict(v) for v in views])
        with _gpu_fp32_oracle(oracle) as og:
            got = _cpu(og(_cuda(views)))
    m = _metrics(got, ref)
    print(f"\n[GPU-fp32 oracle vs CPU oracle, V=2] {_fmt(m)}")
    assert m["depth_rel_p99"] < 1e-4 and m["pts_rel_p99"] < 1e-4 and m["rot_deg"] < 1e-3 and m["scale_rel"] < 1e-4, m
    assert m["depth_rel"] < 2e-3 and m["conf_rel"] < 2e-3, m


def test_forward_full_size_config1_eight_views():
    """BASELINE config[1]: image-only, 8 views 518x518 -- the configuration bench.py times."""
    from oracle.config import mapanything_config

    oracle, model = _build(mapanything_config, seed=0, init="reference")
    views = _views(8, 518, seed=1234)
    with torch.no_grad(), _gpu_fp32_oracle(oracle) as og:
        ref = _cpu(og(_cuda(views)))
        amp = _cpu(og(_cuda(views), amp_bf16=True))
    got = model(_cuda(views))
    _assert_within(_metrics(got, ref), "full-size config[1] (V=8), reference-style init", floor=_metrics(amp, ref))


def test_forward_full_size_config2_multimodal_24_views():
    """BASELINE config[2]: images + intrinsics + poses + depth, 24 views 518x518 (one view without depth, one without pose)."""
    from mapanything_b200.preprocess import preprocess_input_views_for_inference
    from oracle import inference as I
    from oracle.config import mapanything_config

    oracle, model = _build(mapanything_config, seed=0, init="reference")
    views = _multimodal_views(24, 518, seed=61, drop_depth=(5,), drop_pose=(7,))
    on = {"overall_prob": 1.0, "dropout_prob": 0.0, "ray_dirs_prob": 1.0, "depth_prob": 1.0, "cam_prob": 1.0}
    oracle.geometric_input_config.update(on)
    model.geometric_input_config.update(on)
    with torch.no_grad(), _gpu_fp32_oracle(oracle) as og:
        pv_ref = I.preprocess_views(_cuda(views))
        ref = _cpu(og(pv_ref))
        amp = _cpu(og(pv_ref, amp_bf16=True))
        del pv_ref
    got = model(preprocess_input_views_for_inference(_cuda(views)))
    _assert_within(_metrics(got, ref), "full-size config[2] (V=24, multi-modal), reference-style init", floor=_metrics(amp, ref))
