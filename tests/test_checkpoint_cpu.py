"""CPU tests of the checkpoint surface the reference exposes (`MapAnything.from_pretrained`, HF layout config.json +
model.safetensors: reference scripts/gradio_app.py:70-71, mapanything/utils/hf_utils/hf_helpers.py:139-154) and of
tools/verify_checkpoint.py, which reads a checkpoint's tensor names / shapes and reports which App. A assumptions they decide."""
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def _tiny_model():
    from mapanything_b200 import MapAnything, tiny_config

    torch.manual_seed(0)
    return MapAnything(**tiny_config())


def test_save_pretrained_from_pretrained_round_trip(tmp_path):
    from mapanything_b200 import MapAnything

    m = _tiny_model()
    m.save_pretrained(tmp_path)
    assert (tmp_path / "config.json").exists() and (tmp_path / "model.safetensors").exists()
    m2 = MapAnything.from_pretrained(str(tmp_path))
    a, b = m.state_dict(), m2.state_dict()
    assert set(a) == set(b) and len(a) == 318
    assert all(torch.equal(a[k], b[k]) for k in a)
    # the constructor arguments travelled through config.json (the config dicts the ctor mutates are stored post-mutation)
    assert m2.info_sharing.dim == m.info_sharing.dim and m2.encoder.enc_embed_dim == m.encoder.enc_embed_dim
    assert m2.geometric_input_config["ray_dirs_encoder_config"]["patch_size"] == 14


def test_pth_checkpoint_path_ctor_argument(tmp_path):
    """reference model.py:590-620: pretrained_checkpoint_path = a torch.save'd dict with a "model" entry."""
    from mapanything_b200 import MapAnything, tiny_config

    m = _tiny_model()
    torch.save({"model": m.state_dict()}, tmp_path / "ckpt.pth")
    m2 = MapAnything(**tiny_config(), pretrained_checkpoint_path=str(tmp_path / "ckpt.pth"))
    assert all(torch.equal(v, m2.state_dict()[k]) for k, v in m.state_dict().items())


def test_verify_checkpoint_tool(tmp_path):
    m = _tiny_model()
    m.save_pretrained(tmp_path)
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "strict load OK" in r.stdout and "info.dim" in r.stdout
    # without config.json the configuration is inferred from the tensors alone
    (tmp_path / "config.json").unlink()
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # a checkpoint with a different regressor width is detected and still loads through the inferred config
    from mapanything_b200 import MapAnything, tiny_config

    cfg = tiny_config()
    cfg["pred_head_config"]["regressor_head"]["hidden_dims"] = [128, 32]
    torch.save({"model": MapAnything(**cfg).state_dict()}, tmp_path / "other.pth")
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "verify_checkpoint.py"), str(tmp_path / "other.pth")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "DIFFERS" in r.stdout and "[128, 32]" in r.stdout, r.stdout[-2000:]
