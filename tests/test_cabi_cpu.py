"""CPU checks of the drop-in boundary: the C-ABI shared library loads on a GPU-less box and exports every symbol that
include/mapanything_b200.h declares (no compute calls here), the ctypes table matches the header, and the package refuses
to run without a CUDA device instead of falling back."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "mapanything_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ma_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mapanything_b200 import _lib

    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built (python map-anything_b200/build.py)")
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    lib.ma_abi_version.restype = ctypes.c_int
    assert lib.ma_abi_version() == 1
    lib.ma_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.ma_last_error(), bytes)


def test_ctypes_table_matches_header():
    from mapanything_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()


def test_struct_layouts_match_header_sizes():
    """ma_gemm_epilogue / ma_attn_ext field-for-field sizes (LP64): guards against a header edit without a binding edit."""
    from mapanything_b200 import _lib

    assert ctypes.sizeof(_lib.GemmEpilogue) == 8 + 8 + 4 + 4 + 8 + 8 + 8 + 8 + 4 + 4 + 8 + 8 + 4 * 4 + 3 * 8
    assert ctypes.sizeof(_lib.AttnExt) == 4 + 4 + 64 + 64 + 8 + 8 + 8 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(_lib.DecodeSpec) == 5 * 4 + 4   # ma_decode_spec: five ints + conf_vmin
    # the selector codes of the binding are the header's #defines
    hdr = HEADER.read_text()
    import re

    for name, code in {"MA_REP_POINTMAP": _lib.MA_REP["pointmap"], "MA_REP_RAYMAP_DEPTH": _lib.MA_REP["raymap+depth"],
                       "MA_REP_RAYDIRS_DEPTH_POSE": _lib.MA_REP["raydirs+depth+pose"],
                       "MA_REP_CAMPOINTMAP_POSE": _lib.MA_REP["campointmap+pose"],
                       "MA_REP_POINTMAP_RAYDIRS_DEPTH_POSE": _lib.MA_REP["pointmap+raydirs+depth+pose"],
                       "MA_PTS_LINEAR": _lib.MA_PTS["linear"], "MA_PTS_EXP": _lib.MA_PTS["exp"],
                       "MA_PTS_Z_EXP": _lib.MA_PTS["z_exp"]}.items():
        m = re.search(rf"#define {name} (\d+)", hdr)
        assert m and int(m.group(1)) == code, name


def test_bad_arguments_return_status_not_crash():
    """Argument validation happens on the host before any CUDA call: usable without a GPU."""
    from mapanything_b200 import _lib

    if not _lib.LIB_PATH.exists():
        pytest.skip("library not built")
    lib = _lib.load()
    rc = lib.ma_layernorm(None, 1, 0, None, 0, 0, None, None, 0, 0, 1e-6, 0, 0, 0, 0, 0, None)
    assert rc == -1
    assert b"ma_layernorm" in lib.ma_last_error()
    ep = _lib.GemmEpilogue()
    rc = lib.ma_gemm_bf16(None, 0, None, 0, 0, 0, 0, ctypes.byref(ep), 0, None)
    assert rc == -1
    spec = _lib.DecodeSpec(rep=7)
    rc = lib.ma_decode_scene(ctypes.byref(spec), None, 8, None, None, 1, 4, None, None, None, None, None, None, None, None, None,
                             None, None, None)
    assert rc == -1 and b"ma_decode_scene" in lib.ma_last_error()
    assert lib.ma_set_stream_k(-1) in (0, 1)   # query only


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from mapanything_b200 import MapAnything, tiny_config

    model = MapAnything(**tiny_config()).eval()
    views = [{"img": torch.randn(1, 3, 70, 70), "data_norm_type": ["dinov2"]} for _ in range(2)]
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(views)


def test_info_sharing_variants_construct_with_reference_key_layout():
    """3 tap indices (aat_ifr_48_layers*.yaml wiring, reference model.py:304-313) and no_ref_view: the parameter container
    builds on the CPU with the oracle's state-dict keys; unsupported variants raise the reference's ValueError."""
    import pytest

    from mapanything_b200 import MapAnything, tiny_config
    from mapanything_b200.config import mapanything_variant_config
    from oracle.config import tiny_config as oracle_tiny
    from oracle.model import MapAnythingOracle

    kw = dict(info_depth=6, indices=(1, 3, 4))
    m, o = MapAnything(**tiny_config(**kw)), MapAnythingOracle(**oracle_tiny(**kw))
    assert not m.use_encoder_features_for_dpt and set(m.state_dict()) == set(o.state_dict())
    with pytest.raises(ValueError, match="Please provide 2 or 3 indices"):
        MapAnything(**tiny_config(info_depth=6, indices=(1, 2, 3, 4)))
    cfg = mapanything_variant_config("aat_ifr_48_layers_no_ref_view")
    ma = cfg["info_sharing_config"]["module_args"]
    assert (ma["depth"], ma["dim"], ma["num_heads"], ma["indices"]) == (48, 1024, 16, [11, 23, 35])
    assert ma["distinguish_ref_and_non_ref_views"] is False
    gat = mapanything_variant_config("gat_ifr_24_layers")   # round 2: global attention in every block + view-index PE
    assert gat["info_sharing_config"]["model_type"] == "global_attention"
    m = MapAnything(**tiny_config())
    assert m.info_sharing.is_global(0) and not m.info_sharing.is_global(1)
    with pytest.raises(ValueError, match="info_sharing must be one of"):
        mapanything_variant_config("cat_ifr_dust3r")


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a gcc-built client (tests/cabi_host_check.c) links against the
    shared library and runs its host entry points."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest

        pytest.skip("no gcc")
    lib_dir = root / "map-anything_b200" / "mapanything_b200"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    str(root / "include" / "mapanything_b200.h")], check=True)
    exe = tmp_path / "cabi_host_check"
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", str(root / "include"), str(root / "tests" / "cabi_host_check.c"), "-o", str(exe),
                    "-L", str(lib_dir), "-l:libmapanything_b200.so", f"-Wl,-rpath,{lib_dir}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert "cabi host check ok: ksize 25" in res.stdout
