"""Builds the C-ABI CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python map-anything_b200/build.py [--force] [--verbose]

Output: map-anything_b200/mapanything_b200/libmapanything_b200.so  (git-ignored; travels with gpurun).
Objects are cached under map-anything_b200/build/ and rebuilt when a source or header is newer.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT = ROOT / "mapanything_b200" / "libmapanything_b200.so"
OBJ = ROOT / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_header() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "mapanything_b200.h"]
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdr_time = _newest_header()
    jobs = []
    for src in srcs:
        obj = OBJ / (src.stem + ".o")
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time)
        if stale:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for src, res in pool.map(compile_one, jobs):
                if verbose or res.returncode != 0:
                    sys.stderr.write(f"--- nvcc {src.name}\n{res.stdout}{res.stderr}\n")
                if res.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
                (OBJ / (src.stem + ".ptxas.log")).write_text(res.stderr)

    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if jobs or not OUT.exists():
        cmd = [_nvcc(), "-shared", "-o", str(OUT), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
