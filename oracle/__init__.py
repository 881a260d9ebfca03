"""CPU oracle for the MapAnything feed-forward inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `map-anything_b200/` may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` do, and there
only as the checker / the timed CPU reference, never as the product path.

What it is: a plain-PyTorch fp32 restatement of the reference's algorithm
(`mapanything/models/mapanything/model.py`, `mapanything/utils/{geometry,inference}.py`, the vendored
DINOv2 under `mapanything/models/external/dinov2/`, and the un-vendored `uniception` modules as specified
in SURVEY.md App. A).  Every function cites the reference file:line it follows.

Pinning status (see DESIGN.md "Oracle"):
  * PINNED against the reference's own code, imported from /root/reference by `oracle/make_golden.py`
    (fixtures in tests/golden/): the DINOv2 ViT (`oracle/vit.py`), all geometry math
    (`oracle/geometry.py`) and the infer pre/post-processing (`oracle/inference.py`).
  * PARITY UNPINNED: the `uniception` modules (`oracle/uniception_modules.py`) -- the package is not
    vendored, not installed and not downloadable here, and the reference ships no tests or golden
    vectors.  They follow the call contracts in model.py + the YAML hyper-parameters + SURVEY App. A.
"""
