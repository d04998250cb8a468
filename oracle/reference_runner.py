"""TEST INFRASTRUCTURE ONLY -- load the UNMODIFIED reference modules.

Looked up, in this order: $NAF_REFERENCE_ROOT, `oracle/_ref/` (the byte-for-byte copy made by
`oracle/build_ref.py`, git-ignored, which travels to the GPU box with the gpurun snapshot -- so
bench.py never reads /root/reference at run time) and /root/reference (the build container).  Used by `oracle/gen_golden*.py` to produce the committed
fixtures under ``tests/golden/``, by the ``-m "not gpu"`` tests that pin `oracle/naf_oracle.py`
against the reference, and by `bench.py`'s reference arm / `cpu_baseline` leg.  The reference's
NATTEN dependency is replaced by `oracle/natten_stub.py` (see that file for what is and is not
pinned).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    cands = [os.environ.get("NAF_REFERENCE_ROOT"), os.path.join(_HERE, "_ref"), "/root/reference"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "model", "naf.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "model", "naf.py"))


def is_vendored_copy() -> bool:
    return os.path.abspath(REFERENCE_ROOT) == os.path.join(_HERE, "_ref")


_cache = {}


def load():
    """Returns a namespace with the reference's NAF, CrossAttention, RoPE, encoder classes."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from oracle import natten_stub

    natten_stub.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # src/model/__init__.py imports every competitor model; optional deps print warnings.
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        import src.layers as ref_layers  # noqa: WPS433
        import src.layers.attentions as ref_att
        import src.layers.rope as ref_rope
        import src.model.naf as ref_naf

    class NS:
        NAF = ref_naf.NAF
        ImageEncoder = ref_naf.ImageEncoder
        CrossAttention = ref_att.CrossAttention
        RoPE = ref_rope.RoPE
        encoder = ref_layers.encoder
        natten_recent = ref_att.NATTEN_RECENT

    assert NS.natten_recent is False, "stub must select the legacy (na2d_qk/na2d_av) branch"
    _cache["ns"] = NS
    return NS
