"""Import the UNMODIFIED reference (`scl`) for oracle pinning and for the bench's Python-reference leg.

TEST / BENCH INFRASTRUCTURE ONLY (the product never imports this).  The reference tree is read from
/root/reference in the build container.  /root/reference does not exist on the GPU box, so
`stage_reference()` -- called by __graft_entry__.build() while the tree is present -- leaves a
byte-identical, read-only copy of its `scl/` package under oracle/_ref/pyref/ (git-ignored like the
rest of oracle/_ref, so never in history, but it travels with the repo snapshot): that copy is what
bench.py's `cpu_baseline.python_reference` leg times on the GPU box's host cores.  The `-m gpu`
tests and smoke() do not use it.
"""
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "bitarray_shim")
_SOURCE_ROOT = os.environ.get("SCL_REFERENCE_ROOT", "/root/reference")
_STAGED_ROOT = os.path.join(_HERE, "_ref", "pyref")


def _has_tree(root):
    return os.path.isdir(os.path.join(root, "scl", "compressors"))


REFERENCE_ROOT = _SOURCE_ROOT if _has_tree(_SOURCE_ROOT) else _STAGED_ROOT


def stage_reference() -> str:
    """Copy <reference>/scl (unmodified) to oracle/_ref/pyref/scl when the source tree is present.
    Returns the staged root, or "" if there is nothing to stage from."""
    if not _has_tree(_SOURCE_ROOT):
        return _STAGED_ROOT if _has_tree(_STAGED_ROOT) else ""
    dst = os.path.join(_STAGED_ROOT, "scl")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(_STAGED_ROOT, exist_ok=True)
    shutil.copytree(os.path.join(_SOURCE_ROOT, "scl"), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return _STAGED_ROOT


def reference_available() -> bool:
    return _has_tree(REFERENCE_ROOT)


def import_reference():
    """Return the reference's `scl` package (with the bitarray stand-in if needed)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    try:
        import bitarray  # noqa: F401  (a real install wins if one ever appears)
    except ImportError:
        if _SHIM not in sys.path:
            sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import scl  # noqa: F401
    import scl.compressors.rANS  # noqa: F401
    import scl.compressors.tANS  # noqa: F401
    import scl.compressors.arithmetic_coding  # noqa: F401
    import scl.compressors.range_coder  # noqa: F401
    import scl.compressors.probability_models  # noqa: F401

    return scl
