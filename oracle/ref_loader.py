"""Import the UNMODIFIED reference (`scl`, read-only at /root/reference) for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Only usable in the build container: /root/reference
does not exist on the GPU box, so nothing in `-m gpu` tests, smoke() or bench.py
calls this.  It is used by oracle/gen_golden.py (writes tests/golden/*.npz) and by
the `not gpu` tests that cross-check the C restatement against the live reference
when the reference tree is present (they skip otherwise).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SCL_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bitarray_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scl", "compressors"))


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
