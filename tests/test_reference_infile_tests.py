"""GPU tier (-m gpu): the reference's OWN in-file test functions executed against the
drop-in classes.

Each `test_*` function that sits at the bottom of scl/compressors/{rANS,tANS,arithmetic_coding,range_coder}.py
(rANS.py:303-401, tANS.py:285-452, arithmetic_coding.py:290-463, range_coder.py:320-374) is re-bound -- same code
object, untouched -- to a copy of its module's namespace in which every library name this package also provides
(rANSEncoder, rANSParams, Frequencies, DataBlock, BitArray, try_lossless_compression, ...) is OUR object.  The
reference's asserts (known-answer bitstreams, lossless round trips with trailing garbage, exact bits consumed,
code length close to the entropy) then judge the CUDA backend.  The reference's source is read from
/root/reference in the build container and from the copy staged under oracle/_ref/pyref on the GPU box
(oracle/ref_loader.py); without either the tests skip.
"""
import importlib
import types

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.gpu

INFILE_TESTS = [
    ("rANS", "test_check_encoded_bitarray"),
    ("rANS", "test_rANS_coding"),
    ("tANS", "test_generated_lookup_tables"),
    ("tANS", "test_check_encoded_bitarray"),
    ("tANS", "test_tANS_coding"),
    ("arithmetic_coding", "test_bitarray_for_specific_input"),
    ("arithmetic_coding", "test_arithmetic_coding"),
    ("arithmetic_coding", "test_adaptive_arithmetic_coding"),
    ("arithmetic_coding", "test_adaptive_order_k_arithmetic_coding"),
    ("range_coder", "test_range_coding"),
]


def _our_names():
    import stanford_compression_library_b200 as pkg
    from stanford_compression_library_b200.compressors import arithmetic_coding, probability_models, rANS, range_coder, tANS
    from stanford_compression_library_b200.core import data_block, prob_dist
    from stanford_compression_library_b200.utils import bitarray_utils, test_utils

    names = {}
    for mod in (pkg, data_block, prob_dist, bitarray_utils, test_utils, probability_models, rANS, tANS, arithmetic_coding, range_coder):
        for k, v in vars(mod).items():
            if not k.startswith("_") and (isinstance(v, type) or callable(v)) and getattr(v, "__module__", "").startswith("stanford_compression_library_b200"):
                names[k] = v
    return names


@pytest.fixture(scope="module")
def ours():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_loader.reference_available():
        pytest.skip("no reference tree (neither /root/reference nor the staged oracle/_ref/pyref copy)")
    torch.cuda.set_device(0)
    ref_loader.import_reference()
    return _our_names()


@pytest.mark.parametrize("modname,fname", INFILE_TESTS, ids=["%s.%s" % t for t in INFILE_TESTS])
def test_reference_infile_test_on_dropin_classes(ours, modname, fname):
    mod = importlib.import_module("scl.compressors." + modname)
    f = getattr(mod, fname)
    ns = dict(vars(mod))
    swapped = [k for k in ns if k in ours and ns[k] is not ours[k]]
    for k in swapped:
        ns[k] = ours[k]
    assert swapped, "nothing to swap in %s" % modname
    # the classes the test instantiates must be ours, or the test would only prove the reference against itself
    for k in ("rANSEncoder", "tANSEncoder", "ArithmeticEncoder", "RangeEncoder"):
        if k in vars(mod) and isinstance(vars(mod)[k], type) and vars(mod)[k].__module__ == mod.__name__:
            assert k in swapped, "%s was not replaced by the drop-in class" % k
    g = types.FunctionType(f.__code__, ns, f.__name__, f.__defaults__, f.__closure__)
    g()
