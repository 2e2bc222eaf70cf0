"""Test harness helpers with the reference's names and semantics (scl/utils/test_utils.py)."""
import numpy as np

from ..core.data_block import DataBlock
from ..core.prob_dist import Frequencies, ProbabilityDist, get_avg_neg_log_prob
from .bitarray_utils import BitArray, get_random_bitarray


def get_random_data_block(prob_dist: ProbabilityDist, size: int, seed: int = None) -> DataBlock:
    # same generator and call as test_utils.py:26-28, so seeds reproduce the reference's blocks
    rng = np.random.default_rng(seed)
    return DataBlock(rng.choice(prob_dist.alphabet, size=size, p=prob_dist.prob_list).tolist())


def are_blocks_equal(a: DataBlock, b: DataBlock) -> bool:
    return a.size == b.size and all(x == y for x, y in zip(a.data_list, b.data_list))


def try_lossless_compression(data_block, encoder, decoder, add_extra_bits_to_encoder_output=False, verbose=False):
    """encode -> (optionally append 0..99 random bits) -> decode; the decoder must consume exactly
    the encoder's bits (test_utils.py:73-108).  Returns (is_lossless, num_bits, encoded)."""
    encoded = encoder.encode_block(data_block)
    stream = BitArray(encoded)
    if add_extra_bits_to_encoder_output:
        stream += get_random_bitarray(int(np.random.randint(100)))
    decoded, used = decoder.decode_block(stream)
    assert used == len(encoded), "Decoder did not consume all bits"
    return are_blocks_equal(data_block, decoded), used, encoded


def lossless_entropy_coder_test(encoder, decoder, freq: Frequencies, data_size: int, encoding_optimality_precision=None, seed: int = 0):
    """test_utils.py:138-180: random i.i.d. block, lossless round trip, optional optimality bound."""
    prob_dist = freq.get_prob_dist()
    block = get_random_data_block(prob_dist, data_size, seed=seed)
    avg_log_prob = get_avg_neg_log_prob(prob_dist, block)
    ok, nbits, _ = try_lossless_compression(block, encoder, decoder, add_extra_bits_to_encoder_output=True)
    avg_codelen = nbits / block.size
    if encoding_optimality_precision is not None:
        assert abs(avg_codelen - avg_log_prob) < encoding_optimality_precision, (avg_codelen, avg_log_prob)
    assert ok


def lossless_test_against_expected_bitrate(encoder, decoder, data_block, expected_bitrate, encoding_optimality_precision):
    ok, nbits, _ = try_lossless_compression(data_block, encoder, decoder, add_extra_bits_to_encoder_output=True)
    assert abs(nbits / data_block.size - expected_bitrate) < encoding_optimality_precision
    assert ok
