"""tANS ("cached rANS", scl/compressors/tANS.py) on the GPU.

The reference builds Python dict lookup tables at construction (tANS.py:88-110, 208-226) and its
per-symbol work is three table reads.  Here the same tables are built on the device by one
kernel (one thread per state), staged into shared memory with a TMA bulk copy when they fit, and
walked by tans_encode_lane / tans_decode_lane (csrc/scl_lane.cuh).  For identical parameters
tANS and rANS emit identical bitstreams, which the tests check.
"""
from dataclasses import dataclass

from .. import _cabi
from ..core.data_block import DataBlock
from ..core.data_encoder_decoder import DataDecoder, DataEncoder
from ..utils.bitarray_utils import BitArray, get_bit_width
from ..utils.misc_utils import is_power_of_two
from .rANS import _RansCoder, rANSParams


@dataclass
class tANSParams(rANSParams):
    """rANSParams restricted as in tANS.py:31-53."""

    def __post_init__(self):
        super().__post_init__()
        assert is_power_of_two(self.M), "Please normalize self.M parameter (sum of frequencies) to be a power of two"
        assert self.NUM_BITS_OUT == 1, "only NUM_OUT_BITS = 1 supported for now"
        if self.RANGE_FACTOR > (1 << 16):
            print("WARNING: RANGE_FACTOR > 2^16 --> the lookup tables could be huge")


class _TansCoder(_RansCoder):
    _CODER = _cabi.CODER_TANS

    def _tables(self):
        L = int(self.params.L)
        enc, dec = self.device_coder().tans_tables(L)
        return L, enc, dec


class tANSEncoder(_TansCoder, DataEncoder):
    def __init__(self, tans_params: tANSParams):
        super().__init__(tans_params)

    # the reference's three encoder tables, read back from the device on demand
    @property
    def base_encode_step_table(self):
        """{(s, x_shrunk): x_next} (tANS.py:88-99)."""
        L, enc, _ = self._tables()
        out, row = {}, 0
        for s in self.params.freqs.alphabet:
            lo, hi = self.params.min_shrunk_state[s], self.params.max_shrunk_state[s]
            for j, x in enumerate(range(lo, hi + 1)):
                out[(s, x)] = int(enc[row + j])
            row += hi - lo + 1
        return out

    @property
    def shrink_state_num_out_bits_base_table(self):
        """{s: n} (tANS.py:74-86,101-110)."""
        return {s: self.params.NUM_STATE_BITS - get_bit_width(self.params.max_shrunk_state[s]) for s in self.params.freqs.alphabet}

    @property
    def shrink_state_thresh_table(self):
        base = self.shrink_state_num_out_bits_base_table
        return {s: (self.params.max_shrunk_state[s] + 1) << base[s] for s in self.params.freqs.alphabet}

    def encode_block(self, data_block: DataBlock) -> BitArray:
        return self._encode_one(data_block)


class tANSDecoder(_TansCoder, DataDecoder):
    def __init__(self, tans_params: tANSParams):
        super().__init__(tans_params)

    @property
    def base_decode_step_table(self):
        """{x: (s, x_shrunk)} for x in [L, H] (tANS.py:208-215)."""
        L, _, dec = self._tables()
        self.device_coder()
        return {L + i: (self._byte2sym[int(e) & 0xFF], int(e) >> 8) for i, e in enumerate(dec)}

    @property
    def expand_state_num_bits_table(self):
        """{x_shrunk: num_bits} (tANS.py:217-226)."""
        out = {}
        for s in self.params.freqs.alphabet:
            for x in range(self.params.min_shrunk_state[s], self.params.max_shrunk_state[s] + 1):
                out[x] = self.params.NUM_STATE_BITS - get_bit_width(x)
        return out

    def decode_block(self, encoded_bitarray: BitArray):
        return self._decode_one(encoded_bitarray)
