"""B200-native entropy-coding backend for the Stanford Compression Library API.

Drop-in for SCL's per-block coders (rANS, tANS, arithmetic, range) behind the reference's own
DataEncoder / DataDecoder + Frequencies interface; the per-symbol loops run as hand-written
sm_100a CUDA kernels reached through a C-ABI shared library (include/scl_b200.h).
"""
from .core.data_block import DataBlock
from .core.data_encoder_decoder import DataDecoder, DataEncoder
from .core.prob_dist import Frequencies, ProbabilityDist, get_avg_neg_log_prob
from .utils.bitarray_utils import BitArray

__all__ = ["DataBlock", "DataEncoder", "DataDecoder", "Frequencies", "ProbabilityDist", "get_avg_neg_log_prob", "BitArray"]
__version__ = "0.1.0"
