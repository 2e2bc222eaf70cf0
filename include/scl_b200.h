/*
 * scl_b200.h -- C-ABI of the B200-native entropy-coding backend for the Stanford
 * Compression Library's encode_block / decode_block hot path.
 *
 * The reference (kedartatwawadi/stanford_compression_library) is pure Python and has no
 * FFI; its boundary for this path is the class API
 *     DataEncoder.encode_block(DataBlock) -> BitArray            scl/core/data_encoder_decoder.py:29-41
 *     DataDecoder.decode_block(BitArray) -> (DataBlock, nbits)   scl/core/data_encoder_decoder.py:102-116
 * implemented by rANSEncoder/Decoder (scl/compressors/rANS.py:123-297), tANSEncoder/Decoder
 * (tANS.py:56-279), ArithmeticEncoder/Decoder (arithmetic_coding.py:41-287) and
 * RangeEncoder/Decoder (range_coder.py:79-317).  Each entry point below replaces the
 * per-symbol loop of one of those methods for a BATCH of independent blocks; the Python
 * classes of the same names in stanford_compression_library_b200/compressors/ bind them
 * with ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *  - Plain C, no C++/torch types.  Every `d_*` pointer is DEVICE memory owned by the caller;
 *    `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *    asynchronous and stream-ordered; the library keeps no global state and allocates nothing
 *    in the data path (tables are owned by the handle).
 *  - Symbols are bytes.  A coder handle is built from the reference's `Frequencies` in dict
 *    insertion order (prob_dist.py:169-205): `alphabet[i]` is the byte value of the i-th key,
 *    `freq[i]` its count.
 *  - Bit streams are MSB-first exactly like BitArray.tobytes().  Block b's stream occupies
 *    bits [bit_offset[b], bit_offset[b] + bit_len[b]) of the byte buffer.  Encoders write
 *    into slot b = bytes [b*out_stride, (b+1)*out_stride): the LIFO coders (rANS, tANS) fill
 *    their slot back-to-front, so their stream ENDS at the slot end; the forward coders
 *    (arithmetic, range) start at the slot start.  Either way the encoder reports
 *    bit_offset[b] and bit_len[b]; scl_pack_blocks() compacts slots into the contiguous
 *    left-aligned form (== concatenated BitArray.tobytes()).
 *  - Decoders accept any bit offset, ignore trailing bits after a block (test_utils.py:97-105)
 *    and report num_bits_consumed exactly as the reference does.
 *  - Return value: 0 on success, else an SCL_E_* code for call-level failures (bad arguments,
 *    CUDA launch error).  Per-block outcomes go to d_status[b] (SCL_ST_*), which the Python
 *    wrappers turn into the exception the reference would raise.
 */
#ifndef SCL_B200_H
#define SCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* call-level return codes */
#define SCL_E_OK 0
#define SCL_E_INVALID 1     /* bad argument / unsupported parameter combination */
#define SCL_E_CUDA 2        /* CUDA runtime error (see scl_last_cuda_error) */
#define SCL_E_UNSUPPORTED 3 /* valid in the reference but outside this backend's limits */

/* per-block status words (d_status) and the reference exception each one stands for */
#define SCL_ST_OK 0
#define SCL_ST_BAD_SYMBOL 1     /* KeyError: symbol not in freq_dict (prob_dist.py:207-208) */
#define SCL_ST_STATE_MISMATCH 2 /* AssertionError: rANS.py:295 / tANS.py:277 end state != INITIAL_STATE */
#define SCL_ST_OVERFLOW 3       /* OverflowError: size does not fit DATA_BLOCK_SIZE_BITS (rANS.py:206), or slot too small */
#define SCL_ST_TRUNCATED 4      /* ValueError: stream ended inside a block */
#define SCL_ST_TOTAL_FREQ 6     /* AssertionError: arithmetic_coding.py:110-112; or an order-k count reached
                                   max_allowed_total_freq (the reference raises, probability_models.py:164-168) */
#define SCL_ST_EMPTY_BLOCK 7    /* arithmetic decoder given size 0: the reference loops forever (arithmetic_coding.py:232-243) */

/* coder kinds */
#define SCL_CODER_RANS 0
#define SCL_CODER_TANS 1
#define SCL_CODER_RANGE 2
#define SCL_CODER_AEC 3

/* frequency models for the arithmetic coder (scl/compressors/probability_models.py) */
#define SCL_MODEL_FIXED 0        /* FixedFreqModel        :57-67 */
#define SCL_MODEL_ADAPTIVE_IID 1 /* AdaptiveIIDFreqModel  :70-92 */
#define SCL_MODEL_ORDER_K 2      /* AdaptiveOrderKFreqModel :95-168, order = scl_params.model_order.  The table of a block
                                    lives in shared memory while n_sym^k * (n_sym + 1) <= 1600 words; larger ones (a byte
                                    alphabet at k = 1) are worked on in place in HBM -- d_model is then mandatory --
                                    for up to n_sym^k = 512 contexts (else SCL_E_UNSUPPORTED) */

typedef struct scl_coder scl_coder; /* opaque: parameters + device tables */

/* Parameters, mirroring the reference dataclasses field for field. */
typedef struct scl_params {
    int32_t coder;                 /* SCL_CODER_* */
    uint32_t data_block_size_bits; /* DATA_BLOCK_SIZE_BITS (all four coders) */
    /* rANSParams / tANSParams (rANS.py:78-120, tANS.py:31-53) */
    uint32_t num_bits_out;   /* NUM_BITS_OUT */
    uint64_t range_factor;   /* RANGE_FACTOR */
    uint32_t num_state_bits; /* NUM_STATE_BITS = get_bit_width(H), evaluated by the caller with the
                                reference's float formula (bitarray_utils.py:8-20) */
    /* AECParams (arithmetic_coding.py:20-38) / RangeCoderParams (range_coder.py:55-76) */
    uint32_t precision; /* PRECISION */
    /* FreqModelBase (probability_models.py:39-44) -- arithmetic coder only */
    int32_t model;                   /* SCL_MODEL_* */
    uint32_t model_order;            /* k of AdaptiveOrderKFreqModel (:104-107); 0 for the other models */
    uint64_t max_allowed_total_freq; /* halving threshold of AdaptiveIIDFreqModel (:90-92) */
} scl_params;

/* ---- lifetime -------------------------------------------------------------------------- */

/* Build a coder for `n_sym` symbols (1..256).  alphabet may be NULL (identity 0..n_sym-1).
 * Replaces rANSParams.__post_init__ + the tANS table builders (tANS.py:88-110,208-226) +
 * RangeEncoder.__init__ checks (range_coder.py:80-86).  Tables are built on the current
 * device, on `stream`. */
int scl_coder_create(const scl_params *params, const uint8_t *alphabet, const uint64_t *freq, uint32_t n_sym,
                     void *stream, scl_coder **out);
void scl_coder_destroy(scl_coder *c);

/* Worst-case encoded size in BYTES of one block of `block_len` symbols, rounded up to the
 * sector size (32).  Use it as out_stride. */
uint64_t scl_coder_max_encoded_bytes(const scl_coder *c, uint64_t block_len);

/* Length, in uint64 words, of ONE block's model table as passed in d_model: n_sym for the fixed / IID
 * models (freqs_current in alphabet order); n_sym^(k+1) + 1 for the order-k model (freqs_kplus1_tuple
 * flattened row-major, then the context index = past_k read as a base-n_sym number).  0 = not an
 * arithmetic coder. */
uint64_t scl_coder_model_words(const scl_coder *c);

/* Which kernel family the handle selected: 0 = 32-bit-state fast path, 1 = generic 64-bit. */
int scl_coder_path(const scl_coder *c, int decode);

/* ---- the hot path ------------------------------------------------------------------------ */

/* encode_block for n_blocks independent DataBlocks.
 *   d_sym      [n_blocks][sym_stride] bytes; block b has d_sizes[b] symbols (d_sizes NULL: block_len each)
 *   d_out      n_blocks slots of out_stride bytes (out_stride % 16 == 0, base 16-byte aligned)
 *   d_out_bit_offset / d_out_bit_len  [n_blocks]  where block b's stream lies (bits)
 *   d_model    arithmetic coder only: [n_blocks][scl_coder_model_words()] uint64 model tables, read as the
 *              initial model state and overwritten with the final one (the reference mutates its model in
 *              place, arithmetic_coding.py:118); NULL = every block starts from the creation-time table
 * Replaces rANSEncoder.encode_block (rANS.py:186-210), tANSEncoder.encode_block (tANS.py:159-193),
 * ArithmeticEncoder.encode_block (arithmetic_coding.py:80-161), RangeEncoder.encode_block
 * (range_coder.py:188-207). */
int scl_encode_blocks(const scl_coder *c, const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes,
                      uint32_t block_len, uint64_t n_blocks, uint8_t *d_out, uint64_t out_stride,
                      uint64_t *d_out_bit_offset, uint64_t *d_out_bit_len, uint64_t *d_model, uint32_t *d_status,
                      void *stream);

/* decode_block for n_blocks independent streams.
 *   d_in / in_bytes      byte buffer holding all streams (base 16-byte aligned)
 *   d_bit_offset[b]      first bit of block b's stream
 *   d_bit_len[b]         bits available to block b (stream + any trailing bits); NULL = up to in_bytes*8
 *   d_sym                [n_blocks][sym_stride] decoded bytes; d_sizes[b] = decoded size (from the header)
 *   d_bits_consumed[b]   the reference's num_bits_consumed
 * Replaces rANSDecoder.decode_block (rANS.py:270-297), tANSDecoder.decode_block (tANS.py:252-279),
 * ArithmeticDecoder.decode_block (arithmetic_coding.py:203-287), RangeDecoder.decode_block
 * (range_coder.py:269-317). */
int scl_decode_blocks(const scl_coder *c, const uint8_t *d_in, uint64_t in_bytes, const uint64_t *d_bit_offset,
                      const uint64_t *d_bit_len, uint64_t n_blocks, uint8_t *d_sym, uint64_t sym_stride,
                      uint32_t *d_sizes, uint64_t *d_bits_consumed, uint64_t *d_model, uint32_t *d_status,
                      void *stream);

/* Compact per-block streams into one contiguous buffer: block b is copied to byte offset
 * d_dst_byte_offset[b] of d_dst, left-aligned, zero-padded to a whole byte -- the bytes of
 * BitArray.tobytes() (bitarray_utils.py:25).  d_dst_byte_offset is caller-computed (exclusive
 * prefix sum of ceil(bit_len/8), or any layout with room).
 * A 4-byte aligned d_src takes the fast kernel (16-byte chunks per lane), which reads whole
 * aligned 32-bit words: the source must be readable up to the next 4-byte boundary after the last
 * stream bit (slot buffers sized with scl_coder_max_encoded_bytes are).  Other sources take the
 * byte-wise kernel. */
#define SCL_PACK_BYTEWISE 1u /* flags: take the byte-wise kernel even for an aligned source (tests) */
int scl_pack_blocks(const uint8_t *d_src, const uint64_t *d_src_bit_offset, const uint64_t *d_bit_len,
                    uint64_t n_blocks, uint8_t *d_dst, const uint64_t *d_dst_byte_offset, uint32_t flags,
                    void *stream);

/* Same, but each block is written in the reference's on-disk framing
 * (EncodedBlockWriter.write_block, scl/core/encoded_stream.py:150-175):
 *   [u32 BE payload bytes][3-bit pad count][pad zeros][stream]  (Padder :22-46, HeaderHandler :93-103).
 * Block b occupies 4 + ceil((bit_len+3)/8) bytes at d_dst_byte_offset[b]. */
int scl_frame_blocks(const uint8_t *d_src, const uint64_t *d_src_bit_offset, const uint64_t *d_bit_len,
                     uint64_t n_blocks, uint8_t *d_dst, const uint64_t *d_dst_byte_offset, uint32_t flags,
                     void *stream);

/* Record offsets of the contiguous layout: d_byte_offset[b] = sum over j < b of size(j), d_byte_offset[n_blocks] =
 * total bytes, with size(j) = ceil(bit_len[j] / 8) (framed == 0) or 4 + ceil((bit_len[j] + 3) / 8) (framed != 0);
 * the running file position of EncodedBlockWriter (scl/core/encoded_stream.py:150-175), as a device scan.
 * d_status (may be NULL): a block whose status is not SCL_ST_OK takes no room.  d_bit_offset (may be NULL, [n_blocks]):
 * the first stream bit of block b inside the packed buffer, i.e. what scl_decode_blocks wants. */
int scl_packed_offsets(const uint64_t *d_bit_len, const uint32_t *d_status, uint64_t n_blocks, uint32_t framed,
                       uint64_t *d_byte_offset, uint64_t *d_bit_offset, void *stream);

/* encode_block for n_blocks DataBlocks with the CONTIGUOUS output the reference's writer produces
 * (b"".join(encode_block(b).tobytes()), or, framed != 0, the records of EncodedBlockWriter.write_block,
 * scl/core/encoded_stream.py:150-175) -- one call, symbols in, finished stream out.
 *   d_scratch / scratch_stride  slots as for scl_encode_blocks (out_stride rules apply): a LIFO stream's start is
 *                               only known when its block is finished, so a block is coded into its slot first
 *   d_dst / dst_bytes           the packed buffer and its capacity; a record that would end past it is dropped and
 *                               its block gets SCL_ST_OVERFLOW
 *   d_byte_offset [n_blocks+1]  record offsets, [n_blocks] = total bytes written
 *   d_bit_offset  [n_blocks]    first stream bit of block b in d_dst (feed it to scl_decode_blocks)
 *   d_workspace                 scl_encode_packed_workspace_bytes() bytes of device scratch
 * Blocks with a non-OK status take no room in d_dst (their d_bit_offset is unspecified).
 * The second-generation rANS / tANS kernels do all of this in ONE launch: per-warp sums of the 32 record sizes,
 * a decoupled look-back across CTAs for the exclusive prefix, and the copy of each finished task to its final
 * offset overlapped with the coding of the next; the other kernel families run encode, scl_packed_offsets and
 * the copy kernel back to back. */
uint64_t scl_encode_packed_workspace_bytes(const scl_coder *c, uint64_t n_blocks);
int scl_encode_blocks_packed(const scl_coder *c, const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes,
                             uint32_t block_len, uint64_t n_blocks, uint8_t *d_scratch, uint64_t scratch_stride,
                             uint8_t *d_dst, uint64_t dst_bytes, uint32_t framed, uint64_t *d_byte_offset,
                             uint64_t *d_bit_offset, uint64_t *d_bit_len, uint64_t *d_model, uint32_t *d_status,
                             void *d_workspace, uint64_t workspace_bytes, void *stream);

/* Byte counts of n_blocks blocks: DataBlock.get_counts (scl/core/data_block.py:37-64) in bulk -- the
 * step before the coders (build a Frequencies table from the data).  d_counts [n_blocks][256]
 * uint32 per-block counts and/or d_total [256] uint64 ACCUMULATED grid-wide totals (zero it first);
 * either may be NULL. */
int scl_histogram_blocks(const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes, uint32_t block_len,
                         uint64_t n_blocks, uint32_t *d_counts, uint64_t *d_total, void *stream);

/* ---- introspection (tests pin the tANS tables against tANS.py:285-337) -------------------- */
int scl_tans_tables_to_host(const scl_coder *c, uint32_t *enc_table, uint32_t *dec_packed, uint64_t n_entries,
                            void *stream);

/* Test hook, per handle (no process-global state): 1 = route this coder's fast paths to the
 * first-generation kernels (per-lane direct global access) instead of the TMA/ring kernels; 2 =
 * second-generation decode with per-lane sector stores instead of TMA tile stores; 3 / 4 =
 * second-generation decode always / never in its pipe-balanced instruction selection (normally
 * chosen by batch size); 5 = the arithmetic coder keeps 16-bit counters where it would take 8-bit ones; 0 = default.
 * Higher bits vary the packed encoder's copy pool: 32 = no shared-memory staging rings (every copy through registers),
 * 64 / 128 = staged pieces of at most 512 / 1024 bytes (default: the largest of 2048 / 1024 / 512 that fits twice),
 * bits 12-15 = number of dedicated copy warps (0 = the default 4); 256 = split batches at 64 MiB of input instead of
 * 2^30 blocks.  Keeps every code path parity-tested.  Set it before issuing work on the handle; it is read at launch time. */
void scl_coder_debug_path(scl_coder *c, int mode);

/* Diagnostic hook, per handle: with a device buffer of n_words uint64 (>= SMs * 32 * 40, zeroed by the caller) the
 * fused packed encoder leaves per-warp timestamps there (%globaltimer, ns): start, the end of each coding round,
 * and the copy pool's task count / first start / last end / busy and waiting time (tools/trace_packed.py prints
 * the timeline).  NULL switches it off (the default); the kernels test one pointer per round, nothing per symbol. */
void scl_coder_debug_trace(scl_coder *c, uint64_t *d_trace, uint64_t n_words);

/* Diagnostic: re-run only the copy stage of a finished scl_encode_blocks_packed call (same buffers, left as that call
 * left them) with `warps_per_cta` warps on each SM and no coder beside them -- how fast is the copy pool alone?
 * ring_stages = 0: every warp copies through registers (the form coding warps use when they help); >= 2: through a
 * shared-memory staging ring of that many 592-byte stages filled by bulk async copies (the dedicated copy warps' form).
 * (tools/measure_copy_only.py).  Second-generation rANS / tANS handles only. */
int scl_debug_copy_only(const scl_coder *c, uint64_t n_blocks, uint8_t *d_scratch, uint64_t scratch_stride, uint8_t *d_dst,
                        uint64_t dst_bytes, uint32_t framed, uint64_t *d_byte_offset, uint64_t *d_bit_offset,
                        uint64_t *d_bit_len, uint32_t *d_status, uint32_t warps_per_cta, uint32_t ring_stages, void *stream);

const char *scl_last_cuda_error(void);
const char *scl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCL_B200_H */
