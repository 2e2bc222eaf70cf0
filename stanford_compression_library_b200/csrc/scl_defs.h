// scl_defs.h -- shared plain-data definitions (device tables, per-coder constants).
// Included by host table builders, CUDA kernels and the CPU lane-emulation test harness.
#pragma once
#include <stdint.h>

#include "../../include/scl_b200.h"

#if defined(__CUDACC__)
#define SCL_HD __host__ __device__ __forceinline__
#else
#define SCL_HD inline
#endif
#define SCL_UNLIKELY(x) __builtin_expect(!!(x), 0)

namespace scl {

// ---- rANS, 32-bit-state fast path ---------------------------------------------------------
// One entry per BYTE VALUE (not per alphabet index).  Encode step (rANS.py:138-161):
//   k  = nb0 + (x > thresh_m1 ? NBO : 0)       closed form of the shrink_state while-loop.  For NBO == 1 the
//                                              table holds ~thresh_m1, so the test is the carry of x + key and
//                                              k = nb0 + carry (IADD3 + IMAD.X); for NBO > 1 it holds thresh_m1
//   x >>= k  (emit the low k bits)
//   q  = umulhi(x, rcp) >> shift               exact x / f  (checked on the host per symbol)
//   x' = x + bias + q * cmpl                   == (x / f) * M + cum + x % f
struct alignas(16) RansEnc32 {
    uint32_t thresh_key;  // NBO == 1: ~thresh_m1, else thresh_m1
    uint32_t rcp;
    uint32_t bias;
    uint32_t pack;  // cmpl(M - f) << 16 | nb0 << 8 | shift ; 0xFFFFFFFF = byte not in the alphabet
};
static const uint32_t kRansEncInvalid = 0xFFFFFFFFu;

// Decode LUT entry for slot = x mod M (M <= 4096): f << 20 | (slot - cum[s]) << 8 | byte value.
//   x' = f * (x >> log2 M) + bias               (rANS.py:234-249)
typedef uint32_t RansDec32;

// ---- rANS, generic 64-bit path (any M, any RANGE_FACTOR, any NUM_BITS_OUT <= 32) ----------
struct alignas(16) RansGeneric {
    uint64_t freq[256];        // by alphabet index
    uint64_t cum[257];         // exclusive prefix, cum[n_sym] = M
    uint64_t max_shrunk[256];  // RF * f * 2^NBO - 1
    uint16_t sym2idx[256];     // byte value -> alphabet index, 0xFFFF = not in alphabet
    uint8_t idx2sym[256];
};

struct RansConst {
    uint64_t M, L, H, RF;
    uint32_t NBO, NSB, DBSB;
    uint32_t n_sym;
    uint32_t m_log2;  // log2 M when M is a power of two, else 0xFFFFFFFF
    uint32_t l_log2;  // log2 L when L is a power of two, else 0xFFFFFFFF
    uint32_t check_sym;  // 1 when some byte values are not in the alphabet
};

// ---- tANS (tANS.py) -----------------------------------------------------------------------
struct alignas(16) TansSym {  // per byte value
    uint32_t thresh;   // shrink_state_thresh_table (tANS.py:85), saturated to 0xFFFFFFFF
    uint32_t nb0;      // shrink_state_num_out_bits_base_table; 0xFFFFFFFF = invalid byte
    int32_t row;       // row offset into enc_table minus min_shrunk_state: index = row + x_shrunk
    uint32_t pad;
};
// The second-generation encoder's form of the same row, 8 bytes (one LDS.64 instead of LDS.128: the tANS encoder is bound
// by L1/shared wavefronts, profiles/r2n_tans_kernels_ncu_summary.json): w = row << 7 | nb0, so that (int32)w >> 5 is
// the row offset in BYTES and w & 31 the bit count (nb0 <= 24; |row| < 2^23).  w = 0xFFFFFFFF = invalid byte.
struct alignas(8) TansSym8 {
    uint32_t thresh;
    uint32_t w;
};
// dec_packed[x - L] = x_shrunk << 8 | byte value   (base_decode_step_table, tANS.py:208-215)

// ---- range coder (range_coder.py), PRECISION in {24, 32} ----------------------------------
struct alignas(16) RangeTab {
    uint32_t cum[257];      // by alphabet index; cum[n_sym] = T
    uint32_t freq[256];
    uint16_t sym2idx[256];
    uint8_t idx2sym[256];
};
struct RangeConst {
    uint32_t P, DBSB, n_sym, T;
    uint32_t t_shift;  // log2 T when T is a power of two (range // T is then a shift), else 0xFFFFFFFF
};

// ---- arithmetic coder (arithmetic_coding.py), PRECISION <= 32 ------------------------------
struct alignas(16) AecTab {
    uint32_t init_freq[256];  // creation-time freqs_initial by alphabet index
    uint16_t sym2idx[256];
    uint8_t idx2sym[256];
};
constexpr uint32_t kAecCtxMaxWords = 1600;  // order-k model: n_ctx * (n_sym + 1) words per lane = 200 KB of shared memory per warp
constexpr uint32_t kAecCtxGlobalMaxRows = 512;  // larger tables stay in HBM (AecCtxGlobalPolicy); their row totals: 64 KB of shared memory per warp
struct AecConst {
    uint32_t P, DBSB, n_sym, model;
    uint64_t max_total;      // FreqModelBase.max_allowed_total_freq
    uint32_t order_k, n_ctx; // AdaptiveOrderKFreqModel: k and n_sym^k (1 for the other models)
    uint32_t ctx_global;     // order-k table too large for shared memory: the lanes work on the caller's model table in HBM
};

}  // namespace scl
