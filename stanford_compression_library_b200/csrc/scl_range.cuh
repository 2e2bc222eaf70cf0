// scl_range.cuh -- second-generation range-coder lanes (range_coder.py:88-207 encode, :210-317 decode)
// for PRECISION = 32, DATA_BLOCK_SIZE_BITS = 32 and a power-of-two total frequency 16 <= T <= 4096.
//
// The first-generation lanes (scl_lane.cuh) follow the reference loop by loop: 64-bit low/range, a
// `while` normalisation whose trip count differs per lane, one scattered 4-byte global store per
// word.  Measured: ~1000 cycles per warp-symbol (profiles/r1e_range_v1).  Here
//   * low / range / state are 32-bit (low + range <= 2^32 always holds, see `range_norm_mul`);
//   * the normalisation runs two predicated iterations for every lane and then a warp-uniform
//     `while (any lane still has to shift)` loop (rare: a symbol releases > 2 bytes only when the
//     range underflows repeatedly), so the warp never diverges;
//   * range // T is a multiply-high; the decoder's (state - low) // r is an fp32 reciprocal estimate with
//     an exact integer correction, and one LUT read returns symbol, cum and freq;
//   * coded bytes move through the lane-interleaved shared-memory rings of scl_fast.cuh and
//     touch HBM in whole 32-byte sectors; symbols arrive by TMA tile (encode) and leave as
//     32-byte sectors (decode).
// __host__ __device__ throughout: tests/host_emu runs these lanes on the CPU against the oracle.
#pragma once
#include "scl_fast.cuh"

namespace scl {

SCL_HD bool warp_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}

// 1 / x to ~23 bits: one MUFU.RCP (the IEEE-rounded __frcp_rn is a Newton step, a range check and a slow-path
// call on top of it; the quotient estimate is corrected exactly afterwards, so the approximation is enough)
SCL_HD float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

constexpr uint32_t kRangeTop = 1u << 24;     // TOP    = 2^(P-8)   (range_coder.py:72)
constexpr uint32_t kRangeBottom = 1u << 16;  // BOTTOM = 2^(P-16)  (range_coder.py:73)
constexpr uint32_t kRangeMaxExtra = 64;      // cap on the extra normalisation rounds of one symbol (hang protection only)

// One round of normalize() (range_coder.py:107-179 / :240-267): returns whether a byte is shifted out, and
// replaces `range` first when it underflowed.
//   settled  <=>  low and low + range agree above bit 24 (the reference XORs unbounded ints, so a sum of
//                 exactly 2^32 is "not settled")  <=>  adding range does not carry out of low's 24 low bits
//             <=>  range <= 0xFFFFFF - (low & 0xFFFFFF)
// low + range <= 2^32 is an invariant (shrink_range only narrows [low, low + range); a settled shift strips
// the common top byte; the underflow branch takes range = (2^32 - low) mod 2^16), so 32-bit low / range lose
// nothing.  Measured alternatives (profiles/README.md, step r1q): the ALU pipe is the busy one here (93 % against
// 15 % for the FMA pipe), but doing the shifts as multiplications by 256 / 1 on the FMA pipe was slower for the
// encoder (longer latency per round, extra moves); only the decoder's bit cache, which nothing waits for, is
// advanced that way.
SCL_HD bool range_norm_test(uint32_t low, uint32_t &range) {
    const bool ns = range > (~low & 0x00FFFFFFu);  // not settled
    const bool lt = range < kRangeBottom;
    const uint32_t fix = (0u - low) & (kRangeBottom - 1);  // (MASK + 1 - low) & (BOTTOM - 1)
    range = (ns & lt) ? fix : range;
    return lt | !ns;
}
// the same test without touching the state: is another round needed?
SCL_HD bool range_needs_norm(uint32_t low, uint32_t range) { return (range <= (~low & 0x00FFFFFFu)) | (range < kRangeBottom); }

// ------------------------------------------------------------------------------------------------
// encoder lane.  Bytes are appended to a 64-bit window (ahi:alo, newest byte in the low byte of
// alo); whole big-endian words go to the [word][lane] ring, whole sectors from there to HBM.
// ------------------------------------------------------------------------------------------------
struct RangeEncV2 {
    uint32_t low, range;
    uint32_t ahi, alo, nbits;  // byte window and the number of BITS in it (<= 56, a multiple of 8)
    uint32_t wofs, rofs;    // words spilled / drained so far, * 128
    saddr_t ring;
    uint8_t *gbegin, *gend;  // output slot; the stream starts at gbegin
    uint32_t ovf, bad;
    uint32_t rshift, neg1;   // log2 T; and -1 as a run-time value (see DecConst)

    SCL_HD void init(saddr_t ring_, uint8_t *slot_begin, uint8_t *slot_end, uint32_t shift) {
        rshift = shift;  // 4 <= shift <= 12
        neg1 = 0xFFFFFFFFu - (shift >> 8);
        low = 0;
        range = 0xFFFFFFFFu;  // range_coder.py:191-192
        ahi = alo = nbits = 0;
        wofs = rofs = 0;
        ring = ring_;
        gbegin = slot_begin;
        gend = slot_end;
        ovf = bad = 0;
    }
    SCL_HD saddr_t ring_slot(uint32_t ofs) const {
#ifdef __CUDA_ARCH__
        return ring | (ofs & ((kEncRingWords - 1) * 128));  // 2 KiB-aligned ring
#else
        return ring + (ofs & ((kEncRingWords - 1) * 128));
#endif
    }
    SCL_HD void put_byte(uint32_t b) {
        ahi = funnel_l(alo, ahi, 8);
        alo = (alo << 8) | b;
        nbits += 8;
    }
    SCL_HD void put_word(uint32_t w) {  // 4 bytes, most significant first (the size header)
        ahi = alo;
        alo = w;
        nbits += 32;
    }
    // at most 4 new bytes since the last call (two symbols' fixed rounds): at most 7 bytes in the window
    SCL_HD void spill_check() {
        if (nbits >= 32) {
            const uint32_t w = funnel_r(alo, ahi, nbits - 32);  // the oldest four bytes
            sts32(ring_slot(wofs), w);
            wofs += 128;
            nbits -= 32;
        }
    }
    SCL_HD void drain_check() {
        if (wofs - rofs >= 8 * 128) {
            u32x8 s;
            const saddr_t g = ring_slot(rofs);  // rofs is a multiple of 8 words: lower or upper half of the ring
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) s.v[j] = bswap32(lds32(g + j * 128));
            uint8_t *dst = gbegin + (rofs >> 5);  // 4 * words_drained
            if (dst + 32 <= gend)
                st_sector32(dst, s);
            else
                ovf = 1;
            rofs += 8 * 128;
        }
    }
    // one normalisation round as straight-line code: a lane that has nothing to shift multiplies by 1
    SCL_HD void norm_once() {
        const bool go = range_norm_test(low, range);
        const uint32_t k = go ? 8u : 0u;
        ahi = funnel_l(alo, ahi, k);
        alo = funnel_l(low, alo, k);  // (alo << 8) | (low >> 24), or alo unchanged
        nbits += k;
        low <<= k;
        range <<= k;
    }
    // The two fixed rounds of a symbol as one: a round changes `range` (underflow) and shifts low / range, but the bytes
    // it releases are simply the next top byte of `low`, so the window takes both rounds' bytes in ONE funnel-shift pair
    // (the ALU pipe is the busy one: profiles/r2n_range_v2_ncu_summary.json).  A lane that did not shift in the first
    // round does not in the second (state unchanged).
    SCL_HD void norm_twice() {
        const uint32_t k1 = range_norm_test(low, range) ? 8u : 0u;
        const uint32_t low1 = low << k1;
        range <<= k1;
        const uint32_t k2 = range_norm_test(low1, range) ? 8u : 0u;
        range <<= k2;
        const uint32_t kt = k1 + k2;
        ahi = funnel_l(alo, ahi, kt);
        alo = funnel_l(low, alo, kt);  // the top kt bits of low
        nbits += kt;
        low = low1 << k2;
    }
    // e = freq << 16 | cum for the symbol (0xFFFFFFFF: not in the alphabet)
    template <bool CHECK, bool VOTE>
    SCL_HD void step(uint32_t e) {
        if (CHECK && e == 0xFFFFFFFFu) {
            bad = 1;
            e = 1u << 16;  // keep every lane on the same path; the block's status reports the bad symbol
        }
        const uint32_t r = range >> rshift;  // range // T: shrink_range (range_coder.py:88-105)
        low = mad32(e & 0xFFFFu, r, low);
        range = r * mulhi_fma(e, 1u << 16);
        norm_twice();
        // Two rounds are enough for all but ~0.02 % of symbols (Zipf data: 30 % shift 0 bytes, 62 % one, 8 % two),
        // i.e. for ~99.3 % of a warp's steps; whether a third is needed is tested without changing the state.
        const bool need = range_needs_norm(low, range);
        if (SCL_UNLIKELY(VOTE ? warp_any(need) : need)) extra_rounds<VOTE>(need);
    }
    template <bool VOTE>
    SCL_HD void extra_rounds(bool need) {
        uint32_t guard = 0;
#pragma unroll 1
        do {
            spill_check();
            drain_check();
            if (need) norm_once();
            need = range_needs_norm(low, range);
            if (++guard > kRangeMaxExtra) {
                ovf = 1;
                break;
            }
        } while (VOTE ? warp_any(need) : need);
    }
    // flush (range_coder.py:181-186) and write everything still buffered; returns the length in bits
    SCL_HD uint64_t finish() {
        for (uint32_t k = 0; k < 4; ++k) {
            put_byte(low >> 24);
            low <<= 8;
            spill_check();
        }
        const uint32_t words = wofs >> 7;
        for (uint32_t i = rofs >> 7; i < words; ++i) {
            uint8_t *dst = gbegin + 4ull * i;
            if (dst + 4 <= gend)
                st_word(dst, bswap32(lds32(ring_slot(i * 128))));
            else
                ovf = 1;
        }
        if (nbits) {  // < 32: left-align the remaining bytes in one last word
            uint8_t *dst = gbegin + 4ull * words;
            if (dst + 4 <= gend)
                st_word(dst, bswap32(alo << (32 - nbits)));
            else
                ovf = 1;
        }
        return 32ull * words + nbits;
    }
};

// 16 symbols of one lane (cnt < 16 only in the block's last chunk)
template <bool CHECK, bool VOTE>
SCL_HD void range_enc_chunk(RangeEncV2 &L, saddr_t tab, uint32_t sym_stride, const u32x4 &v, uint32_t cnt) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    if (cnt == 16) {
        uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
        uint32_t e0, e1, e2, e3;  // the word's four table entries, fetched one word ahead of their use
        e0 = lds32(tab + (saddr_t)mad32(byte_of(w0, 0), sym_stride, 0u));
        e1 = lds32(tab + (saddr_t)mad32(byte_of(w0, 1), sym_stride, 0u));
        e2 = lds32(tab + (saddr_t)mad32(byte_of(w0, 2), sym_stride, 0u));
        e3 = lds32(tab + (saddr_t)mad32(byte_of(w0, 3), sym_stride, 0u));
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {  // one word = 4 symbols per trip: the code stays resident in the instruction cache
            // (after the last word w1 repeats w3's bytes: four harmless extra loads instead of a branch)
            const uint32_t n0 = lds32(tab + (saddr_t)mad32(byte_of(w1, 0), sym_stride, 0u));
            const uint32_t n1 = lds32(tab + (saddr_t)mad32(byte_of(w1, 1), sym_stride, 0u));
            const uint32_t n2 = lds32(tab + (saddr_t)mad32(byte_of(w1, 2), sym_stride, 0u));
            const uint32_t n3 = lds32(tab + (saddr_t)mad32(byte_of(w1, 3), sym_stride, 0u));
            L.template step<CHECK, VOTE>(e0);
            L.template step<CHECK, VOTE>(e1);
            L.spill_check();
            L.template step<CHECK, VOTE>(e2);
            L.template step<CHECK, VOTE>(e3);
            L.spill_check();
            if (j & 1) L.drain_check();  // <= 4 words per 8 symbols from the fixed rounds: the 16-word ring never overruns
            e0 = n0;
            e1 = n1;
            e2 = n2;
            e3 = n3;
            w1 = w2;
            w2 = w3;
        }
    } else {
        for (uint32_t i = 0; i < cnt; ++i) {
            L.template step<CHECK, VOTE>(lds32(tab + (saddr_t)(((wd[i >> 2] >> (8 * (i & 3))) & 0xFFu) * sym_stride)));
            L.spill_check();
            L.drain_check();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// decoder lane (input through DecLaneV2's sector ring)
// ------------------------------------------------------------------------------------------------
struct RangeDecConst {
    saddr_t lut;     // lut[v] = freq << 20 | cum << 8 | byte value, v in [0, T)
    uint32_t shift;  // log2 T
    uint32_t neg1;   // -1 as a run-time value
    uint32_t T;
    uint32_t last;   // the entry of the LAST alphabet index (numpy's searchsorted(...) - 1 == -1 and beyond-the-end cases)
};

struct RangeDecV2 {
    uint32_t low, range, state;
    uint32_t bits;  // the next unread stream bits, left-aligned (refreshed by the caller every two symbols)
    uint32_t ovf;

    SCL_HD void norm_once(DecLaneV2 &D, uint32_t neg1) {
        (void)neg1;
        const bool go = range_norm_test(low, range);
        const uint32_t k = go ? 8u : 0u, mm = go ? 256u : 1u;
        state = funnel_l(bits, state, k);  // (state << 8) | next byte, or unchanged
        bits *= mm;
        D.bp += k;
        low <<= k;
        range <<= k;
    }
    // both fixed rounds at once (see RangeEncV2::norm_twice): the state takes the rounds' bytes in one funnel shift
    SCL_HD void norm_twice(DecLaneV2 &D) {
        const uint32_t k1 = range_norm_test(low, range) ? 8u : 0u;
        const uint32_t low1 = low << k1;
        range <<= k1;
        const uint32_t k2 = range_norm_test(low1, range) ? 8u : 0u;
        range <<= k2;
        const uint32_t kt = k1 + k2;
        state = funnel_l(bits, state, kt);  // (state << kt) | the next kt stream bits
        bits <<= kt;
        D.bp += kt;
        low = low1 << k2;
    }
    // decode_symbol (range_coder.py:225-238) + normalize (:240-267); returns the LUT entry (byte value in bits 0..7)
    template <bool VOTE>
    SCL_HD uint32_t symbol(DecLaneV2 &D, const RangeDecConst &c) {
        const uint32_t r = range >> c.shift;
        const uint32_t a = state - low;
        // q = a // r: fp32 estimate (error << 1 for q <= T + 1), clamped, then corrected exactly
        float qf = (float)a * rcp_approx((float)r);
        const float cap = (float)(c.T + 1);
        uint32_t q = (uint32_t)(qf < cap ? qf : cap);
        if (q <= c.T) {  // q * r <= T * r <= range < 2^32; |a - q * r| < 2 r <= 2^29 because T >= 16 (host-checked)
            const uint32_t rem = a - q * r;
            if ((int32_t)rem < 0)
                q -= 1;
            else if (rem >= r)
                q += 1;
        }
        const bool last = state < low || q >= c.T;  // searchsorted index -1 -> alphabet[-1]; past the end -> the last symbol
        uint32_t e = lds32(c.lut + (saddr_t)((q & (c.T - 1)) << 2));
        if (last) e = c.last;
        low = mad32(mulhi_fma(e, 1u << 24) & 0xFFFu, r, low);
        range = r * mulhi_fma(e, 1u << 12);
        norm_twice(D);
        const bool need = range_needs_norm(low, range);  // see RangeEncV2::step
        if (SCL_UNLIKELY(VOTE ? warp_any(need) : need)) extra_rounds<VOTE>(D, need, c.neg1);
        return e;
    }
    template <bool VOTE>
    SCL_HD void extra_rounds(DecLaneV2 &D, bool need, uint32_t neg1) {
        uint32_t guard = 0;
#pragma unroll 1
        do {
            if (D.filled - D.bp < 64u) D.refill_now();  // far more extra rounds than the prefetch cadence allows for
            bits = D.peek32();
            if (need) norm_once(D, neg1);
            bits = D.peek32();
            need = range_needs_norm(low, range);
            if (++guard > kRangeMaxExtra) {
                ovf = 1;
                break;
            }
        } while (VOTE ? warp_any(need) : need);
    }
    // 16 symbols -> out[0..16) (4-byte aligned), one word store per 4 symbols
    template <bool VOTE>
    SCL_HD void group16(DecLaneV2 &D, const RangeDecConst &c, uint8_t *out, bool store) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {  // 4 symbols per trip: the code stays resident in the instruction cache
            uint32_t acc = 0;
            bits = D.peek32();  // two symbols' fixed rounds read at most 32 bits
            acc = put_byte<0>(acc, symbol<VOTE>(D, c));
            acc = put_byte<1>(acc, symbol<VOTE>(D, c));
            bits = D.peek32();
            acc = put_byte<2>(acc, symbol<VOTE>(D, c));
            acc = put_byte<3>(acc, symbol<VOTE>(D, c));
            if (store) st_word(out + 4 * j, acc);
        }
    }
};

// RangeDecoder.decode_block (range_coder.py:269-317) for one lane.  `out` is 32-byte aligned with room for
// out_cap bytes; `store` = false runs the lane without writing symbols (padding lanes of a voting warp).
// The header (size, then the 4 priming bytes) is read by the caller: see range_dec_header.
SCL_HD bool range_dec_header(DecLaneV2 &D, uint64_t out_cap, uint32_t &size, RangeDecV2 &R) {
    size = D.get(32);
    R.low = 0;
    R.range = 0xFFFFFFFFu;
    R.state = D.get(32);  // range_coder.py:289-291
    R.ovf = 0;
    R.bits = 0;
    return size <= out_cap;
}

template <bool VOTE>
SCL_HD void range_dec_body(DecLaneV2 &D, RangeDecV2 &R, const RangeDecConst &c, uint8_t *out, uint32_t size, bool store) {
    uint32_t p = 0;
#pragma unroll 1
    while (p + 16 <= size) {
        D.prefetch_begin();
        R.template group16<VOTE>(D, c, out + p, store);
        D.prefetch_end();
        p += 16;
    }
    while (p < size) {
        R.bits = D.peek32();
        const uint32_t e = R.template symbol<VOTE>(D, c);
        if (store) out[p] = (uint8_t)e;
        ++p;
        if ((p & 15) == 0) {
            D.prefetch_begin();
            D.prefetch_end();
        }
    }
}

}  // namespace scl
