// scl_lane.cuh -- the per-lane coder recurrences.  One warp lane == one DataBlock: each lane
// runs the reference's exact integer state machine for its own block (bit-exact
// output forces the block to be the unit of parallelism because every reference coder is a
// single sequential state chain per block).
//
// Everything here is __host__ __device__ so that tests/host_emu can run the very same code on
// the CPU against the oracle (a test-only build; the product only ever runs it inside the CUDA
// kernels of scl_kernels.cu).
#pragma once
#include "scl_defs.h"

namespace scl {

// ------------------------------------------------------------------------------------------------
// small intrinsics with host equivalents
// ------------------------------------------------------------------------------------------------
SCL_HD uint32_t clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
SCL_HD uint32_t clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clzll((long long)x);
#else
    return x ? (uint32_t)__builtin_clzll(x) : 64u;
#endif
}
SCL_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
SCL_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
// (hi:lo) << s, upper 32 bits, s taken mod 32
SCL_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31;
    return s ? ((hi << s) | (lo >> (32 - s))) : hi;
#endif
}
// (hi:lo) >> s, lower 32 bits, s taken mod 32
SCL_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31;
    return s ? ((lo >> s) | (hi << (32 - s))) : lo;
#endif
}
SCL_HD uint32_t mask32(uint32_t k) { return k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u); }
SCL_HD uint64_t mask64(uint32_t k) { return k >= 64 ? ~0ull : ((1ull << k) - 1ull); }

struct u32x4 {
    uint32_t x, y, z, w;
};
struct alignas(8) u32x2 {
    uint32_t x, y;
};

// streaming 16-byte load/store of the lane's own row (bypass L1 allocation: every byte is touched once)
SCL_HD u32x4 ld_stream16(const uint8_t *p) {
    u32x4 r;
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
#else
    const uint32_t *q = (const uint32_t *)p;
    r.x = q[0];
    r.y = q[1];
    r.z = q[2];
    r.w = q[3];
#endif
    return r;
}
SCL_HD void st_stream16(uint8_t *p, const u32x4 &v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
#else
    uint32_t *q = (uint32_t *)p;
    q[0] = v.x;
    q[1] = v.y;
    q[2] = v.z;
    q[3] = v.w;
#endif
}
SCL_HD uint32_t ld_word(const uint8_t *p) { return *(const uint32_t *)p; }
SCL_HD void st_word(uint8_t *p, uint32_t v) { *(uint32_t *)p = v; }

// ------------------------------------------------------------------------------------------------
// bit I/O.  Streams are MSB-first (BitArray.tobytes()), so a 32-bit group of the stream is a
// big-endian word.
// ------------------------------------------------------------------------------------------------

// LIFO writer for rANS/tANS: the reference PREPENDS every symbol's bits (rANS.py:158,196), i.e.
// the payload read as one big-endian integer grows at its most-significant end.  We accumulate
// little-endian in a 64-bit register and spill 32-bit big-endian words from the END of the
// block's slot towards its start; the finished stream is right-aligned in the slot.
struct LifoBitWriter {
    uint8_t *cur;  // next word goes to cur-4
    uint8_t *lo;   // slot begin
    uint64_t acc;
    uint32_t nacc;   // valid bits in acc (< 32 between calls)
    uint64_t words;  // words spilled
    uint32_t ovf;

    SCL_HD void init(uint8_t *slot_begin, uint8_t *slot_end) {
        cur = slot_end;
        lo = slot_begin;
        acc = 0;
        nacc = 0;
        words = 0;
        ovf = 0;
    }
    SCL_HD void spill() {
        cur -= 4;
        if (cur >= lo)
            st_word(cur, bswap32((uint32_t)acc));
        else
            ovf = 1;
        acc >>= 32;
        nacc -= 32;
        ++words;
    }
    // v < 2^k, k <= 32
    SCL_HD void put(uint32_t v, uint32_t k) {
        acc |= (uint64_t)v << nacc;
        nacc += k;
        if (nacc >= 32) spill();
    }
    SCL_HD void put64(uint64_t v, uint32_t k) {  // k <= 64
        if (k > 32) {
            put((uint32_t)v, 32);
            put((uint32_t)(v >> 32), k - 32);
        } else {
            put((uint32_t)v, k);
        }
    }
    // returns the stream length in bits; the stream starts `bits` before the slot end
    SCL_HD uint64_t finish() {
        uint64_t bits = words * 32 + nacc;
        if (nacc) {
            cur -= 4;
            if (cur >= lo)
                st_word(cur, bswap32((uint32_t)acc));
            else
                ovf = 1;
        }
        return bits;
    }
};

// Forward writer for the arithmetic and range coders (bits are appended).
struct FwdBitWriter {
    uint8_t *cur;
    uint8_t *hi;  // slot end
    uint64_t acc;
    uint32_t nacc;
    uint64_t words;
    uint32_t ovf;
    SCL_HD void init(uint8_t *slot_begin, uint8_t *slot_end) {
        cur = slot_begin;
        hi = slot_end;
        acc = 0;
        nacc = 0;
        words = 0;
        ovf = 0;
    }
    SCL_HD void put(uint32_t v, uint32_t k) {  // v < 2^k, k <= 32
        acc = (acc << k) | v;
        nacc += k;
        if (nacc >= 32) {
            uint32_t w = (uint32_t)(acc >> (nacc - 32));
            if (cur + 4 <= hi)
                st_word(cur, bswap32(w));
            else
                ovf = 1;
            cur += 4;
            nacc -= 32;
            ++words;
        }
    }
    SCL_HD void put64(uint64_t v, uint32_t k) {
        if (k > 32) {
            put((uint32_t)(v >> 32), k - 32);
            put((uint32_t)v, 32);
        } else {
            put((uint32_t)v, k);
        }
    }
    SCL_HD void put_run(uint32_t bit, uint64_t count) {  // `count` copies of `bit`
        while (count) {
            uint32_t k = count > 32 ? 32u : (uint32_t)count;
            put(bit ? mask32(k) : 0u, k);
            count -= k;
        }
    }
    SCL_HD uint64_t finish() {
        uint64_t bits = words * 32 + nacc;
        if (nacc) {
            uint32_t w = (uint32_t)(acc << (32 - nacc));
            if (cur + 4 <= hi)
                st_word(cur, bswap32(w));
            else
                ovf = 1;
        }
        return bits;
    }
};

// Forward reader with a 64-bit MSB-aligned window.  Reads past `in_bytes` return zero bits.
struct BitReader {
    const uint8_t *base;
    uint64_t in_bytes;
    uint64_t wi;  // next 32-bit word index to fetch
    uint64_t w;   // window, next bit = bit 63
    uint32_t avail;
    uint64_t used;  // bits handed out

    SCL_HD uint32_t fetch(uint64_t idx) const {
        uint64_t off = idx * 4;
        if (off + 4 <= in_bytes) return bswap32(ld_word(base + off));
        uint32_t v = 0;
        for (uint32_t b = 0; b < 4; ++b) {
            v <<= 8;
            if (off + b < in_bytes) v |= base[off + b];
        }
        return v;
    }
    SCL_HD void init(const uint8_t *base_, uint64_t in_bytes_, uint64_t bit_off) {
        base = base_;
        in_bytes = in_bytes_;
        wi = bit_off >> 5;
        uint32_t sh = (uint32_t)(bit_off & 31);
        uint64_t a = fetch(wi), b = fetch(wi + 1);
        w = ((a << 32) | b) << sh;
        avail = 64 - sh;
        wi += 2;
        used = 0;
        if (avail <= 32) refill();
    }
    SCL_HD void refill() {
        w |= (uint64_t)fetch(wi) << (32 - avail);
        avail += 32;
        ++wi;
    }
    // next k bits as an integer, 0 <= k <= 32
    SCL_HD uint32_t get(uint32_t k) {
        uint32_t hi = (uint32_t)(w >> 32);
        uint32_t v = k ? (hi >> (32 - k)) : 0u;
        w <<= k;  // k <= 32 < 64
        avail -= k;
        used += k;
        if (avail <= 32) refill();
        return v;
    }
    SCL_HD uint64_t get64(uint32_t k) {  // k <= 64
        if (k > 32) {
            uint64_t h = get(k - 32);
            return (h << 32) | get(32);
        }
        return get(k);
    }
    // (x << k) | next k bits, k <= 31, in one funnel shift
    SCL_HD uint32_t shift_in(uint32_t x, uint32_t k) {
        uint32_t hi = (uint32_t)(w >> 32);
        uint32_t r = funnel_l(hi, x, k);
        w <<= k;
        avail -= k;
        used += k;
        if (avail <= 32) refill();
        return r;
    }
};

// ------------------------------------------------------------------------------------------------
// rANS -- 32-bit-state fast path
// ------------------------------------------------------------------------------------------------

// One encode step (rANS.py:163-184 encode_symbol = shrink_state + rans_base_encode_step).
template <bool CHECK>
SCL_HD bool rans32_encode_step(const RansEnc32 *tab, uint32_t nbo, uint32_t s, uint32_t &x, LifoBitWriter &w) {
    const RansEnc32 e = tab[s];
    if (CHECK && e.pack == kRansEncInvalid) return false;
    uint32_t k = ((e.pack >> 8) & 0xFFu) + (x > (nbo == 1 ? ~e.thresh_key : e.thresh_key) ? nbo : 0u);
    w.put(x & mask32(k), k);
    x >>= k;
    uint32_t q = funnel_r(umulhi32(x, e.rcp), 0u, e.pack);  // >> (pack & 31)
    x = x + e.bias + q * (e.pack >> 16);
    return true;
}

// rANSEncoder.encode_block (rANS.py:186-210) for one lane.  Returns the per-block status.
template <bool CHECK>
SCL_HD uint32_t rans32_encode_lane(const RansEnc32 *tab, const RansConst &c, const uint8_t *sym, uint32_t n,
                                   LifoBitWriter &w, uint64_t &bits_out) {
    uint32_t x = (uint32_t)c.L;  // INITIAL_STATE
    uint32_t i = 0;
    bool ok = true;
    const uint32_t nbo = c.NBO;
    if ((((uintptr_t)sym) & 15) == 0) {
        for (; i + 16 <= n; i += 16) {
            u32x4 v = ld_stream16(sym + i);
            uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int b = 0; b < 4; ++b) ok &= rans32_encode_step<CHECK>(tab, nbo, (wd[j] >> (8 * b)) & 0xFFu, x, w);
            }
        }
    }
    for (; i < n; ++i) ok &= rans32_encode_step<CHECK>(tab, nbo, sym[i], x, w);
    // header: [size : DBSB][state : NSB] goes in FRONT of the payload (rANS.py:199,206-208)
    w.put(x, c.NSB);
    uint32_t st = SCL_ST_OK;
    if (c.DBSB < 32 && (n >> c.DBSB)) st = SCL_ST_OVERFLOW;  // uint_to_bitarray OverflowError
    w.put64((uint64_t)n, c.DBSB);
    bits_out = w.finish();
    if (!ok) st = SCL_ST_BAD_SYMBOL;
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// renormalisation bit count after a decode step when L = 2^l (closed form of expand_state,
// rANS.py:251-260): smallest multiple k of NBO with (x << k) >= L.
SCL_HD uint32_t rans32_renorm_bits(uint32_t x, uint32_t l_log2, uint32_t nbo) {
    int32_t d = (int32_t)clz32(x) - (int32_t)(31 - l_log2);  // bits missing to reach bit l
    if (nbo == 1) return (uint32_t)d;                          // d >= 0 because x < 2^(l+1)
    if (d <= 0) return 0;
    uint32_t chunks = ((uint32_t)d + nbo - 1) / nbo;
    return chunks * nbo;
}

// rANSDecoder.decode_block (rANS.py:270-297) for one lane, LUT decode.
SCL_HD uint32_t rans32_decode_lane(const RansDec32 *lut, const RansConst &c, BitReader &r, uint8_t *out,
                                   uint64_t out_cap, uint32_t &size_out, uint64_t &bits_consumed) {
    uint64_t size64 = r.get64(c.DBSB);
    uint32_t x = r.get(c.NSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    const uint32_t size = (uint32_t)size64;
    const uint32_t mmask = (uint32_t)c.M - 1, mlog = c.m_log2, llog = c.l_log2, nbo = c.NBO;
    uint32_t p = size;  // symbols are produced last-first (rANS.py:289-291)
#define SCL_RANS32_DEC_STEP(SYMVAR)                          \
    {                                                        \
        uint32_t e = lut[x & mmask];                         \
        x = (e >> 20) * (x >> mlog) + ((e >> 8) & 0xFFFu);   \
        SYMVAR = e & 0xFFu;                                  \
        uint32_t k = rans32_renorm_bits(x, llog, nbo);       \
        x = r.shift_in(x, k);                                \
    }
    const bool aligned = ((((uintptr_t)out) & 15) == 0);
    while (p > 0 && (!aligned || (p & 15))) {
        uint32_t s;
        SCL_RANS32_DEC_STEP(s);
        out[--p] = (uint8_t)s;
    }
    while (p >= 16) {
        uint32_t wd[4];
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            uint32_t acc = 0;
#pragma unroll
            for (int b = 3; b >= 0; --b) {
                uint32_t s;
                SCL_RANS32_DEC_STEP(s);
                acc |= s << (8 * b);
            }
            wd[j] = acc;
        }
        p -= 16;
        u32x4 v = {wd[0], wd[1], wd[2], wd[3]};
        st_stream16(out + p, v);
    }
#undef SCL_RANS32_DEC_STEP
    size_out = size;
    bits_consumed = r.used;
    return x == (uint32_t)c.L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;  // rANS.py:295
}

// ------------------------------------------------------------------------------------------------
// rANS -- generic path: 64-bit state, any M / RANGE_FACTOR / NUM_BITS_OUT; literal loops
// ------------------------------------------------------------------------------------------------
SCL_HD uint32_t rans64_encode_lane(const RansGeneric &t, const RansConst &c, const uint8_t *sym, uint32_t n,
                                   LifoBitWriter &w, uint64_t &bits_out) {
    uint64_t x = c.L;
    uint32_t st = SCL_ST_OK;
    const uint32_t nbo = c.NBO;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t idx = t.sym2idx[sym[i]];
        if (idx == 0xFFFFu) {
            st = SCL_ST_BAD_SYMBOL;
            break;
        }
        const uint64_t f = t.freq[idx], ms = t.max_shrunk[idx];
        while (x > ms) {  // shrink_state (rANS.py:149-161)
            w.put((uint32_t)(x & mask64(nbo)), nbo);
            x >>= nbo;
        }
        x = (x / f) * c.M + t.cum[idx] + (x % f);  // rans_base_encode_step (rANS.py:138-147)
    }
    if (c.NSB < 64 && (x >> c.NSB)) st = st ? st : SCL_ST_OVERFLOW;
    w.put64(x, c.NSB);
    if (c.DBSB < 32 && (n >> c.DBSB)) st = st ? st : SCL_ST_OVERFLOW;
    w.put64((uint64_t)n, c.DBSB);
    bits_out = w.finish();
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// numpy.searchsorted(cum, v, side="right") - 1 over cum[0..n) (rANS.py:217-232)
template <typename T>
SCL_HD uint32_t find_bin(const T *cum, uint32_t n, T v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cum[mid] <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo - 1;
}

// `avail_bits` = bits the stream may hold from its first bit on.  The reader returns zero bits past the
// end of the buffer, so a malformed stream (a zero state never grows under `x = (x << nbo) + bits`) must
// not be allowed to spin: once the reads have gone past `avail_bits` the reference would have raised
// ValueError from bitarray_to_uint on an empty slice (rANS.py:256, bitarray_utils.py:38) -> TRUNCATED.
SCL_HD uint32_t rans64_decode_lane(const RansGeneric &t, const RansConst &c, BitReader &r, uint64_t avail_bits, uint8_t *out,
                                   uint64_t out_cap, uint32_t &size_out, uint64_t &bits_consumed) {
    uint64_t size64 = r.get64(c.DBSB);
    uint64_t x = r.get64(c.NSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    const uint32_t size = (uint32_t)size64, nbo = c.NBO;
    for (uint32_t p = size; p > 0; --p) {
        uint64_t block_id = x / c.M, slot = x % c.M;  // rans_base_decode_step (rANS.py:234-249)
        uint32_t idx = find_bin<uint64_t>(t.cum, c.n_sym, slot);
        x = block_id * t.freq[idx] + slot - t.cum[idx];
        while (x < c.L) {  // expand_state (rANS.py:251-260)
            x = (x << nbo) + r.get(nbo);
            if (SCL_UNLIKELY(r.used > avail_bits)) {
                bits_consumed = r.used;
                return SCL_ST_TRUNCATED;
            }
        }
        out[p - 1] = t.idx2sym[idx];
    }
    size_out = size;
    bits_consumed = r.used;
    return x == c.L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;
}

// ------------------------------------------------------------------------------------------------
// tANS (table ANS == cached rANS, tANS.py:1-4)
// ------------------------------------------------------------------------------------------------

// One entry of both lookup tables, from state x = L + i: the decode step is evaluated like
// build_rans_base_decode_table (tANS.py:208-215) and, because the rANS step is a bijection
// between states [L, H] and pairs (s, x_shrunk), the same evaluation fills
// base_encode_step_table[(s, x_shrunk)] = x (tANS.py:88-99).
SCL_HD void tans_build_entry(const RansGeneric &t, const RansConst &c, const uint32_t *row_of_idx, uint32_t *enc_table,
                             uint32_t *dec_packed, uint64_t i) {
    uint64_t x = c.L + i;
    uint64_t block_id = x >> c.m_log2, slot = x & (c.M - 1);
    uint32_t idx = find_bin<uint64_t>(t.cum, c.n_sym, slot);
    uint64_t f = t.freq[idx];
    uint64_t shrunk = block_id * f + slot - t.cum[idx];  // rans_base_decode_step(x)
    dec_packed[i] = ((uint32_t)shrunk << 8) | t.idx2sym[idx];
    enc_table[row_of_idx[idx] + (uint32_t)(shrunk - c.RF * f)] = (uint32_t)x;
}

// tANSEncoder.encode_block (tANS.py:159-193); encode_symbol (:126-157) is three table reads.
SCL_HD uint32_t tans_encode_lane(const TansSym *symtab, const uint32_t *enc_table, const RansConst &c,
                                 const uint8_t *sym, uint32_t n, LifoBitWriter &w, uint64_t &bits_out) {
    uint32_t x = (uint32_t)c.L;
    uint32_t st = SCL_ST_OK;
    for (uint32_t i = 0; i < n; ++i) {
        const TansSym e = symtab[sym[i]];
        if (e.nb0 == 0xFFFFFFFFu) {
            st = SCL_ST_BAD_SYMBOL;
            break;
        }
        uint32_t k = e.nb0 + (x >= e.thresh ? 1u : 0u);
        w.put(x & mask32(k), k);
        x >>= k;
        x = enc_table[(int32_t)x + e.row];  // base_encode_step_table[(s, x_shrunk)]
    }
    w.put(x, c.NSB);
    if (c.DBSB < 32 && (n >> c.DBSB)) st = st ? st : SCL_ST_OVERFLOW;
    w.put64((uint64_t)n, c.DBSB);
    bits_out = w.finish();
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// tANSDecoder.decode_block (tANS.py:252-279); decode_symbol (:239-250).
SCL_HD uint32_t tans_decode_lane(const uint32_t *dec_packed, const RansConst &c, BitReader &r, uint8_t *out,
                                 uint64_t out_cap, uint32_t &size_out, uint64_t &bits_consumed) {
    uint64_t size64 = r.get64(c.DBSB);
    uint32_t x = r.get(c.NSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    const uint32_t size = (uint32_t)size64, L = (uint32_t)c.L;
    for (uint32_t p = size; p > 0; --p) {
        uint32_t i = x - L;
        if (i >= L) return SCL_ST_STATE_MISMATCH;  // KeyError on base_decode_step_table
        uint32_t e = dec_packed[i];
        uint32_t xs = e >> 8;
        uint32_t k = c.NSB - (32 - clz32(xs));  // expand_state_num_bits_table (tANS.py:225)
        x = r.shift_in(xs, k);
        out[p - 1] = (uint8_t)(e & 0xFFu);
    }
    size_out = size;
    bits_consumed = r.used;
    return x == L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;
}

// ------------------------------------------------------------------------------------------------
// 16-byte register windows for a lane's own symbol row: one vector access per 16 symbols instead of
// a byte access per symbol (a per-lane byte load costs a whole 32-byte sector and a wavefront per lane).
// ------------------------------------------------------------------------------------------------
struct SymWindow {  // sequential reader
    const uint8_t *row;
    uint64_t cap;  // bytes that may be read from `row` (the row stride)
    uint32_t w0, w1, w2, w3;
    SCL_HD void init(const uint8_t *row_, uint64_t cap_) {
        row = row_;
        cap = cap_;
        w0 = w1 = w2 = w3 = 0;
    }
    SCL_HD uint32_t next(uint32_t i) {  // symbol i; must be called for i = 0, 1, 2, ... in order
        if ((i & 15) == 0) {
            if (((((uintptr_t)row) & 15) == 0) && (uint64_t)i + 16 <= cap) {
                u32x4 v = ld_stream16(row + i);
                w0 = v.x;
                w1 = v.y;
                w2 = v.z;
                w3 = v.w;
            } else {
                uint32_t t[4] = {0, 0, 0, 0};
                for (uint32_t k = 0; k < 16 && (uint64_t)i + k < cap; ++k) t[k >> 2] |= (uint32_t)row[i + k] << (8 * (k & 3));
                w0 = t[0];
                w1 = t[1];
                w2 = t[2];
                w3 = t[3];
            }
        }
        uint32_t s = w0 & 0xFFu;
        w0 = funnel_r(w0, w1, 8);
        w1 = funnel_r(w1, w2, 8);
        w2 = funnel_r(w2, w3, 8);
        w3 >>= 8;
        return s;
    }
};
struct OutWindow {  // sequential writer
    uint8_t *row;
    uint32_t w0, w1, w2, w3;
    SCL_HD void init(uint8_t *row_) {
        row = row_;
        w0 = w1 = w2 = w3 = 0;
    }
    SCL_HD void push(uint32_t i, uint32_t s) {  // symbol i; in order from 0
        w0 = funnel_r(w0, w1, 8);
        w1 = funnel_r(w1, w2, 8);
        w2 = funnel_r(w2, w3, 8);
        w3 = (w3 >> 8) | (s << 24);
        if ((i & 15) == 15) {
            if ((((uintptr_t)row) & 15) == 0) {
                u32x4 v = {w0, w1, w2, w3};
                st_stream16(row + (i - 15), v);
            } else {
                uint32_t t[4] = {w0, w1, w2, w3};
                for (uint32_t k = 0; k < 16; ++k) row[i - 15 + k] = (uint8_t)(t[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
    SCL_HD void flush(uint32_t n) {  // after the last push: write the n % 16 pending symbols
        uint32_t rem = n & 15;
        if (!rem) return;
        uint32_t t[4] = {w0, w1, w2, w3};  // the pending bytes are the TOP `rem` bytes of the 16-byte window
        for (uint32_t k = 0; k < rem; ++k) {
            uint32_t pos = 16 - rem + k;
            row[n - rem + k] = (uint8_t)(t[pos >> 2] >> (8 * (pos & 3)));
        }
    }
};

// floor(a / d) for a < 2^62, 0 < d <= 2^32, given rcp = 1.0 / d: one FP64 multiply plus an exact
// integer correction instead of a 64-bit integer division (~70 instructions on the GPU).  The FP64
// estimate is within +-1 of the true quotient (quotients here are < 2^33, relative error of the
// product < 2^-50), and the remainder test makes the result exact.
SCL_HD uint64_t div_exact_rcp(uint64_t a, uint64_t d, double rcp) {
    uint64_t q = (uint64_t)((double)a * rcp);
    int64_t r = (int64_t)(a - q * d);
    if (r < 0)
        q -= 1;
    else if ((uint64_t)r >= d)
        q += 1;
    return q;
}

// ------------------------------------------------------------------------------------------------
// range coder (range_coder.py), PRECISION in {24, 32}: low/range held in 64 bits so that
// low + range == 2^P is representable (the reference works in unbounded ints)
// ------------------------------------------------------------------------------------------------
SCL_HD uint32_t range_encode_lane(const RangeTab &t, const RangeConst &c, const uint8_t *sym, uint64_t sym_cap, uint32_t n,
                                  FwdBitWriter &w, uint64_t &bits_out) {
    SymWindow sw;
    sw.init(sym, sym_cap);
    const uint32_t P = c.P;
    const uint64_t TOP = 1ull << (P - 8), BOTTOM = 1ull << (P - 16), MASK = (1ull << P) - 1;
    uint64_t low = 0, range = MASK;  // range_coder.py:191-192
    uint32_t st = SCL_ST_OK;
    w.put64((uint64_t)n, c.DBSB);
    if (c.DBSB < 32 && (n >> c.DBSB)) st = SCL_ST_OVERFLOW;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t idx = t.sym2idx[sw.next(i)];
        if (idx == 0xFFFFu) {
            st = SCL_ST_BAD_SYMBOL;
            break;
        }
        // shrink_range (range_coder.py:88-105); values fit 32 bits here: shift or 32-bit division
        uint32_t r = c.t_shift != 0xFFFFFFFFu ? ((uint32_t)range >> c.t_shift) : ((uint32_t)range / c.T);
        low += (uint64_t)t.cum[idx] * r;
        range = (uint64_t)r * t.freq[idx];
        // normalize (range_coder.py:107-179)
        for (;;) {
            bool settled = ((low ^ (low + range)) < TOP);
            if (!settled) {
                if (range >= BOTTOM) break;
                range = (MASK + 1 - low) & (BOTTOM - 1);
            }
            w.put((uint32_t)(low >> (P - 8)), 8);
            low = (low << 8) & MASK;
            range <<= 8;
            if (w.ovf) break;  // a degenerate range of 0 would never terminate in the reference either
        }
    }
    for (uint32_t k = 0; k < P / 8; ++k) {  // flush (range_coder.py:181-186)
        w.put((uint32_t)(low >> (P - 8)), 8);
        low = (low << 8) & MASK;
    }
    bits_out = w.finish();
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// `lut` (may be null) maps v in [0, T) to the alphabet index whose [cum, cum+f) contains v
SCL_HD uint32_t range_decode_lane(const RangeTab &t, const RangeConst &c, const uint8_t *lut, BitReader &r, uint64_t avail_bits,
                                  uint8_t *out, uint64_t out_cap, uint32_t &size_out, uint64_t &bits_consumed) {
    const uint32_t P = c.P;
    const uint64_t TOP = 1ull << (P - 8), BOTTOM = 1ull << (P - 16), MASK = (1ull << P) - 1;
    uint64_t size64 = r.get64(c.DBSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    const uint32_t size = (uint32_t)size64;
    uint64_t low = 0, range = MASK, state = 0;
    for (uint32_t k = 0; k < P / 8; ++k) state = (state << 8) | r.get(8);  // range_coder.py:289-291
    OutWindow ow;
    ow.init(out);
    for (uint32_t i = 0; i < size; ++i) {
        // decode_symbol (range_coder.py:225-238): last i with low + cum_i * (range // T) <= state
        uint32_t rr = c.t_shift != 0xFFFFFFFFu ? ((uint32_t)range >> c.t_shift) : ((uint32_t)range / c.T);
        uint32_t idx;
        if (state < low || rr == 0) {
            idx = c.n_sym - 1;  // searchsorted gives 0 -> alphabet[-1]
        } else {
            uint64_t v = div_exact_rcp(state - low, rr, 1.0 / (double)rr);
            if (v >= c.T)
                idx = c.n_sym - 1;  // beyond the last cumulative value: searchsorted returns the last index
            else
                idx = lut ? lut[v] : find_bin<uint32_t>(t.cum, c.n_sym, (uint32_t)v);
        }
        ow.push(i, t.idx2sym[idx]);
        low += (uint64_t)t.cum[idx] * rr;
        range = (uint64_t)rr * t.freq[idx];
        for (;;) {  // normalize (range_coder.py:240-267)
            bool settled = ((low ^ (low + range)) < TOP);
            if (!settled) {
                if (range >= BOTTOM) break;
                range = (MASK + 1 - low) & (BOTTOM - 1);
            }
            state = ((state << 8) | r.get(8)) & MASK;
            low = (low << 8) & MASK;
            range <<= 8;
            if (r.used > avail_bits) {
                ow.flush(i + 1);
                return SCL_ST_TRUNCATED;
            }
        }
    }
    ow.flush(size);
    size_out = size;
    bits_consumed = r.used;
    return r.used > avail_bits ? SCL_ST_TRUNCATED : SCL_ST_OK;
}

// ------------------------------------------------------------------------------------------------
// arithmetic coder (arithmetic_coding.py) with FixedFreqModel / AdaptiveIIDFreqModel
// (probability_models.py:57-92).  The model's 256 counters live in a Fenwick tree `F`
// (1-based, F[0] unused) supplied by the caller through the accessor type `Tree`
// (shared memory, lane-interleaved, on the device; a plain array in the host harness).
// ------------------------------------------------------------------------------------------------
template <typename Tree>
SCL_HD uint32_t fen_prefix(const Tree &F, uint32_t i) {  // sum of counts[0 .. i)
    uint32_t s = 0;
    while (i) {
        s += F.get(i);
        i &= i - 1;
    }
    return s;
}
template <typename Tree>
SCL_HD void fen_add1(Tree &F, uint32_t idx) {
    for (uint32_t i = idx + 1; i <= 256; i += i & (0u - i)) F.set(i, F.get(i) + 1);
}
template <typename Tree>
SCL_HD void fen_build(Tree &F) {  // counts (in F[1..256]) -> tree, in place
    for (uint32_t i = 1; i <= 256; ++i) {
        uint32_t j = i + (i & (0u - i));
        if (j <= 256) F.set(j, F.get(j) + F.get(i));
    }
}
template <typename Tree>
SCL_HD void fen_unbuild(Tree &F) {  // tree -> counts, in place
    for (uint32_t i = 256; i >= 1; --i) {
        uint32_t j = i + (i & (0u - i));
        if (j <= 256) F.set(j, F.get(j) - F.get(i));
    }
}
// largest idx with prefix(idx) <= v, and that prefix
template <typename Tree>
SCL_HD uint32_t fen_find(const Tree &F, uint32_t v, uint32_t &prefix_out) {
    uint32_t pos = 0, rem = v;
#pragma unroll
    for (uint32_t step = 128; step >= 1; step >>= 1) {
        uint32_t t = F.get(pos + step);
        if (t <= rem) {
            pos += step;
            rem -= t;
        }
    }
    // pos can be 256 only if all 256 inclusive prefixes are <= v, i.e. v >= total
    prefix_out = v - rem;
    return pos;
}

// AdaptiveIIDFreqModel.update_model (probability_models.py:71-92)
template <typename Tree>
SCL_HD void aec_model_update(Tree &F, const AecConst &c, uint32_t idx, uint64_t &total) {
    if (c.model != SCL_MODEL_ADAPTIVE_IID) return;
    fen_add1(F, idx);
    total += 1;
    if (total >= c.max_total) {  // halve everything, keeping each count >= 1
        fen_unbuild(F);
        uint64_t t = 0;
        for (uint32_t i = 1; i <= c.n_sym; ++i) {
            uint32_t h = F.get(i) >> 1;
            h = h > 1 ? h : 1;
            F.set(i, h);
            t += h;
        }
        fen_build(F);
        total = t;
    }
}

// ArithmeticEncoder.encode_block (arithmetic_coding.py:80-161)
template <typename Tree>
SCL_HD uint32_t aec_encode_lane(Tree &F, const AecTab &tab, const AecConst &c, uint64_t total, const uint8_t *sym,
                                uint32_t n, FwdBitWriter &w, uint64_t &bits_out, uint64_t &total_out) {
    const uint32_t P = c.P;
    const uint64_t FULL = 1ull << P, HALF = 1ull << (P - 1), QTR = 1ull << (P - 2);
    uint64_t low = 0, high = FULL, num_mid = 0;
    uint32_t st = SCL_ST_OK;
    if (c.DBSB < 32 && (n >> c.DBSB)) st = SCL_ST_OVERFLOW;
    w.put64((uint64_t)n, c.DBSB);
    for (uint32_t i = 0; i < n && st == SCL_ST_OK; ++i) {
        uint32_t idx = tab.sym2idx[sym[i]];
        if (idx == 0xFFFFu) {
            st = SCL_ST_BAD_SYMBOL;
            break;
        }
        if (!(total < QTR)) {  // arithmetic_coding.py:110-112
            st = SCL_ST_TOTAL_FREQ;
            break;
        }
        // shrink_range (:58-78)
        uint64_t cc = fen_prefix(F, idx), dd = fen_prefix(F, idx + 1);
        uint64_t rng = high - low;
        const double rcp_t = 1.0 / (double)total;
        high = low + div_exact_rcp(rng * dd, total, rcp_t);
        low = low + div_exact_rcp(rng * cc, total, rcp_t);
        aec_model_update(F, c, idx, total);  // :118
        while (high < HALF || low > HALF) {  // :126-143
            if (high < HALF) {
                w.put(0, 1);
                w.put_run(1, num_mid);
                low <<= 1;
                high <<= 1;
            } else {
                w.put(1, 1);
                w.put_run(0, num_mid);
                low = (low - HALF) << 1;
                high = (high - HALF) << 1;
            }
            num_mid = 0;
        }
        while (low > QTR && high < 3 * QTR) {  // :146-150
            num_mid += 1;
            low = (low - QTR) << 1;
            high = (high - QTR) << 1;
        }
        if (w.ovf) break;
    }
    num_mid += 1;  // :153-159
    if (low <= QTR) {
        w.put(0, 1);
        w.put_run(1, num_mid);
    } else {
        w.put(1, 1);
        w.put_run(0, num_mid);
    }
    bits_out = w.finish();
    total_out = total;
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// ArithmeticDecoder.decode_block (arithmetic_coding.py:203-287).  `A` = number of bits in the
// arithmetic part of the stream (arith_bitarray_size); bits past it read as zero (:258-261).
template <typename Tree>
SCL_HD uint32_t aec_decode_lane(Tree &F, const AecTab &tab, const AecConst &c, uint64_t total, BitReader &r,
                                uint64_t avail_bits, uint8_t *out, uint64_t out_cap, uint32_t &size_out,
                                uint64_t &bits_consumed, uint64_t &total_out) {
    const uint32_t P = c.P;
    const uint64_t FULL = 1ull << P, HALF = 1ull << (P - 1), QTR = 1ull << (P - 2);
    uint64_t size64 = r.get64(c.DBSB);
    size_out = 0;
    total_out = total;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    if (size64 == 0) return SCL_ST_EMPTY_BLOCK;
    const uint32_t size = (uint32_t)size64;
    const uint64_t A = avail_bits > c.DBSB ? avail_bits - c.DBSB : 0;
    uint64_t nbc = 0, low = 0, high = FULL, state = 0;
    while (nbc < P && nbc < A) {  // :222-228
        if (r.get(1)) state += 1ull << (P - nbc - 1);
        nbc += 1;
    }
    nbc = P;
    uint32_t st = SCL_ST_OK;
    for (uint32_t i = 0;;) {
        if (!(total < QTR)) {
            st = SCL_ST_TOTAL_FREQ;
            break;
        }
        // decode_step_core (:177-201): last idx with low + cum_idx*rng//T <= state
        //   <=>  cum_idx <= ((state - low + 1) * T - 1) // rng
        uint64_t rng = high - low;
        uint32_t idx;
        uint32_t cc;
        if (state < low) {
            idx = c.n_sym - 1;  // searchsorted -> 0, alphabet[-1]
            cc = fen_prefix(F, idx);
        } else {
            uint64_t v = div_exact_rcp((state - low + 1) * total - 1, rng, 1.0 / (double)rng);
            if (v >= total) v = total - 1;
            idx = fen_find(F, (uint32_t)v, cc);
        }
        uint64_t dd = fen_prefix(F, idx + 1);
        const double rcp_t = 1.0 / (double)total;
        high = low + div_exact_rcp(rng * dd, total, rcp_t);  // shrink_range
        low = low + div_exact_rcp(rng * (uint64_t)cc, total, rcp_t);
        out[i++] = tab.idx2sym[idx];
        aec_model_update(F, c, idx, total);
        if (i == size) break;  // :242-243, before renormalising
        while (high < HALF || low > HALF) {
            if (high < HALF) {
                low <<= 1;
                high <<= 1;
                state <<= 1;
            } else {
                low = (low - HALF) << 1;
                high = (high - HALF) << 1;
                state = (state - HALF) << 1;
            }
            if (nbc < A) state += r.get(1);
            nbc += 1;
        }
        while (low > QTR && high < 3 * QTR) {
            low = (low - QTR) << 1;
            high = (high - QTR) << 1;
            state = (state - QTR) << 1;
            if (nbc < A) state += r.get(1);
            nbc += 1;
        }
    }
    // trailing-bit accounting (:277-282)
    uint32_t extra = 0;
    for (extra = 0; extra < P; ++extra) {
        uint64_t state_low = (state >> extra) << extra;
        uint64_t state_high = state_low + (1ull << extra);
        if (state_low < low || state_high > high) break;
    }
    if (extra == P) extra = P - 1;
    size_out = size;
    total_out = total;
    bits_consumed = (uint64_t)((int64_t)nbc - ((int64_t)extra - 1) + (int64_t)c.DBSB);
    return st;
}

}  // namespace scl
