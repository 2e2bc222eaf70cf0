// scl_fast.cuh -- second-generation lane code for the 32-bit-state rANS fast path.
//
// The profile of the first kernels (profiles/r1a_*) showed this path is bound by ALU-pipe issue
// slots and L1/shared-memory wavefronts, not by DRAM: 47 warp-instructions per symbol, 19-22 of
// 32 lanes active (data-dependent spill/refill branches), 7x DRAM read amplification from
// scattered partial-sector accesses and 2.3x bank conflicts on the table reads.  This version:
//   * does all bit I/O on a FIXED schedule (every 2 symbols / every 16 symbols), predicated, so
//     the warp never diverges on a per-lane renormalisation event;
//   * packs bits with funnel shifts (one SHF inserts the low k bits of the state, no masks);
//   * moves coded words through a lane-interleaved shared-memory ring ([word][lane]: every lane
//     owns a bank, so ring traffic is conflict-free whatever the per-lane positions) and
//     touches HBM only in whole 32-byte sectors (LDG/STG.256);
//   * reads the encode table from 8 bank-rotated replicas so that LDS.128 never conflicts.
// The recurrences are unchanged (scl_lane.cuh documents them against the reference).
//
// Everything stays __host__ __device__: tests/host_emu runs these structs lane by lane on the
// CPU (ring stride 32 words there as well) against the oracle.
#pragma once
#include "scl_lane.cuh"

namespace scl {

constexpr uint32_t kRingStrideWords = 32;  // [word][lane] interleave: one bank per lane
constexpr uint32_t kEncRingWords = 16;     // per-lane capacity of the encoder's output ring
constexpr uint32_t kDecRingWords = 32;     // per-lane capacity of the decoder's input ring (+1 wrap duplicate)
constexpr uint32_t kEncTabCopies = 8;      // bank-rotated replicas of the 16-byte encode entries
constexpr uint32_t kFastMaxBitsPerSym = 16;

SCL_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t s) {  // (hi:lo) >> min(s,32), low word
#ifdef __CUDA_ARCH__
    return __funnelshift_rc(lo, hi, s);
#else
    return s >= 32 ? hi : funnel_r(lo, hi, s);
#endif
}
// x >> s on the FMA pipe (IMAD.HI) instead of the ALU pipe; s in 1..31 (compile-time constant use)
SCL_HD uint32_t shr_fma(uint32_t x, uint32_t s) {
#ifdef __CUDA_ARCH__
    uint32_t r, m = 1u << (32 - s);
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(m));
    return r;
#else
    return x >> s;
#endif
}

struct u32x8 {
    uint32_t v[8];
};
SCL_HD void st_sector32(uint8_t *p, const u32x8 &a) {  // one whole 32-byte sector
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]),
                 "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
                 : "memory");
#else
    for (int i = 0; i < 8; ++i) ((uint32_t *)p)[i] = a.v[i];
#endif
}
SCL_HD u32x8 ld_sector32(const uint8_t *p) {
    u32x8 a;
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.v[0]), "=r"(a.v[1]), "=r"(a.v[2]), "=r"(a.v[3]), "=r"(a.v[4]), "=r"(a.v[5]), "=r"(a.v[6]), "=r"(a.v[7])
                 : "l"(p));
#else
    for (int i = 0; i < 8; ++i) a.v[i] = ((const uint32_t *)p)[i];
#endif
    return a;
}

// ------------------------------------------------------------------------------------------------
// encoder lane
// ------------------------------------------------------------------------------------------------
// Bit accumulator: a 64-bit register pair (hi:lo) filled FROM THE TOP.  Inserting the low k bits
// of v is (hi:lo) = (v:hi:lo) >> k -- two funnel shifts, no mask.  The reference prepends every
// symbol's bits (rANS.py:158,196), i.e. earlier bits are less significant in the payload read as
// a big-endian integer; here earlier bits sink towards the LSB end, which is the same order.
// `room` = 64 - (valid bits); the oldest 32 valid bits are (hi:lo) >> room.
struct EncLaneV2 {
    uint32_t x;        // rANS state
    uint32_t lo, hi;   // accumulator
    uint32_t room;     // 64 - valid bits
    uint32_t wofs;     // words spilled so far * 128 (byte offset into the [word][lane] ring, unwrapped)
    uint32_t rofs;     // words drained so far * 128
    uint32_t *ring;    // this lane's column of the ring: word i at ring[(i % kEncRingWords) * 32]
    uint8_t *gend;     // slot end (32-byte aligned); drained word i lives at gend - 4*(i+1)
    uint8_t *gbegin;   // slot begin
    uint32_t ovf;
    uint32_t bad;

    SCL_HD void init(uint32_t L, uint32_t *ring_, uint8_t *slot_begin, uint8_t *slot_end) {
        x = L;
        lo = hi = 0;
        room = 64;
        wofs = rofs = 0;
        ring = ring_;
        gend = slot_end;
        gbegin = slot_begin;
        ovf = bad = 0;
    }

    // one symbol: shrink_state + rans_base_encode_step (rANS.py:138-161), see scl_lane.cuh
    template <uint32_t NBO, bool CHECK>
    SCL_HD void step(const RansEnc32 &e) {
        if (CHECK && e.pack == kRansEncInvalid) {
            bad = 1;
            return;
        }
        uint32_t k = ((e.pack >> 8) & 0xFFu) + (x > e.thresh_m1 ? NBO : 0u);
        lo = funnel_r(lo, hi, k);  // k <= 16 < 32
        hi = funnel_r(hi, x, k);
        room -= k;
        x >>= k;
        uint32_t q = funnel_r(umulhi32(x, e.rcp), 0u, e.pack);
        x = x + e.bias + q * (e.pack >> 16);
    }

    // after at most 2 symbols (<= 32 new bits): move one whole word to the ring if there is one
    SCL_HD void spill_check() {
        bool full = room <= 32;
        uint32_t w = funnel_rc(lo, hi, room);  // oldest 32 bits when full
        if (full) {
            ring[((wofs >> 7) & (kEncRingWords - 1)) * kRingStrideWords] = w;
            wofs += 128;
            room += 32;
        }
    }

    // after at most 16 symbols (<= 8 new words): move one 32-byte sector to HBM if 8 words wait
    SCL_HD void drain_check() {
        if (wofs - rofs >= 8 * 128) {
            u32x8 s;
            uint32_t base = rofs >> 7;
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j)  // word base+7-j goes to the lowest address first
                s.v[j] = bswap32(ring[((base + 7 - j) & (kEncRingWords - 1)) * kRingStrideWords]);
            uint8_t *dst = gend - 4 * (base + 8);
            if (dst >= gbegin)
                st_sector32(dst, s);
            else
                ovf = 1;
            rofs += 8 * 128;
        }
    }

    // generic insert for the header fields (k <= 32), with its own spill/drain checks
    SCL_HD void put(uint32_t v, uint32_t k) {
        if (k == 32) {
            lo = hi;
            hi = v;
        } else {
            lo = funnel_r(lo, hi, k);
            hi = funnel_r(hi, v, k);
        }
        room -= k;
        spill_check();
        drain_check();
    }
    SCL_HD void put64(uint64_t v, uint32_t k) {
        if (k > 32) {
            put((uint32_t)v, 32);
            put((uint32_t)(v >> 32), k - 32);
        } else {
            put((uint32_t)v, k);
        }
    }

    // write everything still buffered; returns the stream length in bits
    SCL_HD uint64_t finish() {
        uint32_t nvalid = 64 - room;  // < 32 after the last spill_check
        uint64_t words = wofs >> 7;
        uint64_t bits = words * 32 + nvalid;
        for (uint32_t i = rofs >> 7; i < (uint32_t)words; ++i) {
            uint8_t *dst = gend - 4 * ((uint64_t)i + 1);
            if (dst >= gbegin)
                st_word(dst, bswap32(ring[(i & (kEncRingWords - 1)) * kRingStrideWords]));
            else
                ovf = 1;
        }
        if (nvalid) {  // last partial word: the valid bits are its LOW bits (stream is right-aligned)
            uint32_t w = hi >> (room - 32);  // room in (32,64): the valid bits are the top bits of hi
            uint8_t *dst = gend - 4 * (words + 1);
            if (dst >= gbegin)
                st_word(dst, bswap32(w));
            else
                ovf = 1;
        }
        return bits;
    }
};

// Encode `cnt` (<= 16) symbols held in a 16-byte chunk.  The full-chunk path is fully unrolled
// with the spill check after every second symbol.
template <uint32_t NBO, bool CHECK>
SCL_HD void enc_chunk(EncLaneV2 &L, const RansEnc32 *tab, uint32_t tab_stride, const u32x4 &v, uint32_t cnt) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    if (cnt == 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                L.template step<NBO, CHECK>(tab[((wd[j] >> (8 * b)) & 0xFFu) * tab_stride]);
                if (b & 1) L.spill_check();
            }
        }
    } else {
        for (uint32_t i = 0; i < cnt; ++i) {
            L.template step<NBO, CHECK>(tab[((wd[i >> 2] >> (8 * (i & 3))) & 0xFFu) * tab_stride]);
            L.spill_check();
        }
    }
    L.drain_check();
}

// ------------------------------------------------------------------------------------------------
// decoder lane
// ------------------------------------------------------------------------------------------------
// The coded stream is pulled from HBM one whole 32-byte sector at a time into the lane's column
// of a [word][lane] ring (big-endian words, i.e. already byte-swapped), prefetched one 16-symbol
// group ahead.  Positions are in bits from the start of the first sector.
struct DecLaneV2 {
    uint32_t x;
    uint32_t bp;       // bit position of the next unread bit
    uint32_t filled;   // bits stored in the ring so far (multiple of 256)
    uint32_t *ring;    // word i at ring[(i % 32) * 32]; slot 32 duplicates slot 0 (wrap-free lookahead)
    const uint8_t *base;
    uint64_t in_bytes;
    uint64_t next_off;  // byte offset of the next sector to fetch
    u32x8 pf;           // prefetched sector
    uint32_t pf_valid;
    uint32_t start_bp;  // bp of the stream's first bit (for bits-consumed accounting)

    SCL_HD u32x8 load_sector(uint64_t off) const {
        if (off + 32 <= in_bytes) return ld_sector32(base + off);
        u32x8 a;
        for (int i = 0; i < 8; ++i) {
            uint32_t v = 0;
            for (int b = 3; b >= 0; --b) {
                uint64_t p = off + 4 * i + b;
                v = (v << 8) | (p < in_bytes ? base[p] : 0u);
            }
            a.v[i] = v;
        }
        return a;
    }
    SCL_HD void store_sector(const u32x8 &a) {
        uint32_t w0 = (filled >> 5) & (kDecRingWords - 1);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) ring[(w0 + j) * kRingStrideWords] = bswap32(a.v[j]);
        if (w0 == 0) ring[kDecRingWords * kRingStrideWords] = bswap32(a.v[0]);
        filled += 256;
    }
    SCL_HD void init(const uint8_t *base_, uint64_t in_bytes_, uint64_t bit_off, uint32_t *ring_) {
        base = base_;
        in_bytes = in_bytes_;
        ring = ring_;
        uint64_t byte0 = (bit_off >> 3) & ~31ull;
        next_off = byte0;
        bp = (uint32_t)(bit_off - 8 * byte0);
        start_bp = bp;
        filled = 0;
        pf_valid = 0;
        for (int i = 0; i < 3; ++i) {  // 96 bytes = 768 bits up front
            store_sector(load_sector(next_off));
            next_off += 32;
        }
    }
    // top 32 unread bits
    SCL_HD uint32_t peek32() const {
        uint32_t wi = (bp >> 5) & (kDecRingWords - 1);
        uint32_t w0 = ring[wi * kRingStrideWords], w1 = ring[(wi + 1) * kRingStrideWords];
        return funnel_l(w1, w0, bp);
    }
    SCL_HD uint32_t get(uint32_t k) {  // slow path (header fields), k <= 32
        uint32_t b = peek32();
        uint32_t v = k ? (b >> (32 - k)) : 0u;
        bp += k;
        return v;
    }
    SCL_HD uint64_t get64(uint32_t k) {
        if (k > 32) {
            uint64_t h = get(k - 32);
            return (h << 32) | get(32);
        }
        return get(k);
    }
    // group boundary: start fetching the next sector if fewer than 768 bits are buffered ...
    SCL_HD void prefetch_begin() {
        pf_valid = (filled - bp) < 768u;
        if (pf_valid) {
            pf = load_sector(next_off);
            next_off += 32;
        }
    }
    // ... and put it into the ring one group later
    SCL_HD void prefetch_end() {
        if (pf_valid) store_sector(pf);
    }
};

// two symbols from one 32-bit peek (k1 + k2 <= 32 is guaranteed by kFastMaxBitsPerSym)
#define SCL_DEC2_STEP(SYM)                                     \
    {                                                          \
        uint32_t e = lut[x & mmask];                           \
        x = (e >> 20) * (x >> mlog) + ((e >> 8) & 0xFFFu);     \
        SYM = e & 0xFFu;                                       \
        uint32_t k = rans32_renorm_bits(x, llog, NBO);         \
        x = funnel_l(bits, x, k);                              \
        bits <<= k;                                            \
        D.bp += k;                                             \
    }

// decode 16 symbols, last first, into 4 words (byte 3 of w[3] is the first one decoded)
template <uint32_t NBO>
SCL_HD void dec_group16(DecLaneV2 &D, const RansDec32 *lut, uint32_t mmask, uint32_t mlog, uint32_t llog, uint32_t w[4]) {
    uint32_t x = D.x;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        uint32_t acc = 0;
#pragma unroll
        for (int h = 1; h >= 0; --h) {
            uint32_t bits = D.peek32();
            uint32_t s1, s0;
            SCL_DEC2_STEP(s1);
            SCL_DEC2_STEP(s0);
            acc |= (s1 << (16 * h + 8)) | (s0 << (16 * h));
        }
        w[j] = acc;
    }
    D.x = x;
}

// rANSDecoder.decode_block (rANS.py:270-297) for one lane, v2 I/O.  `out` 32-byte aligned.
template <uint32_t NBO>
SCL_HD uint32_t rans32_decode_lane_v2(DecLaneV2 &D, const RansDec32 *lut, const RansConst &c, uint8_t *out, uint64_t out_cap,
                                      uint32_t &size_out, uint64_t &bits_consumed) {
    uint64_t size64 = D.get64(c.DBSB);
    D.x = D.get(c.NSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    const uint32_t size = (uint32_t)size64;
    const uint32_t mmask = (uint32_t)c.M - 1, mlog = c.m_log2, llog = c.l_log2;
    uint32_t p = size;
    // ragged head: bring p down to a multiple of 32 one symbol at a time
    while (p & 31) {
        uint32_t x = D.x, bits = D.peek32(), s;
        SCL_DEC2_STEP(s);
        D.x = x;
        out[--p] = (uint8_t)s;
        if ((p & 15) == 0) {  // keep the ring topped up on the same cadence as the main loop
            D.prefetch_begin();
            D.prefetch_end();
        }
    }
    while (p >= 32) {
        u32x8 o;
        D.prefetch_begin();
        dec_group16<NBO>(D, lut, mmask, mlog, llog, &o.v[4]);
        D.prefetch_end();
        D.prefetch_begin();
        dec_group16<NBO>(D, lut, mmask, mlog, llog, &o.v[0]);
        D.prefetch_end();
        p -= 32;
        st_sector32(out + p, o);
    }
    size_out = size;
    bits_consumed = D.bp - D.start_bp;
    return D.x == (uint32_t)c.L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;
}

}  // namespace scl
