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
constexpr uint32_t kTansTabCopies = 16;    // ... of the tANS encoder's 8-byte symbol rows (the same 128 bytes per byte value)
constexpr uint32_t kFastMaxBitsPerSym = 16;

SCL_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t s) {  // (hi:lo) >> min(s,32), low word
#ifdef __CUDA_ARCH__
    return __funnelshift_rc(lo, hi, s);
#else
    return s >= 32 ? hi : funnel_r(lo, hi, s);
#endif
}
// ---- shared-memory access through plain addresses --------------------------------------------
// On the device `saddr_t` is a 32-bit shared-window address and every access is ONE explicit
// ld/st.shared of the stated width (the compiler otherwise splits a 16-byte entry read into
// LDS.128 + 2x LDS.32, and the extra scalar loads are 4-way bank conflicted: profiles/r1b).
// On the host it is an ordinary pointer value.
#ifdef __CUDA_ARCH__
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t saddr_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32x4 lds128(saddr_t a) {
    u32x4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ u32x2 lds64(saddr_t a) {
    u32x2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ uint32_t lds32(saddr_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts32(saddr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#else
typedef uintptr_t saddr_t;
inline saddr_t saddr_of(const void *p) { return (uintptr_t)p; }
inline u32x4 lds128(saddr_t a) { return *(const u32x4 *)a; }
inline u32x2 lds64(saddr_t a) { return *(const u32x2 *)a; }
inline uint32_t lds32(saddr_t a) { return *(const uint32_t *)a; }
inline void sts32(saddr_t a, uint32_t v) { *(uint32_t *)a = v; }
#endif

// byte `b` of w, zero-extended: one PRMT
SCL_HD uint32_t byte_of(uint32_t w, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __byte_perm(w, 0, 0x4440u + b);
#else
    return (w >> (8 * b)) & 0xFFu;
#endif
}
// a*b + c on the FMA pipe (IMAD)
SCL_HD uint32_t mad32(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a * b + c;
#endif
}
// umulhi(x, m): with m = 2^(32-s) this is x >> s on the FMA pipe (IMAD.HI) instead of the ALU pipe
SCL_HD uint32_t mulhi_fma(uint32_t x, uint32_t m) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(m));
    return r;
#else
    return (uint32_t)(((uint64_t)x * m) >> 32);
#endif
}
// replace byte `pos` of acc with byte 0 of e: one PRMT
template <int POS>
SCL_HD uint32_t put_byte(uint32_t acc, uint32_t e) {
#ifdef __CUDA_ARCH__
    constexpr uint32_t sel = POS == 0 ? 0x3214u : POS == 1 ? 0x3240u : POS == 2 ? 0x3410u : 0x4210u;
    return __byte_perm(acc, e, sel);
#else
    return (acc & ~(0xFFu << (8 * POS))) | ((e & 0xFFu) << (8 * POS));
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
// RAW = the slot is private scratch of the fused packed encoder: words are stored as the numbers they are (no byte
// swap on the way out; the copy pool that reads them back saves the swap on the way in as well).  RAW = false is
// the public slot format: big-endian words, i.e. the bytes of the bit stream.
template <bool RAW>
struct EncLaneV2T {
    uint32_t x;        // rANS state
    uint32_t lo, hi;   // accumulator
    uint32_t room;     // 64 - valid bits
    uint32_t wofs;     // words spilled so far * 128 (byte offset into the [word][lane] ring, unwrapped)
    uint32_t rofs;     // words drained so far * 128
    saddr_t ring;      // this lane's column of the ring: word i at ring + (i % kEncRingWords) * 128
    uint8_t *gend;     // slot end (32-byte aligned); drained word i lives at gend - 4*(i+1)
    uint8_t *gbegin;   // slot begin
    uint32_t ovf;
    uint32_t bad;

    SCL_HD void init(uint32_t L, saddr_t ring_, uint8_t *slot_begin, uint8_t *slot_end) {
        x = L;
        lo = hi = 0;
        room = 64;
        wofs = rofs = 0;
        ring = ring_;
        gend = slot_end;
        gbegin = slot_begin;
        ovf = bad = 0;
    }
    // the device ring is 2 KiB-aligned, so the wrapped offset can be OR-ed in (one LOP3, no add)
    SCL_HD saddr_t ring_slot(uint32_t ofs) const {
#ifdef __CUDA_ARCH__
        return ring | (ofs & ((kEncRingWords - 1) * 128));
#else
        return ring + (ofs & ((kEncRingWords - 1) * 128));
#endif
    }

    // one symbol: shrink_state + rans_base_encode_step (rANS.py:138-161), see scl_lane.cuh.
    // e = {NBO == 1 ? ~thresh_m1 : thresh_m1, rcp, bias, cmpl << 16 | nb0 << 8 | shift}
    template <uint32_t NBO, bool CHECK>
    SCL_HD void step(const u32x4 &e) {
        if (CHECK && e.w == kRansEncInvalid) {
            bad = 1;
            return;
        }
        uint32_t k = byte_of(e.w, 1);
        if (NBO == 1) {
#ifdef __CUDA_ARCH__
            // (x > thresh_m1) is the carry of x + ~thresh_m1 (e.x holds the complement): IADD3 + IMAD.X
            // instead of ISETP + add + select
            asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(k) : "r"(x), "r"(e.x), "r"(k));
#else
            k += (x > ~e.x ? 1u : 0u);
#endif
        } else {
            k += (x > e.x ? NBO : 0u);
        }
        lo = funnel_r(lo, hi, k);  // k <= 16 < 32
        hi = funnel_r(hi, x, k);
        room -= k;
        x >>= k;
        uint32_t q = funnel_r(umulhi32(x, e.y), 0u, e.w);  // >> (shift = e.w & 31)
        x = mad32(q, mulhi_fma(e.w, 1u << 16), x + e.z);   // x + bias + q * cmpl
    }

    // after at most 2 symbols (<= 32 new bits): move one whole word to the ring if there is one
    SCL_HD void spill_check() {
        bool full = room <= 32;
        uint32_t w = funnel_rc(lo, hi, room);  // oldest 32 bits when full
        if (full) {
            sts32(ring_slot(wofs), w);
            wofs += 128;
            room += 32;
        }
    }

    // after at most 16 symbols (<= 8 new words): move one 32-byte sector to HBM if 8 words wait
    SCL_HD void drain_check() {
        if (wofs - rofs >= 8 * 128) {
            u32x8 s;
            // rofs is always a multiple of 8 words, so the group is the lower or upper half of the
            // 16-word ring: one address, eight immediate offsets
            const saddr_t g = ring_slot(rofs);
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j)  // word base+7-j goes to the lowest address first
                s.v[j] = RAW ? lds32(g + (7 - j) * 128) : bswap32(lds32(g + (7 - j) * 128));
            uint8_t *dst = gend - ((rofs >> 5) + 32);  // 4 * (words_drained + 8)
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
                st_word(dst, RAW ? lds32(ring_slot(i * 128)) : bswap32(lds32(ring_slot(i * 128))));
            else
                ovf = 1;
        }
        if (nvalid) {  // last partial word: the valid bits are its LOW bits (stream is right-aligned)
            uint32_t w = hi >> (room - 32);  // room in (32,64): the valid bits are the top bits of hi
            uint8_t *dst = gend - 4 * (words + 1);
            if (dst >= gbegin)
                st_word(dst, RAW ? w : bswap32(w));
            else
                ovf = 1;
        }
        return bits;
    }
};
typedef EncLaneV2T<false> EncLaneV2;

// Encode one full 16-symbol chunk (the hot path: fully unrolled, spill check after every 2nd symbol).
template <uint32_t NBO, bool CHECK, class Lane>
SCL_HD void enc_chunk16(Lane &L, saddr_t tab, uint32_t sym_stride, const u32x4 &v) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    // The table entries depend only on the symbols, not on the coder state: fetch a whole word's four
    // entries one word (4 symbols) ahead of their use.  Left to itself the compiler keeps the loads two
    // symbols ahead, and with the shared-memory pipe 80 % busy the first use of an entry was where 30 %
    // of the warp-stall samples landed (profiles/r1p, source view).
    u32x4 e[4], nx[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) e[b] = lds128(tab + (saddr_t)mad32(byte_of(wd[0], b), sym_stride, 0u));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < 3) {
#pragma unroll
            for (int b = 0; b < 4; ++b) nx[b] = lds128(tab + (saddr_t)mad32(byte_of(wd[j + 1], b), sym_stride, 0u));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            L.template step<NBO, CHECK>(e[b]);
            if (b & 1) L.spill_check();
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) e[b] = nx[b];
    }
    L.drain_check();
}

// Encode `cnt` (<= 16) symbols held in a 16-byte chunk.  `tab` = address of this lane's replica
// of entry 0; entry s is `sym_stride` bytes further per symbol (128 on the device: 8 replicas of
// 16 bytes, so the 8 lanes of a quarter-warp always hit 8 different 16-byte bank groups).
// The full-chunk path is fully unrolled with the spill check after every second symbol.
template <uint32_t NBO, bool CHECK, class Lane>
SCL_HD void enc_chunk(Lane &L, saddr_t tab, uint32_t sym_stride, const u32x4 &v, uint32_t cnt) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    if (cnt == 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                L.template step<NBO, CHECK>(lds128(tab + (saddr_t)mad32(byte_of(wd[j], b), sym_stride, 0u)));
                if (b & 1) L.spill_check();
            }
        }
    } else {
        for (uint32_t i = 0; i < cnt; ++i) {
            L.template step<NBO, CHECK>(lds128(tab + (saddr_t)(((wd[i >> 2] >> (8 * (i & 3))) & 0xFFu) * sym_stride)));
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
    saddr_t ring;      // word i at ring + (i % 32) * 128; slot 32 duplicates slot 0 (wrap-free lookahead)
    const uint8_t *base;
    uint64_t in_bytes;
    uint64_t next_off;  // byte offset of the next sector to fetch
    u32x8 pf;           // prefetched sector
    uint32_t pf_valid;
    uint32_t start_bp;  // bp of the stream's first bit (for bits-consumed accounting)

    SCL_HD u32x8 load_sector(uint64_t off) const {
        if (off + 32 <= in_bytes) return ld_sector32(base + off);
        // the last sectors of the buffer, byte by byte: rolled loops on purpose (this path is inlined at every
        // refill site; unrolled it was ~400 instructions per site and pushed the hot loops out of the instruction cache)
        u32x8 a;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            uint32_t v = 0;
#pragma unroll 1
            for (int b = 3; b >= 0; --b) {
                uint64_t p = off + 4 * i + b;
                v = (v << 8) | (p < in_bytes ? base[p] : 0u);
            }
            a.v[i] = v;
        }
        return a;
    }
    SCL_HD void store_sector(const u32x8 &a) {
        uint32_t o = (filled << 2) & ((kDecRingWords - 1) * 128);  // (filled / 32 mod 32) * 128
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) sts32(ring + o + j * 128, bswap32(a.v[j]));
        if (o == 0) sts32(ring + kDecRingWords * 128, bswap32(a.v[0]));
        filled += 256;
    }
    SCL_HD void init(const uint8_t *base_, uint64_t in_bytes_, uint64_t bit_off, saddr_t ring_) {
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
#ifdef __CUDA_ARCH__
        saddr_t a = mad32(bp & ((kDecRingWords - 1) * 32), 4u, ring);  // LOP3 + IMAD (the compiler's own form is three ops)
#else
        saddr_t a = ring + ((bp & ((kDecRingWords - 1) * 32)) << 2);
#endif
        uint32_t w0 = lds32(a), w1 = lds32(a + 128);
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
        pf_valid = 0;
    }
    // out-of-cadence refill (a consumer that read more than 256 bits since the last group boundary)
    SCL_HD void refill_now() {
        if (pf_valid) {
            store_sector(pf);
            pf_valid = 0;
        } else {
            store_sector(load_sector(next_off));
            next_off += 32;
        }
    }
};

// loop-invariant decode constants, precomputed so the inner loop is shifts-by-multiply on the FMA pipe
struct DecConst {
    saddr_t lut;      // address of LUT entry 0
    uint32_t m4;      // (M - 1) * 4
    uint32_t xq_mul;  // 2^(32 - log2 M): umulhi(x, xq_mul) = x >> log2 M
    uint32_t kbase;   // clz(x) - kbase = bits missing to reach L   (kbase = 31 - log2 L)
    uint32_t mask;    // M - 1
    uint32_t l_log2;
    uint32_t four, neg1;  // 4 and -1 as run-time values: ptxas turns IMAD by a literal power of two back into LEA / IADD3 (ALU pipe)
};

// One decode step (rans_base_decode_step + expand_state, rANS.py:234-260) on lut entry
// e = f << 20 | bias << 8 | byte.  k1 + k2 <= 32 for a pair is guaranteed by kFastMaxBitsPerSym.
// BAL = pipe-balanced form for throughput-bound launches (many warps per SM); with few warps the
// extra cross-pipe hops lengthen the per-symbol dependency chain, so small batches use BAL = false
// (measured: +2 % at 55 tasks per SM, -3 % at 14).
template <uint32_t NBO, int POS, bool BAL = false>
SCL_HD void dec_step(const DecConst &c, uint32_t &x, uint32_t &bits, uint32_t &ksum, uint32_t &acc) {
    // The ALU pipe (LOP3/SHF/PRMT/IADD3, one warp-instruction per 2 cycles) was at 69 % against 22 % for the
    // FMA pipe (profiles/r1h): with BAL every shift / add that has an IMAD form is issued there instead.
    uint32_t e, bias, d;
#ifdef __CUDA_ARCH__
    if (BAL) {
        e = lds32(mad32(x & c.mask, c.four, c.lut));  // LOP3 + IMAD
        bias = mulhi_fma(e, 1u << 24) & 0xFFFu;       // e >> 8 as IMAD.HI, then LOP3
    } else
#endif
    {
        e = lds32(c.lut + (saddr_t)((x << 2) & c.m4));
        bias = (e >> 8) & 0xFFFu;
    }
    uint32_t xq = mulhi_fma(x, c.xq_mul);
    uint32_t f = mulhi_fma(e, 1u << 12);  // e >> 20
    x = mad32(f, xq, bias);
    acc = put_byte<POS>(acc, e);
    // d = bits missing to reach L = log2 L - (position of the top set bit); d >= 1 - NBO because x < 2^(l+NBO)
#ifdef __CUDA_ARCH__
    if (BAL) {
        asm("bfind.u32 %0, %1;" : "=r"(d) : "r"(x));
        d = mad32(d, c.neg1, c.l_log2);
    } else
#endif
        d = clz32(x) - c.kbase;
    uint32_t k;
    if (NBO == 1) {
        k = d;
    } else if ((NBO & (NBO - 1)) == 0) {
        k = (d + (NBO - 1)) & ~(NBO - 1);  // NBO * ceil(d / NBO), and 0 for d <= 0 (d + NBO - 1 >= 0)
    } else {
        int32_t ds = (int32_t)d;
        k = ds <= 0 ? 0u : (((uint32_t)ds + NBO - 1) / NBO) * NBO;
    }
    x = funnel_l(bits, x, k);
    bits <<= k;
    ksum += k;
}

// decode 16 symbols, last first, into 4 words (byte 3 of w[3] is the first one decoded)
template <uint32_t NBO, bool BAL = false>
SCL_HD void dec_group16(DecLaneV2 &D, const DecConst &c, uint32_t w[4]) {
    uint32_t x = D.x;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        uint32_t acc = 0;
        {
            uint32_t bits = D.peek32(), ks = 0;
            dec_step<NBO, 3, BAL>(c, x, bits, ks, acc);
            dec_step<NBO, 2, BAL>(c, x, bits, ks, acc);
            D.bp += ks;
        }
        {
            uint32_t bits = D.peek32(), ks = 0;
            dec_step<NBO, 1, BAL>(c, x, bits, ks, acc);
            dec_step<NBO, 0, BAL>(c, x, bits, ks, acc);
            D.bp += ks;
        }
        w[j] = acc;
    }
    D.x = x;
}

// ------------------------------------------------------------------------------------------------
// tANS on the same I/O machinery (tANS.py:126-157 / :239-250): the steps are pure table reads.
//   symbol entry (16 B, 8 bank-rotated replicas) = {thresh, nb0, row - min_shrunk, pad}
//   enc_table[row + x_shrunk] = next state        dec_packed[x - L] = x_shrunk << 8 | byte
// ------------------------------------------------------------------------------------------------
template <bool CHECK, class Lane>
SCL_HD void tans_enc_step(Lane &L, const u32x2 &e, saddr_t enc_table) {  // e = TansSym8 {thresh, row << 7 | nb0}
    if (CHECK && e.y == 0xFFFFFFFFu) {
        L.bad = 1;
        return;
    }
    uint32_t k = (e.y & 31u) + (L.x >= e.x ? 1u : 0u);  // shrink_state_num_out_bits_base + threshold test
    L.lo = funnel_r(L.lo, L.hi, k);
    L.hi = funnel_r(L.hi, L.x, k);
    L.room -= k;
    L.x = lds32(enc_table + (saddr_t)(int32_t)(((int32_t)e.y >> 5) + (int32_t)((L.x >> k) << 2)));  // base_encode_step_table[(s, x_shrunk)]
}

template <bool CHECK, class Lane>
SCL_HD void tans_enc_chunk16(Lane &L, saddr_t symtab, uint32_t sym_stride, saddr_t enc_table, const u32x4 &v) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    u32x2 e[4], nx[4];  // per-symbol rows fetched one word ahead of their use, as in enc_chunk16
#pragma unroll
    for (int b = 0; b < 4; ++b) e[b] = lds64(symtab + (saddr_t)mad32(byte_of(wd[0], b), sym_stride, 0u));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < 3) {
#pragma unroll
            for (int b = 0; b < 4; ++b) nx[b] = lds64(symtab + (saddr_t)mad32(byte_of(wd[j + 1], b), sym_stride, 0u));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            tans_enc_step<CHECK>(L, e[b], enc_table);
            if (b & 1) L.spill_check();
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) e[b] = nx[b];
    }
    L.drain_check();
}

template <bool CHECK, class Lane>
SCL_HD void tans_enc_chunk(Lane &L, saddr_t symtab, uint32_t sym_stride, saddr_t enc_table, const u32x4 &v, uint32_t cnt) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    if (cnt == 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                tans_enc_step<CHECK>(L, lds64(symtab + (saddr_t)mad32(byte_of(wd[j], b), sym_stride, 0u)), enc_table);
                if (b & 1) L.spill_check();
            }
        }
    } else {
        for (uint32_t i = 0; i < cnt; ++i) {
            tans_enc_step<CHECK>(L, lds64(symtab + (saddr_t)(((wd[i >> 2] >> (8 * (i & 3))) & 0xFFu) * sym_stride)), enc_table);
            L.spill_check();
        }
    }
    L.drain_check();
}

struct TansDecConst {
    saddr_t dec;      // address of dec_packed[0]
    uint32_t L;       // table size (power of two)
    uint32_t lmask4;  // (L - 1) * 4
    uint32_t kbase;   // 32 - NUM_STATE_BITS: k = clz(x_shrunk) - kbase
    uint32_t lmask, nsb_m1, four, neg1;
};

template <int POS, bool BAL = false>
SCL_HD void tans_dec_step(const TansDecConst &c, uint32_t &x, uint32_t &bits, uint32_t &ksum, uint32_t &acc) {
    // a corrupt state (outside [L, 2L)) wraps inside the table instead of faulting; the final
    // state check reports it (the reference raises KeyError from its dict)
    uint32_t e, xs, k;
#ifdef __CUDA_ARCH__
    if (BAL) {  // same pipe balancing as dec_step: address, e >> 8 and the subtraction go to the FMA pipe
        e = lds32(mad32(x & c.lmask, c.four, c.dec));  // (x - L) mod L == x mod L
        xs = mulhi_fma(e, 1u << 24);
        asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(xs));
        k = mad32(k, c.neg1, c.nsb_m1);  // NUM_STATE_BITS - bitwidth(x_shrunk): expand_state_num_bits_table (tANS.py:225)
    } else
#endif
    {
        e = lds32(c.dec + (saddr_t)(((x - c.L) << 2) & c.lmask4));
        xs = e >> 8;
        k = clz32(xs) - c.kbase;  // expand_state_num_bits_table (tANS.py:225)
    }
    acc = put_byte<POS>(acc, e);
    x = funnel_l(bits, xs, k);
    bits <<= k;
    ksum += k;
}

template <bool BAL = false>
SCL_HD void tans_dec_group16(DecLaneV2 &D, const TansDecConst &c, uint32_t w[4]) {
    uint32_t x = D.x;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
        uint32_t acc = 0;
        {
            uint32_t bits = D.peek32(), ks = 0;
            tans_dec_step<3, BAL>(c, x, bits, ks, acc);
            tans_dec_step<2, BAL>(c, x, bits, ks, acc);
            D.bp += ks;
        }
        {
            uint32_t bits = D.peek32(), ks = 0;
            tans_dec_step<1, BAL>(c, x, bits, ks, acc);
            tans_dec_step<0, BAL>(c, x, bits, ks, acc);
            D.bp += ks;
        }
        w[j] = acc;
    }
    D.x = x;
}

// ------------------------------------------------------------------------------------------------
// Stepper policies: the same decode driver (header, ragged head, 16-symbol groups) serves rANS
// and tANS, and the kernels reuse the pieces for their tile-store main loop.
// ------------------------------------------------------------------------------------------------
template <uint32_t NBO, bool BAL = false>
struct RansStepper {
    DecConst dc;
    uint32_t l_log2, m_log2;
    SCL_HD void init(saddr_t lut, const RansConst &c) {
        dc.lut = lut;
        dc.m4 = ((uint32_t)c.M - 1) << 2;
        dc.xq_mul = c.m_log2 ? (1u << (32 - c.m_log2)) : 0u;
        dc.kbase = 31 - c.l_log2;
        dc.mask = (uint32_t)c.M - 1;
        dc.l_log2 = c.l_log2;
        dc.four = 4u + (c.l_log2 >> 8);  // l_log2 < 64 or 0xFFFFFFFF (never on this path): opaque to the compiler
        dc.neg1 = 0xFFFFFFFFu - (c.l_log2 >> 8);
        l_log2 = c.l_log2;
        m_log2 = c.m_log2;
    }
    SCL_HD bool degenerate() const { return m_log2 == 0; }  // M == 1: x >> 0 is not a multiply-high
    SCL_HD uint32_t one(DecLaneV2 &D) const {               // one symbol, returned as a byte
        if (degenerate()) {
            uint32_t e = lds32(dc.lut);
            uint32_t x = (e >> 20) * D.x + ((e >> 8) & 0xFFFu);
            uint32_t k = rans32_renorm_bits(x, l_log2, NBO);
            D.x = funnel_l(D.peek32(), x, k);
            D.bp += k;
            return e & 0xFFu;
        }
        uint32_t x = D.x, bits = D.peek32(), ks = 0, acc = 0;
        dec_step<NBO, 0>(dc, x, bits, ks, acc);
        D.x = x;
        D.bp += ks;
        return acc & 0xFFu;
    }
    SCL_HD void group16(DecLaneV2 &D, uint32_t w[4]) const { dec_group16<NBO, BAL>(D, dc, w); }
};

template <bool BAL = false>
struct TansStepper {
    TansDecConst tc;
    SCL_HD void init(saddr_t dec, const RansConst &c) {
        tc.dec = dec;
        tc.L = (uint32_t)c.L;
        tc.lmask4 = ((uint32_t)c.L - 1) << 2;
        tc.kbase = 32 - c.NSB;
        tc.lmask = (uint32_t)c.L - 1;
        tc.nsb_m1 = c.NSB - 1;
        tc.four = 4u + (c.NSB >> 8);  // opaque 4 / -1, see DecConst
        tc.neg1 = 0xFFFFFFFFu - (c.NSB >> 8);
    }
    SCL_HD bool degenerate() const { return false; }
    SCL_HD uint32_t one(DecLaneV2 &D) const {
        uint32_t x = D.x, bits = D.peek32(), ks = 0, acc = 0;
        tans_dec_step<0>(tc, x, bits, ks, acc);
        D.x = x;
        D.bp += ks;
        return acc & 0xFFu;
    }
    SCL_HD void group16(DecLaneV2 &D, uint32_t w[4]) const { tans_dec_group16<BAL>(D, tc, w); }
};

// header: [size : DBSB][state : NSB]; returns false (with *st set) when the size does not fit
SCL_HD bool dec_read_header(DecLaneV2 &D, const RansConst &c, uint64_t out_cap, uint32_t &size, uint32_t &st) {
    uint64_t size64 = D.get64(c.DBSB);
    D.x = D.get(c.NSB);
    size = 0;
    if (size64 > out_cap) {
        st = SCL_ST_OVERFLOW;
        return false;
    }
    size = (uint32_t)size64;
    return true;
}

// decode single symbols (stored bytewise) until p is a multiple of `align` (16, 32 or 64)
template <class S>
SCL_HD void dec_head(DecLaneV2 &D, const S &s, uint8_t *out, uint32_t &p, uint32_t align) {
    while (p & (align - 1)) {
        out[--p] = (uint8_t)s.one(D);
        if ((p & 15) == 0) {  // keep the ring topped up on the same cadence as the main loop
            D.prefetch_begin();
            D.prefetch_end();
        }
    }
}

// per-lane sector stores for the rest of the block (p a multiple of 32)
template <class S>
SCL_HD void dec_body_sectors(DecLaneV2 &D, const S &s, uint8_t *out, uint32_t &p) {
    while (p >= 32) {
        u32x8 o;
        D.prefetch_begin();
        s.group16(D, &o.v[4]);
        D.prefetch_end();
        D.prefetch_begin();
        s.group16(D, &o.v[0]);
        D.prefetch_end();
        p -= 32;
        st_sector32(out + p, o);
    }
}

template <class S>
SCL_HD uint32_t decode_lane_generic(DecLaneV2 &D, const S &s, const RansConst &c, uint8_t *out, uint64_t out_cap, uint32_t &size_out,
                                    uint64_t &bits_consumed) {
    uint32_t size, st = SCL_ST_OK;
    size_out = 0;
    if (!dec_read_header(D, c, out_cap, size, st)) return st;
    uint32_t p = size;
    if (s.degenerate()) {
        dec_head(D, s, out, p, 1u << 31);  // everything symbol by symbol
        while (p) out[--p] = (uint8_t)s.one(D);
    }
    dec_head(D, s, out, p, 32);
    dec_body_sectors(D, s, out, p);
    size_out = size;
    bits_consumed = D.bp - D.start_bp;
    return D.x == (uint32_t)c.L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;
}

// rANSDecoder.decode_block (rANS.py:270-297) / tANSDecoder.decode_block (tANS.py:252-279) for one
// lane with per-lane sector stores (`out` 32-byte aligned)
template <uint32_t NBO>
SCL_HD uint32_t rans32_decode_lane_v2(DecLaneV2 &D, saddr_t lut, const RansConst &c, uint8_t *out, uint64_t out_cap, uint32_t &size_out,
                                      uint64_t &bits_consumed) {
    RansStepper<NBO> s;
    s.init(lut, c);
    return decode_lane_generic(D, s, c, out, out_cap, size_out, bits_consumed);
}
SCL_HD uint32_t tans_decode_lane_v2(DecLaneV2 &D, saddr_t dec, const RansConst &c, uint8_t *out, uint64_t out_cap, uint32_t &size_out,
                                    uint64_t &bits_consumed) {
    TansStepper<> s;
    s.init(dec, c);
    return decode_lane_generic(D, s, c, out, out_cap, size_out, bits_consumed);
}

}  // namespace scl
