// scl_aec.cuh -- second-generation arithmetic-coder lanes (arithmetic_coding.py:58-161,177-287 with
// FixedFreqModel / AdaptiveIIDFreqModel, probability_models.py:57-92).
//
// The first-generation lanes (scl_lane.cuh) follow the reference loop by loop; on the GPU that means
// data-dependent trip counts everywhere (Fenwick walks, one renormalisation iteration per bit) and
// the profile shows it: 1146 warp-instructions per symbol with 15 of 32 lanes active
// (profiles/r1e).  Here every per-symbol step has a fixed instruction sequence:
//   * the 256 counters live in a two-level radix-16 structure of packed 16-bit pairs
//     (8 words of group totals + 128 words of counts per lane, lane-interleaved in shared memory);
//     cumulative counts are masked dot products (dp2a) over 8 words per level;
//   * the E1/E2 loop (":126-143") and the E3 loop (":146-150") are evaluated in closed form from
//     leading/trailing bit counts, including the reference's strict `<` / `>` boundary cases;
//   * x // T is one FP64 multiply with an exact integer correction (div_exact_rcp).
// Usable when every counter and group total stays below 65536 (host-checked); otherwise the
// first-generation kernel runs.  __host__ __device__ throughout: tests/host_emu checks these
// against the oracle on the CPU.
#pragma once
#include <math.h>

#include "scl_fast.cuh"

#ifndef __CUDACC__
struct ulonglong2 {
    unsigned long long x, y;
};
#endif

namespace scl {

constexpr uint32_t kAecModelWords = 8 + 128;  // G[16] + C[256] as u16 pairs

SCL_HD uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {  // c + a.u16[0]*b.u8[0] + a.u16[1]*b.u8[1]
#ifdef __CUDA_ARCH__
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xFFFFu) * (b & 0xFFu) + (a >> 16) * ((b >> 8) & 0xFFu);
#endif
}
SCL_HD uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {  // c + a.u16[0]*b.u8[2] + a.u16[1]*b.u8[3]
#ifdef __CUDA_ARCH__
    return __dp2a_hi(a, b, c);
#else
    return c + (a & 0xFFFFu) * ((b >> 16) & 0xFFu) + (a >> 16) * (b >> 24);
#endif
}
SCL_HD uint32_t ctz32(uint32_t x) {  // 32 for x == 0
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)__brev(x));
#else
    return x ? (uint32_t)__builtin_ctz(x) : 32u;
#endif
}

// The model of one lane.  `w` = address of word 0 of this lane; word i is `stride` bytes further
// (128 on the device: [word][lane] interleave, one bank per lane; 4 on the host).
// `masks` = address of this lane's replica of the prefix-mask table: entry t (0..16) is 16 bytes
// whose first t bytes are 1 (entry stride `mstride`: 128 on the device = 8 replicas x 16 B).
struct AecModel {
    saddr_t w;
    uint32_t stride;
    saddr_t masks;
    uint32_t mstride;
    SCL_HD uint32_t word(uint32_t i) const { return lds32(w + (saddr_t)(i * stride)); }
    SCL_HD void set_word(uint32_t i, uint32_t v) const { sts32(w + (saddr_t)(i * stride), v); }

    // sum of the first t (0..16) of the 16 packed values in words [first, first+8)
    SCL_HD uint32_t masked_sum(uint32_t first, uint32_t t) const {
        const u32x4 m = lds128(masks + (saddr_t)(t * mstride));
        const saddr_t a = w + (saddr_t)(first * stride);
        uint32_t s = 0;
        s = dp2a_lo(lds32(a), m.x, s);
        s = dp2a_hi(lds32(a + (saddr_t)(1 * stride)), m.x, s);
        s = dp2a_lo(lds32(a + (saddr_t)(2 * stride)), m.y, s);
        s = dp2a_hi(lds32(a + (saddr_t)(3 * stride)), m.y, s);
        s = dp2a_lo(lds32(a + (saddr_t)(4 * stride)), m.z, s);
        s = dp2a_hi(lds32(a + (saddr_t)(5 * stride)), m.z, s);
        s = dp2a_lo(lds32(a + (saddr_t)(6 * stride)), m.w, s);
        s = dp2a_hi(lds32(a + (saddr_t)(7 * stride)), m.w, s);
        return s;
    }
    SCL_HD uint32_t count(uint32_t idx) const {  // freq of alphabet index idx
        uint32_t wv = word(8 + (idx >> 1));
        return (idx & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    }
    // cumulative count below idx (cumulative_freq_dict, prob_dist.py:193-205) and the count itself
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const {
        const uint32_t hi = idx >> 4, lo = idx & 15;
        cum = masked_sum(0, hi) + masked_sum(8 + hi * 8, lo);
        f = count(idx);
    }
    SCL_HD void add1(uint32_t idx) const {  // freq_dict[s] += 1
        const uint32_t cw = 8 + (idx >> 1), gw = idx >> 5;
        set_word(cw, word(cw) + ((idx & 1) ? 0x10000u : 1u));
        set_word(gw, word(gw) + (((idx >> 4) & 1) ? 0x10000u : 1u));
    }
    // last index whose cumulative count is <= v (numpy.searchsorted(side="right") - 1); v < total
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        uint32_t pre = 0, hi = 0, base = 0;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            uint32_t wv = word(j);
            uint32_t t1 = pre + (wv & 0xFFFFu), t2 = t1 + (wv >> 16);
            bool c1 = t1 <= v, c2 = t2 <= v;  // inclusive prefixes are non-decreasing: true..true false..false
            hi += (c1 ? 1u : 0u) + (c2 ? 1u : 0u);
            base = c2 ? t2 : (c1 ? t1 : base);
            pre = t2;
        }
        hi = hi > 15 ? 15 : hi;
        uint32_t lo = 0, b2 = base;
        pre = base;
        const uint32_t first = 8 + hi * 8;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            uint32_t wv = word(first + j);
            uint32_t t1 = pre + (wv & 0xFFFFu), t2 = t1 + (wv >> 16);
            bool c1 = t1 <= v, c2 = t2 <= v;
            lo += (c1 ? 1u : 0u) + (c2 ? 1u : 0u);
            b2 = c2 ? t2 : (c1 ? t1 : b2);
            pre = t2;
        }
        lo = lo > 15 ? 15 : lo;
        const uint32_t idx = hi * 16 + lo;
        cum = b2;
        f = count(idx);
        return idx;
    }
    SCL_HD void load(const uint32_t *init_freq, const uint64_t *model, uint32_t n_sym, uint64_t &total) const {
        total = 0;
        for (uint32_t g = 0; g < 16; g += 2) {
            uint32_t gs[2] = {0, 0};
            for (uint32_t h = 0; h < 2; ++h)
                for (uint32_t j = 0; j < 16; j += 2) {
                    uint32_t i0 = (g + h) * 16 + j;
                    uint32_t f0 = i0 < n_sym ? (model ? (uint32_t)model[i0] : init_freq[i0]) : 0u;
                    uint32_t f1 = i0 + 1 < n_sym ? (model ? (uint32_t)model[i0 + 1] : init_freq[i0 + 1]) : 0u;
                    set_word(8 + (i0 >> 1), f0 | (f1 << 16));
                    gs[h] += f0 + f1;
                }
            set_word(g >> 1, gs[0] | (gs[1] << 16));
            total += gs[0] + gs[1];
        }
    }
    SCL_HD void store(uint64_t *model, uint32_t n_sym) const {
        for (uint32_t i = 0; i < n_sym; ++i) model[i] = count(i);
    }
    // AdaptiveIIDFreqModel's halving (probability_models.py:90-92): f = max(f // 2, 1) for every symbol
    SCL_HD void halve(uint32_t n_sym, uint64_t &total) const {
        total = 0;
        for (uint32_t g = 0; g < 16; ++g) {
            uint32_t gsum = 0;
            for (uint32_t j = 0; j < 16; ++j) {
                uint32_t i = g * 16 + j;
                if (i < n_sym) {
                    uint32_t h = count(i) >> 1;
                    h = h > 1 ? h : 1;
                    uint32_t cw = 8 + (i >> 1), wv = word(cw);
                    set_word(cw, (i & 1) ? ((wv & 0xFFFFu) | (h << 16)) : ((wv & 0xFFFF0000u) | h));
                    gsum += h;
                }
            }
            uint32_t gw = g >> 1, wv = word(gw);
            set_word(gw, (g & 1) ? ((wv & 0xFFFFu) | (gsum << 16)) : ((wv & 0xFFFF0000u) | gsum));
            total += gsum;
        }
    }
};

// ---- model policies: what the coder lanes ask of a frequency model ---------------------------
//   total()            sum of freqs_current
//   query(idx, c, f)   cumulative count below idx and the count of idx
//   find(v, c, f)      last idx with cumulative count <= v  (v < total)
//   update(idx)        update_model(s); returns a status word (the reference raises)

// FixedFreqModel / AdaptiveIIDFreqModel (probability_models.py:57-92) over the two-level structure
struct AecIidPolicy {
    AecModel M;
    uint32_t tot, adaptive, max_total, n_sym;
    SCL_HD uint32_t total() const { return tot; }
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const { M.query(idx, cum, f); }
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        uint32_t idx = M.find(v, cum, f);
        if (idx >= n_sym) {
            idx = n_sym - 1;
            M.query(idx, cum, f);
        }
        return idx;
    }
    SCL_HD uint32_t update(uint32_t idx) {
        if (adaptive) {
            M.add1(idx);
            tot += 1;
            if (tot >= max_total) {  // :90-92
                uint64_t t64 = tot;
                M.halve(n_sym, t64);
                tot = (uint32_t)t64;
            }
        }
        return SCL_ST_OK;
    }
};

// ---- the same model at half the shared memory: 8-bit counters ----------------------------------------------
// profiles/r1s: the arithmetic coder is latency-bound with 11-12 resident warps per SM, and what caps the residency
// is the model (136 words = 17 KiB per warp).  With one BYTE per counter it is 8 + 64 words (9 KiB per warp, 20
// warps per SM).  A counter that passes 255 wraps and its high part moves to a short per-lane list held in a
// register (`big`: up to kAecBigMax entries of sym : 8 | high : 4) -- exact, and almost never populated: at cfg4
// (1 KiB blocks from 256 ones) the most frequent Zipf symbol reaches ~205.  The host selects this model only when
// the list cannot overflow: every initial count <= 255 and (sum of initial counts + block length) / 256 <= kAecBigMax
// symbols can ever reach 256.  Group totals stay 16-bit (the whole table sums to < 65 536 under that bound).
// All paths that involve `big` are slow generic loops on purpose.
constexpr uint32_t kAecModel8Words = 8 + 64;  // G[16] as u16 pairs + C[256] as bytes
constexpr uint32_t kAecBigMax = 5;

SCL_HD uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) {  // c + sum of a.u8[i] * b.u8[i]
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFFu) * ((b >> (8 * i)) & 0xFFu);
    return c;
#endif
}

struct AecModel8 {
    saddr_t w;
    uint32_t stride;
    saddr_t masks;
    uint32_t mstride;
    uint64_t big;  // kAecBigMax x 12 bits: sym | high << 8, high >= 1 for a live entry
    uint32_t ovf;
    SCL_HD uint32_t word(uint32_t i) const { return lds32(w + (saddr_t)(i * stride)); }
    SCL_HD void set_word(uint32_t i, uint32_t v) const { sts32(w + (saddr_t)(i * stride), v); }
    SCL_HD uint32_t hi_of(uint32_t idx) const {
        if (big == 0) return 0;
        uint32_t h = 0;
        for (uint32_t k = 0; k < kAecBigMax; ++k) {
            const uint32_t e = (uint32_t)(big >> (12 * k)) & 0xFFFu;
            if (e && (e & 0xFFu) == idx) h = e >> 8;
        }
        return h;
    }
    SCL_HD void big_inc(uint32_t idx) {  // the counter of idx wrapped: its high part grows by one
        int free_k = -1;
        for (uint32_t k = 0; k < kAecBigMax; ++k) {
            const uint32_t e = (uint32_t)(big >> (12 * k)) & 0xFFFu;
            if (e && (e & 0xFFu) == idx) {
                big += 1ull << (12 * k + 8);
                return;
            }
            if (!e && free_k < 0) free_k = (int)k;
        }
        if (free_k < 0) {
            ovf = 1;  // cannot happen under the host's bound
            return;
        }
        big |= (uint64_t)(idx | 0x100u) << (12 * free_k);
    }
    SCL_HD uint32_t lo_of(uint32_t idx) const { return (word(8 + (idx >> 2)) >> (8 * (idx & 3))) & 0xFFu; }
    SCL_HD uint32_t count(uint32_t idx) const { return lo_of(idx) + (hi_of(idx) << 8); }
    // sum of the first t (0..16) group totals (16-bit pairs in words 0..7): as AecModel::masked_sum
    SCL_HD uint32_t group_prefix(uint32_t t) const {
        const u32x4 m = lds128(masks + (saddr_t)(t * mstride));
        uint32_t s = 0;
        s = dp2a_lo(word(0), m.x, s);
        s = dp2a_hi(word(1), m.x, s);
        s = dp2a_lo(word(2), m.y, s);
        s = dp2a_hi(word(3), m.y, s);
        s = dp2a_lo(word(4), m.z, s);
        s = dp2a_hi(word(5), m.z, s);
        s = dp2a_lo(word(6), m.w, s);
        s = dp2a_hi(word(7), m.w, s);
        return s;
    }
    // sum of the first t (0..16) byte counters of group g (words 8 + 4g .. +3), low parts only
    SCL_HD uint32_t byte_prefix(uint32_t g, uint32_t t) const {
        const u32x4 m = lds128(masks + (saddr_t)(t * mstride));
        const saddr_t a = w + (saddr_t)((8 + 4 * g) * stride);
        uint32_t s = 0;
        s = dp4a_u(lds32(a), m.x, s);
        s = dp4a_u(lds32(a + (saddr_t)stride), m.y, s);
        s = dp4a_u(lds32(a + (saddr_t)(2 * stride)), m.z, s);
        s = dp4a_u(lds32(a + (saddr_t)(3 * stride)), m.w, s);
        return s;
    }
    SCL_HD uint32_t big_below(uint32_t g, uint32_t lo) const {  // high parts of the symbols of group g below position lo
        uint32_t s = 0;
        for (uint32_t k = 0; k < kAecBigMax; ++k) {
            const uint32_t e = (uint32_t)(big >> (12 * k)) & 0xFFFu, sy = e & 0xFFu;
            if (e && (sy >> 4) == g && (sy & 15u) < lo) s += (e >> 8) << 8;
        }
        return s;
    }
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const {
        const uint32_t g = idx >> 4, lo = idx & 15;
        cum = group_prefix(g) + byte_prefix(g, lo);
        f = lo_of(idx);
        if (big) {
            cum += big_below(g, lo);
            f += hi_of(idx) << 8;
        }
    }
    SCL_HD void add1(uint32_t idx) {
        const uint32_t cw = 8 + (idx >> 2), sh = 8 * (idx & 3), gw = idx >> 5;
        const uint32_t wv = word(cw);
        if (((wv >> sh) & 0xFFu) == 0xFFu) {
            set_word(cw, wv - (0xFFu << sh));  // 255 -> 0, carry into the list instead of the neighbour byte
            big_inc(idx);
        } else {
            set_word(cw, wv + (1u << sh));
        }
        set_word(gw, word(gw) + (((idx >> 4) & 1) ? 0x10000u : 1u));
    }
    // last index whose cumulative count is <= v (numpy.searchsorted(side="right") - 1); v < total.
    // Two levels, each searched hierarchically: the inclusive prefixes are non-decreasing, so "how many are <= v" and
    // "the largest one <= v" locate the group (then the counter) -- first by whole words (one dot product per word
    // gives the running prefix), then inside the one word that holds the answer.
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        uint32_t g, base;
        {
            uint32_t P = 0, nw = 0, bw = 0;  // words (pairs of groups) wholly <= v, and their prefix
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
                P = dp2a_lo(word(j), 0x0101u, P);  // + both 16-bit group totals of word j
                const bool c = P <= v;
                nw += c ? 1u : 0u;
                bw = c ? P : bw;
            }
            const uint32_t wv = word(nw < 8 ? nw : 7);  // the word the answer lies in (v < total: nw <= 7)
            const uint32_t t1 = bw + (wv & 0xFFFFu);
            const bool c1 = nw < 8 && t1 <= v;
            g = 2 * nw + (c1 ? 1u : 0u);
            base = c1 ? t1 : bw;
            g = g > 15 ? 15 : g;
        }
        uint32_t lo = 0, b2 = base;
        if (big == 0) {
            const saddr_t a = w + (saddr_t)((8 + 4 * g) * stride);
            const uint32_t w0 = lds32(a), w1 = lds32(a + (saddr_t)stride), w2 = lds32(a + (saddr_t)(2 * stride)), w3 = lds32(a + (saddr_t)(3 * stride));
            const uint32_t s0 = dp4a_u(w0, 0x01010101u, base), s1 = dp4a_u(w1, 0x01010101u, s0), s2 = dp4a_u(w2, 0x01010101u, s1),
                           s3 = dp4a_u(w3, 0x01010101u, s2);
            const bool c0 = s0 <= v, c1 = s1 <= v, c2 = s2 <= v, c3 = s3 <= v;
            const uint32_t kw = (c0 ? 1u : 0u) + (c1 ? 1u : 0u) + (c2 ? 1u : 0u) + (c3 ? 1u : 0u);  // words of four counters wholly <= v
            uint32_t pre = c2 ? s2 : (c1 ? s1 : (c0 ? s0 : base));
            pre = c3 ? s3 : pre;
            const uint32_t wv = c2 ? w3 : (c1 ? w2 : (c0 ? w1 : w0));  // kw == 4 (cannot happen for v < total): w3 again, nothing counted
            lo = 4 * kw;
            b2 = pre;
            if (kw < 4) {
#pragma unroll
                for (uint32_t bb = 0; bb < 4; ++bb) {
                    const uint32_t t = pre + ((wv >> (8 * bb)) & 0xFFu);
                    const bool c = t <= v;
                    lo += c ? 1u : 0u;
                    b2 = c ? t : b2;
                    pre = t;
                }
            }
        } else {
            uint32_t pre = base;
            for (uint32_t j = 0; j < 16; ++j) {
                const uint32_t t = pre + count(g * 16 + j);
                const bool c = t <= v;
                lo += c ? 1u : 0u;
                b2 = c ? t : b2;
                pre = t;
            }
        }
        lo = lo > 15 ? 15 : lo;
        const uint32_t idx = g * 16 + lo;
        cum = b2;
        f = count(idx);
        return idx;
    }
    SCL_HD void load(const uint32_t *init_freq, uint32_t n_sym, uint64_t &total) {  // every count <= 255 (host-checked)
        big = 0;
        ovf = 0;
        total = 0;
        for (uint32_t g = 0; g < 16; g += 2) {
            uint32_t gs[2] = {0, 0};
            for (uint32_t h = 0; h < 2; ++h)
                for (uint32_t j = 0; j < 16; j += 4) {
                    uint32_t wv = 0;
                    for (uint32_t b = 0; b < 4; ++b) {
                        const uint32_t i = (g + h) * 16 + j + b;
                        const uint32_t fr = i < n_sym ? init_freq[i] : 0u;
                        wv |= (fr & 0xFFu) << (8 * b);
                        gs[h] += fr;
                    }
                    set_word(8 + (((g + h) * 16 + j) >> 2), wv);
                }
            set_word(g >> 1, gs[0] | (gs[1] << 16));
            total += gs[0] + gs[1];
        }
    }
    // AdaptiveIIDFreqModel's halving (probability_models.py:90-92): f = max(f // 2, 1) for every symbol
    SCL_HD void halve(uint32_t n_sym, uint64_t &total) {
        uint64_t nbig = 0;
        uint32_t nk = 0;
        total = 0;
        for (uint32_t g = 0; g < 16; ++g) {
            uint32_t gsum = 0;
            for (uint32_t j = 0; j < 16; ++j) {
                const uint32_t i = g * 16 + j;
                if (i < n_sym) {
                    uint32_t h = count(i) >> 1;  // reads the OLD list and this symbol's own, still unmodified, byte
                    h = h > 1 ? h : 1;
                    const uint32_t cw = 8 + (i >> 2), sh = 8 * (i & 3), wv = word(cw);
                    set_word(cw, (wv & ~(0xFFu << sh)) | ((h & 0xFFu) << sh));
                    if (h >> 8) {
                        nbig |= (uint64_t)(i | ((h >> 8) << 8)) << (12 * nk);
                        ++nk;  // halving never creates more big symbols than there were
                    }
                    gsum += h;
                }
            }
            const uint32_t gw = g >> 1, wv = word(gw);
            set_word(gw, (g & 1) ? ((wv & 0xFFFFu) | (gsum << 16)) : ((wv & 0xFFFF0000u) | gsum));
            total += gsum;
        }
        big = nbig;
    }
};

struct AecIid8Policy {
    AecModel8 M;
    uint32_t tot, adaptive, max_total, n_sym;
    SCL_HD uint32_t total() const { return tot; }
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const { M.query(idx, cum, f); }
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        uint32_t idx = M.find(v, cum, f);
        if (idx >= n_sym) {
            idx = n_sym - 1;
            M.query(idx, cum, f);
        }
        return idx;
    }
    SCL_HD uint32_t update(uint32_t idx) {
        if (adaptive) {
            M.add1(idx);
            tot += 1;
            if (tot >= max_total) {  // :90-92
                uint64_t t64 = tot;
                M.halve(n_sym, t64);
                tot = (uint32_t)t64;
            }
        }
        return M.ovf ? SCL_ST_OVERFLOW : SCL_ST_OK;
    }
};

// AdaptiveOrderKFreqModel (probability_models.py:95-168).  Per lane: n_ctx = n_sym^k rows of n_sym
// 32-bit counters (freqs_kplus1_tuple, row-major) followed by the n_ctx row totals; word i is
// `stride` bytes after word i-1 like AecModel.  The row index is the base-n_sym number of the past k
// alphabet indices (past_k, oldest digit first), which is how numpy indexes freqs_kplus1_tuple.
// Rows are scanned linearly: meant for the small alphabets context models are used with.
struct AecCtxPolicy {
    saddr_t w;
    uint32_t stride, n_sym, n_ctx, ctx, max_total;
    SCL_HD uint32_t word(uint32_t i) const { return lds32(w + (saddr_t)(i * stride)); }
    SCL_HD void set_word(uint32_t i, uint32_t v) const { sts32(w + (saddr_t)(i * stride), v); }
    SCL_HD uint32_t n_words() const { return n_ctx * (n_sym + 1); }
    SCL_HD uint32_t total() const { return word(n_ctx * n_sym + ctx); }
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const {
        const uint32_t row = ctx * n_sym;
        cum = 0;
        for (uint32_t j = 0; j < idx; ++j) cum += word(row + j);
        f = word(row + idx);
    }
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        const uint32_t row = ctx * n_sym;
        uint32_t idx = 0, below = 0;
        f = word(row);
        while (idx + 1 < n_sym && below + f <= v) {  // inclusive prefix <= v: the symbol lies further right
            below += f;
            idx += 1;
            f = word(row + idx);
        }
        cum = below;
        return idx;
    }
    // :137-168: count of (past_k, s) += 1, slide past_k; a count reaching max_allowed_total_freq makes the
    // reference raise (its np.max(count // 2, 1) takes 1 as an axis) -> SCL_ST_TOTAL_FREQ
    SCL_HD uint32_t update(uint32_t idx) {
        const uint32_t at = ctx * n_sym + idx, cnt = word(at) + 1, tw = n_ctx * n_sym + ctx;
        set_word(at, cnt);
        set_word(tw, word(tw) + 1);
        ctx = (ctx * n_sym + idx) % n_ctx;
        return cnt >= max_total ? SCL_ST_TOTAL_FREQ : SCL_ST_OK;
    }
    // model table: [n_ctx * n_sym counts, row-major][context index]; NULL = fresh (all ones, context 0)
    SCL_HD void load(const uint64_t *model) {
        for (uint32_t r = 0; r < n_ctx; ++r) {
            uint32_t sum = 0;
            for (uint32_t j = 0; j < n_sym; ++j) {
                const uint32_t v = model ? (uint32_t)model[r * n_sym + j] : 1u;
                set_word(r * n_sym + j, v);
                sum += v;
            }
            set_word(n_ctx * n_sym + r, sum);
        }
        ctx = model ? (uint32_t)(model[n_ctx * n_sym] % n_ctx) : 0u;
    }
    SCL_HD void store(uint64_t *model) const {
        for (uint32_t i = 0; i < n_ctx * n_sym; ++i) model[i] = word(i);
        model[n_ctx * n_sym] = ctx;
    }
};

// The same model when the table does not fit shared memory (a byte alphabet at k = 1 is 256 rows of 256 counters):
// the lane works IN PLACE on its block's model table in HBM -- the uint64 [n_ctx * n_sym counts][context] table of
// the C-ABI (scl_coder_model_words), which therefore must be supplied -- and only the n_ctx row totals live in
// shared memory (`tot`, one word per row, [row][lane] interleave like every per-lane structure here).  Cumulative
// counts are linear scans of the current row, two counters per 16-byte load (an L2-resident row at the batch
// sizes this is meant for).  Counts are taken modulo 2^32 (a count >= max_allowed_total_freq <= 2^30 has already
// made the coder stop, as the reference does).
struct AecCtxGlobalPolicy {
    uint64_t *tab;  // this block's table
    saddr_t tot;    // row totals: row r at tot + r * tstride
    uint32_t tstride, n_sym, n_ctx, ctx, max_total;
    SCL_HD uint32_t total() const { return lds32(tot + (saddr_t)(ctx * tstride)); }
    SCL_HD void query(uint32_t idx, uint32_t &cum, uint32_t &f) const {
        const uint64_t *row = tab + (uint64_t)ctx * n_sym;
        uint32_t c = 0, j = 0;
        if (((((uintptr_t)row) & 15) == 0)) {
            for (; j + 2 <= idx; j += 2) {
                const ulonglong2 v = *(const ulonglong2 *)(row + j);
                c += (uint32_t)v.x + (uint32_t)v.y;
            }
        }
        for (; j < idx; ++j) c += (uint32_t)row[j];
        cum = c;
        f = (uint32_t)row[idx];
    }
    SCL_HD uint32_t find(uint32_t v, uint32_t &cum, uint32_t &f) const {
        const uint64_t *row = tab + (uint64_t)ctx * n_sym;
        uint32_t idx = 0, below = 0;
        f = (uint32_t)row[0];
        while (idx + 1 < n_sym && below + f <= v) {  // inclusive prefix <= v: the symbol lies further right
            below += f;
            idx += 1;
            f = (uint32_t)row[idx];
        }
        cum = below;
        return idx;
    }
    SCL_HD uint32_t update(uint32_t idx) {
        uint64_t *at = tab + (uint64_t)ctx * n_sym + idx;
        const uint64_t cnt = *at + 1;
        *at = cnt;
        const saddr_t tw = tot + (saddr_t)(ctx * tstride);
        sts32(tw, lds32(tw) + 1);
        ctx = (uint32_t)(((uint64_t)ctx * n_sym + idx) % n_ctx);
        return cnt >= max_total ? SCL_ST_TOTAL_FREQ : SCL_ST_OK;
    }
    SCL_HD void load() {  // row totals from the table, context from its last word
        for (uint32_t r = 0; r < n_ctx; ++r) {
            const uint64_t *row = tab + (uint64_t)r * n_sym;
            uint32_t sum = 0;
            for (uint32_t j = 0; j < n_sym; ++j) sum += (uint32_t)row[j];
            sts32(tot + (saddr_t)(r * tstride), sum);
        }
        ctx = (uint32_t)(tab[(uint64_t)n_ctx * n_sym] % n_ctx);
    }
    SCL_HD void store() const { tab[(uint64_t)n_ctx * n_sym] = ctx; }
};

// ---- closed-form renormalisation ------------------------------------------------------------
// State: low in [0, 2^P), high in (low, 2^P].  Let hm = high - 1.
// E1/E2 loop (:126-143): every iteration drops a COMMON leading bit of low and hm.  The reference's
// strict tests stop one iteration early in two boundary cases: high == HALF (hm == 0111..1, reached
// at iteration j_A = P-1-cto(hm)) and low == HALF (1000..0, at j_B = P-1-ctz(low)).  So
//   n = min(leading common bits of (low, hm), j_A, j_B).
SCL_HD uint32_t aec_e12_count(uint64_t low, uint64_t high, uint32_t P) {
    const uint32_t lo = (uint32_t)low, hm = (uint32_t)(high - 1);
    const uint32_t sh = 32 - P;
    uint32_t x = (lo ^ hm) << sh;  // P-bit view aligned to bit 31
    uint32_t n = x ? clz32(x) : P;
    // branch-free: hm all ones (P of them) gives ctz32(~hm) >= P, lo == 0 gives ctz32 = 32 -- both candidates then wrap to
    // values far above P and lose the minimum, exactly as if they had been skipped
    const uint32_t jA = P - 1 - ctz32(~hm);  // ctz(~hm) = number of trailing ones
    const uint32_t jB = P - 1 - ctz32(lo);
    n = jA < n ? jA : n;
    n = jB < n ? jB : n;
    return n;
}
// E3 loop (:146-150): every iteration deletes bit P-2 of low and hm (keeping the top bits).  It runs
// while low > QTR and high < 3*QTR, i.e. (top bit of low set, or low = 01.. and not exactly QTR) and
// (top bit of hm clear, or hm = 10.. and not exactly 3*QTR-1): runs of ones / zeros below the top bit,
// shortened by one when the run ends at the lowest set bit of low / lowest clear bit of hm.
SCL_HD uint32_t aec_e3_count(uint64_t low, uint64_t high, uint32_t P) {
    const uint32_t lo = (uint32_t)low, hm = (uint32_t)(high - 1);
    const uint32_t top = 1u << (P - 1);
    const uint32_t sh = 33 - P;  // drops the top bit, aligns bit P-2 to bit 31 (P >= 2; sh == 32 handled for P == 1 never)
    // both candidates are always worked out (straight-line code), the conditions only select
    const uint32_t body1 = sh >= 32 ? 0u : (lo << sh);
    uint32_t r1 = clz32(~body1);  // leading ones
    r1 = r1 > P - 1 ? P - 1 : r1;
    r1 -= (r1 > 0 && ctz32(lo) == P - 1 - r1) ? 1u : 0u;
    const uint32_t body0 = sh >= 32 ? 0u : (hm << sh);
    uint32_t r0 = clz32(body0);  // leading zeros
    r0 = r0 > P - 1 ? P - 1 : r0;
    r0 -= (r0 > 0 && ctz32(~hm) == P - 1 - r0) ? 1u : 0u;
    const uint32_t m1 = (lo & top) ? 0xFFFFFFFFu : r1, m0 = (hm & top) ? r0 : 0xFFFFFFFFu;
    const uint32_t m = m0 < m1 ? m0 : m1;
    return m == 0xFFFFFFFFu ? 0u : m;
}
// n E1/E2 steps: v -> 2^n v - prefix * 2^P  (prefix = top n bits of low); exact in 64-bit integers
SCL_HD void aec_apply_e12(uint64_t &low, uint64_t &high, uint32_t n, uint32_t P, uint32_t &prefix) {
    prefix = n ? (uint32_t)(low >> (P - n)) : 0u;
    const uint64_t sub = (uint64_t)prefix << P;
    low = (low << n) - sub;
    high = (high << n) - sub;
}
// m E3 steps: v -> 2^m (v - HALF) + HALF
SCL_HD uint64_t aec_apply_e3(uint64_t v, uint32_t m, uint32_t P) {
    const uint64_t half = 1ull << (P - 1);
    return ((v - half) << m) + half;  // unsigned wrap-around == the signed identity (results are in [0, 2^P])
}

// ---- 32-bit state form ------------------------------------------------------------------------
// The lanes keep low and hm = high - 1 as 32-bit words (high itself can be 2^32).  With P-bit
// precision every update is taken modulo 2^P (`pm` = 2^P - 1): an E1/E2 step drops the common top
// bit (low shifts in 0, hm shifts in 1), an E3 step maps v -> 2(v - QTR), i.e. after m steps
// v -> 2^m (v - HALF) + HALF, and for hm the "+1" of high carries through as + (2^m - 1).
SCL_HD void aec_shift_e12(uint32_t &low, uint32_t &hm, uint32_t n, uint32_t pm) {
    low = (low << n) & pm;
    hm = ((hm << n) | ((1u << n) - 1u)) & pm;
}
SCL_HD void aec_shift_e3(uint32_t &low, uint32_t &hm, uint32_t m, uint32_t half, uint32_t pm) {
    low = (((low - half) << m) + half) & pm;
    hm = (((hm - half) << m) + half + ((1u << m) - 1u)) & pm;
}
// floor(a / t) for integer-valued doubles a < 2^53, t >= 1, rcp = 1 / t: the product estimate is
// within 1 of the quotient and the fused remainder a - q t is exact, so one correction suffices.
// 1 / x for the quotient ESTIMATE of aec_floor_div (x in [1, 2^32]): the hardware's reciprocal seed (>= 20 bits) and one
// Newton step give > 40 bits, the estimate of a quotient below 2^32 is then off by far less than one, and
// aec_floor_div's exact remainder test settles the rest -- a full IEEE division (seed, two steps, range checks and a
// slow-path call) is not needed.  On the host: the plain division (the corrected quotient is the same).
SCL_HD double aec_rcp(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
#else
    return 1.0 / x;
#endif
}
SCL_HD double aec_floor_div(double a, double t, double rcp) {
    double q = floor(a * rcp);
    double r = fma(-q, t, a);
    if (r < 0.0)
        q -= 1.0;
    else if (r >= t)
        q += 1.0;
    return q;
}
// shrink_range (:58-78): low += rng * c // T, high = low + rng * (c + f) // T.  One FP64 multiply per
// quotient while the products stay below 2^53 (rng <= 2^32, T < 2^20: always for the 16-bit IID
// structure); beyond that (order-k rows that have seen > 2^20 symbols) 64-bit integer division.
SCL_HD void aec_shrink(uint32_t &low, uint32_t &hm, uint32_t cc, uint32_t f, uint32_t total) {
    if (total >> 20) {
        const uint64_t lo64 = low, rng = (uint64_t)hm - lo64 + 1;  // rng * (cc + f) < 2^32 * 2^30
        hm = (uint32_t)(lo64 + rng * (uint64_t)(cc + f) / total - 1);
        low = (uint32_t)(lo64 + rng * (uint64_t)cc / total);
    } else {
        const double low_d = (double)low, t_d = (double)total, rcp_t = aec_rcp(t_d);
        const double rng_d = (double)hm - low_d + 1.0;
        hm = (uint32_t)(low_d + aec_floor_div(rng_d * (double)(cc + f), t_d, rcp_t) - 1.0);
        low = (uint32_t)(low_d + aec_floor_div(rng_d * (double)cc, t_d, rcp_t));
    }
}
// decode_step_core's target (:177-201): ((state - low + 1) * T - 1) // rng
SCL_HD uint32_t aec_target(uint32_t state, uint32_t low, uint32_t hm, uint32_t total) {
    if (total >> 20) {
        const uint64_t rng = (uint64_t)hm - low + 1;
        return (uint32_t)((((uint64_t)state - low + 1) * total - 1) / rng);
    }
    const double low_d = (double)low, rng_d = (double)hm - low_d + 1.0;
    return (uint32_t)aec_floor_div(((double)state - low_d + 1.0) * (double)total - 1.0, rng_d, aec_rcp(rng_d));
}

// ArithmeticEncoder.encode_block (arithmetic_coding.py:80-161)
// PFIX = 0: PRECISION from c.P; PFIX = 32: the reference's default as a compile-time constant (masks and shift amounts fold)
template <class Policy, uint32_t PFIX = 0>
SCL_HD uint32_t aec2_encode_lane(Policy &M, const AecTab &tab, const AecConst &c, const uint8_t *sym, uint64_t sym_cap, uint32_t n,
                                 FwdBitWriter &w, uint64_t &bits_out) {
    SymWindow sw;
    sw.init(sym, sym_cap);
    const uint32_t P = PFIX ? PFIX : c.P;
    const uint32_t pm = P == 32 ? 0xFFFFFFFFu : ((1u << P) - 1u), HALF = 1u << (P - 1), QTR = 1u << (P - 2);
    uint32_t low = 0, hm = pm, num_mid = 0;  // high = FULL
    uint32_t st = SCL_ST_OK;
    if (c.DBSB < 32 && (n >> c.DBSB)) st = SCL_ST_OVERFLOW;
    w.put64((uint64_t)n, c.DBSB);
    for (uint32_t i = 0; i < n && st == SCL_ST_OK; ++i) {
        const uint32_t idx = tab.sym2idx[sw.next(i)];
        if (idx == 0xFFFFu) {
            st = SCL_ST_BAD_SYMBOL;
            break;
        }
        const uint32_t total = M.total();
        if (!(total < QTR)) {  // :110-112
            st = SCL_ST_TOTAL_FREQ;
            break;
        }
        uint32_t cc, f;
        M.query(idx, cc, f);
        aec_shrink(low, hm, cc, f, total);
        st = M.update(idx);  // update_model (:118)
        if (st != SCL_ST_OK) break;
        // Renormalisation without branches on the counts: in a warp some lane almost always has bits to release, so a
        // branch would run both sides anyway.  With ne == 0 (me == 0) every step below is an identity and zero bits go out.
        const uint32_t ne = aec_e12_count(low, (uint64_t)hm + 1, P);  // <= P - 1
        {
            const uint32_t prefix = (low >> 1) >> (P - 1 - ne);  // low >> (P - ne); 0 for ne == 0
            aec_shift_e12(low, hm, ne, pm);
            // released bits: first prefix bit, then num_mid copies of its complement, then the rest
            const uint32_t nem1 = (ne - 1) & 31u;
            const uint32_t b0 = ne ? (prefix >> nem1) & 1u : 0u;
            const uint32_t tot = ne ? num_mid + ne : 0u;
            if (SCL_UNLIKELY(tot > 32)) {
                w.put(b0, 1);
                w.put_run(b0 ^ 1u, num_mid);
                if (ne > 1) w.put(prefix & mask32(ne - 1), ne - 1);
            } else {
                const uint32_t run = b0 ? 0u : mask32(ne ? num_mid : 0u);
                w.put(ne ? ((b0 << ((tot - 1) & 31u)) | (run << nem1) | (prefix & mask32(nem1))) : 0u, tot);
            }
            num_mid = ne ? 0u : num_mid;
        }
        const uint32_t me = aec_e3_count(low, (uint64_t)hm + 1, P);
        num_mid += me;
        aec_shift_e3(low, hm, me, HALF, pm);
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
    if (w.ovf) st = SCL_ST_OVERFLOW;
    return st;
}

// next k (<= 32) bits of the arithmetic stream with zero fill past its end (`A` bits, :258-261)
SCL_HD uint32_t aec_get_bits(BitReader &r, uint64_t &nbc, uint64_t A, uint32_t k) {
    uint32_t v = 0;
    if (k) {
        if (nbc + k <= A) {
            v = r.get(k);
        } else {
            uint32_t have = nbc < A ? (uint32_t)(A - nbc) : 0u;  // < k
            v = have ? (r.get(have) << (k - have)) : 0u;
        }
    }
    nbc += k;
    return v;
}

// ArithmeticDecoder.decode_block (arithmetic_coding.py:203-287)
template <class Policy, uint32_t PFIX = 0>
SCL_HD uint32_t aec2_decode_lane(Policy &M, const AecTab &tab, const AecConst &c, BitReader &r, uint64_t avail_bits, uint8_t *out,
                                 uint64_t out_cap, uint32_t &size_out, uint64_t &bits_consumed) {
    const uint32_t P = PFIX ? PFIX : c.P;
    const uint32_t pm = P == 32 ? 0xFFFFFFFFu : ((1u << P) - 1u), HALF = 1u << (P - 1), QTR = 1u << (P - 2);
    uint64_t size64 = r.get64(c.DBSB);
    size_out = 0;
    if (size64 > out_cap) return SCL_ST_OVERFLOW;
    if (size64 == 0) return SCL_ST_EMPTY_BLOCK;
    const uint32_t size = (uint32_t)size64;
    const uint64_t A = avail_bits > c.DBSB ? avail_bits - c.DBSB : 0;
    uint64_t nbc = 0;
    uint32_t low = 0, hm = pm;
    uint32_t state = aec_get_bits(r, nbc, A, P);  // :222-229 (MSB first, zero fill)
    uint32_t st = SCL_ST_OK;
    OutWindow ow;
    ow.init(out);
    for (uint32_t i = 0;;) {
        const uint32_t total = M.total();
        if (!(total < QTR)) {
            st = SCL_ST_TOTAL_FREQ;
            break;
        }
        uint32_t idx, cc, f;
        if (state < low) {
            idx = c.n_sym - 1;  // searchsorted -> 0, alphabet[-1]
            M.query(idx, cc, f);
        } else {
            // decode_step_core (:177-201): last idx with cum <= ((state - low + 1) * T - 1) // rng
            uint32_t v = aec_target(state, low, hm, total);
            if (v >= total) v = total - 1;
            idx = M.find(v, cc, f);
        }
        aec_shrink(low, hm, cc, f, total);
        ow.push(i, tab.idx2sym[idx]);
        ++i;
        st = M.update(idx);
        if (st != SCL_ST_OK) break;
        if (i == size) break;  // :242-243
        // (no branches on the counts: with ne == 0 / me == 0 the steps are identities and no bits are read -- see the encoder)
        const uint32_t ne = aec_e12_count(low, (uint64_t)hm + 1, P);
        aec_shift_e12(low, hm, ne, pm);
        state = ((state << ne) & pm) | aec_get_bits(r, nbc, A, ne);  // the dropped top bits are the common prefix (:252-256)
        const uint32_t me = aec_e3_count(low, (uint64_t)hm + 1, P);
        aec_shift_e3(low, hm, me, HALF, pm);
        state = ((((state - HALF) << me) + HALF) & pm) + aec_get_bits(r, nbc, A, me);
    }
    ow.flush(st == SCL_ST_OK ? size : 0);
    const uint64_t low64 = low, high64 = (uint64_t)hm + 1, state64 = state;
    uint32_t extra = 0;  // :277-282
    for (extra = 0; extra < P; ++extra) {
        uint64_t state_low = (state64 >> extra) << extra;
        uint64_t state_high = state_low + (1ull << extra);
        if (state_low < low64 || state_high > high64) break;
    }
    if (extra == P) extra = P - 1;
    size_out = size;
    bits_consumed = (uint64_t)((int64_t)nbc - ((int64_t)extra - 1) + (int64_t)c.DBSB);
    return st;
}

}  // namespace scl
