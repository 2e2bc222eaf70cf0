// scl_tables.hpp -- HOST-side derivation of coder constants and lookup tables from the
// reference's parameter classes.  Pure C++ (no CUDA); used by the C-ABI (scl_capi.cu) and by
// the CPU lane-emulation harness in tests/host_emu.
//
// Follows rANSParams.__post_init__ (scl/compressors/rANS.py:97-120), tANSEncoder's
// shrink tables (tANS.py:74-110), RangeEncoder.__init__ (range_coder.py:80-86) and
// AECParams.__post_init__ (arithmetic_coding.py:33-38).
#pragma once
#include <cstring>
#include <vector>

#include "scl_defs.h"

namespace scl {

typedef unsigned __int128 u128_t;

static inline bool is_pow2_u64(uint64_t x) { return x && !(x & (x - 1)); }
static inline uint32_t log2_u64(uint64_t x) {
    uint32_t l = 0;
    while ((x >> l) > 1) ++l;
    return l;
}
static inline uint32_t bit_length_u64(uint64_t x) {  // Python int.bit_length(); 0 -> 0
    uint32_t w = 0;
    while (x) {
        ++w;
        x >>= 1;
    }
    return w;
}

struct Alphabet {
    uint32_t n_sym = 0;
    uint8_t idx2sym[256];
    uint16_t sym2idx[256];
    bool covers_all_bytes = false;
    int init(const uint8_t *alphabet, uint32_t n) {
        if (n == 0 || n > 256) return SCL_E_UNSUPPORTED;
        n_sym = n;
        for (int i = 0; i < 256; ++i) sym2idx[i] = 0xFFFF;
        memset(idx2sym, 0, sizeof(idx2sym));
        for (uint32_t i = 0; i < n; ++i) {
            uint8_t v = alphabet ? alphabet[i] : (uint8_t)i;
            if (sym2idx[v] != 0xFFFF) return SCL_E_INVALID;  // duplicate key cannot occur in a dict
            sym2idx[v] = (uint16_t)i;
            idx2sym[i] = v;
        }
        covers_all_bytes = (n == 256);
        return SCL_E_OK;
    }
};

// ---------------------------------------------------------------------------------------------
// rANS
// ---------------------------------------------------------------------------------------------
struct RansHost {
    RansConst c;
    Alphabet a;
    std::vector<uint64_t> freq, cum;
    bool enc32 = false;  // 32-bit encode fast path usable
    bool dec32 = false;  // 32-bit LUT decode fast path usable
    std::vector<RansEnc32> enc_tab;  // 256 entries (by byte value)
    std::vector<RansDec32> dec_lut;  // M entries
    RansGeneric gen;
    uint32_t max_bits_per_symbol = 0;  // worst-case emitted bits for one symbol

    int init(const scl_params &p, const uint8_t *alphabet, const uint64_t *f, uint32_t n_sym) {
        int rc = a.init(alphabet, n_sym);
        if (rc) return rc;
        if (p.num_bits_out == 0 || p.num_bits_out > 32 || p.range_factor == 0) return SCL_E_INVALID;
        if (p.data_block_size_bits > 64 || p.num_state_bits == 0 || p.num_state_bits > 64) return SCL_E_INVALID;
        freq.assign(f, f + n_sym);
        cum.assign(n_sym + 1, 0);
        u128_t tot = 0;
        for (uint32_t i = 0; i < n_sym; ++i) {
            if (f[i] == 0) return SCL_E_INVALID;  // the reference would divide by zero (rANS.py:143)
            cum[i] = (uint64_t)tot;
            tot += f[i];
            if (tot >> 62) return SCL_E_UNSUPPORTED;
        }
        cum[n_sym] = (uint64_t)tot;
        memset(&c, 0, sizeof(c));
        c.M = (uint64_t)tot;
        c.RF = p.range_factor;
        u128_t L = (u128_t)c.RF * c.M;
        u128_t H = L * ((u128_t)1 << p.num_bits_out) - 1;
        if (H >> 63) return SCL_E_UNSUPPORTED;  // state must fit 63 bits on this backend
        c.L = (uint64_t)L;
        c.H = (uint64_t)H;
        c.NBO = p.num_bits_out;
        c.NSB = p.num_state_bits;
        c.DBSB = p.data_block_size_bits;
        c.n_sym = n_sym;
        c.m_log2 = is_pow2_u64(c.M) ? log2_u64(c.M) : 0xFFFFFFFFu;
        c.l_log2 = is_pow2_u64(c.L) ? log2_u64(c.L) : 0xFFFFFFFFu;
        c.check_sym = a.covers_all_bytes ? 0 : 1;
        // uint_to_bitarray(state, NUM_STATE_BITS) must be able to hold H, else the reference raises
        if (c.NSB < 64 && (c.H >> c.NSB)) return SCL_E_INVALID;

        // generic tables
        memset(&gen, 0, sizeof(gen));
        for (uint32_t i = 0; i < n_sym; ++i) {
            gen.freq[i] = freq[i];
            gen.cum[i] = cum[i];
            u128_t ms = (u128_t)c.RF * freq[i] * ((u128_t)1 << c.NBO) - 1;  // <= H
            gen.max_shrunk[i] = (uint64_t)ms;
        }
        gen.cum[n_sym] = c.M;
        for (uint32_t i = n_sym + 1; i < 257; ++i) gen.cum[i] = c.M;
        memcpy(gen.sym2idx, a.sym2idx, sizeof(gen.sym2idx));
        memcpy(gen.idx2sym, a.idx2sym, sizeof(gen.idx2sym));

        // worst-case bits per symbol: chunks c0+1 of NBO bits where c0 = smallest c with (L >> c*NBO) <= max_shrunk
        max_bits_per_symbol = 0;
        for (uint32_t i = 0; i < n_sym; ++i) {
            uint32_t c0 = 0;
            while ((c.L >> (c0 * c.NBO)) > gen.max_shrunk[i]) ++c0;
            uint32_t b = (c0 + 1) * c.NBO;
            if (b > max_bits_per_symbol) max_bits_per_symbol = b;
        }
        build_enc32();
        build_dec32();
        return SCL_E_OK;
    }

    // 32-bit encode fast path: H < 2^32, M <= 65536, per-symbol reciprocal proven exact.
    void build_enc32() {
        enc32 = false;
        if (c.H >> 32) return;
        if (c.M > 65536) return;
        enc_tab.assign(256, RansEnc32{0, 0, 0, kRansEncInvalid});
        for (uint32_t i = 0; i < c.n_sym; ++i) {
            uint64_t f = freq[i];
            uint64_t max_shrunk = gen.max_shrunk[i];  // < 2^32
            uint32_t c0 = 0;
            while ((c.L >> (c0 * c.NBO)) > max_shrunk) ++c0;
            uint32_t nb0 = c0 * c.NBO;
            if (nb0 + c.NBO > 31) return;
            u128_t thresh = ((u128_t)max_shrunk + 1) << nb0;
            RansEnc32 e;
            e.thresh_key = (thresh - 1 > 0xFFFFFFFFull) ? 0xFFFFFFFFu : (uint32_t)(thresh - 1);
            if (c.NBO == 1) e.thresh_key = ~e.thresh_key;
            uint32_t shift;
            uint64_t bias = cum[i];
            uint64_t cmpl = c.M - f;
            if (f == 1) {
                // umulhi(x, 2^32-1) = x - 1 for 1 <= x < 2^32; fold the missing 1 into the bias:
                // x + bias + (x-1)(M-1) = x*M + cum  <=>  bias = cum + M - 1
                e.rcp = 0xFFFFFFFFu;
                shift = 0;
                bias = cum[i] + c.M - 1;
            } else if (is_pow2_u64(f)) {
                e.rcp = 0x80000000u;  // umulhi(x, 2^31) = x >> 1
                shift = log2_u64(f) - 1;
            } else {
                uint32_t l = log2_u64(f) + 1;  // ceil(log2 f) for non powers of two
                uint32_t k = 31 + l;
                u128_t two_k = (u128_t)1 << k;
                u128_t m = (two_k + f - 1) / f;  // ceil(2^k / f) < 2^32 because f > 2^(l-1)
                if (m >> 32) return;
                u128_t err = m * f - two_k;  // in [0, f)
                // floor(x*m / 2^k) == floor(x / f) for all x <= X when err * X < 2^k
                if (err * (u128_t)max_shrunk >= two_k) return;
                e.rcp = (uint32_t)m;
                shift = l - 1;
            }
            if (bias > 0xFFFFFFFFull || cmpl > 0xFFFF) return;
            e.bias = (uint32_t)bias;
            e.pack = ((uint32_t)cmpl << 16) | (nb0 << 8) | shift;
            enc_tab[a.idx2sym[i]] = e;
        }
        enc32 = true;
    }

    // 32-bit decode fast path: H < 2^32, M and L powers of two, M <= 4096, f <= 4095.
    void build_dec32() {
        dec32 = false;
        if (c.H >> 32) return;
        if (c.m_log2 == 0xFFFFFFFFu || c.l_log2 == 0xFFFFFFFFu) return;
        if (c.M > 4096) return;
        if (c.NSB != c.l_log2 + c.NBO) return;  // the clz renorm count assumes NSB == bit_length(H)
        if (c.NSB > 32) return;
        dec_lut.assign(c.M, 0);
        for (uint32_t i = 0; i < c.n_sym; ++i) {
            if (freq[i] > 4095) return;
            for (uint64_t s = cum[i]; s < cum[i + 1]; ++s)
                dec_lut[s] = ((uint32_t)freq[i] << 20) | ((uint32_t)(s - cum[i]) << 8) | a.idx2sym[i];
        }
        dec32 = true;
    }

    uint64_t max_encoded_bits(uint64_t n) const { return (uint64_t)c.DBSB + c.NSB + n * max_bits_per_symbol; }
};

// ---------------------------------------------------------------------------------------------
// tANS (tables themselves are built on the device by tans_build_kernel; here: per-symbol rows)
// ---------------------------------------------------------------------------------------------
struct TansHost {
    RansHost r;
    std::vector<TansSym> sym_tab;  // 256 by byte value
    std::vector<TansSym8> sym_tab8;  // the same rows in the second-generation encoder's 8-byte form
    std::vector<uint32_t> row_of_idx;  // enc_table row offset per alphabet index
    int init(const scl_params &p, const uint8_t *alphabet, const uint64_t *f, uint32_t n_sym) {
        int rc = r.init(p, alphabet, f, n_sym);
        if (rc) return rc;
        // tANSParams asserts (tANS.py:38-49)
        if (!is_pow2_u64(r.c.M) || r.c.NBO != 1) return SCL_E_INVALID;
        if (r.c.L > (1ull << 23)) return SCL_E_UNSUPPORTED;  // dec_packed keeps x_shrunk in 24 bits
        sym_tab.assign(256, TansSym{0, 0xFFFFFFFFu, 0, 0});
        sym_tab8.assign(256, TansSym8{0, 0xFFFFFFFFu});
        row_of_idx.assign(n_sym, 0);
        uint64_t row = 0;
        for (uint32_t i = 0; i < n_sym; ++i) {
            uint64_t mn = r.c.RF * r.freq[i];
            uint64_t mx = 2 * mn - 1;
            uint32_t y = bit_length_u64(mx);  // get_bit_width(max_shrunk_state) (tANS.py:81)
            if (y > r.c.NSB) return SCL_E_INVALID;
            TansSym e;
            e.nb0 = r.c.NSB - y;
            u128_t th = ((u128_t)mx + 1) << e.nb0;
            e.thresh = th > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)th;
            e.row = (int32_t)((int64_t)row - (int64_t)mn);
            e.pad = 0;
            sym_tab[r.a.idx2sym[i]] = e;
            sym_tab8[r.a.idx2sym[i]] = TansSym8{e.thresh, ((uint32_t)e.row << 7) | e.nb0};  // L <= 2^23: |row| < 2^23, nb0 <= 24
            row_of_idx[i] = (uint32_t)row;
            row += mn;  // RF*f entries per symbol; total = L
        }
        return SCL_E_OK;
    }
};

// ---------------------------------------------------------------------------------------------
// range coder
// ---------------------------------------------------------------------------------------------
struct RangeHost {
    RangeConst c;
    RangeTab t;
    int init(const scl_params &p, const uint8_t *alphabet, const uint64_t *f, uint32_t n_sym) {
        Alphabet a;
        int rc = a.init(alphabet, n_sym);
        if (rc) return rc;
        if (p.precision % 8 != 0) return SCL_E_INVALID;  // range_coder.py:66
        if (p.precision != 24 && p.precision != 32) return SCL_E_UNSUPPORTED;
        if (p.data_block_size_bits > 64) return SCL_E_INVALID;
        memset(&t, 0, sizeof(t));
        uint64_t tot = 0;
        for (uint32_t i = 0; i < n_sym; ++i) {
            if (f[i] == 0) return SCL_E_INVALID;  // range_coder.py:84
            t.cum[i] = (uint32_t)tot;
            t.freq[i] = (uint32_t)f[i];
            tot += f[i];
            if (tot > (1ull << (p.precision - 16))) return SCL_E_INVALID;  // range_coder.py:85
        }
        for (uint32_t i = n_sym; i < 257; ++i) t.cum[i] = (uint32_t)tot;
        memcpy(t.sym2idx, a.sym2idx, sizeof(t.sym2idx));
        memcpy(t.idx2sym, a.idx2sym, sizeof(t.idx2sym));
        c.P = p.precision;
        c.DBSB = p.data_block_size_bits;
        c.n_sym = n_sym;
        c.T = (uint32_t)tot;
        c.t_shift = is_pow2_u64(tot) ? log2_u64(tot) : 0xFFFFFFFFu;
        // v -> alphabet index for the decoder's symbol search (replaces an 8-step binary search)
        lut.assign((size_t)tot, 0);
        for (uint32_t i = 0; i < n_sym; ++i)
            for (uint64_t v = t.cum[i]; v < (uint64_t)t.cum[i] + t.freq[i]; ++v) lut[v] = (uint8_t)i;
        build_v2(n_sym);
        return SCL_E_OK;
    }
    std::vector<uint8_t> lut;
    // second-generation lanes (scl_range.cuh): PRECISION 32, 32-bit size header, power-of-two 16 <= T <= 4096, f <= 4095
    bool v2 = false;
    std::vector<uint32_t> enc_tab;  // by BYTE VALUE: freq << 16 | cum; 0xFFFFFFFF = byte not in the alphabet
    std::vector<uint32_t> dec_lut;  // by v in [0, T): freq << 20 | cum << 8 | byte value
    uint32_t last_entry = 0;        // dec_lut-format entry of the last alphabet index
    void build_v2(uint32_t n_sym) {
        v2 = false;
        if (c.P != 32 || c.DBSB != 32 || c.t_shift == 0xFFFFFFFFu || c.T < 16 || c.T > 4096) return;
        enc_tab.assign(256, 0xFFFFFFFFu);
        dec_lut.assign(c.T, 0);
        for (uint32_t i = 0; i < n_sym; ++i) {
            if (t.freq[i] > 4095) return;
            enc_tab[t.idx2sym[i]] = (t.freq[i] << 16) | t.cum[i];
            const uint32_t e = (t.freq[i] << 20) | (t.cum[i] << 8) | t.idx2sym[i];
            for (uint32_t v = t.cum[i]; v < t.cum[i] + t.freq[i]; ++v) dec_lut[v] = e;
            last_entry = e;
        }
        v2 = true;
    }
    // every symbol can release at most ceil(P/8) bytes... bound: normalise emits a byte only while
    // range < 2^(P-8) effectively; one symbol shrinks range by at most T <= 2^(P-16) => <= 3 bytes
    uint64_t max_encoded_bits(uint64_t n) const { return (uint64_t)c.DBSB + 8ull * (3 * n + c.P / 8); }
};

// ---------------------------------------------------------------------------------------------
// arithmetic coder
// ---------------------------------------------------------------------------------------------
struct AecHost {
    AecConst c;
    AecTab t;
    int init(const scl_params &p, const uint8_t *alphabet, const uint64_t *f, uint32_t n_sym) {
        Alphabet a;
        int rc = a.init(alphabet, n_sym);
        if (rc) return rc;
        if (p.precision < 4 || p.precision > 32) return SCL_E_UNSUPPORTED;
        if (p.data_block_size_bits > 64) return SCL_E_INVALID;
        if (p.model != SCL_MODEL_FIXED && p.model != SCL_MODEL_ADAPTIVE_IID && p.model != SCL_MODEL_ORDER_K) return SCL_E_UNSUPPORTED;
        if (p.model != SCL_MODEL_ORDER_K && p.model_order != 0) return SCL_E_INVALID;
        c.order_k = p.model_order;
        c.n_ctx = 1;
        c.ctx_global = 0;
        if (p.model == SCL_MODEL_ORDER_K) {
            // per-lane table of n_sym^k * (n_sym + 1) words in shared memory when it fits; else the lanes work on the
            // caller's table in HBM with the row totals in shared memory (n_sym^k <= kAecCtxGlobalMaxRows rows)
            for (uint32_t j = 0; j < p.model_order && n_sym > 1; ++j) {
                if ((uint64_t)c.n_ctx * n_sym > kAecCtxGlobalMaxRows) return SCL_E_UNSUPPORTED;
                c.n_ctx *= n_sym;
            }
            c.ctx_global = (uint64_t)c.n_ctx * (n_sym + 1) > kAecCtxMaxWords ? 1u : 0u;
        }
        memset(&t, 0, sizeof(t));
        for (uint32_t i = 0; i < n_sym; ++i) {
            if (f[i] == 0 || f[i] >> 31) return SCL_E_INVALID;
            t.init_freq[i] = (uint32_t)f[i];
        }
        memcpy(t.sym2idx, a.sym2idx, sizeof(t.sym2idx));
        memcpy(t.idx2sym, a.idx2sym, sizeof(t.idx2sym));
        c.P = p.precision;
        c.DBSB = p.data_block_size_bits;
        c.n_sym = n_sym;
        c.model = (uint32_t)p.model;
        c.max_total = p.max_allowed_total_freq;
        return SCL_E_OK;
    }
    // each symbol narrows the range by at most a factor T < 2^(P-2): <= P bits per symbol, plus
    // the size header and the <= P+1 termination bits
    uint64_t max_encoded_bits(uint64_t n) const { return (uint64_t)c.DBSB + (uint64_t)c.P * n + c.P + 2; }
    uint64_t model_words() const { return c.model == SCL_MODEL_ORDER_K ? (uint64_t)c.n_ctx * c.n_sym + 1 : c.n_sym; }
    // may blocks of up to `block_len` symbols use the 8-bit-counter model (scl_aec.cuh AecModel8)?  Every initial count must
    // fit a byte, and at most kAecBigMax symbols may ever reach 256 (a count only grows by one per coded symbol; the
    // halving rule only shrinks them), which also keeps every group total below 65 536.
    bool model8_ok(uint64_t block_len) const {
        if (c.model == SCL_MODEL_ORDER_K) return false;
        uint64_t sum = 0, mx = 0;
        for (uint32_t i = 0; i < c.n_sym; ++i) {
            sum += t.init_freq[i];
            mx = t.init_freq[i] > mx ? t.init_freq[i] : mx;
        }
        if (mx > 255) return false;
        if (c.model == SCL_MODEL_FIXED) return true;
        return (sum + block_len) / 256 <= kAecBigMax;
    }
};

}  // namespace scl
