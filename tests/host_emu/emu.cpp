// emu.cpp -- CPU lane-emulation harness (TEST INFRASTRUCTURE ONLY; never loaded by the product).
//
// Compiles the product's __host__ __device__ per-lane recurrences (csrc/scl_lane.cuh) and the
// host table builders (csrc/scl_tables.hpp) with plain g++ and runs them one "lane" at a time
// with the same I/O contract as scl_encode_blocks / scl_decode_blocks.  It lets the `not gpu`
// test tier check the integer logic the kernels execute against the oracle without a GPU; the
// `-m gpu` tier then checks the real kernels through the C-ABI.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../stanford_compression_library_b200/csrc/scl_aec.cuh"
#include "../../stanford_compression_library_b200/csrc/scl_fast.cuh"
#include "../../stanford_compression_library_b200/csrc/scl_lane.cuh"
#include "../../stanford_compression_library_b200/csrc/scl_range.cuh"
#include "../../stanford_compression_library_b200/csrc/scl_tables.hpp"

using namespace scl;

namespace {
struct HostTree {
    uint32_t f[257];
    uint32_t get(uint32_t i) const { return f[i]; }
    void set(uint32_t i, uint32_t v) { f[i] = v; }
};
struct Emu {
    scl_params p;
    RansHost *rans = nullptr;
    TansHost *tans = nullptr;
    RangeHost *range = nullptr;
    AecHost *aec = nullptr;
    std::vector<uint32_t> tenc, tdec;
    int aec2 = 0;  // 1 = second-generation arithmetic-coder lanes (16-bit counters), 2 = the same with 8-bit counters
    ~Emu() {
        delete rans;
        delete tans;
        delete range;
        delete aec;
    }
};
uint64_t load_model(HostTree &F, const AecTab &tab, const AecConst &c, const uint64_t *model) {
    uint64_t total = 0;
    F.set(0, 0);
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t f = 0;
        if (i < c.n_sym) f = model ? (uint32_t)model[i] : tab.init_freq[i];
        F.set(i + 1, f);
        total += f;
    }
    fen_build(F);
    return total;
}
void store_model(HostTree &F, const AecConst &c, uint64_t *model) {
    if (!model) return;
    fen_unbuild(F);
    for (uint32_t i = 0; i < c.n_sym; ++i) model[i] = F.get(i + 1);
}
struct MaskTable {
    alignas(16) uint8_t m[17][16];
    MaskTable() {
        for (int t = 0; t <= 16; ++t)
            for (int k = 0; k < 16; ++k) m[t][k] = k < t ? 1 : 0;
    }
};
const MaskTable g_mask_table;
const uint8_t *g_aec_masks = &g_mask_table.m[0][0];
}  // namespace

extern "C" {

void emu_set_aec2(void *h, int on) { ((Emu *)h)->aec2 = on; }
int emu_aec_model8_ok(void *h, uint64_t block_len) { return ((Emu *)h)->aec && ((Emu *)h)->aec->model8_ok(block_len) ? 1 : 0; }

// closed-form renormalisation counts, exposed for a direct check against the literal loops
void emu_aec_renorm_counts(uint32_t P, uint64_t low, uint64_t high, uint32_t *n_e12, uint32_t *m_e3, uint64_t *low_out, uint64_t *high_out) {
    uint32_t n = aec_e12_count(low, high, P), prefix;
    aec_apply_e12(low, high, n, P, prefix);
    uint32_t m = aec_e3_count(low, high, P);
    low = aec_apply_e3(low, m, P);
    high = aec_apply_e3(high, m, P);
    *n_e12 = n;
    *m_e3 = m;
    *low_out = low;
    *high_out = high;
}

int emu_create(const scl_params *params, const uint8_t *alphabet, const uint64_t *freq, uint32_t n_sym, void **out) {
    Emu *e = new Emu();
    e->p = *params;
    int rc;
    switch (params->coder) {
    case SCL_CODER_RANS:
        e->rans = new RansHost();
        rc = e->rans->init(*params, alphabet, freq, n_sym);
        break;
    case SCL_CODER_TANS: {
        e->tans = new TansHost();
        rc = e->tans->init(*params, alphabet, freq, n_sym);
        if (rc) break;
        uint64_t L = e->tans->r.c.L;
        e->tenc.assign(L, 0);
        e->tdec.assign(L, 0);
        for (uint64_t i = 0; i < L; ++i)
            tans_build_entry(e->tans->r.gen, e->tans->r.c, e->tans->row_of_idx.data(), e->tenc.data(), e->tdec.data(), i);
        break;
    }
    case SCL_CODER_RANGE:
        e->range = new RangeHost();
        rc = e->range->init(*params, alphabet, freq, n_sym);
        break;
    case SCL_CODER_AEC:
        e->aec = new AecHost();
        rc = e->aec->init(*params, alphabet, freq, n_sym);
        break;
    default:
        rc = SCL_E_INVALID;
    }
    if (rc) {
        delete e;
        return rc;
    }
    *out = e;
    return 0;
}
void emu_destroy(void *h) { delete (Emu *)h; }

int emu_path(void *h, int decode) {
    Emu *e = (Emu *)h;
    if (e->rans) return decode ? (e->rans->dec32 ? 0 : 1) : (e->rans->enc32 ? 0 : 1);
    return 0;
}
// force the generic 64-bit rANS path (to test it on parameter sets the fast path would take)
void emu_force_generic(void *h) {
    Emu *e = (Emu *)h;
    if (e->rans) e->rans->enc32 = e->rans->dec32 = false;
}
uint64_t emu_max_encoded_bytes(void *h, uint64_t block_len) {
    Emu *e = (Emu *)h;
    uint64_t bits = 0;
    if (e->rans) bits = e->rans->max_encoded_bits(block_len);
    if (e->tans) bits = e->tans->r.max_encoded_bits(block_len);
    if (e->range) bits = e->range->max_encoded_bits(block_len);
    if (e->aec) bits = e->aec->max_encoded_bits(block_len);
    uint64_t bytes = (bits + 7) / 8 + 4;
    return (bytes + 31) & ~31ull;
}
int emu_tans_tables(void *h, uint32_t *enc, uint32_t *dec, uint64_t n) {
    Emu *e = (Emu *)h;
    if (!e->tans || n != e->tans->r.c.L) return SCL_E_INVALID;
    memcpy(enc, e->tenc.data(), n * 4);
    memcpy(dec, e->tdec.data(), n * 4);
    return 0;
}

}  // extern "C"
static AecIidPolicy make_iid_policy(uint32_t *words, const AecTab &t, const AecConst &c, const uint64_t *mm) {
    AecIidPolicy pol;
    pol.M = AecModel{saddr_of(words), 4, saddr_of(g_aec_masks), 16};
    uint64_t total = 0;
    pol.M.load(t.init_freq, mm, c.n_sym, total);
    pol.tot = (uint32_t)total;
    pol.adaptive = c.model == SCL_MODEL_ADAPTIVE_IID;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    pol.n_sym = c.n_sym;
    return pol;
}
static AecIid8Policy make_iid8_policy(uint32_t *words, const AecTab &t, const AecConst &c) {
    AecIid8Policy pol;
    pol.M.w = saddr_of(words);
    pol.M.stride = 4;
    pol.M.masks = saddr_of(g_aec_masks);
    pol.M.mstride = 16;
    uint64_t total = 0;
    pol.M.load(t.init_freq, c.n_sym, total);
    pol.tot = (uint32_t)total;
    pol.adaptive = c.model == SCL_MODEL_ADAPTIVE_IID;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    pol.n_sym = c.n_sym;
    return pol;
}
static AecCtxPolicy make_ctx_policy(uint32_t *words, const AecConst &c) {
    AecCtxPolicy pol;
    pol.w = saddr_of(words);
    pol.stride = 4;
    pol.n_sym = c.n_sym;
    pol.n_ctx = c.n_ctx;
    pol.ctx = 0;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    return pol;
}
static AecCtxGlobalPolicy make_ctx_global_policy(uint32_t *totals, const AecConst &c, uint64_t *table) {
    AecCtxGlobalPolicy pol;
    pol.tab = table;
    pol.tot = saddr_of(totals);
    pol.tstride = 4;
    pol.n_sym = c.n_sym;
    pol.n_ctx = c.n_ctx;
    pol.ctx = 0;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    return pol;
}
extern "C" {
int emu_encode_blocks(void *h, const uint8_t *sym, uint64_t sym_stride, const uint32_t *sizes, uint32_t block_len, uint64_t n_blocks,
                      uint8_t *out, uint64_t out_stride, uint64_t *bit_off, uint64_t *bit_len, uint64_t *model, uint32_t *status) {
    Emu *e = (Emu *)h;
    if ((out_stride & 15) || (((uintptr_t)out) & 15)) return SCL_E_INVALID;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        uint32_t n = sizes ? sizes[b] : block_len;
        uint8_t *slot = out + b * out_stride;
        const uint8_t *row = sym + b * sym_stride;
        uint64_t bits = 0;
        uint32_t st;
        bool lifo = true;
        if (e->rans || e->tans) {
            LifoBitWriter w;
            w.init(slot, slot + out_stride);
            if (e->rans) {
                RansHost &r = *e->rans;
                if (r.enc32)
                    st = r.c.check_sym ? rans32_encode_lane<true>(r.enc_tab.data(), r.c, row, n, w, bits)
                                       : rans32_encode_lane<false>(r.enc_tab.data(), r.c, row, n, w, bits);
                else
                    st = rans64_encode_lane(r.gen, r.c, row, n, w, bits);
            } else {
                st = tans_encode_lane(e->tans->sym_tab.data(), e->tenc.data(), e->tans->r.c, row, n, w, bits);
            }
        } else {
            lifo = false;
            FwdBitWriter w;
            w.init(slot, slot + out_stride);
            if (e->range) {
                st = range_encode_lane(e->range->t, e->range->c, row, sym_stride, n, w, bits);
            } else if (e->aec->c.model == SCL_MODEL_ORDER_K && e->aec->c.ctx_global) {
                if (!model) return SCL_E_INVALID;
                std::vector<uint32_t> totals(e->aec->c.n_ctx);
                AecCtxGlobalPolicy pol = make_ctx_global_policy(totals.data(), e->aec->c, model + b * e->aec->model_words());
                pol.load();
                st = aec2_encode_lane(pol, e->aec->t, e->aec->c, row, sym_stride, n, w, bits);
                pol.store();
            } else if (e->aec->c.model == SCL_MODEL_ORDER_K) {
                alignas(16) uint32_t words[kAecCtxMaxWords];
                AecCtxPolicy pol = make_ctx_policy(words, e->aec->c);
                uint64_t *mm = model ? model + b * e->aec->model_words() : nullptr;
                pol.load(mm);
                st = aec2_encode_lane(pol, e->aec->t, e->aec->c, row, sym_stride, n, w, bits);
                if (mm) pol.store(mm);
            } else if (e->aec2 == 2 && !model && e->aec->model8_ok(n)) {
                alignas(16) uint32_t words[kAecModel8Words];
                AecIid8Policy pol = make_iid8_policy(words, e->aec->t, e->aec->c);
                st = aec2_encode_lane(pol, e->aec->t, e->aec->c, row, sym_stride, n, w, bits);
            } else if (e->aec2) {
                alignas(16) uint32_t words[kAecModelWords];
                uint64_t *mm = model ? model + b * e->aec->c.n_sym : nullptr;
                AecIidPolicy pol = make_iid_policy(words, e->aec->t, e->aec->c, mm);
                st = aec2_encode_lane(pol, e->aec->t, e->aec->c, row, sym_stride, n, w, bits);
                if (mm) pol.M.store(mm, e->aec->c.n_sym);
            } else {
                HostTree F;
                uint64_t *mm = model ? model + b * e->aec->c.n_sym : nullptr;
                uint64_t total = load_model(F, e->aec->t, e->aec->c, mm), tout = 0;
                st = aec_encode_lane(F, e->aec->t, e->aec->c, total, row, n, w, bits, tout);
                store_model(F, e->aec->c, mm);
            }
        }
        bit_len[b] = bits;
        bit_off[b] = lifo ? (b + 1) * out_stride * 8 - bits : b * out_stride * 8;
        status[b] = st;
    }
    return 0;
}

int emu_decode_blocks(void *h, const uint8_t *in, uint64_t in_bytes, const uint64_t *bit_off, const uint64_t *bit_len, uint64_t n_blocks,
                      uint8_t *sym, uint64_t sym_stride, uint32_t *sizes, uint64_t *consumed, uint64_t *model, uint32_t *status) {
    Emu *e = (Emu *)h;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        BitReader r;
        uint64_t off = bit_off[b];
        r.init(in, in_bytes, off);
        uint64_t avail = bit_len ? bit_len[b] : (in_bytes * 8 > off ? in_bytes * 8 - off : 0);
        uint8_t *row = sym + b * sym_stride;
        uint32_t size = 0, st;
        uint64_t used = 0;
        if (e->rans) {
            RansHost &rh = *e->rans;
            st = rh.dec32 ? rans32_decode_lane(rh.dec_lut.data(), rh.c, r, row, sym_stride, size, used)
                          : rans64_decode_lane(rh.gen, rh.c, r, avail, row, sym_stride, size, used);
            if (st == SCL_ST_OK && used > avail) st = SCL_ST_TRUNCATED;
        } else if (e->tans) {
            st = tans_decode_lane(e->tdec.data(), e->tans->r.c, r, row, sym_stride, size, used);
            if (st == SCL_ST_OK && used > avail) st = SCL_ST_TRUNCATED;
        } else if (e->range) {
            st = range_decode_lane(e->range->t, e->range->c, e->range->lut.data(), r, avail, row, sym_stride, size, used);
        } else if (e->aec->c.model == SCL_MODEL_ORDER_K && e->aec->c.ctx_global) {
            if (!model) return SCL_E_INVALID;
            std::vector<uint32_t> totals(e->aec->c.n_ctx);
            AecCtxGlobalPolicy pol = make_ctx_global_policy(totals.data(), e->aec->c, model + b * e->aec->model_words());
            pol.load();
            st = aec2_decode_lane(pol, e->aec->t, e->aec->c, r, avail, row, sym_stride, size, used);
            pol.store();
        } else if (e->aec->c.model == SCL_MODEL_ORDER_K) {
            alignas(16) uint32_t words[kAecCtxMaxWords];
            AecCtxPolicy pol = make_ctx_policy(words, e->aec->c);
            uint64_t *mm = model ? model + b * e->aec->model_words() : nullptr;
            pol.load(mm);
            st = aec2_decode_lane(pol, e->aec->t, e->aec->c, r, avail, row, sym_stride, size, used);
            if (mm) pol.store(mm);
        } else if (e->aec2 == 2 && !model && e->aec->model8_ok(sym_stride)) {
            alignas(16) uint32_t words[kAecModel8Words];
            AecIid8Policy pol = make_iid8_policy(words, e->aec->t, e->aec->c);
            st = aec2_decode_lane(pol, e->aec->t, e->aec->c, r, avail, row, sym_stride, size, used);
        } else if (e->aec2) {
            alignas(16) uint32_t words[kAecModelWords];
            uint64_t *mm = model ? model + b * e->aec->c.n_sym : nullptr;
            AecIidPolicy pol = make_iid_policy(words, e->aec->t, e->aec->c, mm);
            st = aec2_decode_lane(pol, e->aec->t, e->aec->c, r, avail, row, sym_stride, size, used);
            if (mm) pol.M.store(mm, e->aec->c.n_sym);
        } else {
            HostTree F;
            uint64_t *mm = model ? model + b * e->aec->c.n_sym : nullptr;
            uint64_t total = load_model(F, e->aec->t, e->aec->c, mm), tout = 0;
            st = aec_decode_lane(F, e->aec->t, e->aec->c, total, r, avail, row, sym_stride, size, used, tout);
            store_model(F, e->aec->c, mm);
        }
        sizes[b] = size;
        consumed[b] = used;
        status[b] = st;
    }
    return 0;
}

// ---- second-generation lane structs (scl_fast.cuh), driven like the v2 kernels drive them ----
int emu_v2_eligible(void *h) {
    Emu *e = (Emu *)h;
    if (e->tans) {
        const RansHost &r = e->tans->r;
        return r.max_bits_per_symbol <= kFastMaxBitsPerSym && r.c.L * 4 <= 64 * 1024 && r.c.NSB <= 32;
    }
    if (e->range) return e->range->v2 ? 1 : 0;
    if (!e->rans) return 0;
    const RansHost &r = *e->rans;
    return r.enc32 && r.dec32 && r.max_bits_per_symbol <= kFastMaxBitsPerSym && (r.c.NBO == 1 || r.c.NBO == 8);
}

}  // extern "C"

template <uint32_t NBO>
static void enc_v2_block(const RansHost &r, const uint8_t *row, uint32_t n, uint8_t *slot, uint64_t out_stride, uint64_t b,
                         uint64_t *bit_off, uint64_t *bit_len, uint32_t *status) {
    static thread_local uint32_t ring[kEncRingWords * kRingStrideWords];
    EncLaneV2 L;
    L.init((uint32_t)r.c.L, saddr_of(ring), slot, slot + out_stride);
    for (uint32_t i = 0; i < n; i += 16) {
        uint32_t cnt = n - i >= 16 ? 16u : n - i;
        uint8_t tmp[16] = {0};
        memcpy(tmp, row + i, cnt);
        u32x4 v;
        memcpy(&v, tmp, 16);
        if (r.c.check_sym)
            enc_chunk<NBO, true>(L, saddr_of(r.enc_tab.data()), 16, v, cnt);
        else
            enc_chunk<NBO, false>(L, saddr_of(r.enc_tab.data()), 16, v, cnt);
    }
    L.put(L.x, r.c.NSB);
    uint32_t st = SCL_ST_OK;
    if (r.c.DBSB < 32 && (n >> r.c.DBSB)) st = SCL_ST_OVERFLOW;
    L.put64((uint64_t)n, r.c.DBSB);
    uint64_t bits = L.finish();
    if (L.bad) st = SCL_ST_BAD_SYMBOL;
    if (L.ovf) st = SCL_ST_OVERFLOW;
    bit_len[b] = bits;
    bit_off[b] = (b + 1) * out_stride * 8 - bits;
    status[b] = st;
}

extern "C" {

static void tans_enc_v2_block(const Emu &e, const uint8_t *row, uint32_t n, uint8_t *slot, uint64_t out_stride, uint64_t b, uint64_t *bit_off,
                              uint64_t *bit_len, uint32_t *status) {
    static thread_local uint32_t ring[kEncRingWords * kRingStrideWords];
    const RansHost &r = e.tans->r;
    EncLaneV2 L;
    L.init((uint32_t)r.c.L, saddr_of(ring), slot, slot + out_stride);
    for (uint32_t i = 0; i < n; i += 16) {
        uint32_t cnt = n - i >= 16 ? 16u : n - i;
        uint8_t tmp[16] = {0};
        memcpy(tmp, row + i, cnt);
        u32x4 v;
        memcpy(&v, tmp, 16);
        if (r.c.check_sym)
            tans_enc_chunk<true>(L, saddr_of(e.tans->sym_tab8.data()), 8, saddr_of(e.tenc.data()), v, cnt);
        else
            tans_enc_chunk<false>(L, saddr_of(e.tans->sym_tab8.data()), 8, saddr_of(e.tenc.data()), v, cnt);
    }
    L.put(L.x, r.c.NSB);
    uint32_t st = SCL_ST_OK;
    if (r.c.DBSB < 32 && (n >> r.c.DBSB)) st = SCL_ST_OVERFLOW;
    L.put64((uint64_t)n, r.c.DBSB);
    uint64_t bits = L.finish();
    if (L.bad) st = SCL_ST_BAD_SYMBOL;
    if (L.ovf) st = SCL_ST_OVERFLOW;
    bit_len[b] = bits;
    bit_off[b] = (b + 1) * out_stride * 8 - bits;
    status[b] = st;
}

// range coder, second generation (scl_range.cuh), driven like range_encode_v2_kernel / range_decode_v2_kernel drive it
static void range_enc_v2_block(const RangeHost &rh, const uint8_t *row, uint32_t n, uint8_t *slot, uint64_t out_stride, uint64_t b,
                               uint64_t *bit_off, uint64_t *bit_len, uint32_t *status) {
    static thread_local uint32_t ring[kEncRingWords * kRingStrideWords];
    RangeEncV2 L;
    L.init(saddr_of(ring), slot, slot + out_stride, rh.c.t_shift);
    L.put_word(n);
    L.spill_check();
    for (uint32_t i = 0; i < n; i += 16) {
        uint32_t cnt = n - i >= 16 ? 16u : n - i;
        uint8_t tmp[16] = {0};
        memcpy(tmp, row + i, cnt);
        u32x4 v;
        memcpy(&v, tmp, 16);
        range_enc_chunk<true, true>(L, saddr_of(rh.enc_tab.data()), 4, v, cnt);
    }
    uint64_t bits = L.finish();
    uint32_t st = SCL_ST_OK;
    if (L.bad) st = SCL_ST_BAD_SYMBOL;
    if (L.ovf) st = SCL_ST_OVERFLOW;
    bit_len[b] = bits;
    bit_off[b] = b * out_stride * 8;
    status[b] = st;
}

static uint32_t range_dec_v2_block(const RangeHost &rh, DecLaneV2 &D, uint8_t *out, uint64_t out_cap, uint32_t &size_out, uint64_t &used) {
    RangeDecConst dc;
    dc.lut = saddr_of(rh.dec_lut.data());
    dc.shift = rh.c.t_shift;
    dc.neg1 = 0xFFFFFFFFu;
    dc.T = rh.c.T;
    dc.last = rh.last_entry;
    RangeDecV2 R;
    uint32_t size = 0;
    size_out = 0;
    if (!range_dec_header(D, out_cap, size, R)) return SCL_ST_OVERFLOW;
    range_dec_body<true>(D, R, dc, out, size, true);
    size_out = size;
    used = D.bp - D.start_bp;
    return R.ovf ? SCL_ST_OVERFLOW : SCL_ST_OK;
}

int emu_encode_blocks_v2(void *h, const uint8_t *sym, uint64_t sym_stride, uint32_t block_len, uint64_t n_blocks, uint8_t *out,
                         uint64_t out_stride, uint64_t *bit_off, uint64_t *bit_len, uint32_t *status) {
    Emu *e = (Emu *)h;
    if (!emu_v2_eligible(h) || (out_stride & 31) || (((uintptr_t)out) & 31)) return SCL_E_INVALID;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        if (e->range)
            range_enc_v2_block(*e->range, sym + b * sym_stride, block_len, out + b * out_stride, out_stride, b, bit_off, bit_len, status);
        else if (e->tans)
            tans_enc_v2_block(*e, sym + b * sym_stride, block_len, out + b * out_stride, out_stride, b, bit_off, bit_len, status);
        else if (e->rans->c.NBO == 1)
            enc_v2_block<1>(*e->rans, sym + b * sym_stride, block_len, out + b * out_stride, out_stride, b, bit_off, bit_len, status);
        else
            enc_v2_block<8>(*e->rans, sym + b * sym_stride, block_len, out + b * out_stride, out_stride, b, bit_off, bit_len, status);
    }
    return 0;
}

int emu_decode_blocks_v2(void *h, const uint8_t *in, uint64_t in_bytes, const uint64_t *bit_off, const uint64_t *bit_len, uint64_t n_blocks,
                         uint8_t *sym, uint64_t sym_stride, uint32_t *sizes, uint64_t *consumed, uint32_t *status) {
    Emu *e = (Emu *)h;
    if (!emu_v2_eligible(h) || (sym_stride & 31) || (((uintptr_t)sym) & 31) || (((uintptr_t)in) & 31)) return SCL_E_INVALID;
    static thread_local uint32_t ring[(kDecRingWords + 1) * kRingStrideWords];
    if (e->range) {
        for (uint64_t b = 0; b < n_blocks; ++b) {
            DecLaneV2 D;
            D.init(in, in_bytes, bit_off[b], saddr_of(ring));
            uint32_t size = 0;
            uint64_t used = 0;
            uint32_t st = range_dec_v2_block(*e->range, D, sym + b * sym_stride, sym_stride, size, used);
            uint64_t avail = bit_len ? bit_len[b] : (in_bytes * 8 > bit_off[b] ? in_bytes * 8 - bit_off[b] : 0);
            if (st == SCL_ST_OK && used > avail) st = SCL_ST_TRUNCATED;
            sizes[b] = st == SCL_ST_OK ? size : 0;
            consumed[b] = used;
            status[b] = st;
        }
        return 0;
    }
    const RansHost &r = e->tans ? e->tans->r : *e->rans;
    for (uint64_t b = 0; b < n_blocks; ++b) {
        DecLaneV2 D;
        D.init(in, in_bytes, bit_off[b], saddr_of(ring));
        uint32_t size = 0;
        uint64_t used = 0;
        uint32_t st = e->tans ? tans_decode_lane_v2(D, saddr_of(e->tdec.data()), r.c, sym + b * sym_stride, sym_stride, size, used)
                      : r.c.NBO == 1 ? rans32_decode_lane_v2<1>(D, saddr_of(r.dec_lut.data()), r.c, sym + b * sym_stride, sym_stride, size, used)
                                   : rans32_decode_lane_v2<8>(D, saddr_of(r.dec_lut.data()), r.c, sym + b * sym_stride, sym_stride, size, used);
        uint64_t avail = bit_len ? bit_len[b] : (in_bytes * 8 > bit_off[b] ? in_bytes * 8 - bit_off[b] : 0);
        if (st == SCL_ST_OK && used > avail) st = SCL_ST_TRUNCATED;
        sizes[b] = size;
        consumed[b] = used;
        status[b] = st;
    }
    return 0;
}
}
