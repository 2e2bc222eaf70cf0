/*
 * scl_oracle.c -- CPU restatement of the Stanford Compression Library's entropy-coder
 * hot path (rANS, tANS, arithmetic coder, range coder) in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (stanford_compression_library_b200/) imports, links or executes this file.  It is
 * used by tests/ as the parity checker for the CUDA kernels, by
 * __graft_entry__.smoke() as the checker, and by bench.py's `cpu_baseline` /
 * `--impl reference` legs as the timed CPU implementation.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here against
 *   (a) the reference's own known-answer vectors (rANS.py:303-360, tANS.py:285-415),
 *   (b) tests/golden/ (.npz files), produced by oracle/gen_golden.py from the UNMODIFIED
 *       reference imported from /root/reference (with oracle/bitarray_shim), and
 *   (c) the live reference, when /root/reference is present (build container only).
 *
 * Every function follows the reference's control flow literally (while-loops, binary
 * search with numpy.searchsorted(side="right") semantics, LIFO prepends), not the
 * closed forms the CUDA kernels use -- that independence is the point of an oracle.
 * File:line citations are relative to /root/reference/.
 *
 * Conventions: symbols are *indices* into the Frequencies alphabet in dict-insertion
 * order (prob_dist.py:169-171,176-178); bit streams are packed MSB-first, bit 0 = MSB of
 * byte 0, right-padded with zeros exactly like bitarray.tobytes().
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

#define SCL_OK 0
#define SCL_ERR_BAD_SYMBOL 1     /* KeyError in the reference (symbol not in freq_dict) */
#define SCL_ERR_STATE_MISMATCH 2 /* rANS.py:295 assert state == INITIAL_STATE */
#define SCL_ERR_OVERFLOW 3       /* output buffer too small / size does not fit the header */
#define SCL_ERR_TRUNCATED 4      /* ba2int on an empty slice -> ValueError */
#define SCL_ERR_PARAM 5
#define SCL_ERR_TOTAL_FREQ 6     /* arithmetic_coding.py:110-112 assert */

/* ------------------------------------------------------------------------------------ */
/* bit containers                                                                        */
/* ------------------------------------------------------------------------------------ */

/* growable array with one byte per bit; simple on purpose */
typedef struct {
    uint8_t *b;
    uint64_t n, cap;
} bitvec;

static int bv_push(bitvec *v, int bit) {
    if (v->n == v->cap) {
        uint64_t nc = v->cap ? v->cap * 2 : 4096;
        uint8_t *nb = (uint8_t *)realloc(v->b, nc);
        if (!nb) return -1;
        v->b = nb;
        v->cap = nc;
    }
    v->b[v->n++] = (uint8_t)(bit & 1);
    return 0;
}

/* uint_to_bitarray(x, width): MSB first (bitarray_utils.py:28-34) */
static int bv_push_uint(bitvec *v, u128 x, uint32_t width) {
    for (int i = (int)width - 1; i >= 0; --i)
        if (bv_push(v, (int)((x >> i) & 1))) return -1;
    return 0;
}

static void bv_free(bitvec *v) {
    free(v->b);
    v->b = NULL;
    v->n = v->cap = 0;
}

/* pack bits MSB-first into out (zero padded): bitarray.tobytes() */
static int pack_bits(const uint8_t *bits, uint64_t n, uint8_t *out, uint64_t out_cap) {
    uint64_t nbytes = (n + 7) / 8;
    if (nbytes > out_cap) return SCL_ERR_OVERFLOW;
    memset(out, 0, nbytes);
    for (uint64_t i = 0; i < n; ++i)
        if (bits[i]) out[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
    return SCL_OK;
}

static inline int get_bit(const uint8_t *in, uint64_t pos) { return (in[pos >> 3] >> (7 - (pos & 7))) & 1; }

/* bitarray_to_uint(encoded[pos : pos+width]) with Python slice clamping at `nbits`
 * (bitarray_utils.py:37-38).  *got = number of bits actually present. */
static u128 read_uint_clamped(const uint8_t *in, uint64_t nbits, uint64_t pos, uint32_t width, uint32_t *got) {
    u128 v = 0;
    uint32_t g = 0;
    for (uint32_t i = 0; i < width; ++i) {
        if (pos + i >= nbits) break;
        v = (v << 1) | (u128)get_bit(in, pos + i);
        ++g;
    }
    *got = g;
    return v;
}

/* numpy.searchsorted(a, v, side="right") - 1  (rANS.py:231, arithmetic_coding.py:199,
 * range_coder.py:236): index of the last element <= v; -1 if none. */
static int64_t searchsorted_right_minus1(const u128 *a, uint32_t n, u128 v) {
    uint32_t lo = 0, hi = n; /* first index with a[i] > v */
    while (lo < hi) {
        uint32_t mid = (lo + hi) / 2;
        if (a[mid] <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (int64_t)lo - 1;
}

/* ------------------------------------------------------------------------------------ */
/* rANS  (scl/compressors/rANS.py)                                                       */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t n_sym;
    const uint64_t *freq; /* dict order */
    uint64_t *cum;        /* cumulative_freq_dict, prob_dist.py:193-205 */
    uint64_t M, L, H, RF;
    uint32_t DBSB, NBO, NSB;
} rans_params;

/* rANSParams.__post_init__ (rANS.py:97-120).  NUM_STATE_BITS is supplied by the caller,
 * who evaluates the reference's float formula get_bit_width(H) (bitarray_utils.py:8-20). */
static int rans_params_init(rans_params *p, const uint64_t *freq, uint32_t n_sym, uint32_t dbsb, uint32_t nbo,
                            uint64_t rf, uint32_t nsb) {
    if (n_sym == 0 || nbo == 0 || nbo > 32 || nsb == 0 || nsb > 64 || dbsb > 64) return SCL_ERR_PARAM;
    p->n_sym = n_sym;
    p->freq = freq;
    p->cum = (uint64_t *)malloc(sizeof(uint64_t) * n_sym);
    if (!p->cum) return SCL_ERR_PARAM;
    uint64_t s = 0;
    for (uint32_t i = 0; i < n_sym; ++i) {
        p->cum[i] = s;
        s += freq[i];
    }
    p->M = s;
    p->RF = rf;
    u128 L = (u128)rf * s;
    u128 H = L * ((u128)1 << nbo) - 1;
    if (H >> 63) {
        free(p->cum);
        return SCL_ERR_PARAM;
    }
    p->L = (uint64_t)L;
    p->H = (uint64_t)H;
    p->DBSB = dbsb;
    p->NBO = nbo;
    p->NSB = nsb;
    return SCL_OK;
}

static void rans_params_free(rans_params *p) { free(p->cum); }

/* rans_base_encode_step (rANS.py:138-147) */
static inline uint64_t rans_base_encode_step(const rans_params *p, uint32_t s, uint64_t state) {
    uint64_t f = p->freq[s];
    uint64_t block_id = state / f;
    uint64_t slot = p->cum[s] + (state % f);
    return block_id * p->M + slot;
}

/* encode_block (rANS.py:186-210) with shrink_state (rANS.py:149-161) inlined.
 * `rev` collects the payload in reverse order (see header comment in encode for why). */
static int rans_encode_block(const rans_params *p, const uint8_t *sym, uint64_t n, bitvec *out) {
    bitvec rev = {0};
    uint64_t state = p->L; /* INITIAL_STATE, rANS.py:113 */
    int rc = SCL_OK;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = sym[i];
        if (s >= p->n_sym) {
            rc = SCL_ERR_BAD_SYMBOL;
            goto done;
        }
        /* max_shrunk_state[s] = RF*f*2^NBO - 1 (rANS.py:108-109) */
        u128 max_shrunk = (u128)p->RF * p->freq[s] * ((u128)1 << p->NBO) - 1;
        /* shrink_state: each emitted chunk is NBO bits MSB-first, later chunks are placed
         * BEFORE earlier ones (rANS.py:158), and each symbol's bits are placed BEFORE all
         * previous symbols' bits (rANS.py:196).  Reversed payload == chunks in emission
         * order, each chunk LSB-first. */
        while ((u128)state > max_shrunk) {
            uint64_t chunk = state % ((uint64_t)1 << p->NBO);
            for (uint32_t b = 0; b < p->NBO; ++b)
                if (bv_push(&rev, (int)((chunk >> b) & 1))) {
                    rc = SCL_ERR_OVERFLOW;
                    goto done;
                }
            state >>= p->NBO;
        }
        state = rans_base_encode_step(p, s, state);
    }
    /* uint_to_bitarray raises OverflowError when the value does not fit (rANS.py:199,206) */
    if (p->DBSB < 64 && (n >> p->DBSB)) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    if (p->NSB < 64 && (state >> p->NSB)) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    if (bv_push_uint(out, n, p->DBSB) || bv_push_uint(out, state, p->NSB)) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    for (uint64_t i = rev.n; i > 0; --i)
        if (bv_push(out, rev.b[i - 1])) {
            rc = SCL_ERR_OVERFLOW;
            goto done;
        }
done:
    bv_free(&rev);
    return rc;
}

/* decode_block (rANS.py:270-297): rans_base_decode_step (:234-249), expand_state (:251-260) */
static int rans_decode_block(const rans_params *p, const uint8_t *in, uint64_t nbits, uint8_t *out, uint64_t out_cap,
                             uint64_t *n_out, uint64_t *bits_consumed) {
    uint32_t got;
    u128 *cum128 = (u128 *)malloc(sizeof(u128) * p->n_sym);
    if (!cum128) return SCL_ERR_PARAM;
    for (uint32_t i = 0; i < p->n_sym; ++i) cum128[i] = p->cum[i];
    int rc = SCL_OK;
    uint64_t size = (uint64_t)read_uint_clamped(in, nbits, 0, p->DBSB, &got);
    if (got == 0) {
        rc = SCL_ERR_TRUNCATED;
        goto done;
    }
    uint64_t pos = p->DBSB;
    uint64_t state = (uint64_t)read_uint_clamped(in, nbits, pos, p->NSB, &got);
    if (got == 0) {
        rc = SCL_ERR_TRUNCATED;
        goto done;
    }
    pos += p->NSB;
    if (size > out_cap) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    /* symbols come out last-first and are prepended (rANS.py:289-291) */
    for (uint64_t k = 0; k < size; ++k) {
        uint64_t block_id = state / p->M;
        uint64_t slot = state % p->M;
        int64_t idx = searchsorted_right_minus1(cum128, p->n_sym, slot);
        uint32_t s = (uint32_t)idx;
        state = block_id * p->freq[s] + slot - p->cum[s];
        while (state < p->L) { /* expand_state */
            uint64_t rem = (uint64_t)read_uint_clamped(in, nbits, pos, p->NBO, &got);
            if (got == 0) {
                rc = SCL_ERR_TRUNCATED;
                goto done;
            }
            pos += p->NBO;
            state = (state << p->NBO) + rem;
        }
        out[size - 1 - k] = (uint8_t)s;
    }
    if (state != p->L) {
        rc = SCL_ERR_STATE_MISMATCH;
        goto done;
    }
    *n_out = size;
    *bits_consumed = pos;
done:
    free(cum128);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* tANS  (scl/compressors/tANS.py) -- lookup tables built exactly as the reference does   */
/* ------------------------------------------------------------------------------------ */

/* get_bit_width for the small integers used inside tANS (< 2^49, where the reference's
 * float formula equals bit_length; bitarray_utils.py:8-20, SURVEY 8a note) */
static uint32_t bit_width_u64(uint64_t x) {
    if (x == 0) return 1;
    uint32_t w = 0;
    while (x) {
        ++w;
        x >>= 1;
    }
    return w;
}

typedef struct {
    rans_params rp;
    uint64_t *enc_table;     /* base_encode_step_table[(s, x_shrunk)] flattened: row offset + (x_shrunk - min_shrunk) */
    uint64_t *enc_row;       /* row offset per symbol */
    uint32_t *nbits_base;    /* shrink_state_num_out_bits_base_table */
    uint64_t *thresh;        /* shrink_state_thresh_table */
    uint32_t *dec_sym;       /* base_decode_step_table[state - L] -> s */
    uint64_t *dec_shrunk;    /*                                  -> x_shrunk */
} tans_tables;

static void tans_free(tans_tables *t) {
    free(t->enc_table);
    free(t->enc_row);
    free(t->nbits_base);
    free(t->thresh);
    free(t->dec_sym);
    free(t->dec_shrunk);
    rans_params_free(&t->rp);
}

static int tans_build(tans_tables *t, const uint64_t *freq, uint32_t n_sym, uint32_t dbsb, uint32_t nbo, uint64_t rf,
                      uint32_t nsb) {
    memset(t, 0, sizeof(*t));
    int rc = rans_params_init(&t->rp, freq, n_sym, dbsb, nbo, rf, nsb);
    if (rc) return rc;
    rans_params *p = &t->rp;
    /* tANSParams asserts (tANS.py:38-49): M power of two, NUM_BITS_OUT == 1 */
    if ((p->M & (p->M - 1)) != 0 || nbo != 1 || p->L > ((uint64_t)1 << 26)) {
        rans_params_free(p);
        return SCL_ERR_PARAM;
    }
    uint64_t L = p->L;
    t->enc_table = (uint64_t *)malloc(sizeof(uint64_t) * L);
    t->enc_row = (uint64_t *)malloc(sizeof(uint64_t) * n_sym);
    t->nbits_base = (uint32_t *)malloc(sizeof(uint32_t) * n_sym);
    t->thresh = (uint64_t *)malloc(sizeof(uint64_t) * n_sym);
    t->dec_sym = (uint32_t *)malloc(sizeof(uint32_t) * L);
    t->dec_shrunk = (uint64_t *)malloc(sizeof(uint64_t) * L);
    if (!t->enc_table || !t->enc_row || !t->nbits_base || !t->thresh || !t->dec_sym || !t->dec_shrunk) {
        tans_free(t);
        return SCL_ERR_PARAM;
    }
    u128 *cum128 = (u128 *)malloc(sizeof(u128) * n_sym);
    for (uint32_t i = 0; i < n_sym; ++i) cum128[i] = p->cum[i];
    uint64_t row = 0;
    for (uint32_t s = 0; s < n_sym; ++s) {
        uint64_t mn = rf * freq[s], mx = rf * freq[s] * 2 - 1; /* min/max_shrunk_state, NBO == 1 */
        /* build_base_encode_step_table (tANS.py:88-99) */
        t->enc_row[s] = row;
        for (uint64_t x = mn; x <= mx; ++x) t->enc_table[row + (x - mn)] = rans_base_encode_step(p, s, x);
        row += mx - mn + 1;
        /* shrink_state_num_out_bits_base (tANS.py:74-86) */
        uint32_t y = bit_width_u64(mx);
        t->nbits_base[s] = nsb - y;
        t->thresh[s] = (mx + 1) << t->nbits_base[s];
    }
    /* build_rans_base_decode_table (tANS.py:208-215) */
    for (uint64_t x = L; x <= p->H; ++x) {
        uint64_t block_id = x / p->M, slot = x % p->M;
        uint32_t s = (uint32_t)searchsorted_right_minus1(cum128, n_sym, slot);
        t->dec_sym[x - L] = s;
        t->dec_shrunk[x - L] = block_id * freq[s] + slot - p->cum[s];
    }
    free(cum128);
    return SCL_OK;
}

/* tANSEncoder.encode_block (tANS.py:159-193) / encode_symbol (:126-157) */
static int tans_encode_block(const tans_tables *t, const uint8_t *sym, uint64_t n, bitvec *out) {
    const rans_params *p = &t->rp;
    bitvec rev = {0};
    uint64_t state = p->L;
    int rc = SCL_OK;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = sym[i];
        if (s >= p->n_sym) {
            rc = SCL_ERR_BAD_SYMBOL;
            goto done;
        }
        uint32_t nb = t->nbits_base[s];
        if (state >= t->thresh[s]) nb += 1;
        /* out_bits = uint_to_bitarray(state)[-nb:] : low nb bits, MSB-first; reversed = LSB-first */
        for (uint32_t b = 0; b < nb; ++b)
            if (bv_push(&rev, (int)((state >> b) & 1))) {
                rc = SCL_ERR_OVERFLOW;
                goto done;
            }
        state >>= nb;
        state = t->enc_table[t->enc_row[s] + (state - p->RF * p->freq[s])];
    }
    if ((p->DBSB < 64 && (n >> p->DBSB)) || (p->NSB < 64 && (state >> p->NSB))) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    if (bv_push_uint(out, n, p->DBSB) || bv_push_uint(out, state, p->NSB)) {
        rc = SCL_ERR_OVERFLOW;
        goto done;
    }
    for (uint64_t i = rev.n; i > 0; --i)
        if (bv_push(out, rev.b[i - 1])) {
            rc = SCL_ERR_OVERFLOW;
            goto done;
        }
done:
    bv_free(&rev);
    return rc;
}

/* tANSDecoder.decode_block (tANS.py:252-279) / decode_symbol (:239-250) */
static int tans_decode_block(const tans_tables *t, const uint8_t *in, uint64_t nbits, uint8_t *out, uint64_t out_cap,
                             uint64_t *n_out, uint64_t *bits_consumed) {
    const rans_params *p = &t->rp;
    uint32_t got;
    uint64_t size = (uint64_t)read_uint_clamped(in, nbits, 0, p->DBSB, &got);
    if (got == 0) return SCL_ERR_TRUNCATED;
    uint64_t pos = p->DBSB;
    uint64_t state = (uint64_t)read_uint_clamped(in, nbits, pos, p->NSB, &got);
    if (got == 0) return SCL_ERR_TRUNCATED;
    pos += p->NSB;
    if (size > out_cap) return SCL_ERR_OVERFLOW;
    for (uint64_t k = 0; k < size; ++k) {
        if (state < p->L || state > p->H) return SCL_ERR_STATE_MISMATCH; /* KeyError in the reference's dict */
        uint32_t s = t->dec_sym[state - p->L];
        uint64_t shrunk = t->dec_shrunk[state - p->L];
        /* expand_state_num_bits_table (tANS.py:217-226) */
        uint32_t nb = p->NSB - bit_width_u64(shrunk);
        uint64_t rem = 0;
        if (nb) {
            rem = (uint64_t)read_uint_clamped(in, nbits, pos, nb, &got);
            if (got == 0) return SCL_ERR_TRUNCATED;
        }
        state = (shrunk << nb) + rem;
        pos += nb;
        out[size - 1 - k] = (uint8_t)s;
    }
    if (state != p->L) return SCL_ERR_STATE_MISMATCH;
    *n_out = size;
    *bits_consumed = pos;
    return SCL_OK;
}

/* ------------------------------------------------------------------------------------ */
/* frequency models (scl/compressors/probability_models.py)                              */
/* ------------------------------------------------------------------------------------ */

#define SCL_MODEL_FIXED 0        /* FixedFreqModel (:57-67) */
#define SCL_MODEL_ADAPTIVE_IID 1 /* AdaptiveIIDFreqModel (:70-92) */
#define SCL_MODEL_ORDER_K 2      /* AdaptiveOrderKFreqModel (:95-160), k in bits 8.. of the kind word */

/* Order-k state: `freq` is then the whole count array freqs_kplus1_tuple flattened as
 * [context][symbol] (n_ctx = n_sym^k rows) FOLLOWED by one word holding the current context index
 * (past_k read as a base-n_sym number, oldest symbol most significant; :114-126,146-148). */
static uint64_t ipow_u64(uint64_t a, uint32_t k) {
    uint64_t r = 1;
    while (k--) r *= a;
    return r;
}
static uint64_t *model_row(int kind, uint64_t *freq, uint32_t n_sym) {
    if ((kind & 0xFF) != SCL_MODEL_ORDER_K) return freq;
    uint64_t n_ctx = ipow_u64(n_sym, (uint32_t)kind >> 8);
    return freq + freq[n_ctx * n_sym] * n_sym; /* freqs_current (:128-140) */
}

static int model_update(int kind, uint64_t *freq, uint32_t n_sym, uint32_t s, uint64_t max_total) {
    if ((kind & 0xFF) == SCL_MODEL_ORDER_K) { /* update_model (:142-160) */
        uint32_t k = (uint32_t)kind >> 8;
        uint64_t n_ctx = ipow_u64(n_sym, k);
        uint64_t *ctx = &freq[n_ctx * n_sym];
        uint64_t *cnt = &freq[*ctx * n_sym + s];
        *cnt += 1;
        if (k > 0) *ctx = (*ctx * n_sym + s) % n_ctx; /* past_k = past_k[1:] + [idx] */
        /* :156-160 tests ONE count against max_allowed_total_freq and then calls np.max(x // 2, 1),
         * which raises (axis 1 of a scalar): report it instead of guessing */
        if (*cnt >= max_total) return SCL_ERR_TOTAL_FREQ;
        return SCL_OK;
    }
    if (kind != SCL_MODEL_ADAPTIVE_IID) return SCL_OK;
    freq[s] += 1; /* :86 */
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n_sym; ++i) tot += freq[i];
    if (tot >= max_total) /* :90-92 */
        for (uint32_t i = 0; i < n_sym; ++i) {
            uint64_t h = freq[i] / 2;
            freq[i] = h > 1 ? h : 1;
        }
    return SCL_OK;
}

/* ------------------------------------------------------------------------------------ */
/* arithmetic coder (scl/compressors/arithmetic_coding.py)                               */
/* ------------------------------------------------------------------------------------ */

/* ArithmeticEncoder.encode_block (:80-161).  `freq` is the model's CURRENT table and is
 * mutated in place: the reference never resets the model between blocks
 * (data_encoder_decoder.py:23-27 reset() is a no-op). */
static int aec_encode_block(uint32_t dbsb, uint32_t P, int model, uint64_t *freq, uint32_t n_sym, uint64_t max_total,
                            const uint8_t *sym, uint64_t n, bitvec *out) {
    if (P < 2 || P > 62 || dbsb > 64) return SCL_ERR_PARAM;
    const u128 FULL = (u128)1 << P, HALF = (u128)1 << (P - 1), QTR = (u128)1 << (P - 2);
    const u128 MAX_TOTAL = QTR; /* AECParams.MAX_ALLOWED_TOTAL_FREQ (:37) */
    /* :85 assert size < (1 << MAX_BLOCK_SIZE): never fails for any representable size */
    if (dbsb < 64 && (n >> dbsb)) return SCL_ERR_OVERFLOW; /* uint_to_bitarray OverflowError (:99) */
    u128 low = 0, high = FULL;
    if (bv_push_uint(out, n, dbsb)) return SCL_ERR_OVERFLOW;
    uint64_t num_mid = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = sym[i];
        if (s >= n_sym) return SCL_ERR_BAD_SYMBOL;
        const uint64_t *row = model_row(model, freq, n_sym);
        u128 T = 0, c = 0;
        for (uint32_t j = 0; j < n_sym; ++j) {
            if (j == s) c = T;
            T += row[j];
        }
        if (!(T < MAX_TOTAL)) return SCL_ERR_TOTAL_FREQ; /* :110-112 */
        /* shrink_range (:58-78) */
        u128 rng = high - low, d = c + row[s];
        high = low + (rng * d) / T;
        low = low + (rng * c) / T;
        if (model_update(model, freq, n_sym, s, max_total)) return SCL_ERR_TOTAL_FREQ; /* :118 */
        while (high < HALF || low > HALF) { /* :126 (strict tests) */
            if (high < HALF) {
                if (bv_push(out, 0)) return SCL_ERR_OVERFLOW;
                for (uint64_t k = 0; k < num_mid; ++k)
                    if (bv_push(out, 1)) return SCL_ERR_OVERFLOW;
                low <<= 1;
                high <<= 1;
                num_mid = 0;
            } else if (low > HALF) {
                if (bv_push(out, 1)) return SCL_ERR_OVERFLOW;
                for (uint64_t k = 0; k < num_mid; ++k)
                    if (bv_push(out, 0)) return SCL_ERR_OVERFLOW;
                low = (low - HALF) << 1;
                high = (high - HALF) << 1;
                num_mid = 0;
            }
        }
        while (low > QTR && high < 3 * QTR) { /* :146 */
            num_mid += 1;
            low = (low - QTR) << 1;
            high = (high - QTR) << 1;
        }
    }
    num_mid += 1; /* :153 */
    if (low <= QTR) {
        if (bv_push(out, 0)) return SCL_ERR_OVERFLOW;
        for (uint64_t k = 0; k < num_mid; ++k)
            if (bv_push(out, 1)) return SCL_ERR_OVERFLOW;
    } else {
        if (bv_push(out, 1)) return SCL_ERR_OVERFLOW;
        for (uint64_t k = 0; k < num_mid; ++k)
            if (bv_push(out, 0)) return SCL_ERR_OVERFLOW;
    }
    return SCL_OK;
}

/* ArithmeticDecoder.decode_block (:203-287) with decode_step_core (:177-201) */
static int aec_decode_block(uint32_t dbsb, uint32_t P, int model, uint64_t *freq, uint32_t n_sym, uint64_t max_total,
                            const uint8_t *in, uint64_t nbits, uint8_t *out, uint64_t out_cap, uint64_t *n_out,
                            uint64_t *bits_consumed) {
    if (P < 2 || P > 62 || dbsb > 64) return SCL_ERR_PARAM;
    const u128 FULL = (u128)1 << P, HALF = (u128)1 << (P - 1), QTR = (u128)1 << (P - 2);
    uint32_t got;
    uint64_t size = (uint64_t)read_uint_clamped(in, nbits, 0, dbsb, &got);
    if (got == 0) return SCL_ERR_TRUNCATED;
    if (size > out_cap) return SCL_ERR_OVERFLOW;
    /* the reference's `while True` decodes a symbol before testing the count, so an empty
     * block never terminates (:232-243); we report it instead of hanging */
    if (size == 0) return SCL_ERR_PARAM;
    const uint64_t base = dbsb;                        /* encoded_bitarray = encoded_bitarray[DBSB:] (:206) */
    const uint64_t A = nbits > base ? nbits - base : 0; /* arith_bitarray_size */
    uint64_t nbc = 0;
    u128 low = 0, high = FULL, state = 0;
    while (nbc < P && nbc < A) { /* :222-228 */
        if (get_bit(in, base + nbc)) state += (u128)1 << (P - nbc - 1);
        nbc += 1;
    }
    nbc = P; /* :229 */
    u128 *search = (u128 *)malloc(sizeof(u128) * n_sym);
    if (!search) return SCL_ERR_PARAM;
    uint64_t count = 0;
    int rc = SCL_OK;
    for (;;) {
        const uint64_t *row = model_row(model, freq, n_sym);
        u128 T = 0;
        for (uint32_t j = 0; j < n_sym; ++j) T += row[j];
        u128 rng = high - low, c = 0;
        for (uint32_t j = 0; j < n_sym; ++j) { /* search_list (:196-198) */
            search[j] = low + (c * rng) / T;
            c += row[j];
        }
        int64_t idx = searchsorted_right_minus1(search, n_sym, state);
        if (idx < 0) idx = (int64_t)n_sym - 1; /* Python alphabet[-1] */
        uint32_t s = (uint32_t)idx;
        c = 0;
        for (uint32_t j = 0; j < s; ++j) c += row[j];
        u128 d = c + row[s];
        high = low + (rng * d) / T; /* shrink_range */
        low = low + (rng * c) / T;
        out[count++] = (uint8_t)s;
        if (model_update(model, freq, n_sym, s, max_total)) {
            rc = SCL_ERR_TOTAL_FREQ;
            break;
        }
        if (count == size) break; /* :242-243 -- before renormalisation */
        while (high < HALF || low > HALF) {
            if (high < HALF) {
                low <<= 1;
                high <<= 1;
                state <<= 1;
            } else if (low > HALF) {
                low = (low - HALF) << 1;
                high = (high - HALF) << 1;
                state = (state - HALF) << 1;
            }
            if (nbc < A) state += (u128)get_bit(in, base + nbc);
            nbc += 1;
        }
        while (low > QTR && high < 3 * QTR) {
            low = (low - QTR) << 1;
            high = (high - QTR) << 1;
            state = (state - QTR) << 1;
            if (nbc < A) state += (u128)get_bit(in, base + nbc);
            nbc += 1;
        }
    }
    /* trailing-bit accounting (:277-282); Python leaves the loop variable at P-1 if no break */
    uint32_t extra = 0;
    for (extra = 0; extra < P; ++extra) {
        u128 state_low = (state >> extra) << extra;
        u128 state_high = state_low + ((u128)1 << extra);
        if (state_low < low || state_high > high) break;
    }
    if (extra == P) extra = P - 1;
    /* num_bits_consumed -= extra_bits_read - 1 */
    int64_t nb = (int64_t)nbc - ((int64_t)extra - 1);
    nb += dbsb; /* :285 */
    *n_out = size;
    *bits_consumed = (uint64_t)nb;
    free(search);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* range coder (scl/compressors/range_coder.py)                                          */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t DBSB, P;
    u128 TOP, BOTTOM, MASK;
    uint32_t n_sym;
    const uint64_t *freq;
    uint64_t *cum;
    uint64_t T;
} range_params;

static int range_params_init(range_params *p, uint32_t dbsb, uint32_t P, const uint64_t *freq, uint32_t n_sym) {
    if (P % 8 != 0 || P < 24 || P > 56 || dbsb > 64 || n_sym == 0) return SCL_ERR_PARAM; /* :66 */
    p->DBSB = dbsb;
    p->P = P;
    p->TOP = (u128)1 << (P - 8);
    p->BOTTOM = (u128)1 << (P - 16);
    p->MASK = ((u128)1 << P) - 1;
    p->n_sym = n_sym;
    p->freq = freq;
    p->cum = (uint64_t *)malloc(sizeof(uint64_t) * n_sym);
    if (!p->cum) return SCL_ERR_PARAM;
    uint64_t s = 0;
    for (uint32_t i = 0; i < n_sym; ++i) {
        if (freq[i] == 0) { /* :84 */
            free(p->cum);
            return SCL_ERR_PARAM;
        }
        p->cum[i] = s;
        s += freq[i];
    }
    p->T = s;
    if ((u128)s > p->BOTTOM) { /* :85 */
        free(p->cum);
        return SCL_ERR_PARAM;
    }
    return SCL_OK;
}

static int bv_push_byte(bitvec *v, uint32_t byte) {
    if (byte > 255) return -1; /* bytes([..]) would raise ValueError */
    return bv_push_uint(v, byte, 8);
}

/* RangeEncoder.encode_block (:188-207): shrink_range (:88-105), normalize (:107-179), flush (:181-186) */
static int range_encode_block(const range_params *p, const uint8_t *sym, uint64_t n, bitvec *out) {
    if (p->DBSB < 64 && (n >> p->DBSB)) return SCL_ERR_OVERFLOW;
    u128 low = 0, range = p->MASK;
    if (bv_push_uint(out, n, p->DBSB)) return SCL_ERR_OVERFLOW;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t s = sym[i];
        if (s >= p->n_sym) return SCL_ERR_BAD_SYMBOL;
        u128 c = p->cum[s], d = c + p->freq[s];
        range = range / p->T;
        low += c * range;
        range *= d - c;
        while ((low ^ (low + range)) < p->TOP || range < p->BOTTOM) {
            if ((low ^ (low + range)) < p->TOP) {
                if (bv_push_byte(out, (uint32_t)(low >> (p->P - 8)))) return SCL_ERR_OVERFLOW;
                low <<= 8;
                range <<= 8;
                low &= p->MASK;
                continue;
            }
            if (range < p->BOTTOM) {
                range = (p->MASK + 1 - low) & (p->BOTTOM - 1);
                if (bv_push_byte(out, (uint32_t)(low >> (p->P - 8)))) return SCL_ERR_OVERFLOW;
                low <<= 8;
                range <<= 8;
                low &= p->MASK;
            }
        }
    }
    for (uint32_t k = 0; k < p->P / 8; ++k) { /* flush */
        if (bv_push_byte(out, (uint32_t)(low >> (p->P - 8)))) return SCL_ERR_OVERFLOW;
        low <<= 8;
        low &= p->MASK;
    }
    return SCL_OK;
}

/* RangeDecoder.decode_block (:269-317): decode_symbol (:225-238), normalize (:240-267) */
static int range_decode_block(const range_params *p, const uint8_t *in, uint64_t nbits, uint8_t *out, uint64_t out_cap,
                              uint64_t *n_out, uint64_t *bits_consumed) {
    uint32_t got;
    uint64_t size = (uint64_t)read_uint_clamped(in, nbits, 0, p->DBSB, &got);
    if (got == 0) return SCL_ERR_TRUNCATED;
    if (size > out_cap) return SCL_ERR_OVERFLOW;
    const uint64_t base = p->DBSB;
    uint64_t nbc = 0;
    u128 low = 0, range = p->MASK, state = 0;
    for (uint32_t k = 0; k < p->P / 8; ++k) { /* :289-291 (note: no MASK here, as in the reference) */
        uint64_t byte = (uint64_t)read_uint_clamped(in, nbits, base + nbc, 8, &got);
        if (got == 0) return SCL_ERR_TRUNCATED;
        nbc += 8;
        state = (state << 8) | byte;
    }
    u128 *search = (u128 *)malloc(sizeof(u128) * p->n_sym);
    if (!search) return SCL_ERR_PARAM;
    int rc = SCL_OK;
    uint64_t count = 0;
    if (size != 0) {
        for (;;) {
            u128 r = range / p->T;
            for (uint32_t j = 0; j < p->n_sym; ++j) search[j] = low + (u128)p->cum[j] * r;
            int64_t idx = searchsorted_right_minus1(search, p->n_sym, state);
            if (idx < 0) idx = (int64_t)p->n_sym - 1;
            uint32_t s = (uint32_t)idx;
            out[count++] = (uint8_t)s;
            u128 c = p->cum[s], d = c + p->freq[s];
            range = r;
            low += c * range;
            range *= d - c;
            while ((low ^ (low + range)) < p->TOP || range < p->BOTTOM) {
                if (!((low ^ (low + range)) < p->TOP)) range = (p->MASK + 1 - low) & (p->BOTTOM - 1);
                uint64_t byte = (uint64_t)read_uint_clamped(in, nbits, base + nbc, 8, &got);
                if (got == 0) {
                    rc = SCL_ERR_TRUNCATED;
                    goto done;
                }
                nbc += 8;
                state = ((state << 8) | byte) & p->MASK;
                low = (low << 8) & p->MASK;
                range <<= 8;
            }
            if (count == size) break;
        }
    }
    *n_out = size;
    *bits_consumed = nbc + p->DBSB;
done:
    free(search);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* exported entry points                                                                 */
/* ------------------------------------------------------------------------------------ */

#define SCL_CODER_RANS 0
#define SCL_CODER_TANS 1
#define SCL_CODER_RANGE 2
#define SCL_CODER_AEC 3

/* Generic parameter block shared by all exported functions.
 *   rANS/tANS : p0 = DATA_BLOCK_SIZE_BITS, p1 = NUM_BITS_OUT, p2 = RANGE_FACTOR, p3 = NUM_STATE_BITS
 *   range     : p0 = DATA_BLOCK_SIZE_BITS, p1 = PRECISION
 *   AEC       : p0 = DATA_BLOCK_SIZE_BITS, p1 = PRECISION, p2 = model kind, p3 = max_allowed_total_freq
 */
typedef struct {
    int coder;
    uint64_t p0, p1, p2, p3;
} oracle_cfg;

typedef struct {
    oracle_cfg cfg;
    uint32_t n_sym;
    uint64_t *freq;
    rans_params rp;
    tans_tables tt;
    range_params gp;
} oracle_ctx;

void *scl_oracle_create(int coder, const uint64_t *freq, uint32_t n_sym, uint64_t p0, uint64_t p1, uint64_t p2,
                        uint64_t p3, int *err) {
    oracle_ctx *c = (oracle_ctx *)calloc(1, sizeof(oracle_ctx));
    int rc = SCL_OK;
    if (!c) {
        if (err) *err = SCL_ERR_PARAM;
        return NULL;
    }
    c->cfg.coder = coder;
    c->cfg.p0 = p0;
    c->cfg.p1 = p1;
    c->cfg.p2 = p2;
    c->cfg.p3 = p3;
    c->n_sym = n_sym;
    c->freq = (uint64_t *)malloc(sizeof(uint64_t) * (n_sym ? n_sym : 1));
    memcpy(c->freq, freq, sizeof(uint64_t) * n_sym);
    switch (coder) {
    case SCL_CODER_RANS:
        rc = rans_params_init(&c->rp, c->freq, n_sym, (uint32_t)p0, (uint32_t)p1, p2, (uint32_t)p3);
        break;
    case SCL_CODER_TANS:
        rc = tans_build(&c->tt, c->freq, n_sym, (uint32_t)p0, (uint32_t)p1, p2, (uint32_t)p3);
        break;
    case SCL_CODER_RANGE:
        rc = range_params_init(&c->gp, (uint32_t)p0, (uint32_t)p1, c->freq, n_sym);
        break;
    case SCL_CODER_AEC:
        if (p1 < 2 || p1 > 62 || n_sym == 0) rc = SCL_ERR_PARAM;
        break;
    default:
        rc = SCL_ERR_PARAM;
    }
    if (rc) {
        free(c->freq);
        free(c);
        if (err) *err = rc;
        return NULL;
    }
    if (err) *err = SCL_OK;
    return c;
}

void scl_oracle_destroy(void *h) {
    oracle_ctx *c = (oracle_ctx *)h;
    if (!c) return;
    if (c->cfg.coder == SCL_CODER_RANS) rans_params_free(&c->rp);
    if (c->cfg.coder == SCL_CODER_TANS) tans_free(&c->tt);
    if (c->cfg.coder == SCL_CODER_RANGE) free(c->gp.cum);
    free(c->freq);
    free(c);
}

/* a private model for one block: the creation-time table, or all-ones + context 0 for order-k */
static uint64_t *fresh_model(const oracle_ctx *c) {
    int kind = (int)c->cfg.p2;
    if ((kind & 0xFF) == SCL_MODEL_ORDER_K) {
        uint64_t n = ipow_u64(c->n_sym, (uint32_t)kind >> 8) * c->n_sym;
        uint64_t *m = (uint64_t *)malloc(sizeof(uint64_t) * (n + 1));
        for (uint64_t i = 0; i < n; ++i) m[i] = 1; /* np.ones (:106) */
        m[n] = 0;                                  /* past_k = [0] * k (:112) */
        return m;
    }
    uint64_t *m = (uint64_t *)malloc(sizeof(uint64_t) * c->n_sym);
    memcpy(m, c->freq, sizeof(uint64_t) * c->n_sym);
    return m;
}

/* Encode one block.  `model_freq` (AEC only; may be NULL -> use a private copy of the
 * creation-time table) is the model's current table, updated in place like the reference's
 * freq_model.  Returns status; *out_bits = stream length in bits. */
int scl_oracle_encode_block(void *h, const uint8_t *sym, uint64_t n, uint64_t *model_freq, uint8_t *out,
                            uint64_t out_cap, uint64_t *out_bits) {
    oracle_ctx *c = (oracle_ctx *)h;
    bitvec bv = {0};
    int rc;
    uint64_t *tmp = NULL;
    switch (c->cfg.coder) {
    case SCL_CODER_RANS:
        rc = rans_encode_block(&c->rp, sym, n, &bv);
        break;
    case SCL_CODER_TANS:
        rc = tans_encode_block(&c->tt, sym, n, &bv);
        break;
    case SCL_CODER_RANGE:
        rc = range_encode_block(&c->gp, sym, n, &bv);
        break;
    default:
        if (!model_freq) model_freq = tmp = fresh_model(c);
        rc = aec_encode_block((uint32_t)c->cfg.p0, (uint32_t)c->cfg.p1, (int)c->cfg.p2, model_freq, c->n_sym,
                              c->cfg.p3, sym, n, &bv);
        free(tmp);
    }
    if (rc == SCL_OK) {
        rc = pack_bits(bv.b, bv.n, out, out_cap);
        *out_bits = bv.n;
    }
    bv_free(&bv);
    return rc;
}

/* Decode one block from a stream of `nbits` bits starting at bit `bit_offset` of `in`
 * (trailing bits after the block are allowed, test_utils.py:97-105). */
int scl_oracle_decode_block(void *h, const uint8_t *in, uint64_t bit_offset, uint64_t nbits, uint64_t *model_freq,
                            uint8_t *out, uint64_t out_cap, uint64_t *n_out, uint64_t *bits_consumed) {
    oracle_ctx *c = (oracle_ctx *)h;
    int rc;
    uint8_t *shifted = NULL;
    const uint8_t *src = in;
    if (bit_offset) { /* realign so that the block starts at bit 0 */
        uint64_t nbytes = (nbits + 7) / 8;
        shifted = (uint8_t *)calloc(nbytes + 1, 1);
        for (uint64_t i = 0; i < nbits; ++i)
            if (get_bit(in, bit_offset + i)) shifted[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
        src = shifted;
    }
    uint64_t *tmp = NULL;
    switch (c->cfg.coder) {
    case SCL_CODER_RANS:
        rc = rans_decode_block(&c->rp, src, nbits, out, out_cap, n_out, bits_consumed);
        break;
    case SCL_CODER_TANS:
        rc = tans_decode_block(&c->tt, src, nbits, out, out_cap, n_out, bits_consumed);
        break;
    case SCL_CODER_RANGE:
        rc = range_decode_block(&c->gp, src, nbits, out, out_cap, n_out, bits_consumed);
        break;
    default:
        if (!model_freq) model_freq = tmp = fresh_model(c);
        rc = aec_decode_block((uint32_t)c->cfg.p0, (uint32_t)c->cfg.p1, (int)c->cfg.p2, model_freq, c->n_sym,
                              c->cfg.p3, src, nbits, out, out_cap, n_out, bits_consumed);
        free(tmp);
    }
    free(shifted);
    return rc;
}

/* Batched forms (independent blocks; each AEC block starts from a fresh copy of the
 * creation-time model).  OpenMP across blocks: this is what bench.py times as the CPU
 * baseline.  sym: [n_blocks, sym_stride]; sizes may be NULL (all = block_len);
 * out: [n_blocks, out_stride] left-aligned streams; out_bits/status: [n_blocks]. */
int scl_oracle_encode_batch(void *h, const uint8_t *sym, uint64_t sym_stride, const uint32_t *sizes,
                            uint32_t block_len, uint64_t n_blocks, uint8_t *out, uint64_t out_stride,
                            uint64_t *out_bits, int32_t *status, int n_threads) {
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
#endif
    for (int64_t b = 0; b < (int64_t)n_blocks; ++b) {
        uint64_t n = sizes ? sizes[b] : block_len;
        uint64_t nb = 0;
        int rc = scl_oracle_encode_block(h, sym + (uint64_t)b * sym_stride, n, NULL, out + (uint64_t)b * out_stride,
                                         out_stride, &nb);
        out_bits[b] = nb;
        if (status) status[b] = rc;
        bad |= rc;
    }
    return bad ? 1 : 0;
}

int scl_oracle_decode_batch(void *h, const uint8_t *in, const uint64_t *bit_offsets, const uint64_t *bit_lens,
                            uint64_t n_blocks, uint8_t *out, uint64_t out_stride, uint32_t *sizes,
                            uint64_t *bits_consumed, int32_t *status, int n_threads) {
    int bad = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
#endif
    for (int64_t b = 0; b < (int64_t)n_blocks; ++b) {
        uint64_t n = 0, used = 0;
        int rc = scl_oracle_decode_block(h, in, bit_offsets[b], bit_lens[b], NULL, out + (uint64_t)b * out_stride,
                                         out_stride, &n, &used);
        if (sizes) sizes[b] = (uint32_t)n;
        if (bits_consumed) bits_consumed[b] = used;
        if (status) status[b] = rc;
        bad |= rc;
    }
    return bad ? 1 : 0;
}

/* tANS lookup tables, exposed so tests can pin them against tANS.py:285-337 */
int scl_oracle_tans_tables(void *h, uint64_t *enc_table, uint64_t *enc_row, uint32_t *nbits_base, uint64_t *thresh,
                           uint32_t *dec_sym, uint64_t *dec_shrunk) {
    oracle_ctx *c = (oracle_ctx *)h;
    if (c->cfg.coder != SCL_CODER_TANS) return SCL_ERR_PARAM;
    uint64_t L = c->tt.rp.L;
    memcpy(enc_table, c->tt.enc_table, sizeof(uint64_t) * L);
    memcpy(enc_row, c->tt.enc_row, sizeof(uint64_t) * c->n_sym);
    memcpy(nbits_base, c->tt.nbits_base, sizeof(uint32_t) * c->n_sym);
    memcpy(thresh, c->tt.thresh, sizeof(uint64_t) * c->n_sym);
    memcpy(dec_sym, c->tt.dec_sym, sizeof(uint32_t) * L);
    memcpy(dec_shrunk, c->tt.dec_shrunk, sizeof(uint64_t) * L);
    return SCL_OK;
}

int scl_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
