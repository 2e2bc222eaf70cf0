// scl_pack.cuh -- contiguous ("packed") output: the device side of concatenating the per-block streams the
// way the reference does on the host (b"".join(encode_block(b).tobytes()), or EncodedBlockWriter.write_block's
// framed records, scl/core/encoded_stream.py:150-175).
//
// A LIFO coder only knows where a block's stream starts once the block is finished, and a block's place in
// the packed buffer depends on the sizes of every block before it.  Three pieces, device-only:
//   * pack_block_warp: one warp copies one finished stream, bit-granular source -> byte-aligned destination,
//     16 bytes per lane and step (five aligned source words, four funnel shifts, one 128-bit store);
//   * lookback_exclusive: the cross-CTA exclusive prefix of byte counts by decoupled look-back over one
//     64-bit word (flag : 2 | value : 62) per producer -- the `north_star`'s ballot / prefix-sum compaction at
//     the one place a lane-per-block layout has for it;
//   * the three small scan kernels behind scl_packed_offsets (callers that already hold bit lengths).
#pragma once
#include <cuda_runtime.h>

#include "scl_lane.cuh"

namespace scl {

// bytes block b occupies in the packed / framed layout
SCL_HD uint64_t packed_size(uint64_t nbits, bool framed) { return framed ? 4 + ((nbits + 3 + 7) >> 3) : (nbits + 7) >> 3; }
// first stream bit inside the block's record: 0 packed; framed = 32 header bits + 3-bit pad count + pad zeros
SCL_HD uint32_t packed_lead_bits(uint64_t nbits, bool framed) { return framed ? 32u + 3u + (uint32_t)((8 - (nbits + 3) % 8) % 8) : 0u; }

// ---- global-memory access flavours ---------------------------------------------------------------------
// NC = the source was written by an EARLIER kernel (read-only path); otherwise plain loads, which see this
// kernel's own stores once the writer and the reader have met at a barrier.
template <bool NC>
__device__ __forceinline__ uint32_t pk_ld32(const uint32_t *p) {
    return NC ? __ldg(p) : *p;
}
template <bool NC>
__device__ __forceinline__ uint32_t pk_ld8(const uint8_t *p) {
    return NC ? (uint32_t)__ldg(p) : (uint32_t)*p;
}

template <bool NC>
__device__ __forceinline__ uint32_t src_byte_at_bit(const uint8_t *src, uint64_t pos) {  // 8 bits starting at bit `pos`
    uint64_t by = pos >> 3;
    uint32_t sh = (uint32_t)(pos & 7);
    uint32_t v = (pk_ld8<NC>(src + by) << 8);
    if (sh) v |= pk_ld8<NC>(src + by + 1);
    return (v >> (8 - sh)) & 0xFFu;
}

// Byte i of a block's payload is bits [8i - lead, 8i - lead + 8) of its stream (lead = 0, or the framing's
// 3 + num_pad bits); bytes that straddle the lead or the end of the stream are composed bit by bit.
template <bool FRAMED, bool NC>
__device__ __forceinline__ uint32_t pack_payload_byte(const uint8_t *src, uint64_t off, uint64_t nbits, uint32_t num_pad, uint64_t lead,
                                                       uint64_t i) {
    uint32_t v = 0;
    if (8 * i >= lead && 8 * i - lead + 8 <= nbits) return src_byte_at_bit<NC>(src, off + 8 * i - lead);
    for (uint32_t k = 0; k < 8; ++k) {
        const uint64_t q = 8 * i + k;
        uint32_t bit = 0;
        if (FRAMED && q < 3) {
            bit = (num_pad >> (2 - q)) & 1u;
        } else if (q >= lead && q - lead < nbits) {
            const uint64_t p = off + (q - lead);
            bit = (pk_ld8<NC>(src + (p >> 3)) >> (7 - (p & 7))) & 1u;
        }
        v |= bit << (7 - k);
    }
    return v;
}

// One warp, one block: stream bits [off, off + nbits) of `src` -> record at `d` (packed: the stream, left-aligned,
// zero-padded to a byte == BitArray.tobytes(); FRAMED: [u32 BE payload bytes][3-bit pad count][pad zeros][stream],
// encoded_stream.py:22-46,93-103).  Once the destination is 16-byte aligned and all 128 bits are stream bits a
// chunk is five aligned 32-bit source words, four funnel shifts and one 128-bit store.  `src` must be 4-byte
// aligned and readable up to the next 4-byte boundary after the last stream bit.
template <bool FRAMED, bool NC>
__device__ __forceinline__ void pack_block_warp(const uint8_t *__restrict__ src, uint64_t off, uint64_t nbits, uint8_t *__restrict__ d,
                                                uint32_t lane) {
    const uint32_t *src32 = (const uint32_t *)src;
    const uint32_t num_pad = FRAMED ? (uint32_t)((8 - (nbits + 3) % 8) % 8) : 0u;
    const uint64_t lead = FRAMED ? 3 + num_pad : 0;
    const uint64_t payload_bytes = FRAMED ? (nbits + lead) >> 3 : (nbits + 7) >> 3;
    if (FRAMED) {
        if (lane < 4) d[lane] = (uint8_t)(payload_bytes >> (8 * (3 - lane)));  // u32 big-endian (HeaderHandler)
        d += 4;
    }
    // head: up to the first 16-byte aligned destination byte whose bits are all stream bits
    uint64_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (FRAMED && 8 * head < lead) head += 16;  // lead <= 10 bits
    if (head > payload_bytes) head = payload_bytes;
    // full chunks: 8 * (i0 + 16) - lead <= nbits
    const uint64_t n_chunks = (8 * head + 128 <= nbits + lead) ? ((nbits + lead - 8 * head) >> 7) : 0;
    for (uint64_t i = lane; i < head; i += 32) d[i] = (uint8_t)pack_payload_byte<FRAMED, NC>(src, off, nbits, num_pad, lead, i);
    for (uint64_t ch = lane; ch < n_chunks; ch += 32) {
        const uint64_t i0 = head + 16 * ch;
        const uint64_t S = off + 8 * i0 - lead;  // absolute source bit of the chunk's first bit
        const uint32_t *w = src32 + (S >> 5);
        const uint32_t sh = (uint32_t)(S & 31);
        uint32_t W[5];
#pragma unroll
        for (int j = 0; j < 4; ++j) W[j] = bswap32(pk_ld32<NC>(w + j));
        W[4] = sh ? bswap32(pk_ld32<NC>(w + 4)) : 0u;  // not needed (and possibly past the stream) when the chunk is word aligned
        uint4 o;
        o.x = bswap32(funnel_l(W[1], W[0], sh));
        o.y = bswap32(funnel_l(W[2], W[1], sh));
        o.z = bswap32(funnel_l(W[3], W[2], sh));
        o.w = bswap32(funnel_l(W[4], W[3], sh));
        *(uint4 *)(d + i0) = o;
    }
    for (uint64_t i = head + 16 * n_chunks + lane; i < payload_bytes; i += 32)
        d[i] = (uint8_t)pack_payload_byte<FRAMED, NC>(src, off, nbits, num_pad, lead, i);
}

// ---- the fused encoder's copy: scratch slots (RAW words) -> packed records ----------------------------------------
// Tuned for ONE warp that has to move a whole task (32 streams, ~100 KB) while the SM's other warps are busy coding,
// i.e. for few instructions and few exposed memory latencies per stream:
//   * the scratch slots hold the stream as raw 32-bit words (EncLaneV2T<true>): word i is the number whose big-endian
//     bytes are stream bytes 4i..4i+3, so neither the loads here nor the encoder's drains swap bytes; only the
//     finished 16-byte chunk is swapped into byte order;
//   * the source base is 16-byte aligned, so a chunk's five words come from two aligned 128-bit loads, and the word
//     offset inside the pair is the same for every chunk of a stream (template W0);
//   * up to U chunks per lane are in flight before the first is used; loads are predicated, addresses are one base
//     plus immediates; the loads of the bytes at both ends of the stream go out before the first batch and their
//     stores wait until it is in flight.
// Reads up to 16 bytes past the 16-byte group that holds the stream's last bit.
template <int W0>
__device__ __forceinline__ uint4 pack_chunk_from_pair_raw(const uint4 &A, const uint4 &B, uint32_t sh) {
    const uint32_t x[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
    uint4 o;
    o.x = bswap32(funnel_l(x[W0 + 1], x[W0 + 0], sh));
    o.y = bswap32(funnel_l(x[W0 + 2], x[W0 + 1], sh));
    o.z = bswap32(funnel_l(x[W0 + 3], x[W0 + 2], sh));
    o.w = bswap32(funnel_l(x[W0 + 4], x[W0 + 3], sh));
    return o;
}

// predicated 128-bit load: the destination keeps its (unused) old value when the predicate is off
__device__ __forceinline__ void ld_plain128_if(uint4 &r, const uint4 *p, uint32_t off_bytes, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.global.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)
                 : "l"((const uint8_t *)p + off_bytes), "r"((uint32_t)on)
                 : "memory");
}
__device__ __forceinline__ void st_plain128(uint8_t *p, const uint4 &v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int W0, int U, class Hook>
__device__ __forceinline__ void pack_chunks_a16(const uint4 *__restrict__ a, uint32_t sh, uint8_t *__restrict__ dd, uint32_t n_chunks, uint32_t lane,
                                                Hook first_batch_hook) {  // n_chunks >= 1
    bool first = true;
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 32 * U) {  // c0 is warp-uniform: whole slots of a batch are skipped together
        const uint4 *ab = a + c0 + lane;
        uint8_t *db = dd + 16 * (uint64_t)(c0 + lane);
        uint4 A[U], B[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            A[u] = make_uint4(0, 0, 0, 0);
            B[u] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c0 + 32 * u < n_chunks) {  // uniform
                const bool on = c0 + 32 * u + lane < n_chunks;
                ld_plain128_if(A[u], ab, 512 * u, on);
                ld_plain128_if(B[u], ab, 512 * u + 16, on);
            }
        }
        if (first) {
            first_batch_hook();
            first = false;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c0 + 32 * u < n_chunks) {
                if (c0 + 32 * u + lane < n_chunks) st_plain128(db + 512 * u, pack_chunk_from_pair_raw<W0>(A[u], B[u], sh));
            }
        }
    }
}

// 8 stream bits starting at stream position `pos` (which may be negative, or reach past the end: those bits read
// as 0), straight-line: two aligned RAW words, one funnel shift, two masks.  Needs the word holding bit off + pos and
// the one after it to be readable (a slot keeps >= 32 spare bits in front of its stream, the buffer 16 bytes behind).
__device__ __forceinline__ uint32_t stream_byte_masked_raw(const uint8_t *__restrict__ src, uint64_t off, int32_t pos, uint32_t nbits) {
    const uint64_t S = off + (int64_t)pos;
    const uint32_t *w = (const uint32_t *)src + (S >> 5);
    const uint32_t v = funnel_l(w[1], w[0], (uint32_t)S & 31u) >> 24;
    uint32_t keep = 0xFFu;
    if (pos < 0) keep = pos <= -8 ? 0u : (0xFFu >> (uint32_t)(-pos));
    const int32_t r = (int32_t)nbits - pos;  // stream bits available from `pos` on
    if (r < 8) keep &= r <= 0 ? 0u : ~(0xFFu >> (uint32_t)r);
    return v & keep;
}

// nbits < 2^31 (the fast encoders bound block_len * 16 bits by that): everything inside a stream is 32-bit arithmetic
template <bool FRAMED>
__device__ __forceinline__ void pack_block_warp_a16(const uint8_t *__restrict__ src, uint64_t off, uint32_t nbits, uint8_t *__restrict__ d,
                                                    uint32_t lane) {
    constexpr int U = 4;
    const uint32_t num_pad = FRAMED ? ((8u - (nbits + 3u) % 8u) % 8u) : 0u;
    const uint32_t lead = FRAMED ? 3u + num_pad : 0u;
    const uint32_t payload_bytes = FRAMED ? (nbits + lead) >> 3 : (nbits + 7u) >> 3;
    if (FRAMED) {
        if (lane < 4) d[lane] = (uint8_t)(payload_bytes >> (8 * (3 - lane)));
        d += 4;
    }
    uint32_t head = (16u - ((uint32_t)(uintptr_t)d & 15u)) & 15u;
    if (FRAMED && 8 * head < lead) head += 16;
    if (head > payload_bytes) head = payload_bytes;
    const uint32_t n_chunks = (8 * head + 128 <= nbits + lead) ? ((nbits + lead - 8 * head) >> 7) : 0u;
    const uint32_t tail0 = head + 16 * n_chunks;  // the bytes after the last whole chunk: < 32 (16 left over + a partial chunk)
    // The edge bytes are independent of the chunks: their loads go out first, their stores wait until the first chunk
    // batch is in flight.  Packed records have head <= 15 and at most 16 tail bytes: ONE byte per lane covers both
    // ends (lanes 0-15 the head, lanes 16-31 the tail); framed records (head <= 31) take two.
    uint32_t i0, i1 = 0xFFFFFFFFu;
    if (!FRAMED) {
        i0 = lane < 16 ? (lane < head ? lane : 0xFFFFFFFFu) : (tail0 + lane - 16 < payload_bytes ? tail0 + lane - 16 : 0xFFFFFFFFu);
    } else {
        i0 = lane < head ? lane : 0xFFFFFFFFu;
        i1 = tail0 + lane < payload_bytes ? tail0 + lane : 0xFFFFFFFFu;
    }
    uint32_t v0 = 0, v1 = 0;
    if (i0 != 0xFFFFFFFFu) v0 = stream_byte_masked_raw(src, off, (int32_t)(8 * i0) - (int32_t)lead, nbits) | ((FRAMED && i0 == 0) ? (num_pad << 5) : 0u);
    if (FRAMED && i1 != 0xFFFFFFFFu) v1 = stream_byte_masked_raw(src, off, (int32_t)(8 * i1) - (int32_t)lead, nbits) | (i1 == 0 ? (num_pad << 5) : 0u);
    auto edges = [&]() {
        if (i0 != 0xFFFFFFFFu) d[i0] = (uint8_t)v0;
        if (FRAMED && i1 != 0xFFFFFFFFu) d[i1] = (uint8_t)v1;
    };
    if (n_chunks == 0) {
        edges();
        return;
    }
    const uint64_t S0 = off + 8 * head - lead;  // source bit of chunk 0's first bit; chunk ch starts 128 * ch bits later
    const uint4 *a = (const uint4 *)src + (S0 >> 7);
    const uint32_t sh = (uint32_t)S0 & 31u;
    switch ((uint32_t)(S0 >> 5) & 3u) {  // the same for every lane: no divergence
    case 0: pack_chunks_a16<0, U>(a, sh, d + head, n_chunks, lane, edges); break;
    case 1: pack_chunks_a16<1, U>(a, sh, d + head, n_chunks, lane, edges); break;
    case 2: pack_chunks_a16<2, U>(a, sh, d + head, n_chunks, lane, edges); break;
    default: pack_chunks_a16<3, U>(a, sh, d + head, n_chunks, lane, edges); break;
    }
}

// ---- the same copy out of a shared-memory staging ring (the fused encoder's dedicated copy warps) ---------------
// A copy warp that loads into registers exposes one memory latency per batch of chunks (~2200 cycles per 4 KiB:
// profiles/r2q_copy_only.jsonl).  Here the source arrives by bulk async copy (TMA), several pieces ahead and across
// stream boundaries, so the warp only ever touches shared memory and the destination.
//   region of a stream  = the 16-byte granules from the one holding its first needed bit (g0) to two past the one
//                         holding its last bit;
//   piece p of a region = granules [PC p, PC p + PC + 5), PC = 32 or 64 chunks per piece: chunk c needs granules
//                         c + dlt and c + dlt + 1 (dlt = 0..2, the distance from g0 to chunk 0's granule), and the last
//                         piece also holds the granules the tail bytes are cut from.  Pieces overlap by 5 granules (80
//                         bytes read twice, from L2).
constexpr uint32_t kCopyOverlapBytes = 5 * 16;  // a stage holds its piece's chunks' bytes + 5 granules

struct StreamGeo {  // where a stream of `nbits` goes when its record's payload starts at address d (only d % 16 matters)
    uint32_t num_pad, lead, payload_bytes, head, n_chunks, tail0;
};
template <bool FRAMED>
__device__ __forceinline__ StreamGeo stream_geo(uint32_t nbits, uint32_t d_low4) {  // as pack_block_warp_a16
    StreamGeo g;
    g.num_pad = FRAMED ? ((8u - (nbits + 3u) % 8u) % 8u) : 0u;
    g.lead = FRAMED ? 3u + g.num_pad : 0u;
    g.payload_bytes = FRAMED ? (nbits + g.lead) >> 3 : (nbits + 7u) >> 3;
    g.head = (16u - d_low4) & 15u;
    if (FRAMED && 8 * g.head < g.lead) g.head += 16;
    if (g.head > g.payload_bytes) g.head = g.payload_bytes;
    g.n_chunks = (8 * g.head + 128 <= nbits + g.lead) ? ((nbits + g.lead - 8 * g.head) >> 7) : 0u;
    g.tail0 = g.head + 16 * g.n_chunks;
    return g;
}

// stream_byte_masked_raw out of a staged piece: `sb` = shared address of the piece's first granule, S = distance in
// bits from there to stream position `pos`
__device__ __forceinline__ uint32_t stream_byte_masked_smem(uint32_t sb, uint32_t S, int32_t pos, uint32_t nbits) {
    const uint32_t wa = sb + ((S >> 5) << 2);
    uint32_t w0, w1;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(wa));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(wa + 4));
    const uint32_t v = funnel_l(w1, w0, S & 31u) >> 24;
    uint32_t keep = 0xFFu;
    if (pos < 0) keep = pos <= -8 ? 0u : (0xFFu >> (uint32_t)(-pos));
    const int32_t r = (int32_t)nbits - pos;
    if (r < 8) keep &= r <= 0 ? 0u : ~(0xFFu >> (uint32_t)r);
    return v & keep;
}

__device__ __forceinline__ uint4 lds_plain128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
    return r;
}

// ---- decoupled look-back ----------------------------------------------------------------------------------
// state[i] = flag << 62 | value: flag 0 = not there yet, 1 = the producer's own total, 2 = inclusive prefix up to
// and including producer i.  One 64-bit word carries flag and value together, so relaxed accesses suffice.
constexpr uint64_t kLbAgg = 1ull << 62, kLbPrefix = 2ull << 62, kLbMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint64_t *p, uint64_t v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += u;
    }
    return v;
}

// Sum of the values of producers 0 .. g-1; the whole (converged) warp calls it.  Each step inspects 32
// predecessors: it stops at the nearest published prefix, otherwise adds 32 totals and moves on.  Spins
// (with back-off) only while a predecessor has published nothing at all.
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t *state, int64_t g, uint32_t lane) {
    uint64_t excl = 0;
    int64_t idx = g - 1 - (int64_t)lane;
    while (true) {
        uint64_t v;
        uint32_t spins = 0;
        while (true) {
            v = idx >= 0 ? ld_relaxed_gpu(state + idx) : kLbPrefix;  // before producer 0: the empty prefix
            if (!__any_sync(0xffffffffu, (v >> 62) == 0)) break;
            __nanosleep(spins < 8 ? 40 : 400);
            ++spins;
        }
        const uint32_t pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const uint64_t val = v & kLbMask;
        if (pm) {
            const uint32_t first = (uint32_t)__ffs((int)pm) - 1;
            return excl + warp_sum_u64(lane <= first ? val : 0);
        }
        excl += warp_sum_u64(val);
        idx -= 32;
    }
}

// ---- scl_packed_offsets: exclusive scan of packed sizes in three small launches ------------------------------
// (1) every CTA leaves the total of its kScanTile items in out[first item of the tile] (scratch use of the
// output array), (2) one CTA turns those totals into exclusive tile offsets and writes the grand total to
// out[n], (3) every CTA scans its tile from its offset.  No workspace, no atomics.
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t scan_item(const uint64_t *bit_len, const uint32_t *status, uint64_t i, uint64_t n, bool framed) {
    if (i >= n) return 0;
    if (status && status[i] != SCL_ST_OK) return 0;  // a failed block takes no room
    return packed_size(bit_len[i], framed);
}

__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t *total) {  // kScanThreads threads
    __shared__ uint64_t s_warp[kScanThreads / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t incl = warp_incl_scan_u64(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t base = 0, tot = 0;
#pragma unroll
    for (uint32_t w = 0; w < kScanThreads / 32; ++w) {
        const uint64_t t = s_warp[w];
        if (w < warp) base += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_totals_kernel(const uint64_t *__restrict__ bit_len, const uint32_t *__restrict__ status,
                                                                         uint64_t n, uint32_t framed, uint64_t *__restrict__ out) {
    const uint64_t first = (uint64_t)blockIdx.x * kScanTile;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) v += scan_item(bit_len, status, first + (uint64_t)threadIdx.x * kScanItems + j, n, framed != 0);
    uint64_t tot;
    block_excl_scan_u64(v, &tot);
    if (threadIdx.x == 0) out[first] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(uint64_t n, uint64_t n_tiles, uint64_t *__restrict__ out) {
    uint64_t carry = 0;
    for (uint64_t t0 = 0; t0 < n_tiles; t0 += kScanThreads) {
        const uint64_t t = t0 + threadIdx.x;
        const uint64_t v = t < n_tiles ? out[t * kScanTile] : 0;
        uint64_t tot;
        const uint64_t ex = block_excl_scan_u64(v, &tot);
        if (t < n_tiles) out[t * kScanTile] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_final_kernel(const uint64_t *__restrict__ bit_len, const uint32_t *__restrict__ status,
                                                                        uint64_t n, uint32_t framed, uint64_t *__restrict__ out_bytes,
                                                                        uint64_t *__restrict__ out_bits) {
    const uint64_t first = (uint64_t)blockIdx.x * kScanTile;
    const uint64_t base = out_bytes[first];  // read by every thread before anybody overwrites it
    __syncthreads();
    uint64_t item[kScanItems], v = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        item[j] = scan_item(bit_len, status, first + (uint64_t)threadIdx.x * kScanItems + j, n, framed != 0);
        v += item[j];
    }
    uint64_t tot;
    uint64_t pos = base + block_excl_scan_u64(v, &tot);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        const uint64_t i = first + (uint64_t)threadIdx.x * kScanItems + j;
        if (i < n) {
            out_bytes[i] = pos;
            if (out_bits) out_bits[i] = 8 * pos + packed_lead_bits(bit_len[i], framed != 0);
            pos += item[j];
        }
    }
}

}  // namespace scl
