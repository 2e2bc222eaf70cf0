// scl_kernels.cu -- sm_100a kernels and the C-ABI (include/scl_b200.h).
//
// Execution model: one warp lane == one DataBlock.  A CTA first stages the coder's
// lookup tables into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier), then every
// lane runs the reference's per-block state machine from scl_lane.cuh on its own block.
// There is no dense contraction anywhere on this path, so no tensor-core code: the kernels are
// integer ALU + LSU work bounded by HBM traffic and instruction issue.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <type_traits>
#include <vector>

#include <cuda.h>  // CUtensorMap (types only; the encoder entry point is fetched at run time)

#include "scl_aec.cuh"
#include "scl_fast.cuh"
#include "scl_lane.cuh"
#include "scl_pack.cuh"
#include "scl_range.cuh"
#include "scl_tables.hpp"

namespace scl {

// ------------------------------------------------------------------------------------------------
// TMA table staging: one elected thread issues cp.async.bulk (UBLKCP) global -> shared and the
// whole CTA waits on the mbarrier's transaction count.  bytes % 16 == 0, both sides 16-B aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *mbar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void tma_expect(uint64_t *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    } while (!done);
}
// stage one table; every thread of the CTA must call this
__device__ __forceinline__ void stage_table(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar) {
    mbar_init(mbar);
    if (threadIdx.x == 0) {
        tma_expect(mbar, bytes);
        // a single bulk copy may move at most 2^20-16 bytes; tables here are <= 64 KiB
        tma_bulk_g2s(smem_dst, gsrc, bytes, mbar);
    }
    mbar_wait(mbar, 0);
}

struct BlockIo {  // per-launch I/O description shared by all encode kernels
    const uint8_t *sym;
    uint64_t sym_stride;
    const uint32_t *sizes;
    uint32_t block_len;
    uint64_t n_blocks;
    uint8_t *out;
    uint64_t out_stride;
    uint64_t *bit_off;
    uint64_t *bit_len;
    uint32_t *status;
    uint64_t block0;  // second-generation rANS / tANS launches of a split batch: index of this launch's first block in `out`
                      // (sym, bit_off, bit_len, status already point at that block; slots and reported offsets are global)
};
struct DecodeIo {
    const uint8_t *in;
    uint64_t in_bytes;
    const uint64_t *bit_off;
    const uint64_t *bit_len;
    uint64_t n_blocks;
    uint8_t *sym;
    uint64_t sym_stride;
    uint32_t *sizes;
    uint64_t *consumed;
    uint32_t *status;
};

__device__ __forceinline__ uint64_t avail_bits_of(const DecodeIo &io, uint64_t b, uint64_t off) {
    if (io.bit_len) return io.bit_len[b];
    uint64_t tot = io.in_bytes * 8;
    return tot > off ? tot - off : 0;
}

constexpr int kThreads = 128;

// ------------------------------------------------------------------------------------------------
// rANS kernels
// ------------------------------------------------------------------------------------------------
template <bool CHECK>
__global__ void __launch_bounds__(kThreads) rans32_encode_kernel(const RansEnc32 *__restrict__ g_tab, RansConst c, BlockIo io) {
    __shared__ RansEnc32 s_tab[256];
    __shared__ uint64_t mbar;
    stage_table(s_tab, g_tab, sizeof(s_tab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    LifoBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint64_t bits = 0;
    uint32_t st = rans32_encode_lane<CHECK>(s_tab, c, io.sym + b * io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = (b + 1) * io.out_stride * 8 - bits;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kThreads) rans32_decode_kernel(const RansDec32 *__restrict__ g_lut, uint32_t lut_bytes, RansConst c,
                                                                 DecodeIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint64_t mbar;
    RansDec32 *s_lut = (RansDec32 *)s_dyn;
    stage_table(s_lut, g_lut, lut_bytes, &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint64_t used = 0;
    uint32_t st = rans32_decode_lane(s_lut, c, r, io.sym + b * io.sym_stride, io.sym_stride, size, used);
    if (st == SCL_ST_OK && used > avail_bits_of(io, b, off)) st = SCL_ST_TRUNCATED;
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kThreads) rans64_encode_kernel(const RansGeneric *__restrict__ g_tab, RansConst c, BlockIo io) {
    __shared__ RansGeneric s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(RansGeneric), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    LifoBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint64_t bits = 0;
    uint32_t st = rans64_encode_lane(s_tab, c, io.sym + b * io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = (b + 1) * io.out_stride * 8 - bits;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kThreads) rans64_decode_kernel(const RansGeneric *__restrict__ g_tab, RansConst c, DecodeIo io) {
    __shared__ RansGeneric s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(RansGeneric), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint64_t used = 0;
    const uint64_t avail = avail_bits_of(io, b, off);
    uint32_t st = rans64_decode_lane(s_tab, c, r, avail, io.sym + b * io.sym_stride, io.sym_stride, size, used);
    if (st == SCL_ST_OK && used > avail) st = SCL_ST_TRUNCATED;
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// rANS fast path, second generation (scl_fast.cuh).  Persistent kernels: one CTA per SM, `W`
// warps per CTA (host-chosen so that tasks / (SMs * W) is just below an integer: the work is
// one sequential chain per block, so rounds cannot be split and a partial last round idles SMs).
// A task = 32 consecutive blocks = one warp.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTileCols = 64;                  // bytes of each block's row per TMA tile
constexpr uint32_t kTileBytes = 32 * kTileCols;     // 32 rows (one per lane)
constexpr uint32_t kTileStages = 2;
constexpr uint32_t kEncWarpSmem = kTileStages * kTileBytes + kEncRingWords * 128;  // tiles + ring, per warp
constexpr uint32_t kEncTabBytes = 256 * kEncTabCopies * sizeof(RansEnc32);         // 32 KiB
constexpr uint32_t kDecWarpSmem = (kDecRingWords + 1) * 128;                       // ring + wrap duplicate
constexpr uint32_t kMaxWarps = 28;

__device__ __forceinline__ void tma_tile_2d(void *smem_dst, const CUtensorMap *tmap, int32_t c0, int32_t c1, uint64_t *mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(mbar))
        : "memory");
}

// smem layout: [tiles+rings per warp ...][table 32 KiB][mbarriers]
// KIND 0 = rANS (arithmetic step), KIND 1 = tANS (table step; g_tab2 = enc_table, staged after the
// replicated per-symbol table)
//
// PACKED: the kernel also produces the contiguous output (EncodedBlockWriter's job in the reference,
// encoded_stream.py:150-175).  A LIFO stream's start is only known when its block is finished, and its place in
// the packed buffer depends on every earlier block's size, so a task still encodes into its scratch slots;
// then (i) the warp sums its 32 record sizes (__reduce_add_sync), (ii) the last warp of a CTA to finish a round
// adds up the CTA's tasks -- they are consecutive -- and resolves the cross-CTA exclusive prefix by decoupled
// look-back over one word per (round, CTA) (scl_pack.cuh), (iii) the streams are copied to their final byte
// offsets.  The CTA is WARP-SPECIALISED: the coding warps never copy while they have symbols left; kCopyWarps extra
// warps (the CTA has room for 32, the coder uses at most 28) do nothing else, taking resolved tasks from a ticket
// counter.  A copy warp is one dependent instruction chain among 28 coding warps that keep the issue slots and the
// L1/shared pipe ~75 % busy (profiles/r2w): what it costs is instructions, not bytes.  So it stages its source by bulk
// async copies (TMA) into a small shared-memory ring carved from what the coder leaves free -- pieces of up to 2 KiB,
// requested ahead across stream boundaries -- and its loop is: wait for a piece, LDS / funnel shift / STG per 16 bytes,
// hand the stage back (packed_copy_task_ring).  A task's place in the output is handed over through its first block's
// byte_off entry (one word per task, never reused), so the coder never waits for the pool: what the pool has not moved
// when a coding warp runs out of tasks is moved by that warp too (through registers: pack_block_warp_a16), i.e. the
// rest is copied by the whole CTA at memory speed at the end.
// Depends on the in-order dispatch of CTAs (a CTA only waits for lower-numbered CTAs of the same round, or
// for earlier rounds), like every single-pass scan.
constexpr uint32_t kCopyWarps = 4;
struct PackedOut {
    uint8_t *dst;            // packed destination
    uint64_t dst_bytes;      // its capacity
    uint64_t *byte_off;      // [n_blocks + 1] record offsets, [n_blocks] = total
    uint64_t *cta_state;     // [rounds * gridDim.x] look-back words, zeroed before the launch
    uint32_t framed;
    uint64_t g_base;         // look-back index of this launch's (round 0, CTA 0): a split batch continues the numbering
    uint32_t copy_warps;     // dedicated copy warps of the CTA (kCopyWarps)
    uint32_t copy_stages;    // stages of each copy warp's staging ring (from the shared memory the coder leaves free); 0 = copy through registers
    uint32_t copy_piece_bytes;  // 512, 1024 or 2048: stream bytes per stage
    uint32_t helper_ring;    // 1 = a coding warp that has run out of tasks copies through a ring in its idle tile buffers
    uint64_t *trace;         // scl_coder_debug_trace: NULL, or [gridDim.x][32 warps][kTraceWords] timestamps (tools/trace_packed.py)
};
constexpr uint32_t kTraceWords = 40;  // per warp: [0] start, [1 + r] end of coding round r (r < 19), [20] tasks copied, [21] first copy
                                      // start, [22] last copy end, [23] ns spent copying, [24] ns spent waiting for a resolved round
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct PackCtl {  // per CTA, shared memory
    unsigned long long warp_tot[2][kMaxWarps];   // [round & 1] record bytes of each coding warp's task
    uint32_t arrive[2];                          // [round & 1] coding warps that have finished the round
    uint32_t resolved;                           // rounds whose offsets are known (monotonic)
    uint32_t copy_ticket;                        // next task to copy: ticket T = round * W + warp slot (monotonic)
    uint32_t copy_done;                          // tasks copied so far (monotonic; read by nobody but a debugger)
};

template <bool FRAMED>
__device__ __forceinline__ void packed_copy_task(const BlockIo &io, const PackedOut &po, uint64_t task, uint64_t base, uint32_t lane) {
    const uint64_t b = task * 32 + lane;
    const bool active = b < io.n_blocks;
    uint32_t bits = 0, nb = 0;  // a failed or absent block: no bits, no room
    if (active && io.status[b] == SCL_ST_OK) {
        bits = (uint32_t)io.bit_len[b];  // < 2^31 on this path
        nb = (uint32_t)packed_size(bits, FRAMED);
    }
    uint32_t incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += u;
    }
    if (active) {
        const uint64_t at = base + (incl - nb);
        po.byte_off[b] = at;
        io.bit_off[b] = 8 * at + packed_lead_bits(bits, FRAMED);
        if (nb && at + nb > po.dst_bytes) {
            io.status[b] = SCL_ST_OVERFLOW;
            bits = 0;  // dropped (nb keeps its place in the layout)
        }
    }
    // Stream l of the task: ends at its slot's end in the scratch buffer, goes to `at`; both advance uniformly, so
    // one shuffle per stream (its bit count) is all the lanes exchange.  The NEXT stream is requested into L2 while
    // this one is copied (one line per lane, no registers held): the slots were written ~400 MB of traffic ago.
    const uint64_t stride_bits = io.out_stride * 8;
    const uint8_t *const src = io.out;  // (registers, not the caller's frame: see packed_copy_task_ring)
    uint8_t *const dst = po.dst;
    uint64_t slot_end = (io.block0 + task * 32 + 1) * stride_bits, at = base;
    uint32_t bits_l = __shfl_sync(0xffffffffu, bits, 0), nb_l = __shfl_sync(0xffffffffu, nb, 0);
    for (uint32_t l = 0; l < 32; ++l) {
        const uint32_t bits_n = __shfl_sync(0xffffffffu, bits, (l + 1) & 31), nb_n = __shfl_sync(0xffffffffu, nb, (l + 1) & 31);
        if (l + 1 < 32 && bits_n) {
            const uint8_t *p0 = src + (((slot_end + stride_bits - bits_n) >> 3) & ~127ull);
            for (uint32_t o = lane * 128; o < (bits_n >> 3) + 160; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
        }
        if (bits_l) pack_block_warp_a16<FRAMED>(src, slot_end - bits_l, bits_l, dst + at, lane);
        at += nb_l;
        slot_end += stride_bits;
        bits_l = bits_n;
        nb_l = nb_n;
    }
}

// The dedicated copy warps' form of packed_copy_task: the streams arrive in a per-warp shared-memory ring by bulk async
// copies (scl_pack.cuh: regions, pieces), issued by lane 0 `stages` pieces ahead of the piece being moved -- across
// stream boundaries, so the warp does not wait for memory between streams either.  Same output, byte for byte.
// One warp among 28 coding warps runs a dependent instruction chain: what it costs is instructions per piece, so the
// piece loop is kept to the wait, the chunks (one or two per lane), and the refill of the stage just emptied.
struct CopyRing {          // identical in every lane of the warp
    uint32_t buf;          // shared address of stage 0; stage i at buf + i * (piece_bytes + kCopyOverlapBytes)
    uint32_t bars;         // shared address of mbarrier 0 (right behind the stages); mbarrier i at bars + 8 i
    uint32_t tab;          // shared address of the per-task stream table: 32 x {bits, pieces, region start, region bytes}
    uint32_t piece_bytes;  // 512, 1024 or 2048: 32, 64 or 128 chunks per piece; 0 = this warp has no ring (copies through registers)
    uint32_t stage, bar;   // where the next piece lands / its mbarrier (shared addresses)
    uint32_t par;          // that mbarrier's phase parity
    uint32_t wait_cycles;  // tracing only: clock cycles spent waiting for pieces to land
};
__host__ __device__ constexpr uint32_t copy_ring_bytes(uint32_t stages, uint32_t piece_bytes) {
    return stages ? stages * (piece_bytes + kCopyOverlapBytes) + ((stages * 8 + 15) & ~15u) + 512 : 0;
}
// mem: 16-byte aligned, copy_ring_bytes(); `init` = also initialise the mbarriers (once, before the CTA's first barrier)
__device__ __forceinline__ CopyRing copy_ring_at(uint8_t *mem, uint32_t stages, uint32_t piece_bytes, bool init) {
    CopyRing R;
    R.buf = smem_u32(mem);
    R.bars = R.buf + stages * (piece_bytes + kCopyOverlapBytes);
    R.tab = R.bars + ((stages * 8 + 15) & ~15u);
    R.piece_bytes = stages ? piece_bytes : 0;
    R.stage = R.buf;
    R.bar = R.bars;
    R.par = 0;
    R.wait_cycles = 0;
    if (init) {
        for (uint32_t i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(R.bars + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    return R;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}

template <bool FRAMED>
__device__ __forceinline__ void packed_copy_task_ring(const BlockIo &io, const PackedOut &po, uint64_t task, uint64_t base, uint32_t lane, CopyRing &R) {
    // (io / po live in the caller's frame and every shared-memory / mbarrier asm here is a memory barrier to the compiler:
    // what the loops need is copied to registers once, or it is re-loaded from local memory after every such statement)
    const uint64_t out_stride = io.out_stride;
    uint8_t *const dst = po.dst;
    const bool tracing = po.trace != nullptr;
    const uint32_t PB = R.piece_bytes, SB = PB + kCopyOverlapBytes, PC = PB >> 4;
    const uint32_t pb_log2 = 31u - (uint32_t)__clz(PB), pc_log2 = pb_log2 - 4;  // PB is a power of two: no divisions in the loops
    {
        // lane l describes stream l (the block's record and where its source bytes lie) in the warp's table
        const uint64_t b = task * 32 + lane;
        const bool active = b < io.n_blocks;
        uint32_t bits = 0, nb = 0;
        if (active && io.status[b] == SCL_ST_OK) {
            bits = (uint32_t)io.bit_len[b];
            nb = (uint32_t)packed_size(bits, FRAMED);
        }
        uint32_t incl = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        const uint64_t at_mine = base + (incl - nb);
        if (active) {
            po.byte_off[b] = at_mine;
            io.bit_off[b] = 8 * at_mine + packed_lead_bits(bits, FRAMED);
            if (nb && at_mine + nb > po.dst_bytes) {
                io.status[b] = SCL_ST_OVERFLOW;
                bits = 0x80000000u | nb;  // dropped: no bits to move, but the record keeps its place in the layout
            }
        }
        uint32_t np = 0, rs = 0, rlen = 0;
        if (bits && !(bits >> 31)) {
            const StreamGeo g = stream_geo<FRAMED>(bits, (uint32_t)(uintptr_t)(dst + at_mine + (FRAMED ? 4 : 0)) & 15u);
            const uint32_t off = (uint32_t)(out_stride * 8) - bits;
            const uint32_t a = (off + 8 * g.head - g.lead) >> 7;
            rs = ((off - g.lead) >> 7) * 16;
            uint32_t re = (a + g.n_chunks + 3) * 16;
            if (re > (uint32_t)out_stride + 16) re = (uint32_t)out_stride + 16;
            rlen = re - rs;
            np = g.n_chunks ? (g.n_chunks + PC - 1) >> pc_log2 : 1u;
        }
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(R.tab + 16 * lane), "r"(bits), "r"(np), "r"(rs), "r"(rlen) : "memory");
        __syncwarp();
    }
    // producer cursor (warp-uniform): stream pl, p_np pieces of it still to request, the next one at p_src, p_left bytes to
    // the region's end.  A piece always goes into the stage the consumer has just emptied (or, before the first piece is
    // consumed, into the stages in order), so the producer keeps no stage of its own.
    const uint8_t *task_src = io.out + (io.block0 + task * 32) * out_stride;
    uint32_t pl = 0xFFFFFFFFu, p_np = 0, p_left = 0;
    const uint8_t *p_src = task_src;
    auto next_stream = [&]() {
        while (p_np == 0 && pl + 1 < 32) {
            ++pl;
            const uint4 e = lds_plain128(R.tab + 16 * pl);
            p_np = e.y;
            p_src = task_src + (uint64_t)pl * out_stride + e.z;
            p_left = e.w;
        }
    };
    auto issue = [&](uint32_t stage, uint32_t bar) {
        if (p_np == 0) return;
        const uint32_t nbytes = p_left < SB ? p_left : SB;
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage), "l"(p_src), "r"(nbytes), "r"(bar)
                         : "memory");
        }
        p_src += PB;
        p_left -= PB;
        if (--p_np == 0) next_stream();
    };
    next_stream();
    {
        uint32_t st = R.stage, br = R.bar;
        for (; p_np; ) {  // fill the ring
            issue(st, br);
            st += SB;
            br += 8;
            if (st == R.bars) {
                st = R.buf;
                br = R.bars;
            }
            if (st == R.stage) break;
        }
    }
    uint64_t at = base;
    for (uint32_t l = 0; l < 32; ++l) {
        uint32_t bits_l;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bits_l) : "r"(R.tab + 16 * l));
        if (bits_l >> 31) {  // a dropped record: its place stays
            at += bits_l & 0x7FFFFFFFu;
            continue;
        }
        if (bits_l == 0) continue;
        uint8_t *d = dst + at;
        at += (uint32_t)packed_size(bits_l, FRAMED);
        const StreamGeo g = stream_geo<FRAMED>(bits_l, (uint32_t)(uintptr_t)(d + (FRAMED ? 4 : 0)) & 15u);
        if (FRAMED) {
            if (lane < 4) d[lane] = (uint8_t)(g.payload_bytes >> (8 * (3 - lane)));
            d += 4;
        }
        const uint32_t off = (uint32_t)(out_stride * 8) - bits_l;
        const uint32_t g0 = (off - g.lead) >> 7, S0 = off + 8 * g.head - g.lead, offr = off - 128 * g0;
        const uint32_t lofs = 16 * (lane + (S0 >> 7) - g0), sh = S0 & 31u;
        const uint32_t last_piece = g.n_chunks ? (g.n_chunks - 1) >> pc_log2 : 0u;
        if (tracing) {
            const uint32_t c0 = (uint32_t)clock64();
            mbar_wait_s(R.bar, R.par);
            R.wait_cycles += (uint32_t)clock64() - c0;
        } else {
            mbar_wait_s(R.bar, R.par);  // the stream's first piece
        }
        if (lane < g.head) {
            const int32_t pos = (int32_t)(8 * lane) - (int32_t)g.lead;
            d[lane] = (uint8_t)(stream_byte_masked_smem(R.stage, (uint32_t)((int32_t)offr + pos), pos, bits_l) | ((FRAMED && lane == 0) ? (g.num_pad << 5) : 0u));
        }
        // this lane's tail byte, all worked out except for the two words it is cut from: bit offset inside the last piece
        // (or none), keep-mask | bits to OR in << 8, and its address relative to the lane's chunk pointer in that piece
        uint32_t t_S = 0xFFFFFFFFu, t_ko = 0;
        int32_t t_ofs = 0;
        {
            const uint32_t i = g.tail0 + lane;
            if (i < g.payload_bytes) {
                const int32_t pos = (int32_t)(8 * i) - (int32_t)g.lead;
                t_S = (uint32_t)((int32_t)offr + pos) - (last_piece << (pb_log2 + 3));
                uint32_t keep = 0xFFu;
                if (pos < 0) keep = pos <= -8 ? 0u : (0xFFu >> (uint32_t)(-pos));
                const int32_t r = (int32_t)bits_l - pos;
                if (r < 8) keep &= r <= 0 ? 0u : ~(0xFFu >> (uint32_t)r);
                t_ko = keep | ((FRAMED && i == 0) ? (g.num_pad << 13) : 0u);
                t_ofs = (int32_t)i - (int32_t)(g.head + 16 * lane + (last_piece << pb_log2));
            }
        }
        uint8_t *dp = d + g.head + 16 * lane;
        uint32_t rem = g.n_chunks;  // chunks from the current piece on
        auto pieces = [&](auto w0_tag) {
            constexpr int W0 = decltype(w0_tag)::value;
            while (true) {
                const uint32_t sa = R.stage + lofs;
                // (the loads are unconditional: anything inside the ring may be read, only the stores are guarded)
                {
                    const uint4 A0 = lds_plain128(sa), B0 = lds_plain128(sa + 16);
                    if (PB >= 1024) {
                        const uint4 A1 = lds_plain128(sa + 512), B1 = lds_plain128(sa + 528);
                        if (lane < rem) st_plain128(dp, pack_chunk_from_pair_raw<W0>(A0, B0, sh));
                        if (lane + 32 < rem) st_plain128(dp + 512, pack_chunk_from_pair_raw<W0>(A1, B1, sh));
                    } else {
                        if (lane < rem) st_plain128(dp, pack_chunk_from_pair_raw<W0>(A0, B0, sh));
                    }
                }
                if (PB == 2048 && rem > 64) {
                    const uint4 A2 = lds_plain128(sa + 1024), B2 = lds_plain128(sa + 1040);
                    const uint4 A3 = lds_plain128(sa + 1536), B3 = lds_plain128(sa + 1552);
                    if (lane + 64 < rem) st_plain128(dp + 1024, pack_chunk_from_pair_raw<W0>(A2, B2, sh));
                    if (lane + 96 < rem) st_plain128(dp + 1536, pack_chunk_from_pair_raw<W0>(A3, B3, sh));
                }
                const bool last = rem <= PC;
                if (last && t_S != 0xFFFFFFFFu) {
                    const uint32_t wa = R.stage + ((t_S >> 5) << 2);
                    uint32_t w0, w1;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(wa));
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(wa + 4));
                    dp[t_ofs] = (uint8_t)(((funnel_l(w1, w0, t_S & 31u) >> 24) & t_ko) | (t_ko >> 8));
                }
                __syncwarp();  // every lane has used what it read from the stage: it can be refilled
                issue(R.stage, R.bar);
                R.stage += SB;
                R.bar += 8;
                if (R.stage == R.bars) {
                    R.stage = R.buf;
                    R.bar = R.bars;
                    R.par ^= 1;
                }
                if (last) break;
                rem -= PC;
                dp += PB;
                if (tracing) {
                    const uint32_t c0 = (uint32_t)clock64();
                    mbar_wait_s(R.bar, R.par);
                    R.wait_cycles += (uint32_t)clock64() - c0;
                } else {
                    mbar_wait_s(R.bar, R.par);
                }
            }
        };
        switch ((S0 >> 5) & 3u) {  // the word the chunks start at inside their 16-byte granule: fixed per stream
        case 0: pieces(std::integral_constant<int, 0>{}); break;
        case 1: pieces(std::integral_constant<int, 1>{}); break;
        case 2: pieces(std::integral_constant<int, 2>{}); break;
        default: pieces(std::integral_constant<int, 3>{}); break;
        }
    }
}

// The copy pool.  Ticket T = (round T / W, warp slot T % W); tickets run through the CTA's tasks in order.
// One task: claim the next ticket and copy it.  `block` = wait for the ticket's round to be resolved (the dedicated
// copy warps, and coding warps that have run out of symbols); otherwise only a ticket whose round IS resolved is
// claimed (a coding warp that helps while it waits must never wait for a round it has yet to arrive at).
// Returns 0 = nothing left at all, 1 = copied one task, 2 = nothing claimable right now.
// (__noinline__: one copy of the code for its three call sites, registers allocated apart from the coding loop's.)
// RING = through the calling copy warp's staging ring (its own instantiation: its own register allocation).
template <bool RING>
__device__ __noinline__ uint32_t packed_copy_one(PackCtl &ctl, const BlockIo &io, const PackedOut &po, uint32_t W, uint32_t total_warps,
                                                 uint32_t n_tasks, uint32_t lane, bool block, CopyRing *ring) {
    uint32_t T = 0, got = 1;
    if (lane == 0) {
        if (block) {
            T = atomicAdd(&ctl.copy_ticket, 1u);
        } else {
            T = *(volatile uint32_t *)&ctl.copy_ticket;
            const uint32_t r = T / W;
            const uint64_t task = (uint64_t)r * total_warps + blockIdx.x * W + (T - r * W);
            if (task >= n_tasks)
                got = 0;
            else if (*(volatile uint32_t *)&ctl.resolved <= r || atomicCAS(&ctl.copy_ticket, T, T + 1) != T)
                got = 2;
        }
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got != 1) return got;
    T = __shfl_sync(0xffffffffu, T, 0);
    const uint32_t r = T / W, slot = T - r * W;
    const uint64_t task = (uint64_t)r * total_warps + blockIdx.x * W + slot;
    if (task >= n_tasks) return 0;  // tickets run through the rounds in order: nothing valid after the first invalid one
    uint64_t *tr = po.trace ? po.trace + ((uint64_t)blockIdx.x * 32 + (threadIdx.x >> 5)) * kTraceWords : nullptr;
    uint64_t t0 = 0, t1 = 0;
    if (lane == 0) {
        if (tr) t0 = globaltimer_ns();
        const volatile uint32_t *res = &ctl.resolved;
        for (uint32_t ns = 100; *res <= r; ns = ns < 1600 ? 2 * ns : 1600) __nanosleep(ns);  // (a poll costs issue slots the coder wants)
        if (tr) t1 = globaltimer_ns();
    }
    __syncwarp();
    __threadfence_block();
    const uint64_t base = *(const volatile unsigned long long *)&po.byte_off[task * 32];
    if (RING) {
        // the slots were written through the generic proxy (by warps that fenced at gpu scope before arriving); the
        // bulk copies read them through the async proxy
        asm volatile("fence.proxy.async;" ::: "memory");
        CopyRing R = *ring;
        if (po.framed)
            packed_copy_task_ring<true>(io, po, task, base, lane, R);
        else
            packed_copy_task_ring<false>(io, po, task, base, lane, R);
        ring->stage = R.stage;
        ring->bar = R.bar;
        ring->par = R.par;
        if (po.trace && lane == 0) po.trace[((uint64_t)blockIdx.x * 32 + (threadIdx.x >> 5)) * kTraceWords + 25] += R.wait_cycles;
    } else if (po.framed) {
        packed_copy_task<true>(io, po, task, base, lane);
    } else {
        packed_copy_task<false>(io, po, task, base, lane);
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicAdd(&ctl.copy_done, 1u);
        if (tr) {
            const uint64_t t2 = globaltimer_ns();
            if (tr[20] == 0) tr[21] = t1;
            tr[20] += 1;
            tr[22] = t2;
            tr[23] += t2 - t1;
            tr[24] += t1 - t0;
        }
    }
    return 1;
}

// Wait until *word >= target, copying resolved tasks meanwhile (a coding warp that has to wait for a round's
// look-back -- i.e. for slower CTAs -- spends the wait on the pool's work).  The whole warp calls this.
__device__ __forceinline__ void packed_wait_helping(PackCtl &ctl, const BlockIo &io, const PackedOut &po, uint32_t W, uint32_t total_warps,
                                                    uint32_t n_tasks, uint32_t lane, const uint32_t *word, uint32_t target) {
    while (true) {
        uint32_t v = 0;
        if (lane == 0) v = *(const volatile uint32_t *)word;
        if (__shfl_sync(0xffffffffu, v, 0) >= target) return;
        if (packed_copy_one<false>(ctl, io, po, W, total_warps, n_tasks, lane, false, nullptr) != 1) __nanosleep(200);
    }
}

// Diagnostic (scl_debug_copy_only): the copy pool's work WITHOUT the coder beside it -- `warps` warps per CTA re-copy
// the tasks of a finished fused encode (scratch slots, bit lengths and record offsets as that call left them).
// Separates "a copy warp is slow because it waits for memory" from "... because 28 coding warps take its issue slots".
__global__ void __launch_bounds__(1024, 1) copy_only_kernel(BlockIo io, PackedOut po, uint32_t n_tasks) {
    extern __shared__ __align__(16) uint8_t copy_smem[];
    const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CopyRing R = copy_ring_at(copy_smem + warp * copy_ring_bytes(po.copy_stages, po.copy_piece_bytes), po.copy_stages, po.copy_piece_bytes, lane == 0);
    __syncthreads();
    for (uint32_t task = blockIdx.x * W + warp; task < n_tasks; task += gridDim.x * W) {
        const uint64_t base = po.byte_off[(uint64_t)task * 32];
        if (po.copy_stages) {
            if (po.framed)
                packed_copy_task_ring<true>(io, po, task, base, lane, R);
            else
                packed_copy_task_ring<false>(io, po, task, base, lane, R);
        } else if (po.framed)
            packed_copy_task<true>(io, po, task, base, lane);
        else
            packed_copy_task<false>(io, po, task, base, lane);
    }
}

// RAGGED: the blocks have their own sizes (io.sizes; io.block_len = the row capacity): a warp runs as many tiles as its
// longest block needs and every lane stops at its own end.
template <int KIND, uint32_t NBO, bool CHECK, bool PACKED, bool RAGGED>
__global__ void __launch_bounds__(PACKED ? 1024 : kMaxWarps * 32, 1)
    fast_encode_v2_kernel(const __grid_constant__ CUtensorMap tmap, const void *__restrict__ g_tab8, const uint32_t *__restrict__ g_tab2,
                          uint32_t tab2_bytes, RansConst c, BlockIo io, uint32_t n_tasks, PackedOut po) {
    __shared__ PackCtl ctl;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round the dynamic window up to 2 KiB so that ring addresses can be composed with OR
    uint8_t *smem = smem_raw + ((2048u - (smem_u32(smem_raw) & 2047u)) & 2047u);
    // W = coding warps; a PACKED launch carries kCopyWarps more (warp >= W), which own no tiles and no ring
    const uint32_t W = (blockDim.x >> 5) - (PACKED ? po.copy_warps : 0u), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *tiles = smem + warp * (kTileStages * kTileBytes);
    const saddr_t ring = saddr_of(smem + W * (kTileStages * kTileBytes) + warp * (kEncRingWords * 128)) + lane * 4;
    const uint8_t *s_tab = smem + W * kEncWarpSmem;
    const saddr_t s_tab2 = saddr_of(smem + W * kEncWarpSmem + kEncTabBytes);
    uint64_t *mbars = (uint64_t *)(smem + W * kEncWarpSmem + kEncTabBytes + tab2_bytes);
    uint64_t *tab_bar = mbars + W * kTileStages;
    uint64_t *my_bar = mbars + warp * kTileStages;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(tab_bar)) : "memory");
        for (uint32_t i = 0; i < W * kTileStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (PACKED && threadIdx.x == 0) {
        ctl.arrive[0] = ctl.arrive[1] = 0;
        ctl.resolved = ctl.copy_ticket = ctl.copy_done = 0;
    }
    // the copy warps' staging rings lie behind the mbarriers
    auto copy_mem = [&]() { return (uint8_t *)(((uintptr_t)(mbars + W * kTileStages + 1) + 15) & ~(uintptr_t)15) + (warp - W) * copy_ring_bytes(po.copy_stages, po.copy_piece_bytes); };
    if (PACKED && warp >= W && lane == 0 && po.copy_stages) (void)copy_ring_at(copy_mem(), po.copy_stages, po.copy_piece_bytes, true);
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_expect(tab_bar, kEncTabBytes + tab2_bytes);
        tma_bulk_g2s((void *)s_tab, g_tab8, kEncTabBytes, tab_bar);
        if (KIND == 1) tma_bulk_g2s(smem + W * kEncWarpSmem + kEncTabBytes, g_tab2, tab2_bytes, tab_bar);
    }
    mbar_wait(tab_bar, 0);

    // this lane's bank-rotated replica: 8 x 16-byte entries (rANS, LDS.128: a quarter warp per wavefront) or 16 x 8-byte rows
    // (tANS, LDS.64: half a warp per wavefront) -- 128 bytes per byte value either way
    const saddr_t my_tab = saddr_of(s_tab) + (KIND == 0 ? (lane & (kEncTabCopies - 1)) * 16 : (lane & (kTansTabCopies - 1)) * 8);
    const uint32_t n_uniform = io.block_len;
    const uint32_t n_tiles_uniform = (n_uniform + kTileCols - 1) / kTileCols;
    const uint32_t total_warps = gridDim.x * W;
    uint32_t tile_seq = 0;  // tiles consumed by this warp so far (selects stage and mbarrier parity)
    const uint32_t swz = (lane >> 1) & 3;

    if (PACKED && po.trace && lane == 0) po.trace[((uint64_t)blockIdx.x * 32 + warp) * kTraceWords] = globaltimer_ns();
    if (PACKED && warp >= W) {  // a copy warp has no tasks: it serves the pool until the CTA has nothing left
        if (po.copy_stages) {
            CopyRing R = copy_ring_at(copy_mem(), po.copy_stages, po.copy_piece_bytes, false);
            while (packed_copy_one<true>(ctl, io, po, W, total_warps, n_tasks, lane, true, &R)) {
            }
        } else {
            while (packed_copy_one<false>(ctl, io, po, W, total_warps, n_tasks, lane, true, nullptr)) {
            }
        }
        return;
    }
    uint32_t round = 0;
    for (uint32_t task = blockIdx.x * W + warp; task < n_tasks; task += total_warps, ++round) {
        const uint64_t b = (uint64_t)task * 32 + lane;
        const bool active = b < io.n_blocks;
        uint32_t n = n_uniform, n_tiles = n_tiles_uniform;  // RAGGED: n is this lane's, n_tiles the warp's
        bool too_long = false;  // RAGGED: a size beyond the row's capacity (a caller error) codes nothing and is reported
        if (RAGGED) {
            n = active ? io.sizes[b] : 0u;
            too_long = n > n_uniform;
            n = too_long ? 0u : n;
            n_tiles = (__reduce_max_sync(0xffffffffu, n) + kTileCols - 1) / kTileCols;
        }
        if (lane == 0) {
            for (uint32_t t = 0; t < kTileStages && t < n_tiles; ++t) {
                uint32_t st = (tile_seq + t) % kTileStages;
                tma_expect(my_bar + st, kTileBytes);
                tma_tile_2d(tiles + st * kTileBytes, &tmap, (int32_t)(t * kTileCols), (int32_t)(task * 32), my_bar + st);
            }
        }
        EncLaneV2T<PACKED> L;  // PACKED: the slot is private scratch, written as raw words (no byte swaps, here or in the copy pool)
        uint8_t *slot = io.out + (io.block0 + (active ? b : 0)) * io.out_stride;
        L.init((uint32_t)c.L, ring, slot, slot + io.out_stride);
        for (uint32_t t = 0; t < n_tiles; ++t, ++tile_seq) {
            const uint32_t st = tile_seq % kTileStages;
            mbar_wait(my_bar + st, (tile_seq / kTileStages) & 1);
            const uint8_t *row = tiles + st * kTileBytes + lane * kTileCols;
            const uint32_t left = RAGGED ? (n > t * kTileCols ? n - t * kTileCols : 0u) : n - t * kTileCols;
            if (RAGGED ? __all_sync(0xffffffffu, !active || left >= kTileCols) : left >= kTileCols) {
                // full tile: four whole chunks, no per-chunk bookkeeping
                if (active) {
#pragma unroll 1
                    for (uint32_t ch = 0; ch < kTileCols / 16; ++ch) {
                        const uint4 q = *(const uint4 *)(row + ((ch ^ swz) << 4));  // 64-byte TMA swizzle: conflict-free LDS.128
                        const u32x4 v = {q.x, q.y, q.z, q.w};
                        if (KIND == 0)
                            enc_chunk16<NBO, CHECK>(L, my_tab, kEncTabCopies * 16, v);
                        else
                            tans_enc_chunk16<CHECK>(L, my_tab, kEncTabCopies * 16, s_tab2, v);
                    }
                }
            } else {
#pragma unroll 1
                for (uint32_t ch = 0; ch < kTileCols / 16; ++ch) {
                    if (ch * 16 >= left) break;
                    const uint32_t cnt = left - ch * 16 >= 16 ? 16u : left - ch * 16;
                    const uint4 q = *(const uint4 *)(row + ((ch ^ swz) << 4));
                    if (active) {
                        u32x4 v = {q.x, q.y, q.z, q.w};
                        if (KIND == 0)
                            enc_chunk<NBO, CHECK>(L, my_tab, kEncTabCopies * 16, v, cnt);
                        else
                            tans_enc_chunk<CHECK>(L, my_tab, kEncTabCopies * 16, s_tab2, v, cnt);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && t + kTileStages < n_tiles) {
                tma_expect(my_bar + st, kTileBytes);
                tma_tile_2d(tiles + st * kTileBytes, &tmap, (int32_t)((t + kTileStages) * kTileCols), (int32_t)(task * 32), my_bar + st);
            }
        }
        uint32_t rec_bytes = 0;  // PACKED: bytes of this lane's record in the contiguous output
        if (active) {
            L.put(L.x, c.NSB);  // header in front of the payload (rANS.py:199,206-208)
            uint32_t st = SCL_ST_OK;
            if (c.DBSB < 32 && (n >> c.DBSB)) st = SCL_ST_OVERFLOW;
            L.put64((uint64_t)n, c.DBSB);
            uint64_t bits = L.finish();
            if (L.bad) st = SCL_ST_BAD_SYMBOL;
            if (L.ovf || (RAGGED && too_long)) st = SCL_ST_OVERFLOW;
            io.bit_len[b] = bits;
            if (!PACKED) io.bit_off[b] = (io.block0 + b + 1) * io.out_stride * 8 - bits;
            io.status[b] = st;
            if (PACKED && st == SCL_ST_OK) rec_bytes = (uint32_t)packed_size(bits, po.framed != 0);
        }
        __syncwarp();
        if (PACKED) {
            // (i) this task's total; (ii) arrive at the CTA's round; the last warp to arrive resolves the round
            const uint32_t T = __reduce_add_sync(0xffffffffu, rec_bytes);
            if (po.trace && lane == 0 && round < 19) po.trace[((uint64_t)blockIdx.x * 32 + warp) * kTraceWords + 1 + round] = globaltimer_ns();
            const uint32_t first = round * total_warps + blockIdx.x * W;  // first task of this CTA's round
            const uint32_t nvalid = n_tasks - first < W ? n_tasks - first : W;
            const uint32_t par = round & 1;
            uint32_t old = 0;
            // warp_tot[par] / arrive[par] still belong to round - 2 until that round is resolved (its last warp may be
            // in the look-back, waiting for lower-numbered CTAs): nobody arrives at this round before.
            if (round >= 2) packed_wait_helping(ctl, io, po, W, total_warps, n_tasks, lane, &ctl.resolved, round - 1);
            // One fence for both directions: (acquire) what the resolver of round - 2 wrote before it published `resolved` --
            // the reset of arrive[par] -- is ordered before this warp's arrival; (release) this task's slots are visible at gpu
            // scope before anybody (generic or async proxy) is told to read them.
            // (These hand-overs -- warp_tot / arrive, resolved -- are flag-and-fence message passing, not barriers: racecheck,
            // which only knows barriers, reports them: profiles/r4d_racecheck.log.)
            __threadfence();
            if (lane == 0) {
                *(volatile unsigned long long *)&ctl.warp_tot[par][warp] = T;
                __threadfence_block();
                old = atomicAdd(&ctl.arrive[par], 1u);
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == nvalid - 1) {
                if (lane == 0) ctl.arrive[par] = 0;
                __threadfence_block();
                const uint64_t t = lane < nvalid ? *(const volatile unsigned long long *)&ctl.warp_tot[par][lane] : 0ull;
                const uint64_t incl = warp_incl_scan_u64(t, lane);
                const uint64_t A = __shfl_sync(0xffffffffu, incl, 31);
                const int64_t g = (int64_t)po.g_base + (int64_t)round * gridDim.x + blockIdx.x;
                if (lane == 0) st_relaxed_gpu(po.cta_state + g, (g ? kLbAgg : kLbPrefix) | A);
                uint64_t excl = 0;
                if (g) {
                    excl = lookback_exclusive(po.cta_state, g, lane);
                    if (lane == 0) st_relaxed_gpu(po.cta_state + g, kLbPrefix | (excl + A));
                }
                if (lane == 0 && first + nvalid == n_tasks) po.byte_off[io.n_blocks] = excl + A;  // the very last round: grand total
                // Each task's place goes to its first block's entry of byte_off (the copier rewrites it with the same value):
                // one word per task, never reused, so the coder does not wait for the copy pool -- what the pool has not
                // moved when the symbols run out is moved by the whole CTA at the end.
                if (lane < nvalid) *(volatile unsigned long long *)&po.byte_off[(uint64_t)(first + lane) * 32] = excl + incl - t;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *(volatile uint32_t *)&ctl.resolved = round + 1;
            }
        }
    }
    if (PACKED) {  // a coding warp that is out of symbols joins the pool
        if (po.helper_ring) {
            // its tile buffers are idle from here on (every tile load it issued has been waited for): 4 KiB = a ring of three
            // 1 KiB stages, its mbarriers and the stream table
            static_assert(copy_ring_bytes(3, 1024) <= kTileStages * kTileBytes, "the tile buffers hold a 3 x 1 KiB staging ring");
            __syncwarp();
            CopyRing Rw = copy_ring_at(tiles, 3, 1024, lane == 0);
            __syncwarp();
            while (packed_copy_one<true>(ctl, io, po, W, total_warps, n_tasks, lane, true, &Rw)) {
            }
        } else {
            while (packed_copy_one<false>(ctl, io, po, W, total_warps, n_tasks, lane, true, nullptr)) {
            }
        }
    }
}

// Decode.  Output goes through a per-warp 32 x 64-byte tile in shared memory (64-byte swizzle, so
// each lane's STS.128 into its own row is conflict-free) that one lane hands to the TMA
// (cp.async.bulk.tensor store, UTMASTG): the scattered per-lane sector stores leave the L1 data
// pipe, which is the decoder's tightest resource (profiles/r1c).  Needs every block of the warp to
// have the same size (it is read from the stream headers, so this is checked per task with
// __all_sync); otherwise the warp falls back to per-lane sector stores.
constexpr uint32_t kDecTileBytes = 32 * kTileCols;

template <int KIND, uint32_t NBO, bool BAL>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
    fast_decode_v2_kernel(const __grid_constant__ CUtensorMap out_map, uint32_t use_tiles, const uint32_t *__restrict__ g_lut,
                          uint32_t lut_bytes, RansConst c, DecodeIo io, uint32_t n_tasks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *tile = smem + warp * kDecTileBytes;
    const saddr_t ring = saddr_of(smem + W * kDecTileBytes + warp * kDecWarpSmem) + lane * 4;
    const uint32_t *s_lut = (const uint32_t *)(smem + W * (kDecTileBytes + kDecWarpSmem));
    stage_table((void *)s_lut, g_lut, lut_bytes, &mbar);
    const uint32_t total_warps = gridDim.x * W;
    const uint32_t swz = (lane >> 1) & 3;
    const saddr_t my_row = saddr_of(tile) + lane * kTileCols;
    typename std::conditional<KIND == 0, RansStepper<NBO, BAL>, TansStepper<BAL>>::type S;
    S.init(saddr_of(s_lut), c);

    for (uint32_t task = blockIdx.x * W + warp; task < n_tasks; task += total_warps) {
        const uint64_t b = (uint64_t)task * 32 + lane;
        const bool active = b < io.n_blocks;
        DecLaneV2 D;
        uint32_t size = 0, st = SCL_ST_OK, p = 0;
        uint64_t off = 0;
        uint8_t *out = io.sym + (active ? b : 0) * io.sym_stride;
        bool ok = false;
        if (active) {
            off = io.bit_off[b];
            D.init(io.in, io.in_bytes, off, ring);
            ok = dec_read_header(D, c, io.sym_stride, size, st);
            p = size;
        }
        // warp-uniform tile path?
        const uint32_t size0 = __shfl_sync(0xffffffffu, size, 0);
        const bool tiles = use_tiles && !S.degenerate() && __all_sync(0xffffffffu, !active || (ok && size == size0));
        if (tiles) {
            if (active) dec_head(D, S, out, p, 64);
            const uint32_t n_it = (size0 & ~63u) >> 6;  // same for every lane
            for (uint32_t it = 0; it < n_it; ++it) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous store has read the tile
                __syncwarp();
                if (active) {
#pragma unroll 1
                    for (int g = 3; g >= 0; --g) {
                        uint32_t w[4];
                        D.prefetch_begin();
                        S.group16(D, w);
                        D.prefetch_end();
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(my_row + ((g ^ swz) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                                     : "memory");
                    }
                    p -= 64;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA
                __syncwarp();
                if (lane == 0) {
                    const int32_t col = (int32_t)((size0 & ~63u) - 64 * (it + 1));
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&out_map), "r"(col),
                                 "r"((int32_t)(task * 32)), "r"(saddr_of(tile))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        } else if (active && ok) {
            if (S.degenerate())
                while (p) out[--p] = (uint8_t)S.one(D);
            dec_head(D, S, out, p, 32);
            dec_body_sectors(D, S, out, p);
        }
        if (active) {
            uint64_t used = 0;
            if (ok) {
                used = D.bp - D.start_bp;
                st = D.x == (uint32_t)c.L ? SCL_ST_OK : SCL_ST_STATE_MISMATCH;
                if (st == SCL_ST_OK && used > avail_bits_of(io, b, off)) st = SCL_ST_TRUNCATED;
            }
            io.sizes[b] = size;
            io.consumed[b] = used;
            io.status[b] = st;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// tANS kernels.  Tables (L entries each) are built on the device, one thread per state, by
// inverting the rANS decode step exactly as tANS.py:88-99 / :208-215 cache it.
// ------------------------------------------------------------------------------------------------
__global__ void tans_build_kernel(const RansGeneric *__restrict__ t, RansConst c, const uint32_t *__restrict__ row_of_idx,
                                  uint32_t *__restrict__ enc_table, uint32_t *__restrict__ dec_packed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.L) return;
    tans_build_entry(*t, c, row_of_idx, enc_table, dec_packed, i);
}

template <bool SMEM_TABLE>
__global__ void __launch_bounds__(kThreads) tans_encode_kernel(const TansSym *__restrict__ g_sym, const uint32_t *__restrict__ g_enc,
                                                               uint32_t table_bytes, RansConst c, BlockIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ TansSym s_sym[256];
    __shared__ uint64_t mbar;
    mbar_init(&mbar);
    if (threadIdx.x == 0) {
        tma_expect(&mbar, (uint32_t)sizeof(s_sym) + (SMEM_TABLE ? table_bytes : 0u));
        tma_bulk_g2s(s_sym, g_sym, sizeof(s_sym), &mbar);
        if (SMEM_TABLE) tma_bulk_g2s(s_dyn, g_enc, table_bytes, &mbar);
    }
    mbar_wait(&mbar, 0);
    const uint32_t *enc = SMEM_TABLE ? (const uint32_t *)s_dyn : g_enc;
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    LifoBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint64_t bits = 0;
    uint32_t st = tans_encode_lane(s_sym, enc, c, io.sym + b * io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = (b + 1) * io.out_stride * 8 - bits;
    io.status[b] = st;
}

template <bool SMEM_TABLE>
__global__ void __launch_bounds__(kThreads) tans_decode_kernel(const uint32_t *__restrict__ g_dec, uint32_t table_bytes, RansConst c,
                                                               DecodeIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint64_t mbar;
    if (SMEM_TABLE) stage_table(s_dyn, g_dec, table_bytes, &mbar);
    const uint32_t *dec = SMEM_TABLE ? (const uint32_t *)s_dyn : g_dec;
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint64_t used = 0;
    uint32_t st = tans_decode_lane(dec, c, r, io.sym + b * io.sym_stride, io.sym_stride, size, used);
    if (st == SCL_ST_OK && used > avail_bits_of(io, b, off)) st = SCL_ST_TRUNCATED;
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// range coder kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) range_encode_kernel(const RangeTab *__restrict__ g_tab, RangeConst c, BlockIo io) {
    __shared__ RangeTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(RangeTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint64_t bits = 0;
    uint32_t st = range_encode_lane(s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kThreads) range_decode_kernel(const RangeTab *__restrict__ g_tab, const uint8_t *__restrict__ g_lut,
                                                                uint32_t lut_bytes, RangeConst c, DecodeIo io) {
    extern __shared__ __align__(16) uint8_t s_lut[];
    __shared__ RangeTab s_tab;
    __shared__ uint64_t mbar;
    mbar_init(&mbar);
    if (threadIdx.x == 0) {
        tma_expect(&mbar, (uint32_t)sizeof(RangeTab) + lut_bytes);
        tma_bulk_g2s(&s_tab, g_tab, sizeof(RangeTab), &mbar);
        if (lut_bytes) tma_bulk_g2s(s_lut, g_lut, lut_bytes, &mbar);
    }
    mbar_wait(&mbar, 0);
    uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint64_t used = 0;
    uint32_t st = range_decode_lane(s_tab, c, lut_bytes ? s_lut : nullptr, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used);
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// range coder, second generation (scl_range.cuh) on the v2 machinery: persistent CTAs, one warp per
// task of 32 blocks, symbols in by 2-D TMA tiles, coded bytes through sector rings.  Every lane of
// a warp runs the same instruction stream (the normalisation's extra rounds are a warp vote), so
// lanes past the end of the batch run as padding with a zero-capacity slot.
// smem layout as in fast_encode_v2_kernel: [tiles per warp][rings per warp][table 32 KiB][mbarriers];
// the table is [byte value][lane] so that every lane reads its own bank.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRangeEncTabBytes = 256 * 32 * 4;

template <bool CHECK>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
    range_encode_v2_kernel(const __grid_constant__ CUtensorMap tmap, const uint32_t *__restrict__ g_tab, uint32_t shift, BlockIo io,
                           uint32_t n_tasks) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((2048u - (smem_u32(smem_raw) & 2047u)) & 2047u);
    const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *tiles = smem + warp * (kTileStages * kTileBytes);
    const saddr_t ring = saddr_of(smem + W * (kTileStages * kTileBytes) + warp * (kEncRingWords * 128)) + lane * 4;
    const uint8_t *s_tab = smem + W * kEncWarpSmem;
    uint64_t *mbars = (uint64_t *)(smem + W * kEncWarpSmem + kRangeEncTabBytes);
    uint64_t *tab_bar = mbars + W * kTileStages;
    uint64_t *my_bar = mbars + warp * kTileStages;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(tab_bar)) : "memory");
        for (uint32_t i = 0; i < W * kTileStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_expect(tab_bar, kRangeEncTabBytes);
        tma_bulk_g2s((void *)s_tab, g_tab, kRangeEncTabBytes, tab_bar);
    }
    mbar_wait(tab_bar, 0);

    const saddr_t my_tab = saddr_of(s_tab) + lane * 4;
    const uint32_t n = io.block_len;
    const uint32_t n_tiles = (n + kTileCols - 1) / kTileCols;
    const uint32_t total_warps = gridDim.x * W;
    uint32_t tile_seq = 0;
    const uint32_t swz = (lane >> 1) & 3;

    for (uint32_t task = blockIdx.x * W + warp; task < n_tasks; task += total_warps) {
        const uint64_t b = (uint64_t)task * 32 + lane;
        const bool active = b < io.n_blocks;
        if (lane == 0) {
            for (uint32_t t = 0; t < kTileStages && t < n_tiles; ++t) {
                uint32_t st = (tile_seq + t) % kTileStages;
                tma_expect(my_bar + st, kTileBytes);
                tma_tile_2d(tiles + st * kTileBytes, &tmap, (int32_t)(t * kTileCols), (int32_t)(task * 32), my_bar + st);
            }
        }
        RangeEncV2 L;
        uint8_t *slot = io.out + (active ? b : 0) * io.out_stride;
        L.init(ring, slot, active ? slot + io.out_stride : slot, shift);  // padding lanes: zero capacity, nothing is stored
        L.put_word(n);  // [size : 32] (range_coder.py:201-205)
        L.spill_check();
        for (uint32_t t = 0; t < n_tiles; ++t, ++tile_seq) {
            const uint32_t st = tile_seq % kTileStages;
            mbar_wait(my_bar + st, (tile_seq / kTileStages) & 1);
            const uint8_t *row = tiles + st * kTileBytes + lane * kTileCols;
            const uint32_t left = n - t * kTileCols;  // the same for every lane
#pragma unroll 1
            for (uint32_t ch = 0; ch < kTileCols / 16; ++ch) {
                if (ch * 16 >= left) break;
                const uint32_t cnt = left - ch * 16 >= 16 ? 16u : left - ch * 16;
                const uint4 q = *(const uint4 *)(row + ((ch ^ swz) << 4));
                const u32x4 v = {q.x, q.y, q.z, q.w};
                range_enc_chunk<CHECK, true>(L, my_tab, 128, v, cnt);
            }
            __syncwarp();
            if (lane == 0 && t + kTileStages < n_tiles) {
                tma_expect(my_bar + st, kTileBytes);
                tma_tile_2d(tiles + st * kTileBytes, &tmap, (int32_t)((t + kTileStages) * kTileCols), (int32_t)(task * 32), my_bar + st);
            }
        }
        const uint64_t bits = L.finish();
        if (active) {
            uint32_t st = SCL_ST_OK;
            if (L.bad) st = SCL_ST_BAD_SYMBOL;
            if (L.ovf) st = SCL_ST_OVERFLOW;
            io.bit_len[b] = bits;
            io.bit_off[b] = b * io.out_stride * 8;
            io.status[b] = st;
        }
        __syncwarp();
    }
}

// smem: [input rings per warp][LUT]
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
    range_decode_v2_kernel(const uint32_t *__restrict__ g_lut, uint32_t lut_bytes, uint32_t shift, uint32_t T, uint32_t last_entry, DecodeIo io,
                           uint32_t n_tasks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    saddr_t ring = saddr_of(smem + warp * kDecWarpSmem) + lane * 4;
    asm volatile("" : "+r"(ring));  // keep it in a register: the compiler otherwise recomputes it from threadIdx at every peek (8 instructions)
    const uint32_t *s_lut = (const uint32_t *)(smem + W * kDecWarpSmem);
    stage_table((void *)s_lut, g_lut, lut_bytes, &mbar);
    const uint32_t total_warps = gridDim.x * W;
    RangeDecConst dc;
    dc.lut = saddr_of(s_lut);
    dc.shift = shift;
    dc.neg1 = 0xFFFFFFFFu - (shift >> 8);  // opaque -1
    dc.T = T;
    dc.last = last_entry;

    for (uint32_t task = blockIdx.x * W + warp; task < n_tasks; task += total_warps) {
        const uint64_t b = (uint64_t)task * 32 + lane;
        const bool active = b < io.n_blocks;
        const uint64_t bb = active ? b : io.n_blocks - 1;  // padding lanes decode the last block again, without storing
        const uint64_t off = io.bit_off[bb];
        uint8_t *out = io.sym + bb * io.sym_stride;
        DecLaneV2 D;
        D.init(io.in, io.in_bytes, off, ring);
        RangeDecV2 R;
        uint32_t size = 0;
        const bool ok = range_dec_header(D, io.sym_stride, size, R);
        const uint32_t size0 = __shfl_sync(0xffffffffu, size, 0);
        if (__all_sync(0xffffffffu, ok && size == size0)) {
            range_dec_body<true>(D, R, dc, out, size, active);  // warp-uniform: the extra normalisation rounds are a vote
        } else if (active && ok) {
            range_dec_body<false>(D, R, dc, out, size, true);   // ragged sizes: per-lane control flow
        }
        if (active) {
            uint32_t st = SCL_ST_OK;
            uint64_t used = 0;
            if (!ok) {
                st = SCL_ST_OVERFLOW;
            } else {
                used = D.bp - D.start_bp;
                if (R.ovf) st = SCL_ST_OVERFLOW;
                if (st == SCL_ST_OK && used > avail_bits_of(io, b, off)) st = SCL_ST_TRUNCATED;
            }
            io.sizes[b] = st == SCL_ST_OK ? size : 0;
            io.consumed[b] = used;
            io.status[b] = st;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// arithmetic coder kernels: 32 lanes per CTA, each lane's 256-counter model is a Fenwick tree in
// shared memory, lane-interleaved (element i of lane l at word i*32 + l => every access of a
// warp hits 32 distinct banks whatever the per-lane index).
// ------------------------------------------------------------------------------------------------
constexpr int kAecThreads = 32;
struct SmemTree {  // 32-bit counters
    uint32_t *base;  // &smem[lane]
    static constexpr uint32_t kBytes = 257 * kAecThreads * 4;
    __device__ __forceinline__ explicit SmemTree(uint8_t *smem) : base((uint32_t *)smem + threadIdx.x) {}
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return base[i * kAecThreads]; }
    __device__ __forceinline__ void set(uint32_t i, uint32_t v) { base[i * kAecThreads] = v; }
};
// 16-bit counters, two tree nodes of the SAME lane per 32-bit word: still one bank per lane, half
// the shared memory (twice the resident warps).  Usable when no count can reach 65536.
struct SmemTree16 {
    uint32_t *base;
    static constexpr uint32_t kBytes = 129 * kAecThreads * 4;
    __device__ __forceinline__ explicit SmemTree16(uint8_t *smem) : base((uint32_t *)smem + threadIdx.x) {}
    __device__ __forceinline__ uint32_t get(uint32_t i) const {
        uint32_t w = base[(i >> 1) * kAecThreads];
        return (i & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    __device__ __forceinline__ void set(uint32_t i, uint32_t v) {
        uint32_t *p = base + (i >> 1) * kAecThreads;
        uint32_t w = *p;
        *p = (i & 1) ? ((w & 0xFFFFu) | (v << 16)) : ((w & 0xFFFF0000u) | (v & 0xFFFFu));
    }
};

template <typename Tree>
__device__ __forceinline__ uint64_t aec_load_model(Tree &F, const AecTab &tab, const AecConst &c, const uint64_t *model) {
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
template <typename Tree>
__device__ __forceinline__ void aec_store_model(Tree &F, const AecConst &c, uint64_t *model) {
    if (!model) return;
    fen_unbuild(F);
    for (uint32_t i = 0; i < c.n_sym; ++i) model[i] = F.get(i + 1);
}

template <typename Tree>
__global__ void __launch_bounds__(kAecThreads) aec_encode_kernel(const AecTab *__restrict__ g_tab, AecConst c, BlockIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kAecThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    Tree F(s_dyn);
    uint64_t *my_model = model ? model + b * c.n_sym : nullptr;
    uint64_t total = aec_load_model(F, s_tab, c, my_model);
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint64_t bits = 0, total_out = 0;
    uint32_t st = aec_encode_lane(F, s_tab, c, total, io.sym + b * io.sym_stride, n, w, bits, total_out);
    aec_store_model(F, c, my_model);
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

template <typename Tree>
__global__ void __launch_bounds__(kAecThreads) aec_decode_kernel(const AecTab *__restrict__ g_tab, AecConst c, DecodeIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * kAecThreads + threadIdx.x;
    if (b >= io.n_blocks) return;
    Tree F(s_dyn);
    uint64_t *my_model = model ? model + b * c.n_sym : nullptr;
    uint64_t total = aec_load_model(F, s_tab, c, my_model);
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint64_t used = 0, total_out = 0;
    uint32_t st = aec_decode_lane(F, s_tab, c, total, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size,
                                  used, total_out);
    aec_store_model(F, c, my_model);
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// arithmetic coder, second generation (scl_aec.cuh): 4 warps per CTA, per-lane two-level count
// structure in shared memory ([word][lane] interleave), prefix-mask table in 8 bank-rotated replicas.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kAec2Warps = 4;
constexpr uint32_t kAec2ModelBytes = kAecModelWords * 128;      // per warp
constexpr uint32_t kAec2MaskBytes = 17 * kEncTabCopies * 16;    // 17 entries x 8 replicas x 16 B

__device__ __forceinline__ AecModel aec2_setup(uint8_t *smem, const AecTab *g_tab, AecTab *s_tab, uint64_t *mbar) {
    // smem: [mask table][models per warp]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *masks = smem;
    for (uint32_t i = threadIdx.x; i < 17 * kEncTabCopies * 16; i += blockDim.x) {
        uint32_t t = i / (kEncTabCopies * 16), k = i % 16;
        masks[i] = k < t ? 1 : 0;
    }
    stage_table(s_tab, g_tab, sizeof(AecTab), mbar);  // includes a __syncthreads
    __syncthreads();
    AecModel M;
    M.w = saddr_of(smem + kAec2MaskBytes + warp * kAec2ModelBytes) + lane * 4;
    M.stride = 128;
    M.masks = saddr_of(masks) + (lane & (kEncTabCopies - 1)) * 16;
    M.mstride = kEncTabCopies * 16;
    return M;
}

__device__ __forceinline__ AecIidPolicy aec2_iid_policy(const AecModel &M, const AecTab &tab, const AecConst &c) {
    AecIidPolicy pol;
    pol.M = M;
    uint64_t total = 0;
    M.load(tab.init_freq, nullptr, c.n_sym, total);
    pol.tot = (uint32_t)total;
    pol.adaptive = c.model == SCL_MODEL_ADAPTIVE_IID;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    pol.n_sym = c.n_sym;
    return pol;
}

__global__ void __launch_bounds__(kAec2Warps * 32) aec2_encode_kernel(const AecTab *__restrict__ g_tab, AecConst c, BlockIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    const AecModel M = aec2_setup(s_dyn, g_tab, &s_tab, &mbar);
    uint64_t b = (uint64_t)blockIdx.x * (kAec2Warps * 32) + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecIidPolicy pol = aec2_iid_policy(M, s_tab, c);
    uint64_t bits = 0;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint32_t st = c.P == 32 ? aec2_encode_lane<decltype(pol), 32>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits)
                            : aec2_encode_lane<decltype(pol), 0>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kAec2Warps * 32) aec2_decode_kernel(const AecTab *__restrict__ g_tab, AecConst c, DecodeIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    const AecModel M = aec2_setup(s_dyn, g_tab, &s_tab, &mbar);
    uint64_t b = (uint64_t)blockIdx.x * (kAec2Warps * 32) + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecIidPolicy pol = aec2_iid_policy(M, s_tab, c);
    uint64_t used = 0;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint32_t st = c.P == 32 ? aec2_decode_lane<decltype(pol), 32>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used)
                            : aec2_decode_lane<decltype(pol), 0>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used);
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// The same kernels on the 8-bit-counter model (AecModel8): 9 KiB of shared memory per warp instead of 17, so 20
// instead of 12 resident warps per SM for a coder that is latency-bound (profiles/r1s).  Selected by the host when
// the model cannot overflow its escape list (AecHost::model8_ok).
constexpr uint32_t kAec8ModelBytes = kAecModel8Words * 128;  // per warp

__device__ __forceinline__ AecIid8Policy aec8_setup(uint8_t *smem, const AecTab *g_tab, AecTab *s_tab, uint64_t *mbar, const AecConst &c) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *masks = smem;
    for (uint32_t i = threadIdx.x; i < 17 * kEncTabCopies * 16; i += blockDim.x) {
        uint32_t t = i / (kEncTabCopies * 16), k = i % 16;
        masks[i] = k < t ? 1 : 0;
    }
    stage_table(s_tab, g_tab, sizeof(AecTab), mbar);  // includes a __syncthreads
    __syncthreads();
    AecIid8Policy pol;
    pol.M.w = saddr_of(smem + kAec2MaskBytes + warp * kAec8ModelBytes) + lane * 4;
    pol.M.stride = 128;
    pol.M.masks = saddr_of(masks) + (lane & (kEncTabCopies - 1)) * 16;
    pol.M.mstride = kEncTabCopies * 16;
    uint64_t total = 0;
    pol.M.load(s_tab->init_freq, c.n_sym, total);
    pol.tot = (uint32_t)total;
    pol.adaptive = c.model == SCL_MODEL_ADAPTIVE_IID;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    pol.n_sym = c.n_sym;
    return pol;
}

constexpr uint32_t kAec8MaxWarps = 24;  // 24 x 9 KiB of models + masks + table: one CTA fills an SM's shared memory
__global__ void __launch_bounds__(kAec8MaxWarps * 32) aec8_encode_kernel(const AecTab *__restrict__ g_tab, AecConst c, BlockIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    AecIid8Policy pol = aec8_setup(s_dyn, g_tab, &s_tab, &mbar, c);  // (lanes past the batch load a model nobody uses)
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint64_t bits = 0;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint32_t st = c.P == 32 ? aec2_encode_lane<decltype(pol), 32>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits)
                            : aec2_encode_lane<decltype(pol), 0>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits);
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kAec8MaxWarps * 32) aec8_decode_kernel(const AecTab *__restrict__ g_tab, AecConst c, DecodeIo io) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    AecIid8Policy pol = aec8_setup(s_dyn, g_tab, &s_tab, &mbar, c);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    uint64_t used = 0;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint32_t st = c.P == 32 ? aec2_decode_lane<decltype(pol), 32>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used)
                            : aec2_decode_lane<decltype(pol), 0>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used);
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// arithmetic coder with the order-k context model (AecCtxPolicy): per lane n_ctx rows of n_sym
// counters + n_ctx row totals, 32-bit, [word][lane] in shared memory (as many warps per CTA as
// fit); the model table (d_model) carries the counts and the context from one call to the next
// like the reference's model object does.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kAecCtxMaxWarps = 4;
static uint32_t aec_ctx_warps(const AecConst &c) {
    uint32_t per_warp = c.n_ctx * (c.n_sym + 1) * 128, fit = (200u * 1024u) / per_warp;
    return fit > kAecCtxMaxWarps ? kAecCtxMaxWarps : (fit ? fit : 1);
}

__device__ __forceinline__ AecCtxPolicy aec_ctx_setup(uint8_t *smem, const AecConst &c) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AecCtxPolicy pol;
    pol.n_sym = c.n_sym;
    pol.n_ctx = c.n_ctx;
    pol.w = saddr_of(smem + warp * (pol.n_words() * 128)) + lane * 4;
    pol.stride = 128;
    pol.ctx = 0;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    return pol;
}

__global__ void __launch_bounds__(kAecCtxMaxWarps * 32) aec_ctx_encode_kernel(const AecTab *__restrict__ g_tab, AecConst c, BlockIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecCtxPolicy pol = aec_ctx_setup(s_dyn, c);
    uint64_t *my_model = model ? model + b * ((uint64_t)c.n_ctx * c.n_sym + 1) : nullptr;
    pol.load(my_model);
    uint64_t bits = 0;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint32_t st = c.P == 32 ? aec2_encode_lane<decltype(pol), 32>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits)
                            : aec2_encode_lane<decltype(pol), 0>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits);
    if (my_model) pol.store(my_model);
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kAecCtxMaxWarps * 32) aec_ctx_decode_kernel(const AecTab *__restrict__ g_tab, AecConst c, DecodeIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecCtxPolicy pol = aec_ctx_setup(s_dyn, c);
    uint64_t *my_model = model ? model + b * ((uint64_t)c.n_ctx * c.n_sym + 1) : nullptr;
    pol.load(my_model);
    uint64_t used = 0;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint32_t st = c.P == 32 ? aec2_decode_lane<decltype(pol), 32>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used)
                            : aec2_decode_lane<decltype(pol), 0>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used);
    if (my_model) pol.store(my_model);
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// The order-k model with its table in HBM (AecCtxGlobalPolicy): one warp per CTA-slot of 32 blocks, the lanes'
// row totals in shared memory ([row][lane]); `model` is mandatory and is the table itself.
constexpr uint32_t kAecCtxGlobalWarps = 2;
__device__ __forceinline__ AecCtxGlobalPolicy aec_ctx_global_setup(uint8_t *smem, const AecConst &c, uint64_t *my_model) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AecCtxGlobalPolicy pol;
    pol.tab = my_model;
    pol.tot = saddr_of(smem + warp * (c.n_ctx * 128)) + lane * 4;
    pol.tstride = 128;
    pol.n_sym = c.n_sym;
    pol.n_ctx = c.n_ctx;
    pol.ctx = 0;
    pol.max_total = c.max_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c.max_total;
    return pol;
}

__global__ void __launch_bounds__(kAecCtxGlobalWarps * 32) aec_ctx_global_encode_kernel(const AecTab *__restrict__ g_tab, AecConst c, BlockIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecCtxGlobalPolicy pol = aec_ctx_global_setup(s_dyn, c, model + b * ((uint64_t)c.n_ctx * c.n_sym + 1));
    pol.load();
    uint64_t bits = 0;
    uint32_t n = io.sizes ? io.sizes[b] : io.block_len;
    FwdBitWriter w;
    uint8_t *slot = io.out + b * io.out_stride;
    w.init(slot, slot + io.out_stride);
    uint32_t st = c.P == 32 ? aec2_encode_lane<decltype(pol), 32>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits)
                            : aec2_encode_lane<decltype(pol), 0>(pol, s_tab, c, io.sym + b * io.sym_stride, io.sym_stride, n, w, bits);
    pol.store();
    io.bit_len[b] = bits;
    io.bit_off[b] = b * io.out_stride * 8;
    io.status[b] = st;
}

__global__ void __launch_bounds__(kAecCtxGlobalWarps * 32) aec_ctx_global_decode_kernel(const AecTab *__restrict__ g_tab, AecConst c, DecodeIo io, uint64_t *model) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ AecTab s_tab;
    __shared__ uint64_t mbar;
    stage_table(&s_tab, g_tab, sizeof(AecTab), &mbar);
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= io.n_blocks) return;
    AecCtxGlobalPolicy pol = aec_ctx_global_setup(s_dyn, c, model + b * ((uint64_t)c.n_ctx * c.n_sym + 1));
    pol.load();
    uint64_t used = 0;
    BitReader r;
    uint64_t off = io.bit_off[b];
    r.init(io.in, io.in_bytes, off);
    uint32_t size = 0;
    uint32_t st = c.P == 32 ? aec2_decode_lane<decltype(pol), 32>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used)
                            : aec2_decode_lane<decltype(pol), 0>(pol, s_tab, c, r, avail_bits_of(io, b, off), io.sym + b * io.sym_stride, io.sym_stride, size, used);
    pol.store();
    io.sizes[b] = size;
    io.consumed[b] = used;
    io.status[b] = st;
}

// ------------------------------------------------------------------------------------------------
// stream packing: bit-granular copy of each block's stream to a byte-aligned destination (scl_pack.cuh)
// ------------------------------------------------------------------------------------------------
// Optional arguments of both kernels: `status` (a failed block takes no room and is not copied), `dst_bytes`
// (a record that would end past the destination is dropped and its status set to OVERFLOW), `new_bit_off`
// (where the block's first stream bit now lies in dst; may alias src_bit_off).
struct PackIo {
    const uint8_t *src;
    const uint64_t *src_bit_off;
    const uint64_t *bit_len;
    uint64_t n_blocks;
    uint8_t *dst;
    uint64_t dst_bytes;  // 0 = unchecked
    const uint64_t *dst_byte_off;
    uint32_t *status;
    uint64_t *new_bit_off;
};

// first generation: one CTA per block, byte-wise (any source alignment)
template <bool FRAMED>
__global__ void __launch_bounds__(kThreads) pack_kernel(PackIo io) {
    const uint64_t b = blockIdx.x;
    if (io.status && io.status[b] != SCL_ST_OK) return;
    const uint64_t off = io.src_bit_off[b], nbits = io.bit_len[b], at = io.dst_byte_off[b];
    __syncthreads();  // everybody has read src_bit_off[b] before it may be overwritten below
    if (io.dst_bytes && at + packed_size(nbits, FRAMED) > io.dst_bytes) {
        if (threadIdx.x == 0 && io.status) io.status[b] = SCL_ST_OVERFLOW;
        return;
    }
    if (threadIdx.x == 0 && io.new_bit_off) io.new_bit_off[b] = 8 * at + packed_lead_bits(nbits, FRAMED);
    uint8_t *d = io.dst + at;
    const uint32_t num_pad = FRAMED ? (uint32_t)((8 - (nbits + 3) % 8) % 8) : 0u;
    const uint64_t lead = FRAMED ? 3 + num_pad : 0;
    const uint64_t payload_bytes = FRAMED ? (nbits + lead) >> 3 : (nbits + 7) >> 3;
    if (FRAMED) {
        if (threadIdx.x < 4) d[threadIdx.x] = (uint8_t)(payload_bytes >> (8 * (3 - threadIdx.x)));  // u32 big-endian
        d += 4;
    }
    for (uint64_t i = threadIdx.x; i < payload_bytes; i += kThreads) d[i] = (uint8_t)pack_payload_byte<FRAMED, true>(io.src, off, nbits, num_pad, lead, i);
}

// second generation: one warp per block, 16 output bytes per lane and step (pack_block_warp)
constexpr int kPackWarps = 8;
template <bool FRAMED>
__global__ void __launch_bounds__(kPackWarps * 32) pack_v2_kernel(PackIo io) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp0 = (uint64_t)blockIdx.x * kPackWarps + (threadIdx.x >> 5), n_warps = (uint64_t)gridDim.x * kPackWarps;
    for (uint64_t b = warp0; b < io.n_blocks; b += n_warps) {
        if (io.status && io.status[b] != SCL_ST_OK) continue;
        const uint64_t off = io.src_bit_off[b], nbits = io.bit_len[b], at = io.dst_byte_off[b];
        __syncwarp();
        if (io.dst_bytes && at + packed_size(nbits, FRAMED) > io.dst_bytes) {
            if (lane == 0 && io.status) io.status[b] = SCL_ST_OVERFLOW;
            continue;
        }
        if (lane == 0 && io.new_bit_off) io.new_bit_off[b] = 8 * at + packed_lead_bits(nbits, FRAMED);
        pack_block_warp<FRAMED, true>(io.src, off, nbits, io.dst + at, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Byte histograms: DataBlock.get_counts (scl/core/data_block.py:37-64) for a batch of blocks, the
// step before the coders.  One warp per block; each lane streams 16-byte
// pieces of the row (coalesced 512 B per warp load) into one of kHistSub lane-interleaved sub-histograms
// in shared memory (cuts same-bin atomic contention on skewed data), then the warp writes the
// block's 256 counts and/or folds them into a grid-wide total.
// ------------------------------------------------------------------------------------------------
constexpr int kHistWarps = 8;
constexpr int kHistSub = 4;  // sub-histograms per warp (dynamic shared memory: kHistWarps * kHistSub KiB)
__global__ void __launch_bounds__(kHistWarps * 32) histogram_kernel(const uint8_t *__restrict__ sym, uint64_t sym_stride,
                                                                    const uint32_t *__restrict__ sizes, uint32_t block_len, uint64_t n_blocks,
                                                                    uint32_t *__restrict__ counts, unsigned long long *__restrict__ total) {
    extern __shared__ __align__(16) uint32_t s_hist_raw[];
    uint32_t(*s_hist)[kHistSub][256] = (uint32_t(*)[kHistSub][256])s_hist_raw;
    __shared__ uint32_t s_total[256];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_total[i] = 0;
    __syncthreads();
    uint32_t(*h)[256] = s_hist[warp];
    for (uint64_t b = (uint64_t)blockIdx.x * kHistWarps + warp; b < n_blocks; b += (uint64_t)gridDim.x * kHistWarps) {
        for (uint32_t i = lane; i < kHistSub * 256; i += 32) (&h[0][0])[i] = 0;
        __syncwarp();
        const uint8_t *row = sym + b * sym_stride;
        const uint32_t n = sizes ? sizes[b] : block_len;
        uint32_t *mine = h[lane & (kHistSub - 1)];
        uint32_t i = 0;
        if ((((uintptr_t)row) & 15) == 0) {
            for (uint32_t base = 0; base + 512 <= n; base += 512) {
                const uint4 q = *(const uint4 *)(row + base + lane * 16);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&mine[w[j] & 0xFFu], 1u);
                    atomicAdd(&mine[(w[j] >> 8) & 0xFFu], 1u);
                    atomicAdd(&mine[(w[j] >> 16) & 0xFFu], 1u);
                    atomicAdd(&mine[w[j] >> 24], 1u);
                }
                i = base + 512;
            }
        }
        for (uint32_t k = i + lane; k < n; k += 32) atomicAdd(&mine[row[k]], 1u);
        __syncwarp();
        for (uint32_t v = lane; v < 256; v += 32) {
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < kHistSub; ++j) c += h[j][v];
            if (counts) counts[b * 256 + v] = c;
            if (total && c) atomicAdd(&s_total[v], c);
        }
        __syncwarp();
    }
    __syncthreads();
    if (total)
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
            if (s_total[i]) atomicAdd(&total[i], (unsigned long long)s_total[i]);
}

// Second form: ONE LANE PER BLOCK, like the coders.  Every lane owns 256 32-bit counters in shared memory in the
// [counter][lane] interleave (its bank is its lane number: the increments are fire-and-forget shared-memory reductions
// that never conflict and never wait for each other), reads its own row in whole 32-byte sectors and, at the end of a
// block, writes its column out as the block's counts.  Grid totals are the column sums of what a lane has counted over ALL
// its blocks when no per-block counts are asked for (no zeroing in between), otherwise a second small kernel sums the
// per-block counts.  Rows must be 32-byte aligned.
constexpr uint32_t kHist2Warps = 7;  // 7 x 32 KiB of counters
__global__ void __launch_bounds__(kHist2Warps * 32, 1) histogram_lanes_kernel(const uint8_t *__restrict__ sym, uint64_t sym_stride, const uint32_t *__restrict__ sizes,
                                                                              uint32_t block_len, uint64_t n_blocks, uint32_t *__restrict__ counts,
                                                                              unsigned long long *__restrict__ total) {
    extern __shared__ __align__(128) uint32_t s_cnt[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const saddr_t mine = saddr_of(s_cnt + warp * (256 * 32)) + lane * 4;  // counter v of this lane at mine + 128 v
    auto zero = [&]() {
#pragma unroll 8
        for (uint32_t v = 0; v < 256; ++v) sts32(mine + v * 128, 0u);
    };
    auto bump = [&](uint32_t byte) {  // byte value already multiplied by 128
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(mine + byte) : "memory");
    };
    zero();
    const uint64_t warps_total = (uint64_t)gridDim.x * kHist2Warps;
    for (uint64_t task = (uint64_t)blockIdx.x * kHist2Warps + warp; task * 32 < n_blocks; task += warps_total) {
        const uint64_t b = task * 32 + lane;
        const bool active = b < n_blocks;
        const uint32_t n = active ? (sizes ? sizes[b] : block_len) : 0u;
        const uint8_t *row = sym + (active ? b : 0) * sym_stride;
        uint32_t i = 0;
        for (; i + 32 <= n; i += 32) {
            const u32x8 q = ld_sector32(row + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t w = q.v[j];
                bump((w & 0xFFu) << 7);
                bump((w >> 1) & (0xFFu << 7));
                bump((w >> 9) & (0xFFu << 7));
                bump((w >> 17) & (0xFFu << 7));
            }
        }
        for (; i < n; ++i) bump((uint32_t)row[i] << 7);
        if (counts) {
            if (active) {
                uint32_t *dst = counts + b * 256;
#pragma unroll 4
                for (uint32_t v = 0; v < 256; v += 4) {
                    uint4 o;
                    o.x = lds32(mine + (v + 0) * 128);
                    o.y = lds32(mine + (v + 1) * 128);
                    o.z = lds32(mine + (v + 2) * 128);
                    o.w = lds32(mine + (v + 3) * 128);
                    *(uint4 *)(dst + v) = o;
                }
            }
            zero();
        }
    }
    if (total && !counts) {  // the lanes' running counts: column sums per warp, one atomic per value and warp
        __syncwarp();
        const uint32_t *col = s_cnt + warp * (256 * 32);
        for (uint32_t v = lane; v < 256; v += 32) {
            unsigned long long acc = 0;
            for (uint32_t j = 0; j < 32; ++j) acc += col[v * 32 + ((j + lane) & 31)];  // skewed: one bank per lane
            if (acc) atomicAdd(&total[v], acc);
        }
    }
}
// total[v] += sum over blocks of counts[b][v]
__global__ void __launch_bounds__(256) histogram_total_kernel(const uint32_t *__restrict__ counts, uint64_t n_blocks, unsigned long long *__restrict__ total) {
    const uint64_t per = (n_blocks + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per, hi = lo + per < n_blocks ? lo + per : n_blocks;
    unsigned long long acc = 0;
    for (uint64_t b = lo; b < hi; ++b) acc += counts[b * 256 + threadIdx.x];
    if (acc) atomicAdd(&total[threadIdx.x], acc);
}

}  // namespace scl

// ====================================================================================================
// C-ABI
// ====================================================================================================
using namespace scl;

struct scl_coder {
    scl_params p;
    RansHost *rans = nullptr;
    TansHost *tans = nullptr;
    RangeHost *range = nullptr;
    AecHost *aec = nullptr;
    // device tables
    RansEnc32 *d_enc32 = nullptr;
    RansEnc32 *d_enc32x8 = nullptr;  // [256][kEncTabCopies] bank-rotated replicas for the v2 encoder
    int n_sm = 0;
    bool v2_ok = false;              // parameter set eligible for the second-generation kernels
    RansDec32 *d_dec32 = nullptr;
    uint32_t dec32_bytes = 0;
    RansGeneric *d_gen = nullptr;
    TansSym *d_tsym = nullptr;
    TansSym8 *d_tsymx8 = nullptr;  // bank-rotated replicas (16 x 8 bytes per byte value) for the v2 tANS encoder
    uint32_t *d_tenc = nullptr, *d_tdec = nullptr;
    uint32_t ttab_bytes = 0;
    RangeTab *d_range = nullptr;
    uint8_t *d_range_lut = nullptr;
    uint32_t range_lut_bytes = 0;
    uint32_t *d_range_enc_rep = nullptr;  // [byte value][lane] replicas of RangeHost::enc_tab for range_encode_v2_kernel
    uint32_t *d_range_dec_lut = nullptr;  // RangeHost::dec_lut
    uint32_t range_dec_lut_bytes = 0;
    AecTab *d_aec = nullptr;
    // test hook (scl_coder_debug_path): 1 = first-generation kernels, 2 = v2 decode with per-lane sector stores,
    // 3 / 4 = v2 decode always / never in the pipe-balanced form.  Per handle: no process-global state.
    int debug_mode = 0;
    uint64_t *d_trace = nullptr;  // scl_coder_debug_trace
    uint64_t trace_words = 0;
};

static thread_local char g_cuda_err[256] = "";
static int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return SCL_E_CUDA;
}
#define SCL_CUDA(call)                                     \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

extern "C" const char *scl_last_cuda_error(void) { return g_cuda_err; }
extern "C" const char *scl_version(void) { return "scl_b200 0.1 (sm_100a)"; }

template <typename T>
static int upload(T **dptr, const void *src, size_t bytes, size_t alloc_bytes, cudaStream_t s) {
    SCL_CUDA(cudaMalloc((void **)dptr, alloc_bytes));
    if (alloc_bytes > bytes) SCL_CUDA(cudaMemsetAsync(*dptr, 0, alloc_bytes, s));
    SCL_CUDA(cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, s));
    return SCL_E_OK;
}
static size_t round16(size_t x) { return (x + 15) & ~(size_t)15; }

// smem budget for staging tANS tables next to the static tables (227 KiB per CTA on sm_100a)
static const uint32_t kTansSmemTableMax = 160 * 1024;

extern "C" void scl_coder_destroy(scl_coder *c) {
    if (!c) return;
    cudaFree(c->d_enc32);
    cudaFree(c->d_enc32x8);
    cudaFree(c->d_dec32);
    cudaFree(c->d_gen);
    cudaFree(c->d_tsym);
    cudaFree(c->d_tsymx8);
    cudaFree(c->d_tenc);
    cudaFree(c->d_tdec);
    cudaFree(c->d_range);
    cudaFree(c->d_range_lut);
    cudaFree(c->d_range_enc_rep);
    cudaFree(c->d_range_dec_lut);
    cudaFree(c->d_aec);
    delete c->rans;
    delete c->tans;
    delete c->range;
    delete c->aec;
    delete c;
}

extern "C" int scl_coder_create(const scl_params *params, const uint8_t *alphabet, const uint64_t *freq, uint32_t n_sym, void *stream,
                                scl_coder **out) {
    if (!params || !freq || !out) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    scl_coder *c = new (std::nothrow) scl_coder();
    if (!c) return SCL_E_INVALID;
    c->p = *params;
    int rc = SCL_E_OK;
    switch (params->coder) {
    case SCL_CODER_RANS: {
        c->rans = new RansHost();
        rc = c->rans->init(*params, alphabet, freq, n_sym);
        if (rc) break;
        RansHost &r = *c->rans;
        rc = upload(&c->d_gen, &r.gen, sizeof(RansGeneric), sizeof(RansGeneric), s);
        if (!rc && r.enc32) rc = upload(&c->d_enc32, r.enc_tab.data(), sizeof(RansEnc32) * 256, sizeof(RansEnc32) * 256, s);
        if (!rc && r.enc32) {
            std::vector<RansEnc32> rep(256 * kEncTabCopies);
            for (uint32_t sy = 0; sy < 256; ++sy)
                for (uint32_t j = 0; j < kEncTabCopies; ++j) rep[sy * kEncTabCopies + j] = r.enc_tab[sy];
            rc = upload(&c->d_enc32x8, rep.data(), rep.size() * sizeof(RansEnc32), rep.size() * sizeof(RansEnc32), s);
            if (!rc) {
                cudaError_t e = cudaStreamSynchronize(s);  // `rep` dies here
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
            }
        }
        if (!rc) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, dev);
            c->v2_ok = r.max_bits_per_symbol <= kFastMaxBitsPerSym && (r.c.NBO == 1 || r.c.NBO == 8) && c->n_sm > 0;
        }
        if (!rc && r.dec32) {
            c->dec32_bytes = (uint32_t)round16(r.dec_lut.size() * sizeof(RansDec32));
            rc = upload(&c->d_dec32, r.dec_lut.data(), r.dec_lut.size() * sizeof(RansDec32), c->dec32_bytes, s);
        }
        break;
    }
    case SCL_CODER_TANS: {
        c->tans = new TansHost();
        rc = c->tans->init(*params, alphabet, freq, n_sym);
        if (rc) break;
        TansHost &t = *c->tans;
        uint32_t *d_rows = nullptr;
        rc = upload(&c->d_gen, &t.r.gen, sizeof(RansGeneric), sizeof(RansGeneric), s);
        if (!rc) rc = upload(&c->d_tsym, t.sym_tab.data(), sizeof(TansSym) * 256, sizeof(TansSym) * 256, s);
        if (!rc) rc = upload(&d_rows, t.row_of_idx.data(), sizeof(uint32_t) * n_sym, sizeof(uint32_t) * n_sym, s);
        std::vector<TansSym8> trep(256 * kTansTabCopies);
        for (uint32_t sy = 0; sy < 256; ++sy)
            for (uint32_t j = 0; j < kTansTabCopies; ++j) trep[sy * kTansTabCopies + j] = t.sym_tab8[sy];
        static_assert(256 * kTansTabCopies * sizeof(TansSym8) == kEncTabBytes, "both encoders stage a 32 KiB symbol table");
        if (!rc) rc = upload(&c->d_tsymx8, trep.data(), trep.size() * sizeof(TansSym8), trep.size() * sizeof(TansSym8), s);
        if (rc) {
            cudaStreamSynchronize(s);  // the uploads issued so far read host vectors that die with this scope
            cudaFree(d_rows);
            break;
        }
        {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, dev);
            // v2 needs both L-entry tables in shared memory next to the per-warp buffers
            c->v2_ok = t.r.max_bits_per_symbol <= kFastMaxBitsPerSym && t.r.c.L * 4 <= 64 * 1024 && c->n_sm > 0 && t.r.c.NSB <= 32;
        }
        c->ttab_bytes = (uint32_t)round16(t.r.c.L * 4);
        cudaError_t e = cudaMalloc((void **)&c->d_tenc, c->ttab_bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_tdec, c->ttab_bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_tenc, 0, c->ttab_bytes, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->d_tdec, 0, c->ttab_bytes, s);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "tANS table alloc");
            cudaFree(d_rows);
            break;
        }
        uint32_t grid = (uint32_t)((t.r.c.L + 255) / 256);
        tans_build_kernel<<<grid, 256, 0, s>>>(c->d_gen, t.r.c, d_rows, c->d_tenc, c->d_tdec);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_rows);
        if (e != cudaSuccess) rc = cuda_fail(e, "tans_build_kernel");
        break;
    }
    case SCL_CODER_RANGE: {
        c->range = new RangeHost();
        rc = c->range->init(*params, alphabet, freq, n_sym);
        if (!rc) rc = upload(&c->d_range, &c->range->t, sizeof(RangeTab), sizeof(RangeTab), s);
        if (!rc) {
            c->range_lut_bytes = (uint32_t)round16(c->range->lut.size());
            rc = upload(&c->d_range_lut, c->range->lut.data(), c->range->lut.size(), c->range_lut_bytes, s);
        }
        if (!rc && c->range->v2) {
            const RangeHost &rh = *c->range;
            std::vector<uint32_t> rep(256 * 32);
            for (uint32_t sy = 0; sy < 256; ++sy)
                for (uint32_t l = 0; l < 32; ++l) rep[sy * 32 + l] = rh.enc_tab[sy];
            rc = upload(&c->d_range_enc_rep, rep.data(), rep.size() * 4, rep.size() * 4, s);
            c->range_dec_lut_bytes = (uint32_t)round16(rh.dec_lut.size() * 4);
            if (!rc) rc = upload(&c->d_range_dec_lut, rh.dec_lut.data(), rh.dec_lut.size() * 4, c->range_dec_lut_bytes, s);
            if (!rc) {
                cudaError_t e = cudaStreamSynchronize(s);  // `rep` dies here
                if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
            }
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, dev);
            c->v2_ok = c->n_sm > 0;
        }
        break;
    }
    case SCL_CODER_AEC: {
        c->aec = new AecHost();
        rc = c->aec->init(*params, alphabet, freq, n_sym);
        if (!rc) rc = upload(&c->d_aec, &c->aec->t, sizeof(AecTab), sizeof(AecTab), s);
        break;
    }
    default:
        rc = SCL_E_INVALID;
    }
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(s);  // host staging buffers die with this call
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc) {
        scl_coder_destroy(c);
        return rc;
    }
    *out = c;
    return SCL_E_OK;
}

extern "C" uint64_t scl_coder_max_encoded_bytes(const scl_coder *c, uint64_t block_len) {
    uint64_t bits = 0;
    if (c->rans) bits = c->rans->max_encoded_bits(block_len);
    if (c->tans) bits = c->tans->r.max_encoded_bits(block_len);
    if (c->range) bits = c->range->max_encoded_bits(block_len);
    if (c->aec) bits = c->aec->max_encoded_bits(block_len);
    uint64_t bytes = (bits + 7) / 8 + 4;  // + one spare word: the last partial word is written whole
    return (bytes + 31) & ~31ull;         // whole 32-byte sectors (the v2 encoder drains sector-wise)
}

extern "C" uint64_t scl_coder_model_words(const scl_coder *c) { return c && c->aec ? c->aec->model_words() : 0; }

extern "C" int scl_coder_path(const scl_coder *c, int decode) {
    if (c->rans) return decode ? (c->rans->dec32 ? 0 : 1) : (c->rans->enc32 ? 0 : 1);
    return 0;
}

static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return SCL_E_OK;
}

// ---- v2 launch helpers ---------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

// Warps per CTA for a persistent one-CTA-per-SM launch.  Every warp runs whole tasks, so a launch is a number of
// full rounds (n_sm * W tasks each) plus a partial one, and a round's duration grows with the warps that share the
// SM, but far less than proportionally: measured on the rANS kernels, a round of 14 warps per SM takes 0.31 ms and
// a round of 28 takes 0.44 ms, i.e. t(W) ~ a + b W with a / b ~ 17.  The W that minimises the modelled time wins
// (ties: more warps).  Counting only the waste of the last round -- what this did before -- picks 12 warps for
// 65 536 tasks (37 rounds waste 0.3 % of the last one, 16 rounds of 28 warps waste 1.2 %) and runs 45 % slower.
static void pick_launch(uint32_t n_tasks, int n_sm, uint32_t max_w, uint32_t *grid, uint32_t *warps) {
    if (n_tasks <= (uint32_t)n_sm * 8) {  // tiny batch: spread tasks over SMs, few warps each
        uint32_t w = (n_tasks + n_sm - 1) / n_sm;
        *warps = w ? w : 1;
        *grid = (n_tasks + *warps - 1) / *warps;
        return;
    }
    const uint64_t kFixed = 17;  // a / b of the round-time model
    uint64_t best = ~0ull;
    uint32_t bw = max_w;
    for (uint32_t w = 8; w <= max_w; ++w) {
        const uint64_t slots = (uint64_t)n_sm * w, full = n_tasks / slots, rem = n_tasks % slots;
        // the partial round fills CTAs one after the other (task = CTA * W + warp): as long as one CTA is full it
        // lasts as long as a full round, whatever the number of idle SMs
        const uint64_t cost = full * (kFixed + w) + (rem ? kFixed + (rem < w ? rem : w) : 0);
        if (cost <= best) {
            best = cost;
            bw = w;
        }
    }
    *warps = bw;
    *grid = (uint32_t)n_sm;
}

extern "C" void scl_coder_debug_path(scl_coder *c, int mode) {
    if (c) c->debug_mode = mode;
}
// debug_mode: low 4 bits = the path selection above; bit 8 = split batches at 64 MiB of rows instead of 2^30 blocks (so
// that tests reach the multi-launch path).  Copy pool of the packed encoder (tools/sweep_knobs.py; tests run every
// variant): bit 5 = no staging rings (every copy through registers), bit 6 / bit 7 = pieces of at most 512 / 1024 bytes,
// bits 12-15 = number of dedicated copy warps (0 = kCopyWarps).
static inline int dbg_path(const scl_coder *c) { return c->debug_mode & 15; }
static inline bool force_v1(const scl_coder *c) { return dbg_path(c) == 1; }
// Warps per CTA of the 8-bit-counter arithmetic coder: CTAs of 4 warps (five resident per SM = 20 warps), or ONE CTA of
// 24 warps per SM when the batch is at least eight such waves deep (a wave of big CTAs ends as one: +3-4 % at
// 1 048 576 blocks, -10 % at 262 144: profiles/r3e_aec_warps.jsonl).  debug_mode bits 16-20 override (tools/measure_aec_warps.py).
static inline uint32_t aec8_warps(const scl_coder *c, uint64_t n_blocks) {
    const uint32_t w = (uint32_t)(c->debug_mode >> 16) & 31u;
    if (w >= 1 && w <= 24) return w;
    const uint64_t n_sm = c->n_sm > 0 ? (uint64_t)c->n_sm : 148;
    return n_blocks >= n_sm * 24 * 32 * 8 ? 24u : 4u;
}
extern "C" void scl_coder_debug_trace(scl_coder *c, uint64_t *d_trace, uint64_t n_words) {
    if (c) {
        c->d_trace = d_trace;
        c->trace_words = n_words;
    }
}

static uint32_t max_warps_for(size_t per_warp, size_t fixed) {
    size_t avail = 227 * 1024 - 1024 - fixed;  // 227 KiB per CTA minus slack for static smem / barriers
    uint32_t w = (uint32_t)(avail / per_warp);
    return w > kMaxWarps ? kMaxWarps : w;
}

// look-back words a packed launch may need: one per (round, CTA); rounds * grid <= tasks + grid
static uint64_t packed_state_words(uint64_t n_blocks) { return (n_blocks + 31) / 32 + 4096; }

// The kernels index rows with 32-bit TMA coordinates and 32-bit task numbers: batches of more than 2^30 blocks are
// coded by several launches of whole rounds (the fused encoder's running prefix continues across them).
constexpr uint64_t kLaunchMaxBlocks = 1ull << 30;
static uint64_t tasks_per_launch(const scl_coder *c, uint64_t n_tasks, uint64_t row_bytes, uint64_t tasks_per_round) {
    (void)row_bytes;
    const uint64_t max_tasks = (c->debug_mode & 256) ? (64ull << 20) / row_bytes / 32 : kLaunchMaxBlocks / 32;  // test hook: split at 64 MiB of rows
    if (n_tasks <= max_tasks) return n_tasks;
    const uint64_t t = max_tasks / tasks_per_round * tasks_per_round;
    return t < tasks_per_round ? tasks_per_round : t;
}

template <int KIND, uint32_t NBO>
static int launch_encode_v2(const scl_coder *c, const RansConst &rc, const void *tab8, const uint32_t *tab2, uint32_t tab2_bytes,
                            const BlockIo &io, const PackedOut *packed, cudaStream_t s) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return -1;
    const uint32_t n_tasks_all = (uint32_t)((io.n_blocks + 31) / 32);
    uint32_t grid, warps;
    size_t fixed = kEncTabBytes + tab2_bytes + (kMaxWarps * kTileStages + 1) * sizeof(uint64_t) + 2048 + (packed ? 2048 : 0);  // PACKED: the static PackCtl block
    const uint32_t copy_warps = packed ? ((c->debug_mode >> 12) & 15 ? (uint32_t)((c->debug_mode >> 12) & 15) : kCopyWarps) : 0u;
    uint32_t max_w = max_warps_for(kEncWarpSmem, fixed);
    if (max_w > 32 - copy_warps) max_w = 32 - copy_warps;
    pick_launch(n_tasks_all, c->n_sm, max_w, &grid, &warps);
    const uint64_t per_round = (uint64_t)grid * warps;
    const uint64_t chunk_tasks = tasks_per_launch(c, n_tasks_all, io.sym_stride, per_round);
    size_t smem = (size_t)warps * kEncWarpSmem + kEncTabBytes + tab2_bytes + (warps * kTileStages + 1) * sizeof(uint64_t) + 2048;
    PackedOut po{};
    if (packed) {
        po = *packed;
        po.copy_warps = copy_warps;
        // staging rings for the copy warps out of whatever shared memory the coder leaves (227 KiB - 1 KiB static)
        const size_t free_smem = 226 * 1024 > smem + 16 ? 226 * 1024 - smem - 16 : 0;
        // the largest pieces (fewest instructions per byte) of which at least two fit; 512-byte pieces cost more
        // instructions than they save latency (profiles/r3i_sweep.jsonl: 1.128 vs 1.103 ms per GiB without a ring), so
        // a coder that leaves less than ~9 KB free (tANS with its 16 KB encode table) copies through registers
        const size_t per_warp = free_smem / copy_warps > 528 ? free_smem / copy_warps - 528 : 0;  // minus the stream table and padding
        uint32_t stages = 0;
        const uint32_t pb_min = (c->debug_mode & 64) ? 512u : 1024u;
        for (uint32_t pb = (c->debug_mode & 64) ? 512u : (c->debug_mode & 128) ? 1024u : 2048u; pb >= pb_min; pb >>= 1) {
            po.copy_piece_bytes = pb;
            stages = (uint32_t)(per_warp / (pb + kCopyOverlapBytes + 8));
            if (stages >= 2) break;
        }
        if (stages > 8) stages = 8;
        po.copy_stages = (stages < 2 || (c->debug_mode & 32)) ? 0u : stages;
        po.helper_ring = (c->debug_mode & 32) ? 0u : 1u;
        smem += 16 + (size_t)copy_warps * copy_ring_bytes(po.copy_stages, po.copy_piece_bytes);
        po.trace = c->d_trace && c->trace_words >= (uint64_t)grid * 32 * kTraceWords ? c->d_trace : nullptr;
        // one look-back word per (round, CTA) of the WHOLE batch: the launches of a split batch are whole rounds, so the
        // numbering (and with it the running prefix) simply continues from launch to launch
        const uint64_t rounds = (n_tasks_all + per_round - 1) / per_round;
        cudaError_t e = cudaMemsetAsync(po.cta_state, 0, rounds * grid * sizeof(uint64_t), s);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    }
    for (uint64_t t0 = 0; t0 < n_tasks_all; t0 += chunk_tasks) {
        const uint64_t b0 = t0 * 32;
        BlockIo cio = io;
        cio.sym = io.sym + b0 * io.sym_stride;
        cio.sizes = io.sizes ? io.sizes + b0 : nullptr;
        cio.n_blocks = io.n_blocks - b0 < chunk_tasks * 32 ? io.n_blocks - b0 : chunk_tasks * 32;
        cio.bit_off = io.bit_off + b0;
        cio.bit_len = io.bit_len + b0;
        cio.status = io.status + b0;
        cio.block0 = io.block0 + b0;
        const uint32_t n_tasks = (uint32_t)((cio.n_blocks + 31) / 32);
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {cio.block_len, cio.n_blocks};
        cuuint64_t gstr[1] = {cio.sym_stride};
        cuuint32_t box[2] = {kTileCols, 32};
        cuuint32_t estr[2] = {1, 1};
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)cio.sym, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return t0 ? SCL_E_CUDA : -1;  // (-1: nothing launched yet, the caller may fall back to the first generation)
        if (packed) {
            po.byte_off = packed->byte_off + b0;
            po.g_base = (t0 / per_round) * grid;
        }
        cudaError_t e;
#define SCL_LAUNCH_ENC2(CHK, PK, RG)                                                                                                   \
    do {                                                                                                                               \
        e = cudaFuncSetAttribute(fast_encode_v2_kernel<KIND, NBO, CHK, PK, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");                                                              \
        fast_encode_v2_kernel<KIND, NBO, CHK, PK, RG><<<grid, (warps + copy_warps) * 32, smem, s>>>(tmap, tab8, tab2, tab2_bytes, rc, cio, n_tasks, po); \
    } while (0)
#define SCL_LAUNCH_ENC(CHK, PK)              \
    do {                                     \
        if (cio.sizes)                       \
            SCL_LAUNCH_ENC2(CHK, PK, true);  \
        else                                 \
            SCL_LAUNCH_ENC2(CHK, PK, false); \
    } while (0)
        if (rc.check_sym) {
            if (packed)
                SCL_LAUNCH_ENC(true, true);
            else
                SCL_LAUNCH_ENC(true, false);
        } else {
            if (packed)
                SCL_LAUNCH_ENC(false, true);
            else
                SCL_LAUNCH_ENC(false, false);
        }
#undef SCL_LAUNCH_ENC
#undef SCL_LAUNCH_ENC2
        int rc2 = check_launch("fast_encode_v2_kernel");
        if (rc2) return rc2;
    }
    return SCL_E_OK;
}

template <int KIND, uint32_t NBO>
static int launch_decode_v2(const scl_coder *c, const RansConst &rc, const uint32_t *lut, uint32_t lut_bytes, const DecodeIo &io,
                            cudaStream_t s) {
    const uint32_t n_tasks_all = (uint32_t)((io.n_blocks + 31) / 32);
    uint32_t grid, warps;
    pick_launch(n_tasks_all, c->n_sm, max_warps_for(kDecWarpSmem + kDecTileBytes, lut_bytes), &grid, &warps);
    const uint64_t chunk_tasks = tasks_per_launch(c, n_tasks_all, io.sym_stride, (uint64_t)grid * warps);  // > 2^30 blocks: several launches
    size_t smem = (size_t)warps * (kDecWarpSmem + kDecTileBytes) + lut_bytes;
    // pipe-balanced instruction selection pays when the SMs are full (>= 2 rounds of warps); small batches are
    // latency-bound and keep the shorter dependency chain
    const bool bal = dbg_path(c) == 3 ? true : dbg_path(c) == 4 ? false : n_tasks_all >= 24u * c->n_sm;
    auto kern = bal ? fast_decode_v2_kernel<KIND, NBO, true> : fast_decode_v2_kernel<KIND, NBO, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
    PFN_tmapEncodeTiled enc = tmap_encoder();
    for (uint64_t t0 = 0; t0 < n_tasks_all; t0 += chunk_tasks) {
        const uint64_t b0 = t0 * 32;
        DecodeIo cio = io;  // the coded buffer and its bit offsets stay global; the per-block arrays move to this launch's first block
        cio.n_blocks = io.n_blocks - b0 < chunk_tasks * 32 ? io.n_blocks - b0 : chunk_tasks * 32;
        cio.bit_off = io.bit_off + b0;
        cio.bit_len = io.bit_len ? io.bit_len + b0 : nullptr;
        cio.sym = io.sym + b0 * io.sym_stride;
        cio.sizes = io.sizes + b0;
        cio.consumed = io.consumed + b0;
        cio.status = io.status + b0;
        // output tensor map for the TMA tile stores: rows = blocks, inner = the row capacity
        CUtensorMap omap;
        memset(&omap, 0, sizeof(omap));
        uint32_t use_tiles = 0;
        if (enc && cio.sym_stride >= kTileCols && dbg_path(c) != 2) {
            cuuint64_t gdim[2] = {cio.sym_stride, cio.n_blocks};
            cuuint64_t gstr[1] = {cio.sym_stride};
            cuuint32_t box[2] = {kTileCols, 32};
            cuuint32_t estr[2] = {1, 1};
            use_tiles = enc(&omap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)cio.sym, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        kern<<<grid, warps * 32, smem, s>>>(omap, use_tiles, lut, lut_bytes, rc, cio, (uint32_t)((cio.n_blocks + 31) / 32));
        int rc2 = check_launch("fast_decode_v2_kernel");
        if (rc2) return rc2;
    }
    return SCL_E_OK;
}

static int launch_range_encode_v2(const scl_coder *c, const BlockIo &io, cudaStream_t s) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return -1;
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {io.block_len, io.n_blocks};
    cuuint64_t gstr[1] = {io.sym_stride};
    cuuint32_t box[2] = {kTileCols, 32};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)io.sym, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;
    uint32_t n_tasks = (uint32_t)((io.n_blocks + 31) / 32), grid, warps;
    pick_launch(n_tasks, c->n_sm, max_warps_for(kEncWarpSmem, kRangeEncTabBytes + 2048 + 8 * (kMaxWarps * kTileStages + 1)), &grid, &warps);
    size_t smem = 2048 + (size_t)warps * kEncWarpSmem + kRangeEncTabBytes + 8 * (warps * kTileStages + 1);
    const uint32_t shift = c->range->c.t_shift;
    cudaError_t e;
    if (c->range->c.n_sym < 256) {  // some byte values are not in the alphabet: check every symbol
        e = cudaFuncSetAttribute(range_encode_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        range_encode_v2_kernel<true><<<grid, warps * 32, smem, s>>>(tmap, c->d_range_enc_rep, shift, io, n_tasks);
    } else {
        e = cudaFuncSetAttribute(range_encode_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        range_encode_v2_kernel<false><<<grid, warps * 32, smem, s>>>(tmap, c->d_range_enc_rep, shift, io, n_tasks);
    }
    return check_launch("range_encode_v2_kernel");
}

// `packed` non-NULL: the caller wants the contiguous output; *fused is set when the launched kernel produced it
// itself (second-generation rANS / tANS), otherwise the streams are in their slots as usual
static int encode_blocks_impl(const scl_coder *c, const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes, uint32_t block_len,
                              uint64_t n_blocks, uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_bit_offset,
                              uint64_t *d_out_bit_len, uint64_t *d_model, uint32_t *d_status, const PackedOut *packed, bool *fused,
                              void *stream) {
    if (fused) *fused = false;
    if (!c || !d_out || !d_out_bit_offset || !d_out_bit_len || !d_status) return SCL_E_INVALID;
    if (n_blocks == 0) return SCL_E_OK;
    if (!d_sym && (d_sizes || block_len)) return SCL_E_INVALID;
    if ((out_stride & 15) || (((uintptr_t)d_out) & 15) || out_stride < 16) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    BlockIo io{d_sym, sym_stride, d_sizes, block_len, n_blocks, d_out, out_stride, d_out_bit_offset, d_out_bit_len, d_status};
    uint32_t grid = (uint32_t)((n_blocks + kThreads - 1) / kThreads);
    if (c->rans) {
        const RansHost &r = *c->rans;
        // second-generation kernel: uniform block length, TMA-compatible input, sector-aligned output
        if (r.enc32 && c->v2_ok && !force_v1(c) && block_len >= kTileCols && (sym_stride % 16) == 0 &&
            (((uintptr_t)d_sym) & 15) == 0 && (out_stride % 32) == 0 && (((uintptr_t)d_out) & 31) == 0 && n_blocks < (1ull << 36) &&
            (uint64_t)block_len * kFastMaxBitsPerSym < (1ull << 31)) {
            int rc2 = r.c.NBO == 1 ? launch_encode_v2<0, 1>(c, r.c, c->d_enc32x8, nullptr, 0, io, packed, s)
                                   : launch_encode_v2<0, 8>(c, r.c, c->d_enc32x8, nullptr, 0, io, packed, s);
            if (rc2 >= 0) {  // < 0: tensor map could not be built -> first-generation kernel
                if (fused) *fused = packed != nullptr;
                return rc2;
            }
        }
        if (r.enc32) {
            if (r.c.check_sym)
                rans32_encode_kernel<true><<<grid, kThreads, 0, s>>>(c->d_enc32, r.c, io);
            else
                rans32_encode_kernel<false><<<grid, kThreads, 0, s>>>(c->d_enc32, r.c, io);
        } else {
            rans64_encode_kernel<<<grid, kThreads, 0, s>>>(c->d_gen, r.c, io);
        }
        return check_launch("rans_encode_kernel");
    }
    if (c->tans) {
        const TansHost &t = *c->tans;
        if (c->v2_ok && !force_v1(c) && block_len >= kTileCols && (sym_stride % 16) == 0 && (((uintptr_t)d_sym) & 15) == 0 &&
            (out_stride % 32) == 0 && (((uintptr_t)d_out) & 31) == 0 && n_blocks < (1ull << 36) &&
            (uint64_t)block_len * kFastMaxBitsPerSym < (1ull << 31)) {
            int rc2 = launch_encode_v2<1, 1>(c, t.r.c, c->d_tsymx8, c->d_tenc, c->ttab_bytes, io, packed, s);
            if (rc2 >= 0) {
                if (fused) *fused = packed != nullptr;
                return rc2;
            }
        }
        if (c->ttab_bytes <= kTansSmemTableMax) {
            SCL_CUDA(cudaFuncSetAttribute(tans_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTansSmemTableMax));
            tans_encode_kernel<true><<<grid, kThreads, c->ttab_bytes, s>>>(c->d_tsym, c->d_tenc, c->ttab_bytes, t.r.c, io);
        } else {
            tans_encode_kernel<false><<<grid, kThreads, 0, s>>>(c->d_tsym, c->d_tenc, 0, t.r.c, io);
        }
        return check_launch("tans_encode_kernel");
    }
    if (c->range) {
        // block_len < 2^24: the lanes count ring words in 32 bits (* 128), and a symbol releases at most 3 bytes
        if (c->range->v2 && c->v2_ok && !force_v1(c) && !d_sizes && block_len >= kTileCols && block_len < (1u << 24) && (sym_stride % 16) == 0 &&
            (((uintptr_t)d_sym) & 15) == 0 && (out_stride % 32) == 0 && (((uintptr_t)d_out) & 31) == 0 && n_blocks < (1ull << 36)) {
            int rc2 = launch_range_encode_v2(c, io, s);
            if (rc2 >= 0) return rc2;  // < 0: tensor map could not be built -> first-generation kernel
        }
        range_encode_kernel<<<grid, kThreads, 0, s>>>(c->d_range, c->range->c, io);
        return check_launch("range_encode_kernel");
    }
    if (c->aec && c->aec->c.model == SCL_MODEL_ORDER_K && c->aec->c.ctx_global) {
        if (!d_model) return SCL_E_INVALID;  // the table is the lanes' working storage
        const uint32_t g2 = (uint32_t)((n_blocks + kAecCtxGlobalWarps * 32 - 1) / (kAecCtxGlobalWarps * 32));
        const size_t smem = (size_t)kAecCtxGlobalWarps * c->aec->c.n_ctx * 128;
        SCL_CUDA(cudaFuncSetAttribute(aec_ctx_global_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        aec_ctx_global_encode_kernel<<<g2, kAecCtxGlobalWarps * 32, smem, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_ctx_global_encode_kernel");
    }
    if (c->aec && c->aec->c.model == SCL_MODEL_ORDER_K) {
        const uint32_t warps = aec_ctx_warps(c->aec->c);
        uint32_t g2 = (uint32_t)((n_blocks + warps * 32 - 1) / (warps * 32));
        size_t smem = (size_t)warps * c->aec->c.n_ctx * (c->aec->c.n_sym + 1) * 128;
        SCL_CUDA(cudaFuncSetAttribute(aec_ctx_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        aec_ctx_encode_kernel<<<g2, warps * 32, smem, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_ctx_encode_kernel");
    }
    if (c->aec) {
        uint32_t g = (uint32_t)((n_blocks + kAecThreads - 1) / kAecThreads);
        // 16-bit counters when no count can reach 65536: counts start at init_freq and grow by at most
        // one per coded symbol (a caller-supplied d_model may hold anything, so it takes the 32-bit tree)
        uint64_t max_init = 0;
        for (uint32_t i = 0; i < c->aec->c.n_sym; ++i) max_init = c->aec->t.init_freq[i] > max_init ? c->aec->t.init_freq[i] : max_init;
        // second generation: every counter and group total must stay below 65536
        if (!d_model && !force_v1(c) && dbg_path(c) != 5 && c->aec->model8_ok(block_len)) {  // debug path 5: keep the 16-bit counters
            const uint32_t w8 = aec8_warps(c, n_blocks);
            uint32_t g2 = (uint32_t)((n_blocks + w8 * 32 - 1) / (w8 * 32));
            size_t smem = kAec2MaskBytes + w8 * kAec8ModelBytes;
            SCL_CUDA(cudaFuncSetAttribute(aec8_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            aec8_encode_kernel<<<g2, w8 * 32, smem, s>>>(c->d_aec, c->aec->c, io);
            return check_launch("aec8_encode_kernel");
        }
        if (!d_model && 16 * max_init + block_len < 65536 && !force_v1(c)) {
            uint32_t g2 = (uint32_t)((n_blocks + kAec2Warps * 32 - 1) / (kAec2Warps * 32));
            size_t smem = kAec2MaskBytes + kAec2Warps * kAec2ModelBytes;
            SCL_CUDA(cudaFuncSetAttribute(aec2_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            aec2_encode_kernel<<<g2, kAec2Warps * 32, smem, s>>>(c->d_aec, c->aec->c, io);
            return check_launch("aec2_encode_kernel");
        }
        if (!d_model && max_init + block_len < 65536)
            aec_encode_kernel<SmemTree16><<<g, kAecThreads, SmemTree16::kBytes, s>>>(c->d_aec, c->aec->c, io, d_model);
        else
            aec_encode_kernel<SmemTree><<<g, kAecThreads, SmemTree::kBytes, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_encode_kernel");
    }
    return SCL_E_INVALID;
}

extern "C" int scl_encode_blocks(const scl_coder *c, const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes, uint32_t block_len,
                                 uint64_t n_blocks, uint8_t *d_out, uint64_t out_stride, uint64_t *d_out_bit_offset,
                                 uint64_t *d_out_bit_len, uint64_t *d_model, uint32_t *d_status, void *stream) {
    return encode_blocks_impl(c, d_sym, sym_stride, d_sizes, block_len, n_blocks, d_out, out_stride, d_out_bit_offset, d_out_bit_len, d_model,
                              d_status, nullptr, nullptr, stream);
}

static int pack_launch(const PackIo &io, bool framed, bool bytewise, cudaStream_t s);

extern "C" int scl_packed_offsets(const uint64_t *d_bit_len, const uint32_t *d_status, uint64_t n_blocks, uint32_t framed,
                                  uint64_t *d_byte_offset, uint64_t *d_bit_offset, void *stream) {
    if (!d_bit_len || !d_byte_offset) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_blocks == 0) {
        SCL_CUDA(cudaMemsetAsync(d_byte_offset, 0, sizeof(uint64_t), s));
        return SCL_E_OK;
    }
    const uint64_t n_tiles = (n_blocks + kScanTile - 1) / kScanTile;
    if (n_tiles > 0x7FFFFFFFull) return SCL_E_UNSUPPORTED;
    scan_tile_totals_kernel<<<(uint32_t)n_tiles, kScanThreads, 0, s>>>(d_bit_len, d_status, n_blocks, framed, d_byte_offset);
    scan_tile_offsets_kernel<<<1, kScanThreads, 0, s>>>(n_blocks, n_tiles, d_byte_offset);
    scan_tile_final_kernel<<<(uint32_t)n_tiles, kScanThreads, 0, s>>>(d_bit_len, d_status, n_blocks, framed, d_byte_offset, d_bit_offset);
    return check_launch("scan_tile_kernels");
}

extern "C" uint64_t scl_encode_packed_workspace_bytes(const scl_coder *c, uint64_t n_blocks) {
    (void)c;
    return packed_state_words(n_blocks) * sizeof(uint64_t);
}

extern "C" int scl_encode_blocks_packed(const scl_coder *c, const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes,
                                        uint32_t block_len, uint64_t n_blocks, uint8_t *d_scratch, uint64_t scratch_stride, uint8_t *d_dst,
                                        uint64_t dst_bytes, uint32_t framed, uint64_t *d_byte_offset, uint64_t *d_bit_offset,
                                        uint64_t *d_bit_len, uint64_t *d_model, uint32_t *d_status, void *d_workspace,
                                        uint64_t workspace_bytes, void *stream) {
    if (!c || !d_dst || !d_byte_offset || !d_bit_offset || !d_bit_len || !d_status || !d_scratch) return SCL_E_INVALID;
    if (!d_workspace || workspace_bytes < scl_encode_packed_workspace_bytes(c, n_blocks)) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (n_blocks == 0) {
        SCL_CUDA(cudaMemsetAsync(d_byte_offset, 0, sizeof(uint64_t), s));
        return SCL_E_OK;
    }
    PackedOut po{d_dst, dst_bytes, d_byte_offset, (uint64_t *)d_workspace, framed ? 1u : 0u, 0, kCopyWarps, 0, 0, 0, nullptr};
    bool fused = false;
    int rc = encode_blocks_impl(c, d_sym, sym_stride, d_sizes, block_len, n_blocks, d_scratch, scratch_stride, d_bit_offset, d_bit_len, d_model,
                                d_status, &po, &fused, stream);
    if (rc || fused) return rc;
    // every other kernel family: streams are in their slots -> offsets by scan, then the copy kernel
    rc = scl_packed_offsets(d_bit_len, d_status, n_blocks, framed, d_byte_offset, nullptr, stream);
    if (rc) return rc;
    PackIo pio{d_scratch, d_bit_offset, d_bit_len, n_blocks, d_dst, dst_bytes, d_byte_offset, d_status, d_bit_offset};
    return pack_launch(pio, framed != 0, false, s);
}

extern "C" int scl_debug_copy_only(const scl_coder *c, uint64_t n_blocks, uint8_t *d_scratch, uint64_t scratch_stride, uint8_t *d_dst,
                                   uint64_t dst_bytes, uint32_t framed, uint64_t *d_byte_offset, uint64_t *d_bit_offset, uint64_t *d_bit_len,
                                   uint32_t *d_status, uint32_t warps_per_cta, uint32_t ring_stages, void *stream) {
    if (!c || !d_scratch || !d_dst || !d_byte_offset || !d_bit_offset || !d_bit_len || !d_status || warps_per_cta < 1 || warps_per_cta > 32)
        return SCL_E_INVALID;
    BlockIo io{nullptr, 0, nullptr, 0, n_blocks, d_scratch, scratch_stride, d_bit_offset, d_bit_len, d_status, 0};
    const uint32_t piece_bytes = (ring_stages & 0x200) ? 2048u : (ring_stages & 0x100) ? 1024u : 512u;  // bits 8, 9 of ring_stages: 64- / 128-chunk pieces
    ring_stages &= 0xFF;
    PackedOut po{d_dst, dst_bytes, d_byte_offset, nullptr, framed ? 1u : 0u, 0, kCopyWarps, ring_stages, piece_bytes, 0, nullptr};
    const size_t smem = (size_t)warps_per_cta * copy_ring_bytes(ring_stages, piece_bytes);
    if (ring_stages == 1 || smem > 200 * 1024) return SCL_E_INVALID;
    SCL_CUDA(cudaFuncSetAttribute(copy_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    copy_only_kernel<<<c->n_sm > 0 ? c->n_sm : 148, warps_per_cta * 32, smem, (cudaStream_t)stream>>>(io, po, (uint32_t)((n_blocks + 31) / 32));
    return check_launch("copy_only_kernel");
}

extern "C" int scl_decode_blocks(const scl_coder *c, const uint8_t *d_in, uint64_t in_bytes, const uint64_t *d_bit_offset,
                                 const uint64_t *d_bit_len, uint64_t n_blocks, uint8_t *d_sym, uint64_t sym_stride, uint32_t *d_sizes,
                                 uint64_t *d_bits_consumed, uint64_t *d_model, uint32_t *d_status, void *stream) {
    if (!c || !d_in || !d_bit_offset || !d_sizes || !d_bits_consumed || !d_status) return SCL_E_INVALID;
    if (n_blocks == 0) return SCL_E_OK;
    if (!d_sym && sym_stride) return SCL_E_INVALID;
    if (((uintptr_t)d_in) & 15) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    DecodeIo io{d_in, in_bytes, d_bit_offset, d_bit_len, n_blocks, d_sym, sym_stride, d_sizes, d_bits_consumed, d_status};
    uint32_t grid = (uint32_t)((n_blocks + kThreads - 1) / kThreads);
    if (c->rans) {
        const RansHost &r = *c->rans;
        // sym_stride * kFastMaxBitsPerSym < 2^31: the lanes keep the stream position in 32 bits (as the encoder's guard on block_len)
        if (r.dec32 && c->v2_ok && !force_v1(c) && (((uintptr_t)d_in) & 31) == 0 && (sym_stride % 32) == 0 && (((uintptr_t)d_sym) & 31) == 0 &&
            n_blocks < (1ull << 36) && sym_stride * kFastMaxBitsPerSym < (1ull << 31))
            return r.c.NBO == 1 ? launch_decode_v2<0, 1>(c, r.c, c->d_dec32, c->dec32_bytes, io, s)
                                : launch_decode_v2<0, 8>(c, r.c, c->d_dec32, c->dec32_bytes, io, s);
        if (r.dec32)
            rans32_decode_kernel<<<grid, kThreads, c->dec32_bytes, s>>>(c->d_dec32, c->dec32_bytes, r.c, io);
        else
            rans64_decode_kernel<<<grid, kThreads, 0, s>>>(c->d_gen, r.c, io);
        return check_launch("rans_decode_kernel");
    }
    if (c->tans) {
        const TansHost &t = *c->tans;
        if (c->v2_ok && !force_v1(c) && (((uintptr_t)d_in) & 31) == 0 && (sym_stride % 32) == 0 && (((uintptr_t)d_sym) & 31) == 0 &&
            n_blocks < (1ull << 36) && sym_stride * kFastMaxBitsPerSym < (1ull << 31))
            return launch_decode_v2<1, 1>(c, t.r.c, c->d_tdec, c->ttab_bytes, io, s);
        if (c->ttab_bytes <= kTansSmemTableMax) {
            SCL_CUDA(cudaFuncSetAttribute(tans_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTansSmemTableMax));
            tans_decode_kernel<true><<<grid, kThreads, c->ttab_bytes, s>>>(c->d_tdec, c->ttab_bytes, t.r.c, io);
        } else {
            tans_decode_kernel<false><<<grid, kThreads, 0, s>>>(c->d_tdec, 0, t.r.c, io);
        }
        return check_launch("tans_decode_kernel");
    }
    if (c->range) {
        // sym_stride < 2^24 bounds the decoded size, hence the 32-bit bit position of DecLaneV2 (<= 24 bits per symbol)
        if (c->range->v2 && c->v2_ok && !force_v1(c) && (((uintptr_t)d_in) & 31) == 0 && (sym_stride % 32) == 0 && (((uintptr_t)d_sym) & 31) == 0 &&
            sym_stride < (1ull << 24) && n_blocks < (1ull << 36)) {
            const RangeHost &rh = *c->range;
            uint32_t n_tasks = (uint32_t)((n_blocks + 31) / 32), g2, warps;
            pick_launch(n_tasks, c->n_sm, max_warps_for(kDecWarpSmem, c->range_dec_lut_bytes), &g2, &warps);
            size_t smem = (size_t)warps * kDecWarpSmem + c->range_dec_lut_bytes;
            SCL_CUDA(cudaFuncSetAttribute(range_decode_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            range_decode_v2_kernel<<<g2, warps * 32, smem, s>>>(c->d_range_dec_lut, c->range_dec_lut_bytes, rh.c.t_shift, rh.c.T, rh.last_entry, io, n_tasks);
            return check_launch("range_decode_v2_kernel");
        }
        SCL_CUDA(cudaFuncSetAttribute(range_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->range_lut_bytes));
        range_decode_kernel<<<grid, kThreads, c->range_lut_bytes, s>>>(c->d_range, c->d_range_lut, c->range_lut_bytes, c->range->c, io);
        return check_launch("range_decode_kernel");
    }
    if (c->aec && c->aec->c.model == SCL_MODEL_ORDER_K && c->aec->c.ctx_global) {
        if (!d_model) return SCL_E_INVALID;
        const uint32_t g2 = (uint32_t)((n_blocks + kAecCtxGlobalWarps * 32 - 1) / (kAecCtxGlobalWarps * 32));
        const size_t smem = (size_t)kAecCtxGlobalWarps * c->aec->c.n_ctx * 128;
        SCL_CUDA(cudaFuncSetAttribute(aec_ctx_global_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        aec_ctx_global_decode_kernel<<<g2, kAecCtxGlobalWarps * 32, smem, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_ctx_global_decode_kernel");
    }
    if (c->aec && c->aec->c.model == SCL_MODEL_ORDER_K) {
        const uint32_t warps = aec_ctx_warps(c->aec->c);
        uint32_t g2 = (uint32_t)((n_blocks + warps * 32 - 1) / (warps * 32));
        size_t smem = (size_t)warps * c->aec->c.n_ctx * (c->aec->c.n_sym + 1) * 128;
        SCL_CUDA(cudaFuncSetAttribute(aec_ctx_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        aec_ctx_decode_kernel<<<g2, warps * 32, smem, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_ctx_decode_kernel");
    }
    if (c->aec) {
        uint32_t g = (uint32_t)((n_blocks + kAecThreads - 1) / kAecThreads);
        uint64_t max_init = 0;
        for (uint32_t i = 0; i < c->aec->c.n_sym; ++i) max_init = c->aec->t.init_freq[i] > max_init ? c->aec->t.init_freq[i] : max_init;
        if (!d_model && !force_v1(c) && dbg_path(c) != 5 && c->aec->model8_ok(sym_stride)) {  // decoded size <= sym_stride is enforced by the lane
            const uint32_t w8 = aec8_warps(c, n_blocks);
            uint32_t g2 = (uint32_t)((n_blocks + w8 * 32 - 1) / (w8 * 32));
            size_t smem = kAec2MaskBytes + w8 * kAec8ModelBytes;
            SCL_CUDA(cudaFuncSetAttribute(aec8_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            aec8_decode_kernel<<<g2, w8 * 32, smem, s>>>(c->d_aec, c->aec->c, io);
            return check_launch("aec8_decode_kernel");
        }
        if (!d_model && 16 * max_init + sym_stride < 65536 && !force_v1(c)) {  // decoded size <= sym_stride is enforced by the lane
            uint32_t g2 = (uint32_t)((n_blocks + kAec2Warps * 32 - 1) / (kAec2Warps * 32));
            size_t smem = kAec2MaskBytes + kAec2Warps * kAec2ModelBytes;
            SCL_CUDA(cudaFuncSetAttribute(aec2_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            aec2_decode_kernel<<<g2, kAec2Warps * 32, smem, s>>>(c->d_aec, c->aec->c, io);
            return check_launch("aec2_decode_kernel");
        }
        if (!d_model && max_init + sym_stride < 65536)
            aec_decode_kernel<SmemTree16><<<g, kAecThreads, SmemTree16::kBytes, s>>>(c->d_aec, c->aec->c, io, d_model);
        else
            aec_decode_kernel<SmemTree><<<g, kAecThreads, SmemTree::kBytes, s>>>(c->d_aec, c->aec->c, io, d_model);
        return check_launch("aec_decode_kernel");
    }
    return SCL_E_INVALID;
}

// warps of 32 lanes each take whole blocks, grid-stride: enough CTAs to fill the GPU (8 per SM), no more than the blocks need
static uint32_t pack_grid(uint64_t n_blocks) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    uint64_t want = (n_blocks + kPackWarps - 1) / kPackWarps, cap = (uint64_t)(n_sm > 0 ? n_sm : 1) * 8;
    return (uint32_t)(want < cap ? want : cap);
}

static int pack_launch(const PackIo &io, bool framed, bool bytewise, cudaStream_t s) {
    if (io.n_blocks > 0x7FFFFFFFull) return SCL_E_UNSUPPORTED;
    if (!bytewise && (((uintptr_t)io.src) & 3) == 0) {
        if (framed)
            pack_v2_kernel<true><<<pack_grid(io.n_blocks), kPackWarps * 32, 0, s>>>(io);
        else
            pack_v2_kernel<false><<<pack_grid(io.n_blocks), kPackWarps * 32, 0, s>>>(io);
        return check_launch("pack_v2_kernel");
    }
    if (framed)
        pack_kernel<true><<<(uint32_t)io.n_blocks, kThreads, 0, s>>>(io);
    else
        pack_kernel<false><<<(uint32_t)io.n_blocks, kThreads, 0, s>>>(io);
    return check_launch("pack_kernel");
}

extern "C" int scl_pack_blocks(const uint8_t *d_src, const uint64_t *d_src_bit_offset, const uint64_t *d_bit_len, uint64_t n_blocks,
                               uint8_t *d_dst, const uint64_t *d_dst_byte_offset, uint32_t flags, void *stream) {
    if (!d_src || !d_src_bit_offset || !d_bit_len || !d_dst || !d_dst_byte_offset) return SCL_E_INVALID;
    if (n_blocks == 0) return SCL_E_OK;
    PackIo io{d_src, d_src_bit_offset, d_bit_len, n_blocks, d_dst, 0, d_dst_byte_offset, nullptr, nullptr};
    return pack_launch(io, false, (flags & SCL_PACK_BYTEWISE) != 0, (cudaStream_t)stream);
}

extern "C" int scl_frame_blocks(const uint8_t *d_src, const uint64_t *d_src_bit_offset, const uint64_t *d_bit_len, uint64_t n_blocks,
                                uint8_t *d_dst, const uint64_t *d_dst_byte_offset, uint32_t flags, void *stream) {
    if (!d_src || !d_src_bit_offset || !d_bit_len || !d_dst || !d_dst_byte_offset) return SCL_E_INVALID;
    if (n_blocks == 0) return SCL_E_OK;
    PackIo io{d_src, d_src_bit_offset, d_bit_len, n_blocks, d_dst, 0, d_dst_byte_offset, nullptr, nullptr};
    return pack_launch(io, true, (flags & SCL_PACK_BYTEWISE) != 0, (cudaStream_t)stream);
}

extern "C" int scl_tans_tables_to_host(const scl_coder *c, uint32_t *enc_table, uint32_t *dec_packed, uint64_t n_entries, void *stream) {
    if (!c || !c->tans || n_entries != c->tans->r.c.L) return SCL_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    SCL_CUDA(cudaMemcpyAsync(enc_table, c->d_tenc, n_entries * 4, cudaMemcpyDeviceToHost, s));
    SCL_CUDA(cudaMemcpyAsync(dec_packed, c->d_tdec, n_entries * 4, cudaMemcpyDeviceToHost, s));
    SCL_CUDA(cudaStreamSynchronize(s));
    return SCL_E_OK;
}

extern "C" int scl_histogram_blocks(const uint8_t *d_sym, uint64_t sym_stride, const uint32_t *d_sizes, uint32_t block_len, uint64_t n_blocks,
                                    uint32_t *d_counts, uint64_t *d_total, void *stream) {
    if (!d_sym || (!d_counts && !d_total)) return SCL_E_INVALID;
    if (n_blocks == 0) return SCL_E_OK;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    // lane-per-block form: whole 32-byte sectors of every row, and enough blocks to give every lane one
    if ((sym_stride % 32) == 0 && (((uintptr_t)d_sym) & 31) == 0 && n_blocks >= (uint64_t)n_sm * kHist2Warps * 32 &&
        (n_blocks / ((uint64_t)n_sm * kHist2Warps * 32) + 2) * (uint64_t)block_len < (1ull << 32)) {  // a lane's running 32-bit counts
        const int smem2 = kHist2Warps * 256 * 32 * 4;
        SCL_CUDA(cudaFuncSetAttribute(histogram_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        histogram_lanes_kernel<<<n_sm, kHist2Warps * 32, smem2, (cudaStream_t)stream>>>(d_sym, sym_stride, d_sizes, block_len, n_blocks, d_counts,
                                                                                       (unsigned long long *)d_total);
        int rc = check_launch("histogram_lanes_kernel");
        if (rc || !(d_total && d_counts)) return rc;
        histogram_total_kernel<<<4 * n_sm, 256, 0, (cudaStream_t)stream>>>(d_counts, n_blocks, (unsigned long long *)d_total);
        return check_launch("histogram_total_kernel");
    }
    uint64_t want = (n_blocks + kHistWarps - 1) / kHistWarps;
    const uint64_t per_sm = (227 * 1024) / (kHistWarps * kHistSub * 1024 + 2048);  // resident CTAs per SM by shared memory
    uint32_t grid = (uint32_t)(want < (uint64_t)n_sm * per_sm ? want : (uint64_t)n_sm * per_sm);  // grid-stride over blocks
    // s_total is 32-bit per CTA: bound the bytes one CTA can see
    if ((n_blocks / grid + 1) * kHistWarps * (uint64_t)block_len >= (1ull << 32)) return SCL_E_UNSUPPORTED;
    const int hist_smem = kHistWarps * kHistSub * 256 * 4;
    SCL_CUDA(cudaFuncSetAttribute(histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem));
    histogram_kernel<<<grid, kHistWarps * 32, hist_smem, (cudaStream_t)stream>>>(d_sym, sym_stride, d_sizes, block_len, n_blocks, d_counts,
                                                                       (unsigned long long *)d_total);
    return check_launch("histogram_kernel");
}
