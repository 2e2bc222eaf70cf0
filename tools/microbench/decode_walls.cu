// decode_walls.cu -- where is the wall for a lane-per-block table decoder on this SM?
//
// The rANS / tANS decode kernels (csrc/scl_fast.cuh dec_step, DecLaneV2::peek32) do, per symbol and lane, one random
// 4-byte gather from a 16 KiB table in shared memory, half a read of two ring words, a handful of ALU/FMA
// operations, and per 16 symbols one 16-byte store into the output tile.  profiles/r1v put the L1/shared data
// pipe at 81 % with 5.45 wavefronts per warp-symbol.  These loops isolate the pieces at the same residency (one
// persistent CTA per SM, 28 warps), each as a per-lane DEPENDENT chain like the real thing, so that
// "the shared-memory pipe caps decode near 0.4 of the HBM roofline" is a measurement:
//   lut      : x -> lut[x mod 4096] -> x'          (32 random banks per warp instruction)
//   ring     : bit position -> two ring words -> funnel shift -> next position   (conflict-free, [word][lane])
//   lut_ring : the LUT gather every step, the ring read every second step (the decoder's schedule)
//   alu      : the decoder's arithmetic per symbol with the table entry faked in registers (no shared memory)
//   lut_ring_alu : all three, i.e. the decode step without HBM traffic, tile stores or refills
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o decode_walls decode_walls.cu && ./decode_walls
// Prints one JSON line per loop: ns per warp-step, warp-steps / s / SM, and what fraction of the HBM roofline that
// rate would be at 1.78 algorithmic bytes per symbol (the bench's Zipf table) if nothing else cost anything.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kWarps = 28;
constexpr uint32_t kLutWords = 4096;
constexpr uint32_t kRingWords = 32;  // + 1 wrap duplicate, [word][lane]

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}

template <int MODE>  // 0 lut, 1 ring, 2 lut_ring, 3 alu, 4 lut_ring_alu
__global__ void __launch_bounds__(kWarps * 32, 1) wall_kernel(const uint32_t *g_lut, uint32_t iters, uint32_t *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t *s_lut = (uint32_t *)smem;
    uint32_t *s_ring = (uint32_t *)(smem + kLutWords * 4);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < kLutWords; i += blockDim.x) s_lut[i] = g_lut[i];
    uint32_t *my_ring = s_ring + warp * (kRingWords + 1) * 32 + lane;
    uint32_t seed = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
    for (uint32_t i = 0; i <= kRingWords; ++i) {
        seed = seed * 1664525u + 1013904223u;
        my_ring[i * 32] = seed;
    }
    __syncthreads();
    const uint32_t lut = (uint32_t)__cvta_generic_to_shared(s_lut), ring = (uint32_t)__cvta_generic_to_shared(my_ring);
    uint32_t x = seed | 0x10000000u, bp = seed & 1023u, acc = 0;
    auto peek = [&](uint32_t p) {
        const uint32_t a = ring + ((p & ((kRingWords - 1) * 32)) << 2);
        const uint32_t w0 = lds32(a), w1 = lds32(a + 128);
        return __funnelshift_l(w1, w0, p);
    };
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it += 2) {
        if (MODE == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t e = lds32(lut + ((x << 2) & ((kLutWords - 1) << 2)));
                x = x * 0x9E3779B1u + e;
            }
        } else if (MODE == 1) {
            // one peek per two steps like the decoder; the consumed bit count depends on the bits read
            const uint32_t bits = peek(bp);
            const uint32_t k = 6u + (bits >> 29);
            bp += k + (k >> 1);
            acc ^= bits;
        } else if (MODE == 2) {
            uint32_t bits = peek(bp), ks = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t e = lds32(lut + ((x << 2) & ((kLutWords - 1) << 2)));
                const uint32_t k = 3u + (e & 7u);
                x = __funnelshift_l(bits, x * 0x9E3779B1u + e, k & 7);
                bits <<= k;
                ks += k;
            }
            bp += ks;
        } else {
            // the real step's arithmetic (scl_fast.cuh dec_step, plain form): x' = f * (x >> 12) + bias, byte insert,
            // k = clz-based renormalisation count, funnel in k bits
            uint32_t bits = MODE == 4 ? peek(bp) : (x ^ bp), ks = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t e;
                if (MODE == 4)
                    e = lds32(lut + ((x << 2) & ((kLutWords - 1) << 2)));
                else
                    e = (x * 0x9E3779B1u) | 0x00100000u;  // stand-in for the table entry: one IMAD instead of the LDS
                const uint32_t f = e >> 20, bias = (e >> 8) & 0xFFFu;
                x = f * (x >> 12) + bias;
                acc = __byte_perm(acc, e, h ? 0x3240 : 0x3214);
                x |= 0x00010000u;  // keep the fake state in range so that clz behaves like the decoder's
                const uint32_t k = (uint32_t)__clz((int)x) - 3u;
                x = __funnelshift_l(bits, x, k);
                bits <<= k;
                ks += k;
            }
            bp += ks;
        }
    }
    if ((x ^ bp ^ acc) == 0x12345678u) sink[blockIdx.x * blockDim.x + threadIdx.x] = x;  // keep everything alive
}

template <int MODE>
static void run(const char *name, int n_sm, const uint32_t *d_lut, uint32_t *d_sink, double sm_ghz) {
    const uint32_t iters = 1u << 16;
    const size_t smem = kLutWords * 4 + (size_t)kWarps * (kRingWords + 1) * 128;
    cudaFuncSetAttribute(wall_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        wall_kernel<MODE><<<n_sm, kWarps * 32, smem>>>(d_lut, iters, d_sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double warp_steps = (double)kWarps * iters;  // per SM
    const double ns_per_warp_step = best * 1e6 / warp_steps;
    const double sym_per_s = (double)n_sm * kWarps * 32.0 * iters / (best * 1e-3);
    printf("{\"loop\": \"%s\", \"ms\": %.4f, \"ns_per_warp_step_per_SM\": %.3f, \"cycles_per_warp_step_at_%.3f_GHz\": %.2f, "
           "\"symbols_per_s\": %.4g, \"equivalent_roofline_frac_at_1.78_B_per_symbol\": %.3f, \"cuda_error\": \"%s\"}\n",
           name, best, ns_per_warp_step, sm_ghz, ns_per_warp_step * sm_ghz, sym_per_s, sym_per_s * 1.78 / 6548.5e9, cudaGetErrorString(err));
}

int main() {
    int dev = 0, n_sm = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    std::vector<uint32_t> lut(kLutWords);
    uint32_t s = 1;
    for (auto &v : lut) {
        s = s * 1664525u + 1013904223u;
        v = s;
    }
    uint32_t *d_lut, *d_sink;
    cudaMalloc(&d_lut, kLutWords * 4);
    cudaMalloc(&d_sink, (size_t)n_sm * kWarps * 32 * 4);
    cudaMemcpy(d_lut, lut.data(), kLutWords * 4, cudaMemcpyHostToDevice);
    const double ghz = khz / 1e6;
    run<0>("lut", n_sm, d_lut, d_sink, ghz);
    run<1>("ring", n_sm, d_lut, d_sink, ghz);
    run<2>("lut_ring", n_sm, d_lut, d_sink, ghz);
    run<3>("alu", n_sm, d_lut, d_sink, ghz);
    run<4>("lut_ring_alu", n_sm, d_lut, d_sink, ghz);
    return 0;
}
