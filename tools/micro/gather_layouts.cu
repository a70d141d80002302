// Micro-benchmark (GPU box only): sector-granular gather rate of one packed depth|mask texel per point-view from a
// C2-sized frame set (500 x 1440 x 1920 uint32 texels = 5.5 GB) under different texel layouts.  Pattern = what the fused
// sweep does: a warp's 32 points fall in a ~45 x 38 px window of one frame, consecutive warps look at nearby windows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_layouts tools/micro/gather_layouts.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define W 1920
#define H 1440
#define F 500

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t part1by1(uint32_t x) {
    x &= 0xffffu; x = (x | (x << 8)) & 0x00ff00ffu; x = (x | (x << 4)) & 0x0f0f0f0fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}
// LAYOUT 0 linear, 1 tile 16x16, 2 tile 32x32 of 8x4 lines, 3 morton within 64x64 tiles, 4 tile 8x8 (256 B)
template <int LAYOUT> __device__ __forceinline__ size_t texel(int f, int u, int v) {
    if (LAYOUT == 0) return (size_t)f * (W * H) + (size_t)v * W + u;
    if (LAYOUT == 1) return (size_t)f * (W * H) + (size_t)((v >> 4) * (W / 16) + (u >> 4)) * 256 + ((v & 15) << 4) + (u & 15);
    if (LAYOUT == 2) return (size_t)f * (W * H) + (size_t)((v >> 5) * (W / 32) + (u >> 5)) * 1024 + (((v >> 2) & 7) * 4 + ((u >> 3) & 3)) * 32 + ((v & 3) << 3) + (u & 7);
    if (LAYOUT == 3) return (size_t)f * (W * (H + 32)) + (size_t)((v >> 6) * (W / 64) + (u >> 6)) * 4096 + (part1by1(u & 63) | (part1by1(v & 63) << 1));
    return (size_t)f * (W * H) + (size_t)((v >> 3) * (W / 8) + (u >> 3)) * 64 + ((v & 7) << 3) + (u & 7);
}

// one thread = one point; NB gathers (different frames) in flight per thread, `rounds` batches
template <int LAYOUT, int NB, typename T>
__global__ void __launch_bounds__(256) gather_kernel(const T* __restrict__ tex, const T* __restrict__ tex2, unsigned* out, int rounds, int winw, int winh) {
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    unsigned acc = 0;
    for (int r = 0; r < rounds; ++r) {
        T vals[NB], vals2[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const uint32_t hw = hash32(gw / 8 * 977u + r * 131u + k * 7919u);          // 8 consecutive warps (a tile) share frame + window origin
            const int f = hw % F;
            const uint32_t h2 = hash32(hw + 17u);
            const int bx = (h2 % (W - 4 * winw)) + (gw & 3) * winw, by = ((h2 >> 12) % (H - 2 * winh)) + ((gw >> 2) & 1) * winh;
            const uint32_t hl = hash32(gw * 32u + lane + r * 1000003u + k * 65537u);
            const int u = bx + hl % winw, v = by + (hl >> 10) % winh;
            vals[k] = __ldg(tex + texel<LAYOUT>(f, u, v));
            if (tex2) vals2[k] = __ldg(tex2 + texel<LAYOUT>(f, u, v));
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) acc += (unsigned)vals[k] + (tex2 ? (unsigned)vals2[k] : 0u);
    }
    if (acc == 0xdeadbeefu) out[0] = acc;
}

template <int LAYOUT, int NB, typename T> float run(const T* tex, const T* tex2, unsigned* out, int nwarps, int rounds, int winw, int winh) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = nwarps / 8;
    for (int i = 0; i < 2; ++i) gather_kernel<LAYOUT, NB, T><<<blocks, 256>>>(tex, tex2, out, rounds, winw, winh);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) gather_kernel<LAYOUT, NB, T><<<blocks, 256>>>(tex, tex2, out, rounds, winw, winh);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e_)); }
    return ms / 5;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); fflush(stdout); return 1; } } while (0)
int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const size_t ntex = (size_t)F * W * (H + 32);
    uint32_t* tex; uint16_t *t16, *t16b; unsigned* out;
    CK(cudaMalloc(&tex, ntex * 4)); CK(cudaMemset(tex, 1, ntex * 4));
    CK(cudaMalloc(&t16, ntex * 2)); CK(cudaMemset(t16, 1, ntex * 2));
    CK(cudaMalloc(&t16b, ntex * 2)); CK(cudaMemset(t16b, 1, ntex * 2)); CK(cudaMalloc(&out, 4)); printf("allocated\n");
    const int nwarps = 312504, rounds = 2;   // C2: 312.5 k warps x ~9.2 candidates ~ 2 rounds of 4 + ... ; gathers = nwarps*32*rounds*NB
    const char* names[5] = {"linear", "tile16x16", "tile32x32/8x4", "morton64", "tile8x8"};
    for (int win = 0; win < 2; ++win) {
        const int ww = win ? 90 : 45, wh = win ? 76 : 38;
        printf("window %dx%d px per warp, 4 gathers in flight per thread\n", ww, wh);
        double n = (double)nwarps * 32 * rounds * 4;
        float ms;
        ms = run<0, 4, uint16_t>(t16, t16b, out, nwarps, rounds, ww, wh); printf("  u16+u16 two arrays linear (round-1 pattern): %.3f ms  %.1f G point-views/s (2 sectors each)\n", ms, n / ms / 1e6);
        ms = run<0, 4, uint32_t>(tex, nullptr, out, nwarps, rounds, ww, wh); printf("  u32 %-14s: %.3f ms  %.1f G gathers/s\n", names[0], ms, n / ms / 1e6);
        ms = run<1, 4, uint32_t>(tex, nullptr, out, nwarps, rounds, ww, wh); printf("  u32 %-14s: %.3f ms  %.1f G gathers/s\n", names[1], ms, n / ms / 1e6);
        ms = run<2, 4, uint32_t>(tex, nullptr, out, nwarps, rounds, ww, wh); printf("  u32 %-14s: %.3f ms  %.1f G gathers/s\n", names[2], ms, n / ms / 1e6);
        ms = run<3, 4, uint32_t>(tex, nullptr, out, nwarps, rounds, ww, wh); printf("  u32 %-14s: %.3f ms  %.1f G gathers/s\n", names[3], ms, n / ms / 1e6);
        ms = run<4, 4, uint32_t>(tex, nullptr, out, nwarps, rounds, ww, wh); printf("  u32 %-14s: %.3f ms  %.1f G gathers/s\n", names[4], ms, n / ms / 1e6);
        n = (double)nwarps * 32 * 1 * 8;
        ms = run<0, 8, uint32_t>(tex, nullptr, out, nwarps, 1, ww, wh); printf("  u32 %-14s NB=8: %.3f ms  %.1f G gathers/s\n", names[0], ms, n / ms / 1e6);
        ms = run<1, 8, uint32_t>(tex, nullptr, out, nwarps, 1, ww, wh); printf("  u32 %-14s NB=8: %.3f ms  %.1f G gathers/s\n", names[1], ms, n / ms / 1e6);
    }
    return 0;
}
