// Development harness for the build's radix sort (not part of the product): times the four passes on Morton keys of a
// sphere mesh in mesh order, checks them against std::stable_sort, and (TIMING) records per-tile phase time stamps.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o sortlab sortlab.cu && ./sortlab [n]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <numeric>
#define FULL 0xffffffffu
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#ifndef VARIANT
#define VARIANT 0
#endif
#ifndef TIMING
#define TIMING 1
#endif
#define NSTAMP 8
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_w, uint32_t* total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint32_t x = s_w[i]; if (i < w) base += x; tot += x; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}
#define OS_THREADS 256
#define LB_FLAG_AGG   0x40000000u
#define LB_FLAG_INCL  0x80000000u
#define LB_MASK       0x3FFFFFFFu
#if TIMING
#define STAMP(k) do { if (threadIdx.x == 0) stamps[(size_t)tile * NSTAMP + (k)] = gtime(); } while (0)
#else
#define STAMP(k)
#endif

#if VARIANT == 2
#include "sortlab_hybrid.cuh"
#elif VARIANT == 1
#include "sortlab_coop.cuh"
#else
#include "sortlab_kernels.cuh"
#endif

__global__ void k_hist(const uint32_t* keys, uint32_t n, uint32_t* hist) {
    __shared__ uint32_t s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = keys[i];
        for (int p = 0; p < 4; p++) atomicAdd(&s[p * 256 + ((k >> (8 * p)) & 255)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) if (s[i]) atomicAdd(&hist[i], s[i]);
}

static uint32_t expand10(uint32_t v) { v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v; }

int main(int argc, char** argv) {
    uint32_t nseg = argc > 1 ? atoi(argv[1]) : 708;
    // triangle centres of a lat-long sphere in mesh order
    std::vector<uint32_t> keys;
    for (uint32_t i = 0; i < nseg; i++) for (uint32_t j = 0; j < nseg; j++) for (int t = 0; t < 2; t++) {
        double th = M_PI * (i + 0.5 + 0.2 * t) / nseg, ph = 2 * M_PI * (j + 0.5 + 0.2 * t) / nseg;
        double r = 1.0 + 0.05 * sin(7 * th) * cos(5 * ph);
        double x = r * sin(th) * cos(ph), y = r * cos(th), z = r * sin(th) * sin(ph);
        auto q = [](double v) { double s = (v + 1.06) / 2.12 * 1024.0; return (uint32_t)fmin(fmax(s, 0.0), 1023.0); };
        keys.push_back(expand10(q(x)) * 4 + expand10(q(y)) * 2 + expand10(q(z)));
    }
    const uint32_t n = (uint32_t)keys.size();
    printf("n = %u  variant %d\n", n, VARIANT);
    std::vector<uint32_t> order(n); std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });

    uint32_t *d_k[2], *d_v[2], *d_hist, *d_lb, *d_in; unsigned long long* d_st;
    const uint32_t tiles = (n + SL_TILE - 1) / SL_TILE;
    CK(cudaMalloc(&d_in, n * 4)); for (int i = 0; i < 2; i++) { CK(cudaMalloc(&d_k[i], n * 4 + 64)); CK(cudaMalloc(&d_v[i], n * 4 + 64)); }
    CK(cudaMalloc(&d_hist, 4096 * 4)); CK(cudaMalloc(&d_lb, (size_t)4 * tiles * 256 * 4 + 4096)); CK(cudaMalloc(&d_st, (size_t)4 * (tiles + 1024) * NSTAMP * 8)); CK(cudaMemset(d_st, 0, (size_t)4 * (tiles + 1024) * NSTAMP * 8));
    CK(cudaMemcpy(d_in, keys.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t ev[6]; for (auto& e : ev) cudaEventCreate(&e);
    float best[5] = { 1e9, 1e9, 1e9, 1e9, 1e9 };
    sl_setup();
    for (int rep = 0; rep < 12; rep++) {
        CK(cudaMemcpy(d_k[0], d_in, n * 4, cudaMemcpyDeviceToDevice));
        CK(cudaMemset(d_hist, 0, 4096 * 4)); CK(cudaMemset(d_lb, 0, (size_t)4 * tiles * 256 * 4 + 4096));
        k_hist<<<296, 256>>>(d_k[0], n, d_hist);
        CK(cudaDeviceSynchronize());
        int cur = 0;
        cudaEventRecord(ev[0]);
        for (int pass = 0; pass < 4; pass++) {
            sl_launch(pass, d_k[cur], d_v[cur], d_k[cur ^ 1], d_v[cur ^ 1], n, d_hist, d_lb + (size_t)pass * tiles * 256, d_st + (size_t)pass * tiles * NSTAMP, tiles);
            cur ^= 1;
            cudaEventRecord(ev[pass + 1]);
        }
        CK(cudaDeviceSynchronize());
        float tot = 0;
        for (int p = 0; p < 4; p++) { float ms; cudaEventElapsedTime(&ms, ev[p], ev[p + 1]); best[p] = fminf(best[p], ms); tot += ms; }
        best[4] = fminf(best[4], tot);
    }
    printf("passes (us, best of 12): %.1f %.1f %.1f %.1f | total %.1f\n", best[0] * 1e3, best[1] * 1e3, best[2] * 1e3, best[3] * 1e3, best[4] * 1e3);
    std::vector<uint32_t> ok(n), ov(n);
    CK(cudaMemcpy(ok.data(), d_k[0], n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ov.data(), d_v[0], n * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (uint32_t i = 0; i < n; i++) if (ov[i] != order[i] || ok[i] != keys[order[i]]) bad++;
    printf("check vs std::stable_sort: %s (%zu mismatches)\n", bad ? "FAIL" : "ok", bad);
#if TIMING
    const uint32_t rows = sl_stamp_rows(tiles);
    std::vector<unsigned long long> st((size_t)4 * std::max(tiles, rows) * NSTAMP);
    CK(cudaMemcpy(st.data(), d_st, st.size() * 8, cudaMemcpyDeviceToHost));
    for (int pass = 0; pass < SL_STAMP_SETS; pass++) {
        const unsigned long long* s = st.data() + (size_t)pass * tiles * NSTAMP;
        const uint32_t tiles = rows;
        unsigned long long t0 = ~0ull, t1 = 0;
        for (uint32_t t = 0; t < tiles; t++) { t0 = std::min(t0, s[t * NSTAMP]); t1 = std::max(t1, s[t * NSTAMP + NSTAMP - 1]); }
        printf("pass %d: span %.1f us; phase mean (max) us:", pass, (t1 - t0) / 1e3);
        for (int k = 1; k < NSTAMP; k++) {
            double sum = 0, mx = 0;
            for (uint32_t t = 0; t < tiles; t++) { if (s[t * NSTAMP + k] < s[t * NSTAMP + k - 1]) continue; double d = (double)(s[t * NSTAMP + k] - s[t * NSTAMP + k - 1]) / 1e3; sum += d; mx = std::max(mx, d); }
            printf(" %.2f(%.2f)", sum / tiles, mx);
        }
        double late = 0; for (uint32_t t = 0; t < tiles; t++) late = std::max(late, (double)(s[t * NSTAMP] - t0) / 1e3);
        printf(" | latest start +%.1f us\n", late);
    }
#endif
    return bad != 0;
}
