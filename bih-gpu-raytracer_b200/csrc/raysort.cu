// Ray sorting between bounces ("ray compaction/sorting between bounces"; the reference traces primary rays
// only -- its Color() is a stub, R/src/CUDAKernels.cu:370-389 -- so there is no reference counterpart).  A list of secondary
// rays (diffuse bounces) arrives in the order of the pixels that spawned it, but neighbouring pixels send their bounce rays
// in unrelated directions: the 32 rays of a warp start close together and part at once.  Tracing the list through a
// permutation that groups rays by WHERE they start and WHICH WAY they go restores most of the coherence:
//   key  = 21-bit Morton code of the origin's cell in a 128^3 grid over the scene box  <<  3  |  direction octant
//   sort = the build's onesweep radix sort (csrc/build.cu), 3 passes of 8 bits over (key, ray index)
//   trace= k_trace takes ray perm[i] as its i-th item and writes the result to that ray's own slot, so the outputs
//          stay in list order and equal the unsorted trace's bit for bit (tests/test_gpu_secondary.py).
// MEASURED (round 2, B200, one diffuse bounce off the 1080p primary hits; Mrays/s in list order / sorted, sort included):
// atrium 2.07 M rays 2283 / 2230 (key variants: cell+octant 2121, octant+cell 2157, 64 direction bins+cell 2230); 1 M-triangle
// sphere 0.60 M rays 1470 / 1456.  NEUTRAL: a bounce list is already in pixel order, i.e. sorted by origin, and 2 M rays
// spread over a 4-D ray space share few nodes below the top of the tree whatever the order (36 node visits of 64 B per ray:
// ~5 TB/s through L2 -- the incoherent batch is L2-bound, not divergence-bound).  Kept as an option (trace_sort_rays, off).
#include "bihrt_internal.cuh"

#define RS_PASSES 3
#ifndef RS_VARIANT
#define RS_VARIANT 0
#endif

__device__ __forceinline__ uint32_t rs_expand7(uint32_t v) {        // 7 bits -> every third bit
    v &= 0x7fu;
    v = (v | (v << 8)) & 0x0000700fu;
    v = (v | (v << 4)) & 0x000430c3u;
    v = (v | (v << 2)) & 0x00049249u;
    return v;
}

__global__ void __launch_bounds__(256) k_ray_keys(const bihrt_ray* __restrict__ rays, uint32_t n, const BihHeader* __restrict__ hdr,
                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[RS_PASSES * 256];
    for (int i = threadIdx.x; i < RS_PASSES * 256; i += 256) s_hist[i] = 0;
    const float lo0 = hdr->lo[0], lo1 = hdr->lo[1], lo2 = hdr->lo[2];
    const float s0 = 128.f / fmaxf(hdr->hi[0] - lo0, 1e-30f), s1 = 128.f / fmaxf(hdr->hi[1] - lo1, 1e-30f), s2 = 128.f / fmaxf(hdr->hi[2] - lo2, 1e-30f);
    __syncthreads();
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
        const float* p = reinterpret_cast<const float*>(rays + i);
        const float ox = __ldg(p), oy = __ldg(p + 1), oz = __ldg(p + 2), dx = __ldg(p + 3), dy = __ldg(p + 4), dz = __ldg(p + 5);
        const uint32_t cx = (uint32_t)fminf(fmaxf((ox - lo0) * s0, 0.f), 127.f);
        const uint32_t cy = (uint32_t)fminf(fmaxf((oy - lo1) * s1, 0.f), 127.f);
        const uint32_t cz = (uint32_t)fminf(fmaxf((oz - lo2) * s2, 0.f), 127.f);
        const uint32_t oct = (dx < 0.f ? 4u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 1u : 0u);
#if RS_VARIANT == 0
        const uint32_t key = (((rs_expand7(cx) << 2) | (rs_expand7(cy) << 1) | rs_expand7(cz)) << 3) | oct;
#elif RS_VARIANT == 1       /* octant first, then the origin cell */
        const uint32_t key = (oct << 21) | (rs_expand7(cx) << 2) | (rs_expand7(cy) << 1) | rs_expand7(cz);
#else                       /* direction quantised to 4 bins per axis (6 bits) first, then a 64^3 origin grid */
        const float il = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-38f));
        const uint32_t qx = (uint32_t)fminf((dx * il + 1.f) * 2.f, 3.f), qy = (uint32_t)fminf((dy * il + 1.f) * 2.f, 3.f), qz = (uint32_t)fminf((dz * il + 1.f) * 2.f, 3.f);
        const uint32_t key = (((qx << 4) | (qy << 2) | qz) << 18) | (rs_expand7(cx >> 1) << 2) | (rs_expand7(cy >> 1) << 1) | rs_expand7(cz >> 1);
#endif
        keys[i] = key;
#pragma unroll
        for (int q = 0; q < RS_PASSES; q++) atomicAdd(&s_hist[q * 256 + ((key >> (8 * q)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RS_PASSES * 256; i += 256) { const uint32_t v = s_hist[i]; if (v) atomicAdd(&hist[H_HIST + i], v); }
}

int bihrt_ray_sort_launch(bihrt_ctx* c, const bihrt_ray* rays, int64_t n64, const uint32_t** perm) {
    const uint32_t n = (uint32_t)n64;
    const size_t tiles = (n + SORT_TILE - 1) / SORT_TILE, lb_words = (size_t)RS_PASSES * tiles * 256;
    if ((size_t)n > c->rs_cap) {
        for (int i = 0; i < 2; i++) {
            if (c->d_rs_keys[i]) cudaFree(c->d_rs_keys[i]);
            if (c->d_rs_vals[i]) cudaFree(c->d_rs_vals[i]);
            c->d_rs_keys[i] = c->d_rs_vals[i] = nullptr;
        }
        if (c->d_rs_lookback) { cudaFree(c->d_rs_lookback); c->d_rs_lookback = nullptr; }
        c->rs_cap = 0;
        for (int i = 0; i < 2; i++) {
            BIHRT_CUDA(c, cudaMalloc((void**)&c->d_rs_keys[i], ((size_t)n + 8) * 4));
            BIHRT_CUDA(c, cudaMalloc((void**)&c->d_rs_vals[i], ((size_t)n + 8) * 4));
        }
        BIHRT_CUDA(c, cudaMalloc((void**)&c->d_rs_lookback, (lb_words + 16) * 4));
        if (!c->d_rs_hist) BIHRT_CUDA(c, cudaMalloc((void**)&c->d_rs_hist, H_WORDS * 4));
        if (!c->d_rs_hdr) BIHRT_CUDA(c, cudaMalloc((void**)&c->d_rs_hdr, sizeof(BihHeader)));
        c->rs_cap = n;
    }
    BIHRT_CUDA(c, cudaMemsetAsync(c->d_rs_hist, 0, H_WORDS * 4, c->stream));
    BIHRT_CUDA(c, cudaMemsetAsync(c->d_rs_lookback, 0, lb_words * 4, c->stream));
    BIHRT_CUDA(c, cudaMemsetAsync(c->d_rs_hdr, 0, sizeof(BihHeader), c->stream));
    const int grid = (int)max(1u, min((uint32_t)(c->sm_count * 4), (n + 255u) / 256u));
    k_ray_keys<<<grid, 256, 0, c->stream>>>(rays, n, c->d_hdr, c->d_rs_keys[0], c->d_rs_hist);
    c->kernel_launches += 1;
    int rc = bihrt_sort_pairs_launch(c, c->d_rs_keys, c->d_rs_vals, n, RS_PASSES, c->d_rs_hist, c->d_rs_lookback, c->d_rs_hdr);
    if (rc) return rc;
    *perm = c->d_rs_vals[RS_PASSES & 1];
    return BIHRT_OK;
}
