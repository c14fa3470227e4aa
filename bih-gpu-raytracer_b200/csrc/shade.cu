// Secondary-ray generation on the device: the step AFTER the path.  The reference's Color() is a stub that
// returns a constant per hit (R/src/CUDAKernels.cu:370-389) -- no shading, no shadow or bounce rays -- so this
// stage has no reference counterpart to be bit-exact with; it exists so that BASELINE config 3 (primary +
// shadow + one diffuse bounce) runs end to end on the GPU.  The rays it emits are traced by
// the same k_trace kernel, whose results ARE checked against the oracle on exactly these rays.
//
// Pipeline: primary hits of a camera frame (k_trace MODE 2, per-sample t / slot) -> k_sec_count (hits per
// 256-sample tile) -> k_sec_scan (exclusive scan of the tile counts, one block) -> k_sec_write (rays written
// compacted, in sample order: deterministic).
#include "bihrt_internal.cuh"
#include <float.h>

#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}

__global__ void __launch_bounds__(256) k_sec_count(const int32_t* __restrict__ slot, int64_t n, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t s_w[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const uint32_t m = __ballot_sync(FULL, i < n && slot[i] >= 0);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int k = 0; k < 8; k++) t += s_w[k]; tile_cnt[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) k_sec_scan(uint32_t* __restrict__ tile_cnt, uint32_t ntiles, unsigned long long* __restrict__ total) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? tile_cnt[i] : 0u;
        const uint32_t inc = warp_incl_scan_u32(v, lane);
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) { const uint32_t x = s_w[lane]; const uint32_t xi = warp_incl_scan_u32(x, lane); s_w[lane] = xi - x; }
        __syncthreads();
        const uint32_t excl = s_carry + s_w[w] + inc - v;
        if (i < ntiles) tile_cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

// sample index -> primary ray, exactly as k_trace generates it (R/src/CUDAKernels.cu:414-415, R/src/Camera.cu:18-20)
__device__ __forceinline__ void primary_ray(const bihrt_camera& cam, int w, int h, int spp, uint64_t seed, uint32_t flags, int64_t i,
                                            float o[3], float d[3]) {
    const uint32_t pixel = (uint32_t)(i / spp), s = (uint32_t)(i % spp);
    const int px = (int)(pixel % (uint32_t)w), py = (int)(pixel / (uint32_t)w);
    const float ru = (flags & BIHRT_RENDER_JITTER) ? bihrt_jitter(seed, pixel, s, 0) : 0.5f;
    const float rv = (flags & BIHRT_RENDER_JITTER) ? bihrt_jitter(seed, pixel, s, 1) : 0.5f;
    const float uu = __fdiv_rn(__fadd_rn((float)px, ru), (float)w);
    const float vv = __fdiv_rn(__fadd_rn((float)py, rv), (float)h);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[k] = cam.origin[k];
        d[k] = __fsub_rn(__fadd_rn(__fadd_rn(cam.lower_left[k], __fmul_rn(uu, cam.horizontal[k])), __fmul_rn(vv, cam.vertical[k])), cam.origin[k]);
    }
}

__global__ void __launch_bounds__(256) k_sec_write(const float* __restrict__ t, const int32_t* __restrict__ slot, int64_t n,
                                                   const uint32_t* __restrict__ tile_off, const BihTri* __restrict__ tris,
                                                   bihrt_camera cam, int w, int h, int spp, uint64_t seed, uint32_t flags,
                                                   int kind, float lx, float ly, float lz,
                                                   bihrt_ray* __restrict__ out_rays, int32_t* __restrict__ out_sample) {
    __shared__ uint32_t s_w[8];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool hit = i < n && slot[i] >= 0;
    const uint32_t m = __ballot_sync(FULL, hit);
    if (lane == 0) s_w[wi] = __popc(m);
    __syncthreads();
    uint32_t base = tile_off[blockIdx.x];
    for (int k = 0; k < wi; k++) base += s_w[k];
    if (!hit) return;
    const uint32_t dst = base + __popc(m & ((1u << lane) - 1u));
    float o[3], d[3];
    primary_ray(cam, w, h, spp, seed, flags, i, o, d);
    const BihTri tr = tris[slot[i]];
    // geometric normal e1 x e2; a culled Moller-Trumbore hit is front-facing, so the normal faces the viewer
    float nx = tr.e1y * tr.e2z - tr.e2y * tr.e1z, ny = tr.e1z * tr.e2x - tr.e2z * tr.e1x, nz = tr.e1x * tr.e2y - tr.e2x * tr.e1y;
    const float inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-38f));
    nx *= inv; ny *= inv; nz *= inv;
    const float tt = t[i];
    const float px = o[0] + tt * d[0] + 1e-3f * nx, py = o[1] + tt * d[1] + 1e-3f * ny, pz = o[2] + tt * d[2] + 1e-3f * nz;
    float dx, dy, dz;
    if (kind == BIHRT_SECONDARY_SHADOW) {
        dx = lx - px; dy = ly - py; dz = lz - pz;                         // unnormalised: t = 1 at the light
    } else {
        // cosine-weighted direction about the normal (Malley), counter-based hash on dims 2,3
        const uint32_t pixel = (uint32_t)(i / spp), s = (uint32_t)(i % spp);
        const float r1 = bihrt_jitter(seed, pixel, s, 2), r2 = bihrt_jitter(seed, pixel, s, 3);
        const float rad = sqrtf(r1), phi = 6.28318530718f * r2;
        float sn, cs;
        sincosf(phi, &sn, &cs);
        const float a = rad * cs, b = rad * sn, c2 = sqrtf(fmaxf(0.f, 1.f - r1));
        // orthonormal basis (Frisvad / Duff et al.)
        const float sg = copysignf(1.0f, nz), aa = -1.0f / (sg + nz), bb = nx * ny * aa;
        const float t1x = 1.0f + sg * nx * nx * aa, t1y = sg * bb, t1z = -sg * nx;
        const float t2x = bb, t2y = sg + ny * ny * aa, t2z = -ny;
        dx = a * t1x + b * t2x + c2 * nx; dy = a * t1y + b * t2y + c2 * ny; dz = a * t1z + b * t2z + c2 * nz;
    }
    bihrt_ray r; r.o[0] = px; r.o[1] = py; r.o[2] = pz; r.d[0] = dx; r.d[1] = dy; r.d[2] = dz;
    out_rays[dst] = r;
    if (out_sample) out_sample[dst] = (int32_t)i;
}

int bihrt_secondary_launch(bihrt_ctx* c, const float* t, const int32_t* slot, int64_t n, uint32_t* tile_cnt, unsigned long long* total,
                           const bihrt_camera& cam, int w, int h, int spp, uint64_t seed, uint32_t flags, int kind, const float light[3],
                           bihrt_ray* out_rays, int32_t* out_sample) {
    const uint32_t ntiles = (uint32_t)((n + 255) / 256);
    k_sec_count<<<ntiles, 256, 0, c->stream>>>(slot, n, tile_cnt);
    k_sec_scan<<<1, 1024, 0, c->stream>>>(tile_cnt, ntiles, total);
    k_sec_write<<<ntiles, 256, 0, c->stream>>>(t, slot, n, tile_cnt, c->d_tris, cam, w, h, spp, seed, flags, kind,
                                                light[0], light[1], light[2], out_rays, out_sample);
    c->kernel_launches += 3;
    BIHRT_CUDA(c, cudaGetLastError());
    return BIHRT_OK;
}
