// Ray traversal + ray/triangle intersection on sm_100a.  Replaces cudaRender / Color / TraverseTree /
// FindNearestTriangle / RayTriangleIntersection (R/src/CUDAKernels.cu:17-50,206-423) and
// Camera::GetRay / Ray::Ray (R/src/Camera.cu:18-20, R/src/Ray.cu:3-10).
//
// Kernel shape: persistent warps (grid = SMs x resident blocks) pull 32-ray packets -- 8x4 pixel
// tiles for camera rays -- from one atomic counter; every lane walks its own ray with a 16-byte
// per-entry stack; nodes (16 B) and leaf-ordered triangles (3 x 16 B) are fetched with single
// 128-bit read-only loads.  The walk is a two-phase ("while-while") loop: lanes step through
// internal nodes until they hold a leaf to test, the warp re-converges, leaves are tested, repeat.
//
// Logical per-ray algorithm = oracle/bih_oracle.c:traverse_proper (pruned traversal that returns
// what the reference's TraverseTree returns, including on axis-aligned flat geometry where the
// reference's strict comparisons decide).  Arithmetic is IEEE binary32 without FMA contraction
// (explicit __f*_rn) so t is bit-identical to the oracle's.
#include "bihrt_internal.cuh"
#include <float.h>

#define FULL 0xffffffffu
#define STACK_DEPTH 32          // <= 30 pushes possible: one per Morton bit on a root-to-leaf path
#define NO_NODE 0xFFFFFFFFu

// (double)det < 0.000001 (R/src/CUDAKernels.cu:28)  <=>  det < 0x358637be as binary32
#define DET_EPS __uint_as_float(0x358637beu)

struct Ray {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
    uint32_t sign;   // bit k: invDir[k] < 0
};

__device__ __forceinline__ Ray make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    r.ix = __frcp_rn(dx); r.iy = __frcp_rn(dy); r.iz = __frcp_rn(dz);      // 1 / b, R/src/Ray.cu:6
    r.sign = (r.ix < 0.f ? 1u : 0u) | (r.iy < 0.f ? 2u : 0u) | (r.iz < 0.f ? 4u : 0u);
    return r;
}

struct Hit { float t; int slot; uint32_t prim; };

template <bool COUNTED>
__device__ __forceinline__ void test_leaf(const BihTri* __restrict__ tris, uint32_t slot, const Ray& r, Hit& h, uint32_t& ntris) {
    for (;;) {
        const float4* p = reinterpret_cast<const float4*>(tris + slot);
        const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
        if (COUNTED) ntris++;
        const float e1x = q0.w, e1y = q1.x, e1z = q1.y, e2x = q1.z, e2y = q1.w, e2z = q2.x;
        // pvec = cross(dir, e2)
        const float px = __fsub_rn(__fmul_rn(r.dy, e2z), __fmul_rn(e2y, r.dz));
        const float py = __fsub_rn(__fmul_rn(r.dz, e2x), __fmul_rn(e2z, r.dx));
        const float pz = __fsub_rn(__fmul_rn(r.dx, e2y), __fmul_rn(e2x, r.dy));
        const float det = __fadd_rn(__fadd_rn(__fmul_rn(e1x, px), __fmul_rn(e1y, py)), __fmul_rn(e1z, pz));
        if (!(det < DET_EPS)) {
            const float inv = __frcp_rn(det);        // == (float)(1.0 / (double)det), :31 (53 >= 2*24+2 bits)
            const float tx = __fsub_rn(r.ox, q0.x), ty = __fsub_rn(r.oy, q0.y), tz = __fsub_rn(r.oz, q0.z);
            const float u = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, px), __fmul_rn(ty, py)), __fmul_rn(tz, pz)), inv);
            if (!(u < 0.f || u > 1.f)) {
                const float qx = __fsub_rn(__fmul_rn(ty, e1z), __fmul_rn(e1y, tz));
                const float qy = __fsub_rn(__fmul_rn(tz, e1x), __fmul_rn(e1z, tx));
                const float qz = __fsub_rn(__fmul_rn(tx, e1y), __fmul_rn(e1x, ty));
                const float v = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.dx, qx), __fmul_rn(r.dy, qy)), __fmul_rn(r.dz, qz)), inv);
                if (!(v < 0.f || __fadd_rn(u, v) > 1.f)) {
                    const float t = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(e2x, qx), __fmul_rn(e2y, qy)), __fmul_rn(e2z, qz)), inv);
                    if (t > 0.f && t < h.t) { h.t = t; h.slot = (int)slot; h.prim = __float_as_uint(q2.y); }
                }
            }
        }
        if (__float_as_uint(q2.z) & 1u) break;      // last triangle of the leaf
        slot++;
    }
}

// slab test against the scene box, R/src/CUDAKernels.cu:237-262 (same operation order)
__device__ __forceinline__ bool scene_slab(const float lo[3], const float hi[3], const Ray& r, float& tMin, float& tMax) {
    const float nx = (r.sign & 1u) ? hi[0] : lo[0], fx = (r.sign & 1u) ? lo[0] : hi[0];
    const float ny = (r.sign & 2u) ? hi[1] : lo[1], fy = (r.sign & 2u) ? lo[1] : hi[1];
    const float nz = (r.sign & 4u) ? hi[2] : lo[2], fz = (r.sign & 4u) ? lo[2] : hi[2];
    tMin = __fmul_rn(__fsub_rn(nx, r.ox), r.ix);
    tMax = __fmul_rn(__fsub_rn(fx, r.ox), r.ix);
    const float tymin = __fmul_rn(__fsub_rn(ny, r.oy), r.iy);
    const float tymax = __fmul_rn(__fsub_rn(fy, r.oy), r.iy);
    if ((tMin > tymax) || (tymin > tMax)) return false;
    if (tymin > tMin) tMin = tymin;
    if (tymax < tMax) tMax = tymax;
    const float tzmin = __fmul_rn(__fsub_rn(nz, r.oz), r.iz);
    const float tzmax = __fmul_rn(__fsub_rn(fz, r.oz), r.iz);
    if ((tMin > tzmax) || (tzmin > tMax)) return false;
    if (tzmin > tMin) tMin = tzmin;
    if (tzmax < tMax) tMax = tzmax;
    return true;
}

// One ray.  `active` lanes of the warp call this together (inactive lanes pass active=false and
// only take part in the warp-level re-convergence).
template <bool COUNTED>
__device__ __forceinline__ void trace_ray(const BihNode* __restrict__ nodes, const BihTri* __restrict__ tris,
                                          uint32_t nu, const float lo[3], const float hi[3], bool active,
                                          const Ray& r, Hit& h, uint32_t& nnodes, uint32_t& ntris, uint32_t& maxsp) {
    h.t = FLT_MAX; h.slot = -1; h.prim = 0xFFFFFFFFu;
    float rMin = 0.f, pMin = 0.f, pMax = 0.f;
    uint32_t cur = NO_NODE;
    if (active && nu > 0) {
        float sMax;
        if (scene_slab(lo, hi, r, rMin, sMax)) {
            if (nu == 1) { test_leaf<COUNTED>(tris, 0, r, h, ntris); }
            else { cur = 0; pMin = fmaxf(rMin, 0.f); pMax = sMax; }
        }
    }
    uint4 stack[STACK_DEPTH];
    int sp = 0;
    // pending leaves of the node just visited: A is tested first (near), then B (far)
    uint32_t leafA = NO_NODE, leafB = NO_NODE;
    float loA = 0.f, hiA = 0.f, loB = 0.f, hiB = 0.f;

    while (__any_sync(FULL, cur != NO_NODE)) {
        // ---- phase 1: internal nodes until this lane holds a leaf to test (or runs out of work)
        while (cur != NO_NODE) {
            bool advanced = false;
            if (pMin <= fminf(pMax, h.t)) {                    // entry check (closed interval)
                const float4 nd = __ldg(reinterpret_cast<const float4*>(nodes + cur));
                if (COUNTED) nnodes++;
                const uint32_t rl = __float_as_uint(nd.z), rr = __float_as_uint(nd.w);
                const uint32_t axis = ((rl >> 30) & 1u) | ((rr >> 29) & 2u);
                const float org = axis == 0 ? r.ox : (axis == 1 ? r.oy : r.oz);
                const float inv = axis == 0 ? r.ix : (axis == 1 ? r.iy : r.iz);
                const bool neg = (r.sign >> axis) & 1u;          // near = sign[axis], :286
                const float t0 = __fmul_rn(__fsub_rn(nd.x, org), inv);
                const float t1 = __fmul_rn(__fsub_rn(nd.y, org), inv);
                const float tn = neg ? t1 : t0, tf = neg ? t0 : t1;
                const uint32_t refn = neg ? rr : rl, reff = neg ? rl : rr;
                const bool near_ok = rMin < tn;                  // the reference's strict test, :292
                const float nMax = fminf(pMax, tn);
                const float fMin = fmaxf(pMin, tf);
                const bool near_leaf = (refn & BIH_REF_LEAF) != 0, far_leaf = (reff & BIH_REF_LEAF) != 0;
                if (near_ok && near_leaf && pMin <= nMax) { leafA = refn & BIH_REF_INDEX; loA = pMin; hiA = nMax; }
                if (far_leaf && fMin <= pMax) { leafB = reff & BIH_REF_INDEX; loB = fMin; hiB = pMax; }
                const bool near_i = near_ok && !near_leaf && (pMin <= nMax);
                const bool far_i = !far_leaf && (fMin <= pMax);
                if (near_i) {
                    if (far_i) {
                        stack[sp] = make_uint4(reff & BIH_REF_INDEX, __float_as_uint(tf), __float_as_uint(fMin), __float_as_uint(pMax));
                        sp++;
                        if (COUNTED) maxsp = max(maxsp, (uint32_t)sp);
                    }
                    cur = refn & BIH_REF_INDEX; pMax = nMax; advanced = true;
                } else if (far_i) {
                    cur = reff & BIH_REF_INDEX; rMin = tf; pMin = fMin; advanced = true;
                }
            }
            if (!advanced) {
                if (sp > 0) {
                    sp--;
                    const uint4 e = stack[sp];
                    cur = e.x; rMin = __uint_as_float(e.y); pMin = __uint_as_float(e.z); pMax = __uint_as_float(e.w);
                } else cur = NO_NODE;
            }
            if ((leafA & leafB) != NO_NODE) break;               // something to test
        }
        __syncwarp();
        // ---- phase 2: leaves, near first; each is re-checked against the closest hit so far
        if (leafA != NO_NODE) {
            if (loA <= fminf(hiA, h.t)) test_leaf<COUNTED>(tris, leafA, r, h, ntris);
            leafA = NO_NODE;
        }
        if (leafB != NO_NODE) {
            if (loB <= fminf(hiB, h.t)) test_leaf<COUNTED>(tris, leafB, r, h, ntris);
            leafB = NO_NODE;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// kernels.  MODE 0: ray list -> (t, slot, prim);  1: camera -> packed framebuffer;
//           2: camera -> per-sample (t, slot, prim)
// ------------------------------------------------------------------------------------------
template <int MODE, bool COUNTED>
__global__ void __launch_bounds__(128) k_trace(TraceArgs a) {
    const int lane = threadIdx.x & 31;
    const BihHeader* hdr = a.hdr;
    const uint32_t nu = hdr->nu;
    float lo[3] = { hdr->lo[0], hdr->lo[1], hdr->lo[2] }, hi[3] = { hdr->hi[0], hdr->hi[1], hdr->hi[2] };
    uint32_t nnodes = 0, ntris = 0, maxsp = 0;

    // work units
    uint32_t nunits;
    int tx = 0, my_tiles = 0;
    if (MODE == 0) nunits = (uint32_t)((a.nrays + 31) / 32);
    else {
        tx = (a.w + 31) / 32;
        const int ty = (a.h + 31) / 32, T = tx * ty;
        my_tiles = T > a.shard_index ? (T - a.shard_index + a.shard_count - 1) / a.shard_count : 0;
        nunits = (uint32_t)my_tiles * 32u;
    }
    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(a.work, 1u);
        u = __shfl_sync(FULL, u, 0);
        if (u >= nunits) break;
        if (MODE == 0) {
            const int64_t i = (int64_t)u * 32 + lane;
            const bool active = i < a.nrays;
            Ray r = make_ray(0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
            if (active) {
                const float* p = reinterpret_cast<const float*>(a.rays + i);
                r = make_ray(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3), __ldg(p + 4), __ldg(p + 5));
            }
            Hit h;
            trace_ray<COUNTED>(a.nodes, a.tris, nu, lo, hi, active, r, h, nnodes, ntris, maxsp);
            if (active) {
                if (a.out_t) a.out_t[i] = h.t;
                if (a.out_slot) a.out_slot[i] = h.slot;
                if (a.out_prim) a.out_prim[i] = (int32_t)h.prim;
            }
        } else {
            const int tile = a.shard_index + (int)(u >> 5) * a.shard_count;
            const int sub = (int)(u & 31u);
            const int px = (tile % tx) * 32 + (sub & 3) * 8 + (lane & 7);
            const int py = (tile / tx) * 32 + (sub >> 2) * 4 + (lane >> 3);
            const bool inside = px < a.w && py < a.h;
            const uint32_t pixel = (uint32_t)(py * a.w + px);
            uint32_t hits = 0;
            for (int s = 0; s < a.spp; s++) {
                // u,v per R/src/CUDAKernels.cu:414-415; GetRay per R/src/Camera.cu:18-20
                const float ru = (a.flags & BIHRT_RENDER_JITTER) ? bihrt_jitter(a.seed, pixel, (uint32_t)s, 0) : 0.5f;
                const float rv = (a.flags & BIHRT_RENDER_JITTER) ? bihrt_jitter(a.seed, pixel, (uint32_t)s, 1) : 0.5f;
                const float uu = __fdiv_rn(__fadd_rn((float)px, ru), (float)a.w);
                const float vv = __fdiv_rn(__fadd_rn((float)py, rv), (float)a.h);
                float d[3];
#pragma unroll
                for (int k = 0; k < 3; k++)
                    d[k] = __fsub_rn(__fadd_rn(__fadd_rn(a.cam.lower_left[k], __fmul_rn(uu, a.cam.horizontal[k])),
                                               __fmul_rn(vv, a.cam.vertical[k])), a.cam.origin[k]);
                const Ray r = make_ray(a.cam.origin[0], a.cam.origin[1], a.cam.origin[2], d[0], d[1], d[2]);
                Hit h;
                trace_ray<COUNTED>(a.nodes, a.tris, nu, lo, hi, inside, r, h, nnodes, ntris, maxsp);
                if (MODE == 1) hits += (h.slot >= 0);
                else if (inside) {
                    const int64_t o = (int64_t)pixel * a.spp + s;
                    if (a.out_t) a.out_t[o] = h.t;
                    if (a.out_slot) a.out_slot[o] = h.slot;
                    if (a.out_prim) a.out_prim[o] = (int32_t)h.prim;
                }
            }
            if (MODE == 1 && inside) {
                // Color + rgbToInt, R/src/CUDAKernels.cu:82-88,385-387,420-422 (sums of 255/20/40 are exact)
                const float fh = (float)hits, fm = (float)(a.spp - (int)hits), fs = (float)a.spp;
                float cr = __fdiv_rn(__fadd_rn(__fmul_rn(fh, 255.f), __fmul_rn(fm, 20.f)), fs);
                float cb = __fdiv_rn(__fmul_rn(fm, 40.f), fs);
                cr = fmaxf(0.f, fminf(255.f, cr));
                cb = fmaxf(0.f, fminf(255.f, cb));
                a.fb[pixel] = ((uint32_t)(int)cb << 16) | ((uint32_t)(int)cr << 8) | (uint32_t)(int)cr;
            }
        }
    }
    if (COUNTED) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nnodes += __shfl_xor_sync(FULL, nnodes, o);
            ntris += __shfl_xor_sync(FULL, ntris, o);
            maxsp = max(maxsp, __shfl_xor_sync(FULL, maxsp, o));
        }
        if (lane == 0) {
            atomicAdd(&a.counters[0], (unsigned long long)nnodes);
            atomicAdd(&a.counters[1], (unsigned long long)ntris);
            atomicMax(&a.counters[2], (unsigned long long)maxsp);
        }
    }
}

template <int MODE, bool COUNTED>
static int launch(bihrt_ctx* c, const TraceArgs& a) {
    int per_sm = c->opt_trace_blocks_per_sm;
    if (per_sm <= 0) {
        BIHRT_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<MODE, COUNTED>, 128, 0));
        if (per_sm < 1) per_sm = 1;
    }
    BIHRT_CUDA(c, cudaMemsetAsync(a.work, 0, 4, c->stream));
    k_trace<MODE, COUNTED><<<c->sm_count * per_sm, 128, 0, c->stream>>>(a);
    c->kernel_launches += 1;
    BIHRT_CUDA(c, cudaGetLastError());
    return BIHRT_OK;
}

int bihrt_trace_launch(bihrt_ctx* c, const TraceArgs& a, int mode, bool counted) {
    switch (mode * 2 + (counted ? 1 : 0)) {
        case 0: return launch<0, false>(c, a);
        case 1: return launch<0, true>(c, a);
        case 2: return launch<1, false>(c, a);
        case 3: return launch<1, true>(c, a);
        case 4: return launch<2, false>(c, a);
        case 5: return launch<2, true>(c, a);
    }
    return bihrt_fail(c, BIHRT_ERR_INVALID, "bad trace mode %d", mode);
}
