/*
 * bih_oracle.c -- CPU restatement of the BIH build + traversal + intersection path of
 * rehakvoj1/BIH-GPU-Raytracer.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libbihrt.so) never links, loads or calls it.
 *
 * Parity pin: tests/test_oracle_kat.py checks orc_build() against the one golden vector the
 * reference ships (R/BIH1.txt == R/BIH2.txt == every frame of R/log.txt, a 35-node tree dump;
 * committed in transformed form as tests/golden/dodecahedron_bih.json).  Traversal and
 * intersection results are NOT pinned by any reference fixture (the reference has no tests);
 * they are pinned by construction: orc_trace_ref == orc_brute_force == the reference's own
 * kernels compiled for sm_100a (oracle/ref_harness.cu, GPU box only).
 *
 * R/ = /root/reference/BIH_Raytracer/BIH_Raytracer/.  Every function cites the lines it follows.
 * The reference has no CPU builder; what is restated here is the algorithm of its thrust calls
 * and CUDA kernels, evaluated sequentially in IEEE binary32 (build with -ffp-contract=off).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* PTX min.f32 / max.f32 as used by CUDA's min(float,float)/max(float,float) overloads in
 * FindClipPlanes (R/src/CUDAKernels.cu:523-528): -0 < +0, NaN loses. */
static inline float dev_fminf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float dev_fmaxf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;
    return a > b ? a : b;
}

/* atomicMinFloat / atomicMaxFloat, R/src/CUDAKernels.cu:52-66: integer atomics on the bit
 * pattern == min/max in the sign-magnitude total order (so -0 < +0). Sequential restatement. */
static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline void atomic_min_float(float* addr, float value) {
    if (!signbit(value)) { int32_t o = f2i(*addr), v = f2i(value); *addr = i2f(o < v ? o : v); }
    else { uint32_t o = f2u(*addr), v = f2u(value); *addr = u2f(o > v ? o : v); }
}
static inline void atomic_max_float(float* addr, float value) {
    if (!signbit(value)) { int32_t o = f2i(*addr), v = f2i(value); *addr = i2f(o > v ? o : v); }
    else { uint32_t o = f2u(*addr), v = f2u(value); *addr = u2f(o < v ? o : v); }
}

/* std::minmax(initializer_list) semantics (R/src/App.cpp:123-125,133-135): smallest = leftmost
 * of the equivalent minima, largest = rightmost of the equivalent maxima, compared with <. */
static inline void minmax_list(const float* v, int n, float* mn, float* mx) {
    float lo = v[0], hi = v[0];
    for (int i = 1; i < n; i++) {
        if (v[i] < lo) lo = v[i];
        if (!(v[i] < hi)) hi = v[i];
    }
    *mn = lo; *mx = hi;
}

static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; } /* __clz(0)==32 */

/* ------------------------------------------------------------------------------------------ */
/* scene pre-pass: App::LoadModels, R/src/App.cpp:103-164                                      */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_prep(const float* tri9, int64_t n, float* lo3, float* hi3, float* centre3,
                      float* cnorm3, float* scene_lo, float* scene_hi) {
    if (n <= 0) return;
    /* R/src/App.cpp:103-106: scene box seeded with the first vertex */
    float slo[3] = { tri9[0], tri9[1], tri9[2] };
    float shi[3] = { tri9[0], tri9[1], tri9[2] };
    for (int64_t i = 0; i < n; i++) {
        const float* t = tri9 + 9 * i;
        for (int k = 0; k < 3; k++) {
            float v[3] = { t[k], t[3 + k], t[6 + k] };
            float mn, mx;
            minmax_list(v, 3, &mn, &mx);                       /* :123-125 */
            lo3[3 * i + k] = mn;                               /* :126 */
            hi3[3 * i + k] = mx;                               /* :127 */
            centre3[3 * i + k] = (mn + mx) / 2.0f;             /* :128-131 */
            float s[4] = { mn, mx, slo[k], shi[k] };           /* :133-137 */
            minmax_list(s, 4, &slo[k], &shi[k]);
        }
    }
    for (int64_t i = 0; i < n; i++) {                          /* :144-156 */
        for (int k = 0; k < 3; k++) {
            float cur_minus_min = centre3[3 * i + k] - slo[k];
            float max_minus_min = shi[k] - slo[k];
            cnorm3[3 * i + k] = cur_minus_min / max_minus_min;
        }
    }
    for (int k = 0; k < 3; k++) { scene_lo[k] = slo[k]; scene_hi[k] = shi[k]; }
}

/* ------------------------------------------------------------------------------------------ */
/* Morton codes: expandBits / morton3D, R/src/Renderer.cpp:116-136                             */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3d(float x, float y, float z) {
    /* device min/max: fmaxf(NaN,0)=0, so a flat scene (0/0) lands in cell 0 */
    x = dev_fminf(dev_fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = dev_fminf(dev_fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    z = dev_fminf(dev_fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    uint32_t xx = expand_bits((uint32_t)x);
    uint32_t yy = expand_bits((uint32_t)y);
    uint32_t zz = expand_bits((uint32_t)z);
    return xx * 4 + yy * 2 + zz;
}
ORC_API void orc_morton(const float* cnorm3, int64_t n, uint32_t* codes) {
    for (int64_t i = 0; i < n; i++)
        codes[i] = morton3d(cnorm3[3 * i], cnorm3[3 * i + 1], cnorm3[3 * i + 2]);
}

/* ------------------------------------------------------------------------------------------ */
/* thrust::sequence + stable_sort_by_key, R/src/Renderer.cpp:436-445                           */
/* (stable LSD counting sort, 4 x 8 bits; any stable sort gives the same permutation)          */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_sort(uint32_t* codes, uint32_t* idx, int64_t n) {
    for (int64_t i = 0; i < n; i++) idx[i] = (uint32_t)i;      /* ResetTrisIdxs, GPUArrayManager.cpp:197-200 */
    if (n <= 1) return;
    uint32_t* k2 = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* v2 = (uint32_t*)malloc((size_t)n * 4);
    uint32_t *ka = codes, *va = idx, *kb = k2, *vb = v2;
    for (int pass = 0; pass < 4; pass++) {
        int64_t hist[257];
        memset(hist, 0, sizeof hist);
        int sh = pass * 8;
        for (int64_t i = 0; i < n; i++) hist[((ka[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (int64_t i = 0; i < n; i++) {
            int64_t p = hist[(ka[i] >> sh) & 255]++;
            kb[p] = ka[i]; vb[p] = va[i];
        }
        uint32_t* t;
        t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    /* 4 passes: data is back in (codes, idx) */
    free(k2); free(v2);
}

/* ------------------------------------------------------------------------------------------ */
/* reduce_by_key + unique_by_key_copy, R/src/Renderer.cpp:450-472                              */
/* ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_rle(const uint32_t* codes, int64_t n, uint32_t* umc, uint32_t* cnt, int32_t* first) {
    int64_t nu = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0 || codes[i] != codes[i - 1]) { umc[nu] = codes[i]; cnt[nu] = 1; first[nu] = (int32_t)i; nu++; }
        else cnt[nu - 1]++;
    }
    return nu;
}

/* ------------------------------------------------------------------------------------------ */
/* BuildTree, R/src/CUDAKernels.cu:591-710 (one loop iteration per CUDA thread)                */
/* Arrays follow TreeInternalNode (R/src/Tree.cuh:16-24) as structure-of-arrays:               */
/*   clip[2*i+{0,1}], axis[i], is_leaf[2*i+{0,1}], children[2*i+{0,1}], parent[i]              */
/* Initial values follow GPUArrayManager::AllocateBIHTree, R/src/GPUArrayManager.cpp:58-91.    */
/* Nu < 2: the reference kernel's guard `idx > UMCSize-2` is an unsigned compare and reads out  */
/* of bounds (undefined behaviour); the restatement builds no internal node.                   */
/* ------------------------------------------------------------------------------------------ */
static inline int signum(int v) { return (0 < v) - (v < 0); }

ORC_API void orc_build_tree(const uint32_t* umc, int64_t nu_, float* clip, int32_t* axis,
                            uint8_t* is_leaf, int32_t* children, int32_t* parent,
                            int32_t* leaf_parents) {
    int nu = (int)nu_;
    for (int i = 0; i < nu; i++) leaf_parents[i] = -1;             /* GPUArrayManager.cpp:60-67 */
    for (int i = 0; i + 1 < nu; i++) {                              /* :73-84 */
        parent[i] = -1; children[2 * i] = children[2 * i + 1] = -1; axis[i] = -1;
        clip[2 * i] = -FLT_MAX; clip[2 * i + 1] = FLT_MAX;
        is_leaf[2 * i] = is_leaf[2 * i + 1] = 0;
    }
    if (nu < 2) return;
    for (int idx = 0; idx <= nu - 2; idx++) {
        uint32_t cur = umc[idx];                                    /* :599 */
        uint32_t ncp[2] = { (uint32_t)-1, (uint32_t)-1 };           /* :600 */
        if (idx) ncp[0] = (uint32_t)clz32(cur ^ umc[idx - 1]);      /* :603-607 */
        if (idx < nu - 1) ncp[1] = (uint32_t)clz32(cur ^ umc[idx + 1]); /* :610-614 */
        int d = signum((int)(ncp[1] - ncp[0]));                     /* :616 */
        int lcp_min = (int)ncp[1 - ((d + 1) / 2)];                  /* :620 */
        int l_max = 1, lcp_tmp = -2, l_idx = -1;
        do {                                                        /* :624-633 */
            l_max *= 2;
            l_idx = idx + l_max * d;
            if (l_idx < 0 || l_idx > nu - 1) lcp_tmp = -1;
            else lcp_tmp = clz32(cur ^ umc[l_idx]);
        } while (lcp_tmp > lcp_min);
        int l = 0, tmp_idx = -1;
        for (int t = l_max / 2; t >= 1; t /= 2) {                   /* :638-650 */
            tmp_idx = idx + (l + t) * d;
            if (tmp_idx < 0 || tmp_idx > nu - 1) lcp_tmp = -1;
            else lcp_tmp = clz32(cur ^ umc[tmp_idx]);
            if (lcp_tmp > lcp_min) l = l + t;
        }
        int other_end = idx + l * d;                                /* :651 */
        int lcp_ends = clz32(cur ^ umc[other_end]);                 /* :652 */
        int s = 0, t = l;
        for (;;) {                                                  /* :658-675 */
            t = (int)ceilf((float)t / 2.0f);                        /* __float2int_ru(t / 2.0f) */
            tmp_idx = idx + (s + t) * d;
            if (tmp_idx < 0 || tmp_idx > nu - 1) lcp_tmp = -1;
            else lcp_tmp = clz32(cur ^ umc[tmp_idx]);
            if (lcp_tmp > lcp_ends) s = s + t;
            if (t == 1) break;
        }
        int split = idx + s * d + (d < 0 ? d : 0);                  /* :677 */
        children[2 * idx] = split;                                  /* :680-681 */
        children[2 * idx + 1] = split + 1;
        int lo = idx < other_end ? idx : other_end, hi = idx < other_end ? other_end : idx;
        is_leaf[2 * idx] = (lo == split);                           /* :683 */
        is_leaf[2 * idx + 1] = (hi == split + 1);                   /* :684 */
        if (is_leaf[2 * idx]) leaf_parents[split] = idx; else parent[split] = idx;           /* :686-692 */
        if (is_leaf[2 * idx + 1]) leaf_parents[split + 1] = idx; else parent[split + 1] = idx; /* :694-700 */
        axis[idx] = (clz32(umc[split] ^ umc[split + 1]) + 1) % 3;   /* :702-706 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* FindClipPlanes, R/src/CUDAKernels.cu:497-549 (one loop iteration per CUDA thread; the float  */
/* atomics are order independent, so the sequential order is immaterial)                       */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_clip_planes(int64_t nu, const float* lo3, const float* hi3, const uint32_t* tris_idx,
                             const int32_t* leaf_parents, const uint32_t* cnt, const int32_t* first,
                             const int32_t* axis, const int32_t* children, const int32_t* parent,
                             float* clip) {
    for (int idx = 0; idx < (int)nu; idx++) {
        int first_idx = first[idx];
        uint32_t dup = cnt[idx];
        float blo[3], bhi[3];
        for (int k = 0; k < 3; k++) {                               /* :513-514 */
            blo[k] = lo3[3 * (int64_t)tris_idx[first_idx] + k];
            bhi[k] = hi3[3 * (int64_t)tris_idx[first_idx] + k];
        }
        for (int64_t i = first_idx; i < (int64_t)first_idx + dup; i++) { /* :517-529 */
            int64_t b = tris_idx[i];
            for (int k = 0; k < 3; k++) {
                blo[k] = dev_fminf(blo[k], lo3[3 * b + k]);
                bhi[k] = dev_fmaxf(bhi[k], hi3[3 * b + k]);
            }
        }
        int prev = idx;
        int par = leaf_parents[idx];                                /* :534 */
        while (par != -1) {                                         /* :536-547 */
            int ax = axis[par];
            if (children[2 * par] == prev) atomic_max_float(&clip[2 * par], bhi[ax]);
            if (children[2 * par + 1] == prev) atomic_min_float(&clip[2 * par + 1], blo[ax]);
            prev = par;
            par = parent[par];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* whole build = first half of Renderer::Render, R/src/Renderer.cpp:422-503                    */
/* Caller allocates every array with n entries (nodes: n-1, but n is fine).  Returns Nu.       */
/* ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_build(const float* tri9, int64_t n,
                          float* lo3, float* hi3, float* cnorm3, float* scene_lo, float* scene_hi,
                          uint32_t* codes_sorted, uint32_t* tris_idx,
                          uint32_t* umc, uint32_t* cnt, int32_t* first,
                          float* clip, int32_t* axis, uint8_t* is_leaf, int32_t* children,
                          int32_t* parent, int32_t* leaf_parents) {
    if (n <= 0) return 0;
    float* centre3 = (float*)malloc((size_t)n * 12);
    orc_prep(tri9, n, lo3, hi3, centre3, cnorm3, scene_lo, scene_hi);
    free(centre3);
    orc_morton(cnorm3, n, codes_sorted);
    orc_sort(codes_sorted, tris_idx, n);
    int64_t nu = orc_rle(codes_sorted, n, umc, cnt, first);
    orc_build_tree(umc, nu, clip, axis, is_leaf, children, parent, leaf_parents);
    orc_clip_planes(nu, lo3, hi3, tris_idx, leaf_parents, cnt, first, axis, children, parent, clip);
    return nu;
}

/* ------------------------------------------------------------------------------------------ */
/* rays: Ray::Ray, R/src/Ray.cu:3-10                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float o[3], d[3], inv[3]; int sign[3]; } orc_ray;

static inline void make_ray(orc_ray* r, const float* o, const float* d) {
    for (int k = 0; k < 3; k++) {
        r->o[k] = o[k]; r->d[k] = d[k];
        r->inv[k] = 1 / d[k];                                       /* :6 (IEEE; Release used -use_fast_math) */
        r->sign[k] = (r->inv[k] < 0);                               /* :7-9 */
    }
}

/* RayTriangleIntersection, R/src/CUDAKernels.cu:17-50 (glm::cross / glm::dot written out in the
 * glm 0.9.9.4 evaluation order: cross = (y1*z2 - y2*z1, z1*x2 - z2*x1, x1*y2 - x2*y1),
 * dot = (x*x' + y*y') + z*z'  -- glm/detail/func_geometric.inl compute_dot<vec3>: tmp.x+tmp.y+tmp.z) */
static inline void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - b[1] * a[2];
    c[1] = a[2] * b[0] - b[2] * a[0];
    c[2] = a[0] * b[1] - b[0] * a[1];
}
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static inline int ray_triangle(const float* tri, const orc_ray* r, float* out_t) {
    float e1[3], e2[3], pvec[3], tvec[3], qvec[3];
    for (int k = 0; k < 3; k++) { e1[k] = tri[3 + k] - tri[k]; e2[k] = tri[6 + k] - tri[k]; } /* :18-19 */
    cross3(r->d, e2, pvec);                                         /* :24 */
    float det = dot3(e1, pvec);                                     /* :26 */
    if ((double)det < 0.000001) return 0;                           /* :28 (double literal) */
    float inv_det = (float)(1.0 / (double)det);                     /* :31 */
    for (int k = 0; k < 3; k++) tvec[k] = r->o[k] - tri[k];         /* :33 */
    float u = dot3(tvec, pvec) * inv_det;                           /* :35 */
    if (u < 0 || u > 1) return 0;                                   /* :37 */
    cross3(tvec, e1, qvec);                                         /* :40 */
    float v = dot3(r->d, qvec) * inv_det;                           /* :42 */
    if (v < 0 || u + v > 1) return 0;                               /* :44 */
    *out_t = dot3(e2, qvec) * inv_det;                              /* :47 */
    return 1;
}

/* scene = everything TraverseTree reads */
typedef struct {
    const float* tri9; int64_t n; int64_t nu;
    const uint32_t* tris_idx; const uint32_t* cnt; const int32_t* first;
    const float* clip; const int32_t* axis; const uint8_t* is_leaf; const int32_t* children;
    float scene_lo[3], scene_hi[3];
} orc_scene;

typedef struct { double t; int slot; } orc_hit;                     /* HitRecord, R/src/Tree.cuh:9-14 (t is double) */
typedef struct { uint64_t nodes, tris; int max_stack; } orc_counters;

/* FindNearestTriangle, R/src/CUDAKernels.cu:206-224: records the SORTED SLOT i, not the prim id */
static inline void find_nearest(const orc_scene* s, const orc_ray* r, int leaf, orc_hit* rec, orc_counters* c) {
    float t = FLT_MAX;
    for (int64_t i = s->first[leaf]; i < (int64_t)s->first[leaf] + s->cnt[leaf]; i++) {
        uint32_t tri = s->tris_idx[i];
        c->tris++;
        if (ray_triangle(s->tri9 + 9 * (int64_t)tri, r, &t)) {
            if (t > 0 && t < rec->t) { rec->t = t; rec->slot = (int)i; }
        }
    }
}

/* slab test vs the scene box, R/src/CUDAKernels.cu:237-262 */
static inline int scene_slab(const orc_scene* s, const orc_ray* r, float* tmin_out, float* tmax_out) {
    const float* bb[2] = { s->scene_lo, s->scene_hi };
    float tMin = (bb[r->sign[0]][0] - r->o[0]) * r->inv[0];
    float tMax = (bb[1 - r->sign[0]][0] - r->o[0]) * r->inv[0];
    float tymin = (bb[r->sign[1]][1] - r->o[1]) * r->inv[1];
    float tymax = (bb[1 - r->sign[1]][1] - r->o[1]) * r->inv[1];
    if ((tMin > tymax) || (tymin > tMax)) return 0;
    if (tymin > tMin) tMin = tymin;
    if (tymax < tMax) tMax = tymax;
    float tzmin = (bb[r->sign[2]][2] - r->o[2]) * r->inv[2];
    float tzmax = (bb[1 - r->sign[2]][2] - r->o[2]) * r->inv[2];
    if ((tMin > tzmax) || (tzmin > tMax)) return 0;
    if (tzmin > tMin) tMin = tzmin;
    if (tzmax < tMax) tMax = tzmax;
    *tmin_out = tMin; *tmax_out = tMax;
    return 1;
}

/* TraverseTree, R/src/CUDAKernels.cu:227-368, literal (node pointers -> indices, nullptr -> -1) */
static void traverse_ref(const orc_scene* s, const orc_ray* r, orc_hit* rec, orc_counters* c) {
    float tMin, tMax;
    if (s->nu <= 0) return;
    if (!scene_slab(s, r, &tMin, &tMax)) return;
    if (s->nu == 1) { find_nearest(s, r, 0, rec, c); return; }      /* reference: undefined; see orc_build_tree */
    int cur = 0;                                                    /* :266 */
    float t[2];
    struct { int node; float tMin, tMax; } stack[64];               /* :276 */
    int sp = 0;
    stack[sp].node = -1; stack[sp].tMin = 0; stack[sp].tMax = 0; sp++; /* :278-279 (t fields uninitialised there) */
    while (cur != -1) {                                             /* :280 */
        c->nodes++;
        int ax = s->axis[cur];
        float org = r->o[ax], inv = r->inv[ax];
        int near = r->sign[ax], far = 1 - near;                     /* :286-287 */
        t[0] = (s->clip[2 * cur] - org) * inv;                      /* :288-289 */
        t[1] = (s->clip[2 * cur + 1] - org) * inv;
        int a = (tMin < t[near]);                                   /* :292 */
        int b = (tMax < t[far]);                                    /* :293 */
        const uint8_t* lf = s->is_leaf + 2 * cur;
        const int32_t* ch = s->children + 2 * cur;
#define ORC_POP() do { sp--; cur = stack[sp].node; tMin = stack[sp].tMin; tMax = stack[sp].tMax; } while (0)
        if (!a && b) {                                              /* :300-305 */
            ORC_POP();
        } else if (a && b) {                                        /* :306-319 */
            if (lf[near]) { find_nearest(s, r, ch[near], rec, c); ORC_POP(); }
            else { cur = ch[near]; tMax = t[near]; }
        } else if (!a && !b) {                                      /* :320-333 */
            if (lf[far]) { find_nearest(s, r, ch[far], rec, c); ORC_POP(); }
            else { cur = ch[far]; tMin = t[far]; }
        } else {                                                    /* :334-365 */
            if (lf[near] && lf[far]) {
                find_nearest(s, r, ch[near], rec, c);
                find_nearest(s, r, ch[far], rec, c);
                ORC_POP();
            } else if (!lf[near] && lf[far]) {
                find_nearest(s, r, ch[far], rec, c);
                cur = ch[near]; tMax = t[near];
            } else if (lf[near] && !lf[far]) {
                find_nearest(s, r, ch[near], rec, c);
                cur = ch[far]; tMin = t[far];
            } else {
                stack[sp].node = ch[far]; stack[sp].tMin = t[far]; stack[sp].tMax = tMax; sp++;
                if (sp > c->max_stack) c->max_stack = sp;
                cur = ch[near]; tMax = t[near];
            }
        }
#undef ORC_POP
    }
}

/* Pruned ("proper") BIH traversal over the SAME tree that returns what TraverseTree returns.
 * This is the logical per-ray order the CUDA kernel implements (csrc/trace.cu); its node and
 * triangle counters define V_n and V_t of SURVEY.md 8(d).
 *
 * It carries two intervals: the reference's own tMin (rMin: overwritten on far descents exactly as
 * R/src/CUDAKernels.cu:331,354,359) and a tight interval [pMin,pMax] (interval INTERSECTION, pMin
 * clamped to 0 because hits need t > 0 (:218), pMax also bounded by the closest hit so far).
 *   near child considered <=> reference's strict test (rMin < t[near], :292)  AND  the CLOSED tight
 *                             interval [pMin, min(pMax, t[near])] is non-empty;
 *   far child considered  <=> the CLOSED tight interval [max(pMin, t[far]), pMax] is non-empty
 *                             (this implies the reference's !(tMax < t[far]), :293, because the
 *                             reference's tMax is never below pMax);
 *   a considered child (leaf or node) becomes an item (reference + bounds) and is actually entered
 *   only if its tight interval, cut at the closest hit found so far, is still non-empty when its
 *   turn comes.
 * So the visited leaves are a subset of the reference's, and a leaf is only dropped when its
 * tight interval is empty or starts beyond the closest hit: the result equals TraverseTree's
 * except when Moller-Trumbore's rounded t disagrees with the rounded plane distances at an
 * interval end (the documented tie class).  The closed tests matter for axis-aligned flat
 * geometry (walls, floors) where the two clip planes and the triangle plane coincide bit for bit:
 * there the reference's strict test decides and is reproduced through rMin.
 * Leaf test order is the reference's: near leaf, then far leaf; a far LEAF next to a near
 * INTERNAL child is tested before descending (:344-349), so exact-t ties resolve identically. */
static void traverse_proper(const orc_scene* s, const orc_ray* r, orc_hit* rec, orc_counters* c) {
    float rMin, sMax;
    if (s->nu <= 0) return;
    if (!scene_slab(s, r, &rMin, &sMax)) return;
    if (s->nu == 1) { find_nearest(s, r, 0, rec, c); return; }
    float pMin = dev_fmaxf(rMin, 0.0f), pMax = sMax;
    /* item = child reference (>= 0: internal node, < 0: leaf ~ref) + its three interval bounds;
     * a leaf is an item like a node, so the traversal has one stack and one entry check */
    struct { int ref; float rMin, pMin, pMax; } stack[64];
    int sp = 0;
    int cur = 0;
    for (;;) {
        int next = 0;
        /* entry check: tight interval, cut at the closest hit so far, must be non-empty (closed) */
        if (pMin <= dev_fminf(pMax, (float)rec->t)) {
            if (cur < 0) {
                find_nearest(s, r, ~cur, rec, c);
            } else {
                c->nodes++;
                int ax = s->axis[cur];
                float org = r->o[ax], inv = r->inv[ax];
                int near = r->sign[ax], far = 1 - near;
                float t0 = (s->clip[2 * cur] - org) * inv;
                float t1 = (s->clip[2 * cur + 1] - org) * inv;
                float tn = near ? t1 : t0, tf = near ? t0 : t1;
                const uint8_t* lf = s->is_leaf + 2 * cur;
                const int32_t* ch = s->children + 2 * cur;
                int refn = lf[near] ? ~ch[near] : ch[near], reff = lf[far] ? ~ch[far] : ch[far];
                float nMax = dev_fminf(pMax, tn);                   /* near: [pMin, nMax] */
                float fMin = dev_fmaxf(pMin, tf);                   /* far : [fMin, pMax] */
                int go_near = (rMin < tn) && (pMin <= nMax);        /* reference's strict test (:292) + closed tight interval */
                int go_far = (fMin <= pMax);
                if (go_near && go_far) {
                    /* near before far, except a far LEAF next to a near NODE is tested first (:344-349) */
                    if (refn >= 0 && reff < 0) {
                        stack[sp].ref = refn; stack[sp].rMin = rMin; stack[sp].pMin = pMin; stack[sp].pMax = nMax;
                        cur = reff; rMin = tf; pMin = fMin;
                    } else {
                        stack[sp].ref = reff; stack[sp].rMin = tf; stack[sp].pMin = fMin; stack[sp].pMax = pMax;
                        cur = refn; pMax = nMax;
                    }
                    sp++;
                    if (sp > c->max_stack) c->max_stack = sp;
                    next = 1;
                } else if (go_near) { cur = refn; pMax = nMax; next = 1; }
                else if (go_far) { cur = reff; rMin = tf; pMin = fMin; next = 1; }
            }
        }
        if (next) continue;
        if (sp == 0) break;
        sp--;
        cur = stack[sp].ref; rMin = stack[sp].rMin; pMin = stack[sp].pMin; pMax = stack[sp].pMax;
    }
}

/* mode 3 -- what the shipped kernel does (csrc/trace.cu): traverse_proper plus the BOUNDING BOXES OF THE TWO CHILDREN,
 * which the B200 node record carries next to the reference's two clip planes (csrc/bihrt_internal.cuh:BihNode).  A child
 * whose box the ray misses inside the child's interval is not entered, and an entered child's interval is cut to the
 * box.  The boxes are exact min / max of vertex coordinates, so the leaves entered are a SUBSET of traverse_proper's, in
 * the same order: a leaf is skipped only when the ray cannot hit a triangle in it before the closest hit so far -- the
 * result equals traverse_proper's (and hence TraverseTree's) except where Moller-Trumbore's rounded t disagrees with a
 * rounded box distance by more than the ray's margin (2^-22 x (largest finite |o/d| + larger end of its scene interval)), i.e. on grazing edge / vertex hits (the documented tie
 * class).  Same arithmetic, operation order and NaN behaviour as the kernel, so GPU results and counters are compared
 * bit for bit against this mode; tests/test_oracle_semantics.py ties it back to traverse_ref. */
#define BOX_EPS 2.384185791015625e-07f      /* 2^-22 */
/* plane distance as ONE fused multiply-add: t = fma(plane, 1/d, -(o/d)).  The rounding of o/d is an absolute error of
 * 2^-24 |o/d| in t, so every box interval is widened by pad = 2^-22 x the largest finite |o/d| of the ray, plus 2^-22 x the
 * larger end of the ray's scene interval for the rounding of the FMA itself (see make_boxray). */
typedef struct { float inv[3], n[3], pad; } orc_boxray;
static inline void make_boxray(orc_boxray* q, const orc_ray* r, float sMin, float sMax) {
    float m = 0.0f;
    for (int k = 0; k < 3; k++) {
        q->n[k] = -(r->o[k] * r->inv[k]);
        float a = fabsf(q->n[k]);
        m = dev_fmaxf(m, a <= FLT_MAX ? a : 0.0f);
        /* a zero direction component: fma(plane, inf, -(o * inf)) is NaN or a signed infinity depending on signs, never a
         * distance; NaN as 1 / d makes both plane distances of the axis NaN and the NaN-dropping min / max leave it out */
        q->inv[k] = fabsf(r->inv[k]) <= FLT_MAX ? r->inv[k] : NAN;
    }
    /* + 2^-22 x the larger end of the ray's interval inside the scene box: the relative rounding error of the FMA, bounded once
     * per ray for every distance that can decide a test (those at an end of the current interval) */
    float tabs = dev_fmaxf(fabsf(sMin), fabsf(sMax));
    q->pad = fmaf(BOX_EPS, tabs <= FLT_MAX ? tabs : FLT_MAX, BOX_EPS * m);
}
static inline void box_slab(const float* b /* lo.xyz hi.xyz */, const orc_ray* r, const orc_boxray* q, float* bn, float* bf) {
    (void)r;
    float ax0 = fmaf(b[0], q->inv[0], q->n[0]), ax1 = fmaf(b[3], q->inv[0], q->n[0]);
    float ay0 = fmaf(b[1], q->inv[1], q->n[1]), ay1 = fmaf(b[4], q->inv[1], q->n[1]);
    float az0 = fmaf(b[2], q->inv[2], q->n[2]), az1 = fmaf(b[5], q->inv[2], q->n[2]);
    float n_ = dev_fmaxf(dev_fmaxf(dev_fminf(ax0, ax1), dev_fminf(ay0, ay1)), dev_fminf(az0, az1));
    float f_ = dev_fminf(dev_fminf(dev_fmaxf(ax0, ax1), dev_fmaxf(ay0, ay1)), dev_fmaxf(az0, az1));
    *bn = n_ - q->pad;
    *bf = f_ + q->pad;
}

static void traverse_box(const orc_scene* s, const float* cbox /* 12 floats per node: left box, right box */,
                         const orc_ray* r, orc_hit* rec, orc_counters* c) {
    float rMin, sMax;
    if (s->nu <= 0) return;
    if (!scene_slab(s, r, &rMin, &sMax)) return;
    if (s->nu == 1) { find_nearest(s, r, 0, rec, c); return; }
    float pMin = dev_fmaxf(rMin, 0.0f), pMax = sMax;
    struct { int ref; float rMin, pMin, pMax; } stack[64];
    int sp = 0;
    int cur = 0;
    orc_boxray q;
    make_boxray(&q, r, rMin, sMax);
    for (;;) {
        int next = 0;
        if (pMin <= dev_fminf(pMax, (float)rec->t)) {
            pMax = dev_fminf(pMax, (float)rec->t);
            if (cur < 0) {
                find_nearest(s, r, ~cur, rec, c);
            } else {
                c->nodes++;
                int ax = s->axis[cur];
                float org = r->o[ax], inv = r->inv[ax];
                int near = r->sign[ax], far = 1 - near;
                float t0 = (s->clip[2 * cur] - org) * inv;
                float t1 = (s->clip[2 * cur + 1] - org) * inv;
                float tn = near ? t1 : t0, tf = near ? t0 : t1;
                const uint8_t* lf = s->is_leaf + 2 * cur;
                const int32_t* ch = s->children + 2 * cur;
                int refn = lf[near] ? ~ch[near] : ch[near], reff = lf[far] ? ~ch[far] : ch[far];
                float nMax = dev_fminf(pMax, tn);
                float fMin = dev_fmaxf(pMin, tf);
                float bn[2], bf[2];
                box_slab(cbox + 12 * (int64_t)cur, r, &q, &bn[0], &bf[0]);
                box_slab(cbox + 12 * (int64_t)cur + 6, r, &q, &bn[1], &bf[1]);
                float nLo = dev_fmaxf(pMin, bn[near]), nHi = dev_fminf(nMax, bf[near]);
                float fLo = dev_fmaxf(fMin, bn[far]), fHi = dev_fminf(pMax, bf[far]);
                int go_near = (rMin < tn) && (nLo <= nHi);
                int go_far = (fLo <= fHi);
                if (go_near && go_far) {
                    if (refn >= 0 && reff < 0) {
                        stack[sp].ref = refn; stack[sp].rMin = rMin; stack[sp].pMin = nLo; stack[sp].pMax = nHi;
                        cur = reff; rMin = tf; pMin = fLo; pMax = fHi;
                    } else {
                        stack[sp].ref = reff; stack[sp].rMin = tf; stack[sp].pMin = fLo; stack[sp].pMax = fHi;
                        cur = refn; pMin = nLo; pMax = nHi;
                    }
                    sp++;
                    if (sp > c->max_stack) c->max_stack = sp;
                    next = 1;
                } else if (go_near) { cur = refn; pMin = nLo; pMax = nHi; next = 1; }
                else if (go_far) { cur = reff; rMin = tf; pMin = fLo; pMax = fHi; next = 1; }
            }
        }
        if (next) continue;
        if (sp == 0) break;
        sp--;
        cur = stack[sp].ref; rMin = stack[sp].rMin; pMin = stack[sp].pMin; pMax = stack[sp].pMax;
    }
}

/* children boxes of every internal node: exact min / max over the vertices below each child (what the GPU build derives
 * from its min / max heaps, csrc/build.cu:k_nodes).  cbox: 12 floats per node. */
static void children_boxes(const orc_scene* s, float* cbox) {
    int64_t nu = s->nu, ni = nu > 1 ? nu - 1 : 0;
    float* lbox = (float*)malloc(sizeof(float) * 6 * (size_t)(nu > 0 ? nu : 1));
    float* nbox = (float*)malloc(sizeof(float) * 6 * (size_t)(ni > 0 ? ni : 1));
    char* done = (char*)calloc((size_t)(ni > 0 ? ni : 1), 1);
    for (int64_t l = 0; l < nu; l++) {
        float* b = lbox + 6 * l;
        for (int k = 0; k < 3; k++) { b[k] = INFINITY; b[3 + k] = -INFINITY; }
        for (uint32_t j = 0; j < s->cnt[l]; j++) {
            const float* t = s->tri9 + 9 * (int64_t)s->tris_idx[s->first[l] + j];
            for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) {
                b[k] = dev_fminf(b[k], t[3 * v + k]); b[3 + k] = dev_fmaxf(b[3 + k], t[3 * v + k]);
            }
        }
    }
    /* children have arbitrary indices (Karras numbering): sweep until every node has both children's boxes */
    int64_t remaining = ni;
    while (remaining > 0)
        for (int64_t i = 0; i < ni; i++) if (!done[i]) {
            const float* cb[2]; int ok = 1;
            for (int k = 0; k < 2; k++) {
                int chd = s->children[2 * i + k];
                if (s->is_leaf[2 * i + k]) cb[k] = lbox + 6 * (int64_t)chd;
                else if (done[chd]) cb[k] = nbox + 6 * (int64_t)chd;
                else ok = 0;
            }
            if (!ok) continue;
            float* b = nbox + 6 * i;
            for (int k = 0; k < 3; k++) { b[k] = dev_fminf(cb[0][k], cb[1][k]); b[3 + k] = dev_fmaxf(cb[0][3 + k], cb[1][3 + k]); }
            memcpy(cbox + 12 * i, cb[0], 24); memcpy(cbox + 12 * i + 6, cb[1], 24);
            done[i] = 1; remaining--;
        }
    free(done); free(nbox); free(lbox);
}

/* brute force over every slot in slot order (the reference's commented-out TraverseTriangles,
 * R/src/CUDAKernels.cu:157-202, visits leaves in node order instead; same result except exact ties) */
static void traverse_brute(const orc_scene* s, const orc_ray* r, orc_hit* rec, orc_counters* c) {
    float t = FLT_MAX;
    for (int64_t i = 0; i < s->n; i++) {
        c->tris++;
        if (ray_triangle(s->tri9 + 9 * (int64_t)s->tris_idx[i], r, &t))
            if (t > 0 && t < rec->t) { rec->t = t; rec->slot = (int)i; }
    }
}

/* mode: 0 = reference semantics, 1 = proper traversal, 2 = brute force, 3 = proper traversal + children boxes (the shipped kernel).
 * rays6: o.xyz d.xyz per ray.  Outputs: t (FLT_MAX on miss), slot (-1), prim = tris_idx[slot] (-1).
 * counters3 (may be NULL): total nodes visited, total triangle tests, max stack depth. */
ORC_API void orc_trace(int mode, const float* tri9, int64_t n, int64_t nu,
                       const uint32_t* tris_idx, const uint32_t* cnt, const int32_t* first,
                       const float* clip, const int32_t* axis, const uint8_t* is_leaf,
                       const int32_t* children, const float* scene_lo, const float* scene_hi,
                       const float* rays6, int64_t nrays, float* out_t, int32_t* out_slot,
                       int32_t* out_prim, uint64_t* counters3, int nthreads) {
    orc_scene s = { tri9, n, nu, tris_idx, cnt, first, clip, axis, is_leaf, children, {0, 0, 0}, {0, 0, 0} };
    for (int k = 0; k < 3; k++) { s.scene_lo[k] = scene_lo[k]; s.scene_hi[k] = scene_hi[k]; }
    uint64_t tot_nodes = 0, tot_tris = 0; int max_stack = 0;
    float* cbox = NULL;
    if (mode == 3) { cbox = (float*)malloc(sizeof(float) * 12 * (size_t)(nu > 1 ? nu - 1 : 1)); children_boxes(&s, cbox); }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+:tot_nodes, tot_tris) reduction(max:max_stack)
#endif
    for (int64_t i = 0; i < nrays; i++) {
        orc_ray r;
        make_ray(&r, rays6 + 6 * i, rays6 + 6 * i + 3);
        orc_hit rec = { FLT_MAX, -1 };                              /* Color(), R/src/CUDAKernels.cu:380-382 */
        orc_counters c = { 0, 0, 0 };
        if (mode == 0) traverse_ref(&s, &r, &rec, &c);
        else if (mode == 1) traverse_proper(&s, &r, &rec, &c);
        else if (mode == 3) traverse_box(&s, cbox, &r, &rec, &c);
        else traverse_brute(&s, &r, &rec, &c);
        out_t[i] = (float)rec.t;
        out_slot[i] = rec.slot;
        if (out_prim) out_prim[i] = rec.slot >= 0 ? (int32_t)tris_idx[rec.slot] : -1;
        tot_nodes += c.nodes; tot_tris += c.tris;
        if (c.max_stack > max_stack) max_stack = c.max_stack;
    }
    free(cbox);
    if (counters3) { counters3[0] = tot_nodes; counters3[1] = tot_tris; counters3[2] = (uint64_t)max_stack; }
}

/* ------------------------------------------------------------------------------------------ */
/* camera rays + pixel packing                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* Counter-based jitter replacing the reference's per-pixel XORWOW state (curand_init(1984, pixel,
 * 0), R/src/CUDAKernels.cu:411-419,458): DELIBERATE DIFFERENCE, parity unpinned in the reference.
 * Same integer ops as csrc/common.cuh:bihrt_jitter so CPU and GPU samples are bit-identical.
 * Returns a float in (0,1] like curand_uniform. */
static inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline float jitter01(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) {
    uint32_t h = mix32((uint32_t)seed ^ mix32(pixel + 0x9e3779b9u * (sample * 2u + dim + 1u)) ^ (uint32_t)(seed >> 32));
    return (float)((h >> 8) + 1u) * (1.0f / 16777216.0f);
}

/* Camera::GetRay, R/src/Camera.cu:18-20: dir = llc + u*horizontal + v*vertical - origin.
 * cam12 = origin.xyz, lowerLeftCorner.xyz, horizontal.xyz, vertical.xyz (R/src/Camera.h:14-17).
 * u,v per R/src/CUDAKernels.cu:414-415; jitter=0 gives pixel centres (offset 0.5). */
ORC_API void orc_camera_rays(const float* cam12, int w, int h, int spp, int jitter, uint64_t seed,
                             float* rays6) {
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++)
            for (int s = 0; s < spp; s++) {
                uint32_t pixel = (uint32_t)(j * w + i);
                float ru = jitter ? jitter01(seed, pixel, (uint32_t)s, 0) : 0.5f;
                float rv = jitter ? jitter01(seed, pixel, (uint32_t)s, 1) : 0.5f;
                float u = ((float)i + ru) / (float)w;
                float v = ((float)j + rv) / (float)h;
                float* r = rays6 + 6 * (((int64_t)j * w + i) * spp + s);
                for (int k = 0; k < 3; k++) {
                    r[k] = cam12[k];
                    r[3 + k] = cam12[3 + k] + u * cam12[6 + k] + v * cam12[9 + k] - cam12[k];
                }
            }
}

/* The same rays for a window [x0, x0+ww) x [y0, y0+wh) of the w x h frame (pixel ids and u,v are the FRAME's), in
 * window-row-major order: lets a test check a few tiles of a frame too large to trace on the CPU. */
ORC_API void orc_camera_rays_window(const float* cam12, int w, int h, int spp, int jitter, uint64_t seed,
                                    int x0, int y0, int ww, int wh, float* rays6) {
    for (int j = 0; j < wh; j++)
        for (int i = 0; i < ww; i++)
            for (int s = 0; s < spp; s++) {
                uint32_t pixel = (uint32_t)((y0 + j) * w + (x0 + i));
                float ru = jitter ? jitter01(seed, pixel, (uint32_t)s, 0) : 0.5f;
                float rv = jitter ? jitter01(seed, pixel, (uint32_t)s, 1) : 0.5f;
                float u = ((float)(x0 + i) + ru) / (float)w;
                float v = ((float)(y0 + j) + rv) / (float)h;
                float* r = rays6 + 6 * (((int64_t)j * ww + i) * spp + s);
                for (int k = 0; k < 3; k++) {
                    r[k] = cam12[k];
                    r[3 + k] = cam12[3 + k] + u * cam12[6 + k] + v * cam12[9 + k] - cam12[k];
                }
            }
}

/* cudaRender accumulate + rgbToInt, R/src/CUDAKernels.cu:82-88,385-387,412-422:
 * hit -> (255,255,0), miss -> (20,20,40), mean over spp, pack r | g<<8 | b<<16.
 * hit_slot has w*h*spp entries in the order produced by orc_camera_rays. */
ORC_API void orc_pack_framebuffer(const int32_t* hit_slot, int w, int h, int spp, uint32_t* fb) {
    for (int64_t p = 0; p < (int64_t)w * h; p++) {
        float col[3] = { 0, 0, 0 };
        for (int s = 0; s < spp; s++) {
            int hit = hit_slot[p * spp + s] >= 0;
            col[0] += hit ? 255.0f : 20.0f;
            col[1] += hit ? 255.0f : 20.0f;
            col[2] += hit ? 0.0f : 40.0f;
        }
        for (int k = 0; k < 3; k++) {
            col[k] /= (float)spp;
            col[k] = dev_fmaxf(0.0f, dev_fminf(255.0f, col[k]));
        }
        fb[p] = ((uint32_t)(int)col[2] << 16) | ((uint32_t)(int)col[1] << 8) | (uint32_t)(int)col[0];
    }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
