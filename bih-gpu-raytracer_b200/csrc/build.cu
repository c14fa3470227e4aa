// BIH construction on sm_100a.  Replaces the first half of Renderer::Render
// (R/src/Renderer.cpp:422-503: thrust::transform(morton_functor), thrust::sequence,
// thrust::stable_sort_by_key, thrust::reduce_by_key, thrust::unique_by_key_copy, Launch_BuildTree,
// Launch_FindClipPlanes) and the host pre-pass of App::LoadModels (R/src/App.cpp:110-156).
//
// Pipeline (one stream, no host synchronisation, Nu never leaves the device):
//   k_init        reset histograms / counters / scene-box accumulators
//   k_scene_box   scene AABB                                   36 B read per triangle
//   k_morton      AABB centre -> 30-bit Morton key, fused 4x256 digit histogram   36 R + 4 W
//   k_onesweep x4 stable LSD radix sort, 8-bit digits, decoupled look-back       16 R + 16 W per pass
//   k_rle         head flags -> unique codes, first slot of each leaf, Nu          4 R + <=8 W
//   k_tree        per leaf: gather + reorder triangles (36 R + 48 W), leaf AABB, then bottom-up
//                 agglomerative radix-tree construction that emits each node ONCE, in the reference's
//                 node numbering, with both clip planes                           ~64 R/W scratch + 16 W
// Results are bit-identical to the reference algorithm (oracle/bih_oracle.c): the radix tree over the
// unique sorted codes is unique, node ids follow the Karras numbering rule, clip planes are pure
// max/min of input floats.
#include "bihrt_internal.cuh"
#include <float.h>

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------
// small block-level helpers (256 threads)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}
// exclusive scan over blockDim.x == 256 threads; s_w needs 8 words; returns exclusive prefix, *total = block sum
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

// Stage `cnt` (<=256) triangles starting at triangle `base` (multiple of 256) into shared memory
// with coalesced 128-bit loads; thread t then reads its 9 floats at s[9*t] (stride 9: conflict-free).
__device__ __forceinline__ void stage_tris_256(const float* __restrict__ tri, uint32_t base, uint32_t cnt, float* s) {
    const float4* src = reinterpret_cast<const float4*>(tri + (size_t)base * 9);
    uint32_t nfl = cnt * 9, nv4 = nfl >> 2;
    for (uint32_t i = threadIdx.x; i < nv4; i += 256) reinterpret_cast<float4*>(s)[i] = __ldcs(src + i);
    for (uint32_t i = (nv4 << 2) + threadIdx.x; i < nfl; i += 256) s[i] = tri[(size_t)base * 9 + i];
    __syncthreads();
}

// std::minmax({a,b,c}) of R/src/App.cpp:123-125: leftmost minimum, rightmost maximum under <.
__device__ __forceinline__ void minmax3(float a, float b, float c, float& mn, float& mx) {
    mn = a; mx = a;
    if (b < mn) mn = b;
    if (!(b < mx)) mx = b;
    if (c < mn) mn = c;
    if (!(c < mx)) mx = c;
}

// d_hist layout (uint32 words)
#define H_HIST      0        // 4 x 256 digit counts
#define H_TILECTR   1024     // [0..3] onesweep tile counters, [4] rle tile counter
#define H_WORDS     1040

// ------------------------------------------------------------------------------------------
__global__ void k_init(uint32_t* hist, uint32_t* enc, BihHeader* hdr, uint32_t n) {
    for (int i = threadIdx.x; i < H_WORDS; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x < 3) { enc[threadIdx.x] = 0xFFFFFFFFu; enc[3 + threadIdx.x] = 0u; }
    if (threadIdx.x == 0) { hdr->n = n; hdr->nu = 0; hdr->status = 0; hdr->root_axis = 0; }
}

// ------------------------------------------------------------------------------------------
// scene AABB (R/src/App.cpp:103-106,133-137).  Sign of a zero bound may differ from the host's
// std::minmax order dependence; it cannot change any Morton code or hit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scene_box(const float* __restrict__ tri, uint32_t n, uint32_t* __restrict__ enc) {
    __shared__ __align__(16) float s[256 * 9];
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        uint32_t cnt = min(256u, n - base);
        stage_tris_256(tri, base, cnt, s);
        if (threadIdx.x < cnt) {
            const float* t = s + 9 * threadIdx.x;
#pragma unroll
            for (int v = 0; v < 3; v++)
#pragma unroll
                for (int k = 0; k < 3; k++) { float f = t[3 * v + k]; lo[k] = fminf(lo[k], f); hi[k] = fmaxf(hi[k], f); }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&enc[k], enc_float(lo[k])); atomicMax(&enc[3 + k], enc_float(hi[k])); }
    }
}

// ------------------------------------------------------------------------------------------
// Morton keys (R/src/App.cpp:128-131,144-156 + R/src/Renderer.cpp:116-136) + digit histograms
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton_axis(float mn, float mx, float slo, float shi) {
    float centre = __fdiv_rn(__fadd_rn(mn, mx), 2.0f);
    float nrm = __fdiv_rn(__fsub_rn(centre, slo), __fsub_rn(shi, slo));
    float q = fminf(fmaxf(__fmul_rn(nrm, 1024.0f), 0.0f), 1023.0f);   // fmaxf(NaN,0)=0: flat axis -> cell 0
    return expand_bits10(__float2uint_rz(q));
}

__global__ void __launch_bounds__(256) k_morton(const float* __restrict__ tri, uint32_t n, const uint32_t* __restrict__ enc,
                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ hist, BihHeader* hdr) {
    __shared__ __align__(16) float s[256 * 9];
    __shared__ uint32_t s_hist[4 * 256];
    for (int i = threadIdx.x; i < 1024; i += 256) s_hist[i] = 0;
    float slo[3], shi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { slo[k] = dec_float(enc[k]); shi[k] = dec_float(enc[3 + k]); }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { hdr->lo[k] = slo[k]; hdr->hi[k] = shi[k]; }
    }
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        uint32_t cnt = min(256u, n - base);
        stage_tris_256(tri, base, cnt, s);      // its __syncthreads also orders the s_hist zeroing
        bool valid = threadIdx.x < cnt;
        uint32_t code = 0;
        if (valid) {
            const float* t = s + 9 * threadIdx.x;
            float mn, mx;
            minmax3(t[0], t[3], t[6], mn, mx); uint32_t xx = morton_axis(mn, mx, slo[0], shi[0]);
            minmax3(t[1], t[4], t[7], mn, mx); uint32_t yy = morton_axis(mn, mx, slo[1], shi[1]);
            minmax3(t[2], t[5], t[8], mn, mx); uint32_t zz = morton_axis(mn, mx, slo[2], shi[2]);
            code = xx * 4 + yy * 2 + zz;
            keys[base + threadIdx.x] = code;
        }
        // warp-aggregated histogram: coherent meshes put whole warps into one bin
        uint32_t act = __ballot_sync(FULL, valid);
        if (valid) {
#pragma unroll
            for (int p = 0; p < 4; p++) {
                uint32_t d = (code >> (8 * p)) & 255u;
                uint32_t peers = __match_any_sync(act, d);
                if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[p * 256 + d], __popc(peers));
            }
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256) { uint32_t v = s_hist[i]; if (v) atomicAdd(&hist[H_HIST + i], v); }
}

// ------------------------------------------------------------------------------------------
// One pass of a stable least-significant-digit radix sort, 8-bit digit, single sweep over the data
// with chained-scan decoupled look-back across tiles (replaces thrust::sequence +
// thrust::stable_sort_by_key, R/src/Renderer.cpp:436-445).  Stability: a tile is ranked in index
// order (warp-striped items, per-warp match-any ranking), tiles are ordered by the look-back chain.
// ------------------------------------------------------------------------------------------
#define OS_THREADS 256
#define OS_ITEMS   16
#define OS_TILE    (OS_THREADS * OS_ITEMS)      // 4096 keys
#define LB_FLAG_AGG   0x40000000u
#define LB_FLAG_INCL  0x80000000u
#define LB_MASK       0x3FFFFFFFu
#define SPIN_LIMIT    (1u << 22)

template <bool FIRST>
__global__ void __launch_bounds__(OS_THREADS) k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                         uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                         uint32_t n, int pass, uint32_t* __restrict__ hist,
                                                         uint32_t* __restrict__ lookback, BihHeader* hdr) {
    __shared__ uint32_t s_whist[8][256];
    __shared__ uint32_t s_keys[OS_TILE];
    __shared__ uint32_t s_vals[OS_TILE];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint32_t s_goff[256];
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * 8;
    if (tid == 0) s_tile = atomicAdd(&hist[H_TILECTR + pass], 1u);
    for (int i = tid; i < 8 * 256; i += OS_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * OS_TILE;
    const uint32_t valid = min((uint32_t)OS_TILE, n - tile_base);

    // global base of each digit for this pass = exclusive scan of the whole-array histogram
    uint32_t tot;
    uint32_t gbase = block_excl_scan_256(hist[H_HIST + pass * 256 + tid], s_w, &tot);

    uint32_t key[OS_ITEMS], rank[OS_ITEMS];
    const uint32_t i0 = tile_base + warp * (32 * OS_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
        uint32_t gi = i0 + i * 32;
        key[i] = gi < n ? __ldcs(keys_in + gi) : 0xFFFFFFFFu;
    }
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
        uint32_t d = (key[i] >> shift) & 255u;
        uint32_t peers = __match_any_sync(FULL, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = s_whist[warp][d]; s_whist[warp][d] = old + __popc(peers); }
        old = __shfl_sync(FULL, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // digit `tid`: exclusive scan over the 8 warps, then over digits
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t c = s_whist[w][tid]; s_whist[w][tid] = cnt; cnt += c; }
    uint32_t binstart = block_excl_scan_256(cnt, s_w, &tot);
    s_binstart[tid] = binstart;
    // padding keys (0xFFFFFFFF, digit 255 in every pass) sit at the very end of the tile: not counted
    uint32_t cnt_real = cnt - ((tid == 255) ? ((uint32_t)OS_TILE - valid) : 0u);

    // decoupled look-back: exclusive count of this digit over all previous tiles
    uint32_t excl = 0;
    uint32_t* lb = lookback + (size_t)tile * 256 + tid;
    if (tile == 0) {
        st_relaxed(lb, cnt_real | LB_FLAG_INCL);
    } else {
        st_relaxed(lb, cnt_real | LB_FLAG_AGG);
        const uint32_t* p = lb - 256;
        uint32_t spins = 0;
        for (;;) {
            uint32_t v = ld_relaxed(p);
            uint32_t f = v & ~LB_MASK;
            if (f == 0) { if (++spins > SPIN_LIMIT) { atomicOr(&hdr->status, 1u); break; } continue; }
            excl += v & LB_MASK;
            if (f == LB_FLAG_INCL) break;
            p -= 256;
        }
        st_relaxed(lb, (excl + cnt_real) | LB_FLAG_INCL);
    }
    s_goff[tid] = gbase + excl - binstart;
    __syncthreads();

    // tile-local reorder through shared memory so the global scatter is coalesced per digit
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
        uint32_t d = (key[i] >> shift) & 255u;
        uint32_t pos = s_binstart[d] + s_whist[warp][d] + rank[i];
        s_keys[pos] = key[i];
        uint32_t gi = i0 + i * 32;
        uint32_t v;
        if (FIRST) v = gi;                                   // thrust::sequence fused: value = input index
        else v = gi < n ? __ldcs(vals_in + gi) : 0u;
        s_vals[pos] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OS_ITEMS; k++) {
        uint32_t j = tid + k * OS_THREADS;
        if (j < valid) {
            uint32_t kk = s_keys[j];
            uint32_t dst = s_goff[(kk >> shift) & 255u] + j;
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Run-length encoding of the sorted keys: unique codes, first slot of every leaf, Nu
// (replaces thrust::reduce_by_key + thrust::unique_by_key_copy, R/src/Renderer.cpp:450-472;
// duplicatesCnts[k] = first[k+1] - first[k]).  Single sweep, look-back on one word per tile.
// ------------------------------------------------------------------------------------------
#define RLE_ITEMS 8
#define RLE_TILE  (256 * RLE_ITEMS)

__global__ void __launch_bounds__(256) k_rle(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ umc,
                                             uint32_t* __restrict__ first, uint32_t* __restrict__ hist,
                                             uint32_t* __restrict__ lookback, BihHeader* hdr) {
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_tile, s_excl;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(&hist[H_TILECTR + 4], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t ntiles = (n + RLE_TILE - 1) / RLE_TILE;
    const uint32_t g0 = tile * RLE_TILE + tid * RLE_ITEMS;
    uint32_t key[RLE_ITEMS];
    if (g0 + RLE_ITEMS <= n) {
        uint4 a = *reinterpret_cast<const uint4*>(keys + g0), b = *reinterpret_cast<const uint4*>(keys + g0 + 4);
        key[0] = a.x; key[1] = a.y; key[2] = a.z; key[3] = a.w; key[4] = b.x; key[5] = b.y; key[6] = b.z; key[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < RLE_ITEMS; i++) key[i] = (g0 + i < n) ? keys[g0 + i] : 0u;
    }
    uint32_t prev = (g0 > 0 && g0 < n) ? keys[g0 - 1] : 0u;
    uint32_t heads = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < RLE_ITEMS; i++) {
        uint32_t gi = g0 + i;
        bool h = (gi < n) && (gi == 0 || key[i] != prev);
        prev = key[i];
        heads |= (h ? 1u : 0u) << i;
        cnt += h;
    }
    uint32_t total;
    uint32_t toff = block_excl_scan_256(cnt, s_w, &total);
    if (tid == 0) {
        uint32_t excl = 0;
        if (tile == 0) {
            st_relaxed(lookback, total | LB_FLAG_INCL);
        } else {
            st_relaxed(lookback + tile, total | LB_FLAG_AGG);
            const uint32_t* p = lookback + tile - 1;
            uint32_t spins = 0;
            for (;;) {
                uint32_t v = ld_relaxed(p);
                uint32_t f = v & ~LB_MASK;
                if (f == 0) { if (++spins > SPIN_LIMIT) { atomicOr(&hdr->status, 2u); break; } continue; }
                excl += v & LB_MASK;
                if (f == LB_FLAG_INCL) break;
                p -= 1;
            }
            st_relaxed(lookback + tile, (excl + total) | LB_FLAG_INCL);
        }
        s_excl = excl;
        if (tile == ntiles - 1) { uint32_t nu = excl + total; hdr->nu = nu; first[nu] = n; }
    }
    __syncthreads();
    uint32_t k = s_excl + toff;
#pragma unroll
    for (int i = 0; i < RLE_ITEMS; i++) {
        if (heads & (1u << i)) { umc[k] = key[i]; first[k] = g0 + i; k++; }
    }
}

// ------------------------------------------------------------------------------------------
// Leaves + tree.  One thread per leaf (unique Morton cell):
//   1. gather the leaf's triangles in sorted order, write the 48-byte leaf-ordered records, and take
//      the union of their AABBs (FindClipPlanes' first loop, R/src/CUDAKernels.cu:511-529);
//   2. climb: bottom-up agglomerative construction of the binary radix tree over the unique codes.
//      A range [l,r] of leaves joins the neighbour it shares the longer Morton prefix with; the second
//      child to arrive at a split owns the new node.  That yields exactly the tree BuildTree's
//      per-node binary searches find (R/src/CUDAKernels.cu:591-710) -- the radix tree of a sorted set
//      of distinct keys is unique -- and each node is written once, with both clip planes
//      (clip0 = max hi[axis] of the left subtree, clip1 = min lo[axis] of the right subtree), instead
//      of O(depth) float atomics per leaf that all meet at the root (R/src/CUDAKernels.cu:532-547).
//   Node numbering = the reference's: a node covering leaves [a,b] is stored at index b if it is a left
//   child, a if it is a right child, 0 for the root, so node i here IS
//   TreeInternalNode i there.
// ------------------------------------------------------------------------------------------
struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ float pick(const float v[3], int axis) { return axis == 0 ? v[0] : (axis == 1 ? v[1] : v[2]); }

__global__ void __launch_bounds__(128) k_tree(const float* __restrict__ tri_in, const uint32_t* __restrict__ idx_sorted,
                                              const uint32_t* __restrict__ umc, const uint32_t* __restrict__ first,
                                              BihHeader* hdr, BihNode* __restrict__ nodes,
                                              BihTri* __restrict__ tris, int32_t* __restrict__ arrive,
                                              float4* __restrict__ boxscratch) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nu = hdr->nu;
    if (k >= nu) return;
    const uint32_t s0 = first[k], s1 = first[k + 1];
    Box box;
    for (uint32_t j = s0; j < s1; j++) {
        const uint32_t p = idx_sorted[j];
        const float* t = tri_in + (size_t)p * 9;
        float v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = __ldg(t + i);
        float mn[3], mx[3];
        minmax3(v[0], v[3], v[6], mn[0], mx[0]);
        minmax3(v[1], v[4], v[7], mn[1], mx[1]);
        minmax3(v[2], v[5], v[8], mn[2], mx[2]);
        if (j == s0) {
#pragma unroll
            for (int a = 0; a < 3; a++) { box.lo[a] = mn[a]; box.hi[a] = mx[a]; }
        } else {
#pragma unroll
            for (int a = 0; a < 3; a++) { box.lo[a] = fminf(box.lo[a], mn[a]); box.hi[a] = fmaxf(box.hi[a], mx[a]); }
        }
        float4 q0 = make_float4(v[0], v[1], v[2], __fsub_rn(v[3], v[0]));
        float4 q1 = make_float4(__fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]), __fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]));
        float4 q2 = make_float4(__fsub_rn(v[8], v[2]), __uint_as_float(p), __uint_as_float(j + 1 == s1 ? 1u : 0u), __uint_as_float(j));
        float4* dst = reinterpret_cast<float4*>(tris + j);
        dst[0] = q0; dst[1] = q1; dst[2] = q2;
    }
    if (nu < 2) return;

    uint32_t left = k, right = k;
    bool have_node = false;
    float cl0 = 0.f, cl1 = 0.f;
    uint32_t ref_l = 0, ref_r = 0;
    for (;;) {
        const bool is_root = (left == 0 && right == nu - 1);
        bool am_left = false;
        if (!is_root) {
            if (left == 0) am_left = true;
            else if (right == nu - 1) am_left = false;
            else {
                int dl = __clz(umc[left - 1] ^ umc[left]);
                int dr = __clz(umc[right] ^ umc[right + 1]);
                am_left = dr > dl;       // never equal for distinct sorted keys
            }
        }
        if (have_node) {
            uint32_t idx = is_root ? 0u : (am_left ? right : left);
            BihNode nd; nd.clip0 = cl0; nd.clip1 = cl1; nd.ref_l = ref_l; nd.ref_r = ref_r;
            *reinterpret_cast<float4*>(nodes + idx) = *reinterpret_cast<float4*>(&nd);
        }
        if (is_root) break;
        const uint32_t ps = am_left ? right : left - 1;     // parent splits between leaves ps and ps+1
        float4* mine = boxscratch + (size_t)ps * 4 + (am_left ? 0 : 2);
        __stcg(mine, make_float4(box.lo[0], box.lo[1], box.lo[2], box.hi[0]));
        __stcg(mine + 1, make_float4(box.hi[1], box.hi[2], 0.f, 0.f));
        __threadfence();
        const int other = atomicExch(&arrive[ps], (int)(am_left ? left : right));
        if (other < 0) return;                               // first child to arrive leaves its box behind
        __threadfence();
        const float4* sib = boxscratch + (size_t)ps * 4 + (am_left ? 2 : 0);
        float4 b0 = __ldcg(sib), b1 = __ldcg(sib + 1);
        Box sb; sb.lo[0] = b0.x; sb.lo[1] = b0.y; sb.lo[2] = b0.z; sb.hi[0] = b0.w; sb.hi[1] = b1.x; sb.hi[2] = b1.y;
        const Box& lbox = am_left ? box : sb;
        const Box& rbox = am_left ? sb : box;
        const uint32_t a = am_left ? left : (uint32_t)other;
        const uint32_t b = am_left ? (uint32_t)other : right;
        const int axis = (__clz(umc[ps] ^ umc[ps + 1]) + 1) % 3;    // R/src/CUDAKernels.cu:702-706
        cl0 = pick(lbox.hi, axis);
        cl1 = pick(rbox.lo, axis);
        // a radix-tree node over leaves [l,r] splits on the first bit in which umc[l] and umc[r] differ,
        // so a child's axis follows from its range ends (== (lcp(children of the child) + 1) % 3)
        ref_l = (a == ps) ? BIH_REF_LEAFREF(first[ps]) : BIH_REF_NODE(ps, (__clz(umc[a] ^ umc[ps]) + 1) % 3);
        ref_r = (ps + 1 == b) ? BIH_REF_LEAFREF(first[ps + 1]) : BIH_REF_NODE(ps + 1, (__clz(umc[ps + 1] ^ umc[b]) + 1) % 3);
        if (a == 0 && b == nu - 1) hdr->root_axis = (uint32_t)axis;
        Box u;
#pragma unroll
        for (int i = 0; i < 3; i++) { u.lo[i] = fminf(lbox.lo[i], rbox.lo[i]); u.hi[i] = fmaxf(lbox.hi[i], rbox.hi[i]); }
        box = u;
        left = a; right = b;
        have_node = true;
    }
}

// ------------------------------------------------------------------------------------------
int bihrt_build_launch(bihrt_ctx* c) {
    const uint32_t n = (uint32_t)c->n;
    cudaStream_t st = c->stream;
    const uint32_t os_tiles = (n + OS_TILE - 1) / OS_TILE;
    const uint32_t rle_tiles = (n + RLE_TILE - 1) / RLE_TILE;
    const size_t lb_words = (size_t)4 * os_tiles * 256 + rle_tiles;
    if (lb_words > c->lookback_words) return bihrt_fail(c, BIHRT_ERR_INTERNAL, "look-back buffer too small");

    k_init<<<1, 256, 0, st>>>(c->d_hist, c->d_scenebox_enc, c->d_hdr, n);
    BIHRT_CUDA(c, cudaMemsetAsync(c->d_lookback, 0, lb_words * 4, st));
    BIHRT_CUDA(c, cudaMemsetAsync(c->d_arrive, 0xFF, (size_t)n * 4, st));
    const int stream_grid = (int)min((uint32_t)(c->sm_count * 8), (n + 255) / 256);
    k_scene_box<<<stream_grid, 256, 0, st>>>(c->d_tri_in, n, c->d_scenebox_enc);
    k_morton<<<stream_grid, 256, 0, st>>>(c->d_tri_in, n, c->d_scenebox_enc, c->d_keys[0], c->d_hist, c->d_hdr);
    int cur = 0;
    for (int pass = 0; pass < 4; pass++) {
        uint32_t* lb = c->d_lookback + (size_t)pass * os_tiles * 256;
        if (pass == 0)
            k_onesweep<true><<<os_tiles, OS_THREADS, 0, st>>>(c->d_keys[cur], nullptr, c->d_keys[cur ^ 1], c->d_vals[cur ^ 1], n, pass, c->d_hist, lb, c->d_hdr);
        else
            k_onesweep<false><<<os_tiles, OS_THREADS, 0, st>>>(c->d_keys[cur], c->d_vals[cur], c->d_keys[cur ^ 1], c->d_vals[cur ^ 1], n, pass, c->d_hist, lb, c->d_hdr);
        cur ^= 1;
    }
    // 4 passes: sorted data is back in buffer 0
    k_rle<<<rle_tiles, 256, 0, st>>>(c->d_keys[cur], n, c->d_umc, c->d_first, c->d_hist,
                                     c->d_lookback + (size_t)4 * os_tiles * 256, c->d_hdr);
    k_tree<<<(n + 127) / 128, 128, 0, st>>>(c->d_tri_in, c->d_vals[cur], c->d_umc, c->d_first, c->d_hdr, c->d_nodes, c->d_tris,
                                            c->d_arrive, reinterpret_cast<float4*>(c->d_boxscratch));
    c->kernel_launches += 9;    // k_init, k_scene_box, k_morton, 4 x k_onesweep, k_rle, k_tree
    BIHRT_CUDA(c, cudaGetLastError());
    return BIHRT_OK;
}
