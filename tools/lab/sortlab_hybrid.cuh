// VARIANT 2: hybrid sort = fine histogram (top 14 bits) -> plan (<= 251 buckets of <= 8192 keys, contiguous bin ranges) ->
// ONE stable global partition by bucket (the onesweep pass with digit = bucket) -> a shared-memory LSD sort of every bucket.
#define SL_ITEMS 16
#define SL_TILE (OS_THREADS * SL_ITEMS)
#define HY_BINS 16384
#define HY_BIN_SHIFT 16
#define HY_CAP 8192
#define HY_TARGET 250
struct HyPlan { uint32_t ok, nb, cq, pad; uint32_t boff[260]; uint8_t bmap[HY_BINS]; };

__global__ void k_fine_hist(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ F) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = keys[i] >> HY_BIN_SHIFT;
        const uint32_t act = __activemask();
        const uint32_t peers = __match_any_sync(act, b);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&F[b], __popc(peers));
    }
}

__global__ void __launch_bounds__(1024) k_plan(const uint32_t* __restrict__ F, uint32_t n, HyPlan* __restrict__ plan) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_bsize[256];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 256) s_bsize[tid] = 0;
    if (tid == 0) s_max = 0;
    uint32_t f[16], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 v = reinterpret_cast<const uint4*>(F)[tid * 4 + i];
        f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
        sum += v.x + v.y + v.z + v.w;
    }
    uint32_t inc = warp_incl_scan(sum, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) { uint32_t x = s_w[lane]; uint32_t xi = warp_incl_scan(x, lane); s_w[lane] = xi - x; }
    __syncthreads();
    uint32_t pre = s_w[w] + inc - sum;
    const uint32_t cq = max((n + HY_TARGET - 1) / HY_TARGET, 1u);
    uint32_t bm[4] = { 0, 0, 0, 0 };
    uint32_t runb = pre / cq, runc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t b = pre / cq;
        bm[i >> 2] |= b << (8 * (i & 3));
        if (b != runb) { if (runc) atomicAdd(&s_bsize[runb], runc); runb = b; runc = 0; }
        runc += f[i];
        pre += f[i];
    }
    if (runc) atomicAdd(&s_bsize[runb], runc);
    reinterpret_cast<uint4*>(plan->bmap)[tid] = make_uint4(bm[0], bm[1], bm[2], bm[3]);
    __syncthreads();
    if (tid < 256) {
        const uint32_t v = s_bsize[tid];
        uint32_t mx = v;
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o2));
        if (lane == 0) atomicMax(&s_max, mx);
        uint32_t inc2 = warp_incl_scan(v, lane);
        if (lane == 31) s_w[w] = inc2;
    }
    __syncthreads();
    if (tid < 256) {
        const uint32_t v = s_bsize[tid];
        uint32_t base = 0;
        for (int i = 0; i < w; i++) base += s_w[i];
        const uint32_t inc2 = warp_incl_scan(v, lane);
        plan->boff[tid] = base + inc2 - v;
        if (tid == 255) plan->boff[256] = base + inc2;
    }
    if (tid == 0) { plan->ok = s_max <= HY_CAP ? 1u : 0u; plan->nb = (n + cq - 1) / cq; plan->cq = cq; }
}

// the stable global partition: k_pass of VARIANT 0 with digit = bucket of the key
__global__ void __launch_bounds__(OS_THREADS) k_msd(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                   uint32_t n, uint32_t* __restrict__ tilectr, uint32_t* __restrict__ lookback, const HyPlan* __restrict__ plan) {
    extern __shared__ uint32_t dsm[];
    uint32_t* s_keys = dsm;                       // [SL_TILE]
    uint32_t* s_vals = s_keys + SL_TILE;          // [SL_TILE]
    uint32_t* s_whist = s_vals + SL_TILE;         // [8][256]
    uint8_t* s_bmap = reinterpret_cast<uint8_t*>(s_whist + 8 * 256);   // [HY_BINS]
    __shared__ uint32_t s_binstart[256];
    __shared__ uint32_t s_goff[256];
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_tile;
    if (!plan->ok) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tilectr, 1u);
    for (int i = tid; i < 8 * 256; i += OS_THREADS) s_whist[i] = 0;
    for (int i = tid; i < HY_BINS / 16; i += OS_THREADS) reinterpret_cast<uint4*>(s_bmap)[i] = reinterpret_cast<const uint4*>(plan->bmap)[i];
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * SL_TILE;
    const uint32_t valid = min((uint32_t)SL_TILE, n - tile_base);
    const uint32_t gbase = plan->boff[tid];
    uint32_t key[SL_ITEMS], rank[SL_ITEMS], dig[SL_ITEMS];
    const uint32_t i0 = tile_base + warp * (32 * SL_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) { uint32_t gi = i0 + i * 32; key[i] = gi < n ? __ldcs(keys_in + gi) : ~0u; }
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) {
        const uint32_t d = key[i] == ~0u ? 255u : s_bmap[key[i] >> HY_BIN_SHIFT];
        dig[i] = d;
        uint32_t peers = __match_any_sync(FULL, d);           // neighbouring triangles fall into the same bucket: few distinct values
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = s_whist[warp * 256 + d]; s_whist[warp * 256 + d] = old + __popc(peers); }
        old = __shfl_sync(FULL, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    uint32_t cnt = 0, tot;
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t c = s_whist[w * 256 + tid]; s_whist[w * 256 + tid] = cnt; cnt += c; }
    uint32_t binstart = block_excl_scan_256(cnt, s_w, &tot);
    s_binstart[tid] = binstart;
    uint32_t cnt_real = cnt - ((tid == 255) ? ((uint32_t)SL_TILE - valid) : 0u);
    uint32_t excl = 0;
    uint32_t* lb = lookback + (size_t)tile * 256 + tid;
    if (tile == 0) st_relaxed(lb, cnt_real | LB_FLAG_INCL);
    else {
        st_relaxed(lb, cnt_real | LB_FLAG_AGG);
        int t = (int)tile - 1;
        bool done = false;
        while (!done) {
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = (t - i >= 0) ? ld_relaxed(lb - 256 * (size_t)(tile - (uint32_t)(t - i))) : LB_FLAG_INCL;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (done) break;
                const uint32_t f = v[i] & ~LB_MASK;
                if (f == 0) break;
                excl += v[i] & LB_MASK; t--;
                if (f == LB_FLAG_INCL) done = true;
            }
        }
        st_relaxed(lb, (excl + cnt_real) | LB_FLAG_INCL);
    }
    s_goff[tid] = gbase + excl - binstart;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) {
        const uint32_t d = dig[i];
        const uint32_t pos = s_binstart[d] + s_whist[warp * 256 + d] + rank[i];
        s_keys[pos] = key[i];
        s_vals[pos] = i0 + i * 32;                       // thrust::sequence fused: value = input index
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SL_ITEMS; k++) {
        const uint32_t j = tid + k * OS_THREADS;
        if (j < valid) {
            const uint32_t kk = s_keys[j];
            const uint32_t dst = s_goff[s_bmap[kk >> HY_BIN_SHIFT]] + j;
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[j];
        }
    }
}
#define MSD_SMEM ((2 * SL_TILE + 8 * 256) * 4 + HY_BINS)

// every bucket sorted by its remaining bits in shared memory: <= 8192 keys, 512 threads x 16 keys in registers, LSD passes of 8 bits;
// the payload that moves with a key is its 16-bit position in the bucket (the values are gathered once, at the end); a pass whose
// digit is the same for all keys of the bucket is skipped
#define LS_THREADS 512
#define LS_ITEMS 16
#define LS_WARPS 16
__global__ void __launch_bounds__(LS_THREADS, 2) k_local_sort(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const HyPlan* __restrict__ plan) {
    extern __shared__ uint32_t dsm[];
    uint32_t* s_keys = dsm;                                            // [HY_CAP]
    uint32_t* s_whist = s_keys + HY_CAP;                               // [LS_WARPS][256]
    uint16_t* s_lp0 = reinterpret_cast<uint16_t*>(s_whist + LS_WARPS * 256);   // [HY_CAP] local positions, ping
    uint16_t* s_lp1 = s_lp0 + HY_CAP;                                  // pong
    __shared__ uint32_t s_binstart[256];
    __shared__ uint32_t s_tot[2][256];
    __shared__ uint32_t s_w[16];
    __shared__ uint32_t s_and, s_or;
    if (!plan->ok) return;
    const uint32_t b = blockIdx.x;
    if (b >= plan->nb) return;
    const uint32_t off = plan->boff[b], S = plan->boff[b + 1] - off;
    if (S == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    if (tid == 0) { s_and = ~0u; s_or = 0u; }
    __syncthreads();
    uint32_t key[LS_ITEMS], rank[LS_ITEMS];
    const int nitems = (int)((S + LS_THREADS - 1) / LS_THREADS);        // keys per thread this bucket needs (block-uniform): work follows the bucket size
    const uint32_t p0 = warp * (32 * nitems) + lane;
    uint32_t a = ~0u, o = 0u;
#pragma unroll
    for (int i = 0; i < LS_ITEMS; i++) {
        key[i] = ~0u;
        if (i < nitems) {
            const uint32_t p = p0 + i * 32;
            if (p < S) { key[i] = __ldcg(keys_in + off + p); a &= key[i]; o |= key[i]; }
            s_lp0[p] = (uint16_t)p;
        }
    }
    a = __reduce_and_sync(FULL, a); o = __reduce_or_sync(FULL, o);
    if (lane == 0) { atomicAnd(&s_and, a); atomicOr(&s_or, o); }
    __syncthreads();
    const uint32_t diff = s_and ^ s_or;
    uint16_t* lp_in = s_lp0; uint16_t* lp_out = s_lp1;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * 8;
        if (((diff >> shift) & 255u) == 0u) continue;            // block-uniform
        for (int i = tid; i < LS_WARPS * 256; i += LS_THREADS) s_whist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; i++) {
            if (i >= nitems) break;
            const uint32_t d = (key[i] >> shift) & 255u;
            uint32_t peers = FULL;
#pragma unroll
            for (int bb = 0; bb < 8; bb++) { const uint32_t bit = (d >> bb) & 1u; peers &= __ballot_sync(FULL, bit) ^ (bit - 1u); }
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader) { old = s_whist[warp * 256 + d]; s_whist[warp * 256 + d] = old + __popc(peers); }
            old = __shfl_sync(FULL, old, leader);
            rank[i] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        // digit d: exclusive scan over the 16 warps (2 threads per digit, 8 warps each), then over the digits
        {
            const int d = tid & 255, q = tid >> 8;
            uint32_t part = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = s_whist[(q * 8 + w) * 256 + d]; s_whist[(q * 8 + w) * 256 + d] = part; part += c; }
            s_tot[q][d] = part;
            __syncthreads();
            const uint32_t t0 = s_tot[0][d], total = t0 + s_tot[1][d];
            if (q == 1) {
#pragma unroll
                for (int w = 0; w < 8; w++) s_whist[(8 + w) * 256 + d] += t0;
            }
            const uint32_t v = q == 0 ? total : 0u;
            const uint32_t inc = warp_incl_scan(v, lane);
            if (lane == 31) s_w[warp] = inc;
            __syncthreads();
            if (q == 0) { uint32_t base = 0; for (int i = 0; i < warp; i++) base += s_w[i]; s_binstart[d] = base + inc - v; }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < LS_ITEMS; i++) {
            if (i >= nitems) break;
            const uint32_t d = (key[i] >> shift) & 255u;
            const uint32_t pos = s_binstart[d] + s_whist[warp * 256 + d] + rank[i];
            s_keys[pos] = key[i]; lp_out[pos] = lp_in[p0 + i * 32];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LS_ITEMS; i++) if (i < nitems) key[i] = s_keys[p0 + i * 32];
        uint16_t* t = lp_in; lp_in = lp_out; lp_out = t;
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < LS_ITEMS; i++) {
        const uint32_t p = p0 + i * 32;
        if (i < nitems && p < S) { keys_out[off + p] = key[i]; vals_out[off + p] = __ldcg(vals_in + off + lp_in[p]); }
    }
}
#define LS_SMEM ((HY_CAP + LS_WARPS * 256) * 4 + 2 * HY_CAP * 2)

static uint32_t* g_F; static HyPlan* g_plan; static uint32_t* g_ctr;
#define SL_STAMP_SETS 0
static uint32_t sl_stamp_rows(uint32_t tiles) { return tiles; }
static void sl_setup() {
    CK(cudaMalloc(&g_F, HY_BINS * 4)); CK(cudaMalloc(&g_plan, sizeof(HyPlan))); CK(cudaMalloc(&g_ctr, 4));
    CK(cudaFuncSetAttribute(k_msd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MSD_SMEM));
    CK(cudaFuncSetAttribute(k_local_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM));
}
// (the fine histogram belongs to the key generation kernel in the product; here it is part of "pass 0")
static void sl_launch(int pass, uint32_t* ki, uint32_t* vi, uint32_t* ko, uint32_t* vo, uint32_t n, uint32_t* hist, uint32_t* lb, unsigned long long* st, uint32_t tiles) {
    if (pass == 0) {
        cudaMemsetAsync(g_F, 0, HY_BINS * 4); cudaMemsetAsync(g_ctr, 0, 4);
        k_fine_hist<<<296, 256>>>(ki, n, g_F);
    } else if (pass == 1) {
        k_plan<<<1, 1024>>>(g_F, n, g_plan);
    } else if (pass == 2) {
        k_msd<<<tiles, OS_THREADS, MSD_SMEM>>>(ki, ko, vo, n, g_ctr, lb, g_plan);      // harness buffers: pass 2 reads buffer 0, writes buffer 1
    } else {
        k_local_sort<<<256, LS_THREADS, LS_SMEM>>>(ki, vi, ko, vo, g_plan);              // reads buffer 1, writes buffer 0
    }
}
