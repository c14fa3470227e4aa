// VARIANT 1: all four passes in ONE cooperative persistent kernel.  CTA c owns the contiguous chunk of tiles [c*m, (c+1)*m);
// per pass: per-CTA digit counts -> grid barrier -> column prefix over the CTAs before it (parallel loads, no look-back chain)
// -> rank + scatter -> grid barrier.  With m == 1 (n <= G * tile) the keys stay in registers between the two halves of a pass.
#ifndef CS_THREADS
#define CS_THREADS 512
#endif
#ifndef CS_ITEMS
#define CS_ITEMS 14
#endif
#ifndef CS_BPSM
#define CS_BPSM 1
#endif
#define SL_TILE (CS_THREADS * CS_ITEMS)
#define CS_WARPS (CS_THREADS / 32)
struct GridBar { uint32_t count, gen; };
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void grid_barrier(GridBar* bar, uint32_t G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t gen = ld_relaxed(&bar->gen);
        uint32_t old;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(&bar->count) : "memory");
        if (old == G - 1) { st_relaxed(&bar->count, 0u); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(&bar->gen), "r"(gen + 1u) : "memory"); }
        else { while (ld_acquire(&bar->gen) == gen) { } }
    }
    __syncthreads();
}
// exclusive scan of one value per digit (threads 0..255 hold digits, the others pass 0); s_w: CS_WARPS words
__device__ __forceinline__ uint32_t digit_excl_scan(uint32_t v, uint32_t* s_w) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { uint32_t x = s_w[i]; if (i < w) base += x; }
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(CS_THREADS, CS_BPSM) k_sort_coop(uint32_t* __restrict__ keys0, uint32_t* __restrict__ vals0,
                                                             uint32_t* __restrict__ keys1, uint32_t* __restrict__ vals1, uint32_t n,
                                                             const uint32_t* __restrict__ hist, uint32_t* __restrict__ H, GridBar* bar,
                                                             uint32_t m, unsigned long long* __restrict__ stamps) {
    extern __shared__ uint32_t smem[];
    uint32_t* s_whist = smem;                          // [CS_WARPS][256]
    uint32_t* s_keys = s_whist + CS_WARPS * 256;       // [SL_TILE]
    uint32_t* s_vals = s_keys + SL_TILE;               // [SL_TILE]
    uint32_t* s_binstart = s_vals + SL_TILE;           // [256]
    uint32_t* s_goff = s_binstart + 256;               // [256]
    uint32_t* s_part = s_goff + 256;                   // [256] digit counts of the tile held in registers (m == 1)
    uint32_t* s_col = s_part + 256;                    // [CS_THREADS]
    uint32_t* s_w = s_col + CS_THREADS;                // [CS_WARPS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t G = gridDim.x, c = blockIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t tile = c;   // stamps index
    for (int pass = 0; pass < 4; pass++) {
        const int shift = pass * 8;
        const uint32_t* kin = (pass & 1) ? keys1 : keys0; const uint32_t* vin = (pass & 1) ? vals1 : vals0;
        uint32_t* kout = (pass & 1) ? keys0 : keys1; uint32_t* vout = (pass & 1) ? vals0 : vals1;
        uint32_t* Hp = H + (size_t)pass * G * 256;
        if (pass == 0) STAMP(0);
        uint32_t key[CS_ITEMS], rank[CS_ITEMS];
        uint32_t base = 0;                               // threads < 256: global position of the next key with digit tid
        // ---- first half: digit counts of this CTA's chunk
        if (m == 1) {
            const uint32_t tile_base = c * SL_TILE;
            for (int i = tid; i < CS_WARPS * 256; i += CS_THREADS) s_whist[i] = 0;
            const uint32_t i0 = tile_base + warp * (32 * CS_ITEMS) + lane;
#pragma unroll
            for (int i = 0; i < CS_ITEMS; i++) { uint32_t gi = i0 + i * 32; key[i] = gi < n ? __ldcg(kin + gi) : ~0u; }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CS_ITEMS; i++) {
                uint32_t d = (key[i] >> shift) & 255u;
                uint32_t peers = FULL;
#pragma unroll
                for (int b = 0; b < 8; b++) { const uint32_t bit = (d >> b) & 1u; peers &= __ballot_sync(FULL, bit) ^ (bit - 1u); }
                int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (lane == leader) { old = s_whist[warp * 256 + d]; s_whist[warp * 256 + d] = old + __popc(peers); }
                old = __shfl_sync(FULL, old, leader);
                rank[i] = old + __popc(peers & lt);
                __syncwarp();
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t cnt = 0;
#pragma unroll
                for (int w = 0; w < CS_WARPS; w++) { uint32_t x = s_whist[w * 256 + tid]; s_whist[w * 256 + tid] = cnt; cnt += x; }
                const uint32_t valid = tile_base < n ? min((uint32_t)SL_TILE, n - tile_base) : 0u;
                s_part[tid] = cnt;                       // incl. padding
                Hp[(size_t)c * 256 + tid] = cnt - ((tid == 255) ? ((uint32_t)SL_TILE - valid) : 0u);
            }
        } else {
            for (int i = tid; i < 256; i += CS_THREADS) s_binstart[i] = 0;
            __syncthreads();
            for (uint32_t t = 0; t < m; t++) {
                const uint32_t tile_base = (c * m + t) * SL_TILE;
                if (tile_base >= n) break;
#pragma unroll
                for (int i = 0; i < CS_ITEMS; i++) {
                    uint32_t gi = tile_base + i * CS_THREADS + tid;
                    if (gi < n) atomicAdd(&s_binstart[(__ldcg(kin + gi) >> shift) & 255u], 1u);
                }
            }
            __syncthreads();
            if (tid < 256) Hp[(size_t)c * 256 + tid] = s_binstart[tid];
        }
        if (pass == 0) STAMP(1);
        grid_barrier(bar, G);
        if (pass == 0) STAMP(2);
        // ---- column prefix: keys with digit tid in the CTAs before this one + all keys with smaller digits
        {
            constexpr int P = CS_THREADS / 256;
            const int part = tid >> 8, d = tid & 255;
            uint32_t sum = 0;
            uint32_t cc = part;
            for (; cc + 3 * P < c; cc += 4 * P) {
                const uint32_t a0 = __ldcg(Hp + (size_t)cc * 256 + d), a1 = __ldcg(Hp + (size_t)(cc + P) * 256 + d);
                const uint32_t a2 = __ldcg(Hp + (size_t)(cc + 2 * P) * 256 + d), a3 = __ldcg(Hp + (size_t)(cc + 3 * P) * 256 + d);
                sum += a0 + a1 + a2 + a3;
            }
            for (; cc < c; cc += P) sum += __ldcg(Hp + (size_t)cc * 256 + d);
            s_col[tid] = sum;
            __syncthreads();
            uint32_t colsum = 0;
            if (tid < 256) { for (int p = 0; p < P; p++) colsum += s_col[p * 256 + tid]; }
            const uint32_t gb = digit_excl_scan(tid < 256 ? hist[pass * 256 + tid] : 0u, s_w);
            base = gb + colsum;
        }
        if (pass == 0) STAMP(3);
        // ---- second half: rank, reorder through shared memory, coalesced scatter
        for (uint32_t t = 0; t < m; t++) {
            const uint32_t tile_base = (c * m + t) * SL_TILE;
            if (tile_base >= n) break;
            const uint32_t valid = min((uint32_t)SL_TILE, n - tile_base);
            const uint32_t i0 = tile_base + warp * (32 * CS_ITEMS) + lane;
            uint32_t cnt = 0;
            if (m > 1) {
                for (int i = tid; i < CS_WARPS * 256; i += CS_THREADS) s_whist[i] = 0;
#pragma unroll
                for (int i = 0; i < CS_ITEMS; i++) { uint32_t gi = i0 + i * 32; key[i] = gi < n ? __ldcg(kin + gi) : ~0u; }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < CS_ITEMS; i++) {
                    uint32_t d = (key[i] >> shift) & 255u;
                    uint32_t peers = FULL;
#pragma unroll
                    for (int b = 0; b < 8; b++) { const uint32_t bit = (d >> b) & 1u; peers &= __ballot_sync(FULL, bit) ^ (bit - 1u); }
                    int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if (lane == leader) { old = s_whist[warp * 256 + d]; s_whist[warp * 256 + d] = old + __popc(peers); }
                    old = __shfl_sync(FULL, old, leader);
                    rank[i] = old + __popc(peers & lt);
                    __syncwarp();
                }
                __syncthreads();
                if (tid < 256) {
#pragma unroll
                    for (int w = 0; w < CS_WARPS; w++) { uint32_t x = s_whist[w * 256 + tid]; s_whist[w * 256 + tid] = cnt; cnt += x; }
                }
            } else if (tid < 256) cnt = s_part[tid];
            const uint32_t binstart = digit_excl_scan(cnt, s_w);
            if (tid < 256) {
                s_binstart[tid] = binstart;
                s_goff[tid] = base - binstart;
                base += cnt - ((tid == 255) ? ((uint32_t)SL_TILE - valid) : 0u);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CS_ITEMS; i++) {
                uint32_t d = (key[i] >> shift) & 255u;
                uint32_t pos = s_binstart[d] + s_whist[warp * 256 + d] + rank[i];
                s_keys[pos] = key[i];
                uint32_t gi = i0 + i * 32;
                s_vals[pos] = (pass == 0) ? gi : (gi < n ? __ldcg(vin + gi) : 0u);
            }
            __syncthreads();
            if (pass == 0) STAMP(4);
#pragma unroll
            for (int k = 0; k < CS_ITEMS; k++) {
                uint32_t j = tid + k * CS_THREADS;
                if (j < valid) {
                    uint32_t kk = s_keys[j];
                    uint32_t dst = s_goff[(kk >> shift) & 255u] + j;
                    kout[dst] = kk;
                    vout[dst] = s_vals[j];
                }
            }
            if (m > 1) __syncthreads();
        }
        if (pass == 0) STAMP(5);
        if (pass < 3) grid_barrier(bar, G);
        if (pass == 0) STAMP(6);
    }
    STAMP(7);
}
static GridBar* g_bar; static uint32_t* g_H; static int g_G; static size_t g_smem;
#define SL_STAMP_SETS 1
static uint32_t sl_stamp_rows(uint32_t) { return (uint32_t)g_G; }
static void sl_setup() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    g_G = sms * CS_BPSM;
    CK(cudaMalloc(&g_bar, sizeof(GridBar))); CK(cudaMemset(g_bar, 0, sizeof(GridBar)));
    CK(cudaMalloc(&g_H, (size_t)4 * g_G * 256 * 4));
    g_smem = (size_t)(CS_WARPS * 256 + 2 * SL_TILE + 768 + CS_THREADS + CS_WARPS) * 4;
    CK(cudaFuncSetAttribute(k_sort_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sort_coop, CS_THREADS, g_smem);
    printf("coop: G = %d, %d threads x %d items, smem %zu B, occupancy %d blocks/SM\n", g_G, CS_THREADS, CS_ITEMS, g_smem, occ);
}
static void sl_launch(int pass, uint32_t* ki, uint32_t* vi, uint32_t* ko, uint32_t* vo, uint32_t n, uint32_t* hist, uint32_t* lb, unsigned long long* st, uint32_t tiles) {
    if (pass != 0) return;
    uint32_t m = (tiles + g_G - 1) / g_G;
    void* args[] = { &ki, &vi, &ko, &vo, &n, &hist, &g_H, &g_bar, &m, &st };
    CK(cudaLaunchCooperativeKernel((void*)k_sort_coop, dim3(g_G), dim3(CS_THREADS), args, g_smem, 0));
}
