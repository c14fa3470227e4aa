// VARIANT 0: the shipped onesweep pass (8-bit digits, 4096-key tiles, per-thread serial look-back)
#ifndef SL_ITEMS
#define SL_ITEMS 16
#endif
#define SL_TILE (OS_THREADS * SL_ITEMS)
#ifndef SL_BALLOT
#define SL_BALLOT 1
#endif

template <bool FIRST>
__global__ void __launch_bounds__(OS_THREADS) k_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                     uint32_t n, int pass, uint32_t* __restrict__ hist,
                                                     uint32_t* __restrict__ lookback, unsigned long long* __restrict__ stamps) {
    __shared__ uint32_t s_whist[8][256];
    __shared__ uint32_t s_keys[SL_TILE];
    __shared__ uint32_t s_vals[SL_TILE];
    __shared__ uint32_t s_binstart[256];
    __shared__ uint32_t s_goff[256];
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = pass * 8;
#if TIMING
    const unsigned long long t_start = gtime();
#endif
    if (tid == 0) s_tile = atomicAdd(&hist[2048 + pass], 1u);
    for (int i = tid; i < 8 * 256; i += OS_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
#if TIMING
    if (tid == 0) stamps[(size_t)tile * NSTAMP] = t_start;
#endif
    STAMP(1);
    const uint32_t tile_base = tile * SL_TILE;
    const uint32_t valid = min((uint32_t)SL_TILE, n - tile_base);
    uint32_t tot;
    uint32_t gbase = block_excl_scan_256(hist[pass * 256 + tid], s_w, &tot);
    STAMP(2);
    uint32_t key[SL_ITEMS]; uint32_t rank[SL_ITEMS];
    const uint32_t i0 = tile_base + warp * (32 * SL_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) { uint32_t gi = i0 + i * 32; key[i] = gi < n ? __ldcs(keys_in + gi) : ~0u; }
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) {
        uint32_t d = (key[i] >> shift) & 255u;
#if SL_BALLOT
        uint32_t peers = FULL;
#pragma unroll
        for (int b = 0; b < 8; b++) { const uint32_t bit = (d >> b) & 1u; peers &= __ballot_sync(FULL, bit) ^ (bit - 1u); }
#else
        uint32_t peers = __match_any_sync(FULL, d);
#endif
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) { old = s_whist[warp][d]; s_whist[warp][d] = old + __popc(peers); }
        old = __shfl_sync(FULL, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    STAMP(3);
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t c = s_whist[w][tid]; s_whist[w][tid] = cnt; cnt += c; }
    uint32_t binstart = block_excl_scan_256(cnt, s_w, &tot);
    s_binstart[tid] = binstart;
    uint32_t cnt_real = cnt - ((tid == 255) ? ((uint32_t)SL_TILE - valid) : 0u);
    uint32_t excl = 0;
    uint32_t* lb = lookback + (size_t)tile * 256 + tid;
    if (tile == 0) {
        st_relaxed(lb, cnt_real | LB_FLAG_INCL);
    } else {
        st_relaxed(lb, cnt_real | LB_FLAG_AGG);
        STAMP(4);
        const uint32_t* p = lb - 256;
        for (;;) {
            uint32_t v = ld_relaxed(p);
            uint32_t f = v & ~LB_MASK;
            if (f == 0) continue;
            excl += v & LB_MASK;
            if (f == LB_FLAG_INCL) break;
            p -= 256;
        }
        st_relaxed(lb, (excl + cnt_real) | LB_FLAG_INCL);
    }
    s_goff[tid] = gbase + excl - binstart;
    __syncthreads();
    STAMP(5);
#pragma unroll
    for (int i = 0; i < SL_ITEMS; i++) {
        uint32_t d = (key[i] >> shift) & 255u;
        uint32_t pos = s_binstart[d] + s_whist[warp][d] + rank[i];
        s_keys[pos] = key[i];
        uint32_t gi = i0 + i * 32;
        uint32_t v;
        if (FIRST) v = gi; else v = gi < n ? __ldcs(vals_in + gi) : 0u;
        s_vals[pos] = v;
    }
    __syncthreads();
    STAMP(6);
#pragma unroll
    for (int k = 0; k < SL_ITEMS; k++) {
        uint32_t j = tid + k * OS_THREADS;
        if (j < valid) {
            uint32_t kk = s_keys[j];
            uint32_t dst = s_goff[(kk >> shift) & 255u] + j;
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[j];
        }
    }
    STAMP(7);
}
#define SL_STAMP_SETS 4
static uint32_t sl_stamp_rows(uint32_t tiles) { return tiles; }
static void sl_setup() {}
static void sl_launch(int pass, uint32_t* ki, uint32_t* vi, uint32_t* ko, uint32_t* vo, uint32_t n, uint32_t* hist, uint32_t* lb, unsigned long long* st, uint32_t tiles) {
    if (pass == 0) k_pass<true><<<tiles, OS_THREADS>>>(ki, nullptr, ko, vo, n, pass, hist, lb, st);
    else k_pass<false><<<tiles, OS_THREADS>>>(ki, vi, ko, vo, n, pass, hist, lb, st);
}
