/*
 * bihrt.h -- C ABI of libbihrt.so: B200-native (sm_100a) BIH build + ray traversal + ray/triangle
 * intersection, a drop-in for the hot path of rehakvoj1/BIH-GPU-Raytracer.
 *
 * The reference has no FFI/plugin interface: the path sits behind three C++ classes with
 * thrust-typed members (App, Renderer, GPUArrayManager).  The boundary kept here is the set of
 * entry points and their data contracts; each export cites what it replaces.
 * R/ = BIH_Raytracer/BIH_Raytracer/ of the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns BIHRT_OK (0) or a negative code and
 *     never exits the process (the reference prints and calls exit(99), R/src/Renderer.cpp:63-73);
 *     bihrt_last_error() gives the message of the last failure on that context.
 *   - pointer arguments marked "host or device" are classified with cudaPointerGetAttributes.
 *   - one context = one GPU, one stream; calls are asynchronous on that stream unless they return
 *     data to host memory.  One context per host thread.
 *   - there is NO CPU fallback: every compute entry point fails if no sm_100-class GPU is present.
 */
#ifndef BIHRT_H
#define BIHRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BIHRT_API __attribute__((visibility("default")))
#else
#define BIHRT_API
#endif

#define BIHRT_VERSION 100

#define BIHRT_OK             0
#define BIHRT_ERR_INVALID   -1   /* bad argument */
#define BIHRT_ERR_CUDA      -2   /* CUDA runtime error (message has the CUDA string) */
#define BIHRT_ERR_NOMEM     -3   /* device or host allocation failed (reference: Allocate* -> false) */
#define BIHRT_ERR_STATE     -4   /* call order: no scene loaded / BIH not built */
#define BIHRT_ERR_IO        -5   /* OBJ file unreadable or malformed */
#define BIHRT_ERR_INTERNAL  -6   /* device-side watchdog tripped (never expected) */

typedef struct bihrt_ctx bihrt_ctx;

/* Run-time replacement of the compile-time macros in R/src/Constants.h:4-8. */
typedef struct bihrt_config {
    int32_t  device;        /* CUDA device ordinal */
    uint32_t flags;         /* reserved, 0 */
    int32_t  reserved[6];
} bihrt_config;

/* Camera, same four vectors as R/src/Camera.h:14-17; ray(u,v) = (origin,
 * lower_left + u*horizontal + v*vertical - origin), NOT normalised (R/src/Camera.cu:18-20). */
typedef struct bihrt_camera {
    float origin[3];
    float lower_left[3];
    float horizontal[3];
    float vertical[3];
} bihrt_camera;

/* Ray = origin + direction (R/src/Ray.h:19-20); invDir/sign are derived inside (R/src/Ray.cu:3-10). */
typedef struct bihrt_ray {
    float o[3];
    float d[3];
} bihrt_ray;

/* bihrt_render flags */
#define BIHRT_RENDER_JITTER   1u   /* jittered samples (reference behaviour); otherwise pixel centres */
#define BIHRT_RENDER_COUNTS   2u   /* framebuffer receives per-pixel hit counts instead of packed colours */

/* Reference view of a built BIH: arrays with the exact field meaning of the reference's device
 * arrays.  Caller allocates; any pointer may be NULL to skip that array.
 * Capacities: [n] for the per-triangle arrays, [nu] / [nu-1] for the per-cell / per-node ones
 * (allocating n entries for each is always enough). */
typedef struct bihrt_refview {
    int64_t   n;                    /* out: triangles */
    int64_t   nu;                   /* out: unique Morton cells = leaves (GetUniqueMCSize) */
    float     scene_lo[3];          /* out: AABBs::sceneBBoxLo, R/src/AABB.h:15 */
    float     scene_hi[3];          /* out: AABBs::sceneBBoxHi */
    uint32_t* morton_codes;         /* [n]    m_mortonCodes after the sort */
    uint32_t* tris_indexes;         /* [n]    m_trisIndexes: sorted slot -> input triangle */
    uint32_t* unique_morton_codes;  /* [nu]   m_uniqueMortonCodes */
    uint32_t* duplicates_cnts;      /* [nu]   m_duplicatesCnts */
    int32_t*  first_idxs;           /* [nu]   m_firstIdxs */
    float*    clip_planes;          /* [2*(nu-1)] TreeInternalNode::t_clipPlanes, R/src/Tree.cuh:17 */
    int32_t*  axis;                 /* [nu-1] t_axis */
    uint8_t*  is_leaf;              /* [2*(nu-1)] isLeaf */
    int32_t*  children;             /* [2*(nu-1)] children */
    int32_t*  parent;               /* [nu-1] parent (root: -1) */
    int32_t*  leaf_parents;         /* [nu]   m_leafParents */
} bihrt_refview;

typedef struct bihrt_build_info {
    int64_t n;                /* triangles */
    int64_t nu;               /* leaves */
    int64_t node_bytes;       /* compact nodes, 16 B each */
    int64_t tri_bytes;        /* leaf-ordered triangles, 48 B each */
    float   last_build_ms;    /* device time of the last bihrt_build (CUDA events) */
    int32_t sort_passes;
    int32_t reserved[6];
} bihrt_build_info;

/* ---- lifetime ------------------------------------------------------------------------------ */
BIHRT_API int         bihrt_version(void);
BIHRT_API int         bihrt_create(bihrt_ctx** out, const bihrt_config* cfg);   /* replaces Renderer::Init + GPUArrayManager ctor, R/src/Renderer.cpp:87-105 */
BIHRT_API void        bihrt_destroy(bihrt_ctx* ctx);
BIHRT_API const char* bihrt_last_error(const bihrt_ctx* ctx);                   /* replaces checkCudaErrors' stderr print */
/* Stream of the context.  A new context runs on its OWN private non-blocking stream.  bihrt_set_stream takes a
 * cudaStream_t with CUDA's own meaning -- NULL (0) is the legacy default stream, exactly what
 * torch.cuda.current_stream().cuda_stream is while torch runs on its default stream -- or BIHRT_STREAM_OWN to go back
 * to the private stream.  Work of a caller that runs on ANOTHER stream is not ordered with the library's: either hand
 * that stream to bihrt_set_stream, or order the two with events on the handle bihrt_get_stream returns (the Python
 * mirror does this for torch tensors). */
#define BIHRT_STREAM_OWN ((void*)(intptr_t)-1)
BIHRT_API int         bihrt_set_stream(bihrt_ctx* ctx, void* cuda_stream);
BIHRT_API int         bihrt_get_stream(bihrt_ctx* ctx, void** cuda_stream);     /* the cudaStream_t the context launches on */
BIHRT_API int         bihrt_sync(bihrt_ctx* ctx);                               /* replaces the cudaDeviceSynchronize after each step, R/src/Renderer.cpp:428-503 */
/* Options (DESIGN.md has the full list of tuning knobs).  The two that change WHAT is built:
 *   "morton_bits" 30 (default): the reference's 10-bit grid per axis, R/src/Renderer.cpp:116-136 -- the parity path, bit-exact tree.
 *                 63: QUALITY MODE, not a parity path: 21 bits per axis, ties broken by sorted position,
 *                 every subtree of at most "leaf_cap" (default 4, 1..64) triangles collapsed into one leaf.  Hits equal brute force
 *                 (including on axis-aligned geometry where the reference's strict comparisons lose hits); bihrt_refit and
 *                 bihrt_export_reference_view are not available.  A context that adopts a replicated BIH (bihrt_bih_adopt) must
 *                 have the builder's value; bihrt_bih_copy and bihrt_bih_import carry it over themselves. */
BIHRT_API int         bihrt_set_option(bihrt_ctx* ctx, const char* name, int64_t value);
BIHRT_API int         bihrt_get_stat(bihrt_ctx* ctx, const char* name, int64_t* value);   /* "kernel_launches": kernels of this library launched so far */

/* ---- scene load: App::LoadModels + GPUArrayManager::Allocate*, R/src/App.cpp:65-167,
 *      R/src/GPUArrayManager.cpp:7-91.  xyz9 = n x (v0,v1,v2) floats in model->mesh->face order,
 *      host or device; the library copies. ------------------------------------------------------ */
BIHRT_API int bihrt_scene_load_triangles(bihrt_ctx* ctx, const float* xyz9, int64_t n);
BIHRT_API int bihrt_scene_update_vertices(bihrt_ctx* ctx, const float* xyz9, int64_t n);  /* animation: same n, BIH becomes stale */
BIHRT_API int bihrt_scene_load_obj(bihrt_ctx* ctx, const char* path);          /* minimal Wavefront reader replacing Model(path), R/src/Model.cpp:10-29 */

/* ---- build: first half of Renderer::Render (Morton transform, sort, RLE, Launch_BuildTree,
 *      Launch_FindClipPlanes), R/src/Renderer.cpp:422-503.  No host sync inside. ------------------ */
BIHRT_API int bihrt_build(bihrt_ctx* ctx);
/* NOT a parity path: after bihrt_scene_update_vertices keep the order, leaves and topology of the
 * last full build and recompute scene box, triangle records and clip planes only (about half the time of a build).
 * The BIH stays valid (planes bound their subtrees) but is not the tree the reference would build for the moved vertices. */
BIHRT_API int bihrt_refit(bihrt_ctx* ctx);
BIHRT_API int bihrt_get_build_info(bihrt_ctx* ctx, bihrt_build_info* out);     /* synchronises */
BIHRT_API int bihrt_export_reference_view(bihrt_ctx* ctx, bihrt_refview* view); /* synchronises; see bihrt_refview */

/* ---- trace: Launch_cudaRender -> cudaRender -> TraverseTree, R/src/CUDAKernels.cu:227-447.
 *      Per ray: t in units of |d| (FLT_MAX on miss), slot = index into the Morton-sorted triangle
 *      order (the reference's HitRecord::triangleIdx, R/src/CUDAKernels.cu:215-221; -1 on miss),
 *      prim = input triangle index = tris_indexes[slot].  rays / outputs: host or device; any
 *      output may be NULL. ---------------------------------------------------------------------- */
BIHRT_API int bihrt_trace(bihrt_ctx* ctx, const bihrt_ray* rays, int64_t n, float* t, int32_t* slot, int32_t* prim);
/* Occlusion query (shadow rays): blocker[i] = slot of SOME triangle ray i hits with 0 < t < tmax, or -1 (host or
 * device).  The traversal is cut at tmax and ends at the first such hit, so blocker[i] >= 0 exactly when the closest
 * hit bihrt_trace reports has t < tmax; which blocker is found is unspecified.  With the shadow rays of
 * bihrt_secondary_rays (direction = light - origin) tmax = 1 asks "is the light hidden". */
BIHRT_API int bihrt_trace_any(bihrt_ctx* ctx, const bihrt_ray* rays, int64_t n, float tmax, int32_t* blocker);
/* instrumented trace: counters[0]=internal nodes visited, [1]=triangle tests, [2]=max stack depth,
 * [3]=rays; same results as bihrt_trace (outputs may be NULL). */
BIHRT_API int bihrt_trace_counted(bihrt_ctx* ctx, const bihrt_ray* rays, int64_t n, float* t, int32_t* slot,
                        int32_t* prim, uint64_t counters[4]);

/* Render w x h pixels, spp samples each, into the context's framebuffer with the reference's
 * colouring: hit (255,255,0), miss (20,20,40), mean over samples, packed r | g<<8 | b<<16, row 0 =
 * bottom (R/src/CUDAKernels.cu:82-88,385-387,391-423).  Jitter comes from a counter-based hash of
 * (seed, pixel, sample) instead of per-pixel XORWOW state (documented difference). */
BIHRT_API int bihrt_render(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                 uint64_t seed, uint32_t flags);
/* Same as bihrt_render through the instrumented kernel: counters as in bihrt_trace_counted. */
BIHRT_API int bihrt_render_counted(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                         uint64_t seed, uint32_t flags, uint64_t counters[4]);
/* Multi-GPU: render only the 32x32-pixel tiles with tile_id % shard_count == shard_index; every
 * other pixel of the framebuffer is set to 0, so a sum (or bitwise OR) of the shards' framebuffers
 * is the full image. */
BIHRT_API int bihrt_render_shard(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                       uint64_t seed, uint32_t flags, int32_t shard_index, int32_t shard_count);
/* Multi-GPU, sample sharding: trace only samples [sample_begin, sample_end) of every pixel of a
 * spp-sample frame and store per-pixel HIT COUNTS in the framebuffer.  Summing the counts of all ranks
 * and calling bihrt_framebuffer_resolve(spp) gives exactly the image bihrt_render(spp) produces; every
 * rank walks the whole image, so warp coherence is that of the single-GPU render. */
BIHRT_API int bihrt_render_samples(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                         uint64_t seed, uint32_t flags, int32_t sample_begin, int32_t sample_end);
BIHRT_API int bihrt_framebuffer_resolve(bihrt_ctx* ctx, int32_t spp);   /* counts -> packed colours, in place */
/* Multi-GPU, unit interleave (preferred): every rank walks every 32x32 tile but traces only every count-th
 * 32-ray unit of it (unit = a few neighbouring pixels with ALL their samples), writing per-pixel hit counts and 0
 * for the pixels it does not own.  Balanced by construction, same tile locality as the single-GPU render, and a
 * pixel's samples stay together in consecutive lanes.  count must divide 32 << k, k = log2 of the lane groups. */
BIHRT_API int bihrt_render_interleaved(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                             uint64_t seed, uint32_t flags, int32_t index, int32_t count);
/* Multi-GPU, unit interleave fused with the framebuffer gather: same partition as bihrt_render_interleaved, but the
 * final packed colour of every pixel this rank owns is stored by the trace kernel itself straight into `target_fb`
 * (w*h uint32; NULL = this context's own framebuffer) and no other pixel is touched.  With target_fb = the
 * framebuffer of the gathering GPU (a device pointer of another context of this process, or one opened with
 * bihrt_framebuffer_ipc_open in a one-process-per-GPU job) the stores travel over NVLink while the kernel runs:
 * no per-rank framebuffer clear, no hit-count reduce, no resolve pass.  The frame is complete once every rank's
 * launch has finished (a barrier, not a data collective).  Requires the lane-group layout every rank uses for
 * the same spp, so a pixel's samples never leave one rank. */
BIHRT_API int bihrt_render_interleaved_to(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                                uint64_t seed, uint32_t flags, int32_t index, int32_t count, uint32_t* target_fb);
/* Sharing the gathering GPU's framebuffer with the other processes of the job (CUDA IPC): export allocates the
 * context's w x h framebuffer and writes a 64-byte handle; open maps it in another process (peer access over
 * NVLink) and returns the pointer to pass as target_fb; close unmaps it. */
BIHRT_API int bihrt_framebuffer_ipc_export(bihrt_ctx* ctx, int32_t w, int32_t h, void* handle64);
BIHRT_API int bihrt_framebuffer_ipc_open(bihrt_ctx* ctx, const void* handle64, uint32_t** peer_fb);
BIHRT_API int bihrt_framebuffer_ipc_close(bihrt_ctx* ctx, uint32_t* peer_fb);
/* While a handle is out the exporting context refuses (BIHRT_ERR_STATE) to reallocate its framebuffer for a larger
 * frame -- the peers would keep storing into freed memory.  Call this once every peer has closed its mapping.  (The
 * same care is the caller's for a raw bihrt_framebuffer() pointer handed to another context as target_fb.) */
BIHRT_API int bihrt_framebuffer_ipc_unexport(bihrt_ctx* ctx);
/* Per-sample hit buffers of the same rays bihrt_render traces (index = (j*w+i)*spp + s); device or
 * host outputs, any may be NULL.  Used by the parity tests. */
BIHRT_API int bihrt_render_hits(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                      uint64_t seed, uint32_t flags, float* t, int32_t* slot, int32_t* prim);

/* ---- secondary rays (the stage after the path; the reference's Color() is a stub, R/src/CUDAKernels.cu:370-389).
 *      From the primary hits of a camera frame (the w*h*spp samples bihrt_render traces) build the next ray
 *      batch ON THE DEVICE: every sample that hits emits one ray from the hit point (moved 1e-3 along the
 *      geometric normal) to `light` (SHADOW; direction = light - origin, so t = 1 at the light) or in a
 *      cosine-weighted direction about the normal (DIFFUSE; counter-based hash of seed, pixel, sample).  Rays
 *      are written compacted and in sample order to out_rays (DEVICE, capacity w*h*spp); out_sample[i] (device,
 *      may be NULL) = source sample.  *count (host) receives the number of rays.  Synchronises. */
#define BIHRT_SECONDARY_SHADOW  0
#define BIHRT_SECONDARY_DIFFUSE 1
BIHRT_API int bihrt_secondary_rays(bihrt_ctx* ctx, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp, uint64_t seed,
                         uint32_t flags, int32_t kind, const float light[3], bihrt_ray* out_rays, int32_t* out_sample,
                         int64_t* count);

/* ---- framebuffer: Renderer::m_cudaDestResource, R/src/Renderer.h:46, R/src/Renderer.cpp:762-768 */
BIHRT_API int bihrt_framebuffer(bihrt_ctx* ctx, uint32_t** dev_ptr, int32_t* w, int32_t* h);  /* device pointer, owned by ctx */
BIHRT_API int bihrt_framebuffer_read(bihrt_ctx* ctx, uint32_t* host_dst);                      /* synchronises */

/* ---- BIH replication across GPUs (one broadcast of this blob per build) ------- */
BIHRT_API int bihrt_bih_blob_bytes(bihrt_ctx* ctx, uint64_t* bytes);                     /* synchronises (needs Nu) */
BIHRT_API int bihrt_bih_export(bihrt_ctx* ctx, void* dev_dst, uint64_t bytes);           /* device buffer, D2D on the ctx stream */
BIHRT_API int bihrt_bih_import(bihrt_ctx* ctx, const void* dev_src, uint64_t bytes);     /* adopt a blob built on another GPU */

/* In place: the blob of an n-triangle scene is one contiguous device region whose size follows from n alone, so it can be
 * the buffer of the broadcast on every rank (no export / import copies, no host read of Nu).  bihrt_bih_region lays the
 * context's blob out for n triangles and returns the region (a no-op on the rank that built an n-triangle scene);
 * receivers call bihrt_bih_adopt(n) once the broadcast has been enqueued on the context's stream. */
BIHRT_API int bihrt_bih_region(bihrt_ctx* ctx, int64_t n, void** dev_ptr, uint64_t* bytes);
BIHRT_API int bihrt_bih_adopt(bihrt_ctx* ctx, int64_t n);
/* The same inside ONE process (a C host driving several GPUs, no NCCL): peer copy of src's BIH into dst over NVLink,
 * ordered after the work enqueued on src's stream and before dst's next call. */
BIHRT_API int bihrt_bih_copy(bihrt_ctx* dst, bihrt_ctx* src);

/* ---- several GPUs in ONE process ("Multi-GPU adds bihrt_create_multi(ctx**, int ngpu) which internally holds
 *      one NCCL communicator").  The reference is single-GPU; rays are independent given a replicated tree (8(e)).
 *      bihrt_create_multi fills ctxs[0..ngpu) with one context per device 0..ngpu-1 that share an NCCL communicator set
 *      (ncclCommInitAll; NCCL is bound at run time from libnccl.so.2).  Context 0 is the builder and the gatherer: load the
 *      scene and bihrt_build on it as usual, then per frame
 *          bihrt_multi_broadcast(ctxs[0]);                    one ncclBroadcast of the BIH blob, in place, no host copy
 *          bihrt_multi_render(ctxs[0], cam, w, h, spp, ...);   unit interleave; every device's trace kernel stores its finished
 *                                                             pixels straight into context 0's framebuffer over NVLink
 *          bihrt_framebuffer_read(ctxs[0], host);             ordered after every device's launch (events), same image as
 *                                                             bihrt_render on one GPU, bit for bit
 *      Everything is asynchronous on the contexts' streams; bihrt_multi_sync waits for all of them.  Contexts of a group are
 *      released together with bihrt_destroy_multi (not bihrt_destroy). ----------------------------------------------------- */
BIHRT_API int  bihrt_create_multi(bihrt_ctx** ctxs, int32_t ngpu);
BIHRT_API void bihrt_destroy_multi(bihrt_ctx** ctxs, int32_t ngpu);
BIHRT_API int  bihrt_multi_size(const bihrt_ctx* ctx);                       /* contexts in ctx's group (1 for a plain context) */
BIHRT_API int  bihrt_multi_nccl_version(void);                               /* e.g. 22809; 0 if libnccl.so.2 cannot be bound */
BIHRT_API int  bihrt_multi_broadcast(bihrt_ctx* ctx0);
BIHRT_API int  bihrt_multi_render(bihrt_ctx* ctx0, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp,
                                  uint64_t seed, uint32_t flags);
BIHRT_API int  bihrt_multi_sync(bihrt_ctx* ctx0);

#ifdef __cplusplus
}
#endif
#endif /* BIHRT_H */
