// One process driving several GPUs through the C ABI: bihrt_create_multi makes N contexts
// (one per device) that share ONE NCCL communicator set (ncclCommInitAll), held inside the library.  The reference is
// single-GPU (R/src/Renderer.cpp: one device, one default stream); this is the path's data-parallel extension:
//   bihrt_multi_broadcast   the BIH built on context 0 is replicated into the others by ONE ncclBroadcast of the blob,
//                           in place (blob to blob), on the contexts' own streams -- no host copy, no host read of Nu;
//   bihrt_multi_render      unit interleave: every context traces every N-th run of 32-ray units of every tile and its
//                           trace kernel stores the finished pixels straight into context 0's framebuffer (peer stores over
//                           NVLink; the gather is fused into the kernel), then context 0's stream waits for the others' events;
//   bihrt_multi_sync        host-side frame barrier.
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library neither needs NCCL to load nor brings a second
// copy into a process that already has one (a torch process: the loader hands back the copy torch loaded).
#include "bihrt_internal.cuh"
#include <dlfcn.h>
#include <cstring>
#include <vector>

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
};
const int NCCL_UINT8 = 1;       // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    for (const char* name : { "libnccl.so.2", "libnccl.so" }) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
    api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.lib, "ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
    api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast && api.GetErrorString;
    return api;
}

}  // namespace

// the group shared by the contexts of one bihrt_create_multi call
struct bihrt_group {
    std::vector<bihrt_ctx*> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<cudaEvent_t> done;      // per context: "my launch of this frame has been enqueued up to here"
    cudaEvent_t start = nullptr;        // on context 0's stream: the others may touch its framebuffer / blob after this
    bool peer_ok = false;
};

#define NCCL_CHECK(c, call) do { int r_ = (call); if (r_ != 0) \
    return bihrt_fail((c), BIHRT_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r_)); } while (0)

extern "C" {

int bihrt_create_multi(bihrt_ctx** ctxs, int32_t ngpu) {
    if (!ctxs || ngpu < 1 || ngpu > 64) return BIHRT_ERR_INVALID;
    for (int i = 0; i < ngpu; i++) ctxs[i] = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < ngpu) { cudaGetLastError(); return BIHRT_ERR_CUDA; }
    bihrt_group* g = new (std::nothrow) bihrt_group();
    if (!g) return BIHRT_ERR_NOMEM;
    int rc = BIHRT_OK;
    for (int i = 0; i < ngpu && rc == BIHRT_OK; i++) {
        bihrt_config cfg; memset(&cfg, 0, sizeof cfg); cfg.device = i;
        bihrt_ctx* c = nullptr;
        rc = bihrt_create(&c, &cfg);
        if (rc == BIHRT_OK) { c->group = g; c->group_rank = i; g->ctx.push_back(c); ctxs[i] = c; }
    }
    auto fail = [&](int code) {
        for (size_t i = 0; i < g->ctx.size(); i++) { g->ctx[i]->group = nullptr; bihrt_destroy(g->ctx[i]); ctxs[i] = nullptr; }
        for (auto cm : g->comm) if (cm) nccl().CommDestroy(cm);
        delete g;
        return code;
    };
    if (rc != BIHRT_OK) return fail(rc);
    // peer access both ways with context 0 (framebuffer stores go to device 0)
    g->peer_ok = true;
    for (int i = 1; i < ngpu; i++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, i, 0);
        if (!can) { g->peer_ok = false; continue; }
        cudaSetDevice(i);
        cudaError_t e = cudaDeviceEnablePeerAccess(0, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peer_ok = false;
        cudaGetLastError();
    }
    for (int i = 0; i < ngpu; i++) {
        cudaSetDevice(i);
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return fail(BIHRT_ERR_CUDA);
        g->done.push_back(ev);
    }
    cudaSetDevice(0);
    if (cudaEventCreateWithFlags(&g->start, cudaEventDisableTiming) != cudaSuccess) return fail(BIHRT_ERR_CUDA);
    if (ngpu > 1) {
        if (!nccl().ok) { bihrt_fail(g->ctx[0], BIHRT_ERR_CUDA, "libnccl.so.2 not found (or too old): %s", dlerror() ? dlerror() : "missing symbols"); return fail(BIHRT_ERR_CUDA); }
        g->comm.assign(ngpu, nullptr);
        std::vector<int> devs(ngpu);
        for (int i = 0; i < ngpu; i++) devs[i] = i;
        int r = nccl().CommInitAll(g->comm.data(), ngpu, devs.data());
        if (r != 0) { bihrt_fail(g->ctx[0], BIHRT_ERR_CUDA, "ncclCommInitAll failed: %s", nccl().GetErrorString(r)); return fail(BIHRT_ERR_CUDA); }
    }
    return BIHRT_OK;
}

void bihrt_destroy_multi(bihrt_ctx** ctxs, int32_t ngpu) {
    if (!ctxs || ngpu < 1 || !ctxs[0]) return;
    bihrt_group* g = ctxs[0]->group;
    for (int i = 0; i < ngpu; i++) if (ctxs[i]) { cudaSetDevice(ctxs[i]->device); cudaStreamSynchronize(ctxs[i]->stream); }
    if (g) {
        for (auto cm : g->comm) if (cm) nccl().CommDestroy(cm);
        for (size_t i = 0; i < g->done.size(); i++) { cudaSetDevice((int)i); cudaEventDestroy(g->done[i]); }
        if (g->start) { cudaSetDevice(0); cudaEventDestroy(g->start); }
    }
    for (int i = 0; i < ngpu; i++) if (ctxs[i]) { ctxs[i]->group = nullptr; bihrt_destroy(ctxs[i]); ctxs[i] = nullptr; }
    delete g;
}

int bihrt_multi_size(const bihrt_ctx* c) { return (c && c->group) ? (int)c->group->ctx.size() : (c ? 1 : 0); }

int bihrt_multi_nccl_version(void) {
    int v = 0;
    if (nccl().ok && nccl().GetVersion && nccl().GetVersion(&v) == 0) return v;
    return 0;
}

// Replicate the BIH of context 0 (the root passed in must be context 0 of its group) into every other context.
int bihrt_multi_broadcast(bihrt_ctx* root) {
    if (!root) return BIHRT_ERR_INVALID;
    bihrt_group* g = root->group;
    if (!g || g->ctx[0] != root) return bihrt_fail(root, BIHRT_ERR_STATE, "bihrt_multi_broadcast takes context 0 of a bihrt_create_multi group");
    if (!root->built) return bihrt_fail(root, BIHRT_ERR_STATE, "BIH not built");
    const int ngpu = (int)g->ctx.size();
    if (ngpu == 1) return BIHRT_OK;
    std::vector<void*> ptr(ngpu);
    uint64_t bytes = 0;
    for (int i = 0; i < ngpu; i++) {
        bihrt_ctx* c = g->ctx[i];
        if (i > 0) c->opt_morton_bits = root->built_quality ? 63 : 30;     // the replica is a tree of the builder's kind
        uint64_t b = 0;
        int rc = bihrt_bih_region(c, root->n, &ptr[i], &b);                // (sets the device; a no-op on the builder)
        if (rc) return i == 0 ? rc : bihrt_fail(root, rc, "context %d: %s", i, bihrt_last_error(c));
        bytes = b;
    }
    NCCL_CHECK(root, nccl().GroupStart());
    for (int i = 0; i < ngpu; i++) {
        cudaSetDevice(g->ctx[i]->device);
        int r = nccl().Broadcast(ptr[i], ptr[i], (size_t)bytes, NCCL_UINT8, 0, g->comm[i], g->ctx[i]->stream);
        if (r != 0) { nccl().GroupEnd(); return bihrt_fail(root, BIHRT_ERR_CUDA, "ncclBroadcast failed: %s", nccl().GetErrorString(r)); }
    }
    NCCL_CHECK(root, nccl().GroupEnd());
    for (int i = 1; i < ngpu; i++) { int rc = bihrt_bih_adopt(g->ctx[i], root->n); if (rc) return rc; }
    return BIHRT_OK;
}

// One frame over the whole group: same image as bihrt_render on one GPU, in context 0's framebuffer.
int bihrt_multi_render(bihrt_ctx* root, const bihrt_camera* cam, int32_t w, int32_t h, int32_t spp, uint64_t seed, uint32_t flags) {
    if (!root) return BIHRT_ERR_INVALID;
    bihrt_group* g = root->group;
    if (!g || g->ctx[0] != root) return bihrt_fail(root, BIHRT_ERR_STATE, "bihrt_multi_render takes context 0 of a bihrt_create_multi group");
    const int ngpu = (int)g->ctx.size();
    if (ngpu == 1) return bihrt_render(root, cam, w, h, spp, seed, flags);
    if (!g->peer_ok) return bihrt_fail(root, BIHRT_ERR_CUDA, "the devices of this group cannot address device 0's memory (no peer access)");
    // the interleave needs count | units per tile (32 << lane-group shift); fall back to fewer lane groups is not needed for
    // powers of two, anything else is refused by bihrt_render_interleaved_to with a message
    // the others may start once context 0 has got THIS far (last frame's readers of its framebuffer are done) ...
    BIHRT_CUDA(root, cudaSetDevice(root->device));
    BIHRT_CUDA(root, cudaEventRecord(g->start, root->stream));
    int rc = bihrt_render_interleaved_to(root, cam, w, h, spp, seed, flags, 0, ngpu, nullptr);     // (allocates the framebuffer if needed)
    if (rc) return rc;
    uint32_t* fb0 = nullptr;
    if ((rc = bihrt_framebuffer(root, &fb0, nullptr, nullptr))) return rc;
    for (int i = 1; i < ngpu; i++) {
        bihrt_ctx* c = g->ctx[i];
        BIHRT_CUDA(root, cudaSetDevice(c->device));
        BIHRT_CUDA(root, cudaStreamWaitEvent(c->stream, g->start, 0));
        rc = bihrt_render_interleaved_to(c, cam, w, h, spp, seed, flags, i, ngpu, fb0);
        if (rc) return bihrt_fail(root, rc, "context %d: %s", i, bihrt_last_error(c));
        BIHRT_CUDA(root, cudaSetDevice(c->device));
        BIHRT_CUDA(root, cudaEventRecord(g->done[i], c->stream));
    }
    // ... and whatever context 0 does next (bihrt_framebuffer_read, the next build) comes after all of them
    BIHRT_CUDA(root, cudaSetDevice(root->device));
    for (int i = 1; i < ngpu; i++) BIHRT_CUDA(root, cudaStreamWaitEvent(root->stream, g->done[i], 0));
    return BIHRT_OK;
}

int bihrt_multi_sync(bihrt_ctx* root) {
    if (!root) return BIHRT_ERR_INVALID;
    bihrt_group* g = root->group;
    if (!g) return bihrt_sync(root);
    for (auto c : g->ctx) { int rc = bihrt_sync(c); if (rc) return c == root ? rc : bihrt_fail(root, rc, "context %d: %s", c->group_rank, bihrt_last_error(c)); }
    return BIHRT_OK;
}

}  // extern "C"
