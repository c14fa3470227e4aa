/*
 * bihrt_cli.c -- C host program on top of the C ABI (include/bihrt.h).
 *
 * The headless counterpart of the reference's entry point and frame loop
 * (main(), R/src/Main.cpp:44-68; App::LoadModels, R/src/App.cpp:65-167; App::Run, R/src/App.cpp:170-187;
 * Renderer::Render, R/src/Renderer.cpp:415-672): load a mesh, then per frame rebuild the BIH and render
 * with the reference's camera, resolution and samples per pixel; the packed framebuffer is written as a
 * PPM instead of being presented through OpenGL (R/src/Renderer.cpp:644-670).
 *
 *   bihrt_cli <mesh.obj | mesh.tri9> [-w 640] [-h 480] [-s 4] [-f frames] [-o out.ppm] [-d device] [-g gpus]
 *   (.tri9 = raw little-endian float32, 9 floats per triangle)
 *
 * -g N (N = 2, 4, 8): devices 0 .. N-1 of this process share the frame through bihrt_create_multi -- N
 * contexts and one NCCL communicator held inside the library.  Device 0 rebuilds the BIH, bihrt_multi_broadcast replicates it
 * (one ncclBroadcast, in place), bihrt_multi_render lets every device trace every N-th run of 32-ray units of every tile and
 * store the finished pixels straight into device 0's framebuffer: no reduce, no second framebuffer.  -p selects the
 * NCCL-free variant instead (bihrt_bih_copy = peer copy of the blob + bihrt_render_interleaved_to per device).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "bihrt.h"

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ != BIHRT_OK) { \
    fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? bihrt_last_error(ctx) : "(no context)"); \
    if (ctx) { if (group_in_use) bihrt_destroy_multi(group_ptr, group_n); else bihrt_destroy(ctx); } return 1; } } while (0)

static int group_in_use = 0, group_n = 0;
static bihrt_ctx** group_ptr = NULL;

int main(int argc, char** argv) {
    const char* path = NULL; const char* out = "frame.ppm";
    int w = 640, h = 480, spp = 4, frames = 1, device = 0, gpus = 1, peer_copy = 0;     /* R/src/Constants.h:4-8 */
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-w") && i + 1 < argc) w = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-h") && i + 1 < argc) h = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) spp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-f") && i + 1 < argc) frames = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-p")) peer_copy = 1;
        else path = argv[i];
    }
    if (!path) { fprintf(stderr, "usage: %s <mesh.obj|mesh.tri9> [-w W] [-h H] [-s spp] [-f frames] [-o out.ppm] [-d dev] [-g gpus [-p]]\n", argv[0]); return 2; }
    if (gpus < 1 || gpus > 16) { fprintf(stderr, "-g must be 1..16\n"); return 2; }

    bihrt_ctx* ctx = NULL;
    bihrt_ctx* group[16] = { NULL };             /* -g N through bihrt_create_multi: group[0] is ctx */
    int use_group = gpus > 1 && !peer_copy;
    if (use_group) {
        int rc = bihrt_create_multi(group, gpus);
        if (rc != BIHRT_OK) {
            fprintf(stderr, "bihrt_create_multi(%d) -> %d (NCCL or peer access unavailable): falling back to peer copies\n", gpus, rc);
            use_group = 0;
        } else {
            ctx = group[0]; group_in_use = 1; group_n = gpus; group_ptr = group;
            printf("bihrt_create_multi: %d contexts, NCCL %d\n", bihrt_multi_size(ctx), bihrt_multi_nccl_version());
        }
    }
    if (!use_group) {
        bihrt_config cfg; memset(&cfg, 0, sizeof cfg); cfg.device = device;
        CHECK(bihrt_create(&ctx, &cfg));
    }
#define DESTROY_ALL() do { if (use_group) bihrt_destroy_multi(group, gpus); else if (ctx) bihrt_destroy(ctx); ctx = NULL; } while (0)

    size_t len = strlen(path);
    if (len > 5 && !strcmp(path + len - 5, ".tri9")) {
        FILE* f = fopen(path, "rb");
        if (!f) { fprintf(stderr, "cannot open %s\n", path); DESTROY_ALL(); return 1; }
        fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
        float* tri = (float*)malloc((size_t)bytes);
        if (!tri || fread(tri, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "read failed\n"); fclose(f); DESTROY_ALL(); return 1; }
        fclose(f);
        CHECK(bihrt_scene_load_triangles(ctx, tri, (int64_t)(bytes / 36)));
        CHECK(bihrt_sync(ctx));
        free(tri);
    } else {
        CHECK(bihrt_scene_load_obj(ctx, path));
    }

    /* Camera((2,0,-2), W/H), R/src/Renderer.cpp:99, R/src/Camera.cu:5-9 */
    const float aspect = (float)w / (float)h;
    bihrt_camera cam = { { 2.0f, 0.0f, -2.0f }, { 0.0f, -1.0f, -1.0f }, { aspect * 2.0f, 0.0f, 0.0f }, { 0.0f, 2.0f, 0.0f } };

    /* helper contexts on the other devices (-g N) */
    bihrt_ctx* helper[16] = { NULL };
    for (int g = 1; g < gpus && !use_group; g++) {
        bihrt_config hc; memset(&hc, 0, sizeof hc); hc.device = device + g;
        int rc = bihrt_create(&helper[g], &hc);
        if (rc != BIHRT_OK) { fprintf(stderr, "bihrt_create(device %d) -> %d\n", device + g, rc); for (int k = 1; k < g; k++) bihrt_destroy(helper[k]); bihrt_destroy(ctx); return 1; }
    }
#define CHECKH(g, call) do { int rc_ = (call); if (rc_ != BIHRT_OK) { \
    fprintf(stderr, "device %d: %s -> %d: %s\n", device + (g), #call, rc_, bihrt_last_error(helper[g])); \
    for (int k_ = 1; k_ < gpus; k_++) { bihrt_destroy(helper[k_]); } \
    bihrt_destroy(ctx); return 1; } } while (0)

    bihrt_build_info info;
    for (int f = 0; f < frames; f++) {                         /* the reference rebuilds every frame */
        double t0 = now_ms();
        CHECK(bihrt_build(ctx));
        if (gpus == 1) {
            CHECK(bihrt_render(ctx, &cam, w, h, spp, 1984u + (uint64_t)f, BIHRT_RENDER_JITTER));
        } else if (use_group) {
            CHECK(bihrt_multi_broadcast(ctx));
            CHECK(bihrt_multi_render(ctx, &cam, w, h, spp, 1984u + (uint64_t)f, BIHRT_RENDER_JITTER));
            CHECK(bihrt_multi_sync(ctx));
        } else {
            uint32_t* fb0 = NULL;
            for (int g = 1; g < gpus; g++) CHECKH(g, bihrt_bih_copy(helper[g], ctx));      /* waits for the build only */
            CHECK(bihrt_render_interleaved_to(ctx, &cam, w, h, spp, 1984u + (uint64_t)f, BIHRT_RENDER_JITTER, 0, gpus, NULL));
            CHECK(bihrt_framebuffer(ctx, &fb0, NULL, NULL));
            for (int g = 1; g < gpus; g++)
                CHECKH(g, bihrt_render_interleaved_to(helper[g], &cam, w, h, spp, 1984u + (uint64_t)f, BIHRT_RENDER_JITTER, g, gpus, fb0));
            for (int g = 1; g < gpus; g++) CHECKH(g, bihrt_sync(helper[g]));
        }
        CHECK(bihrt_sync(ctx));
        double t1 = now_ms();
        CHECK(bihrt_get_build_info(ctx, &info));
        printf("frame %d: %lld triangles, %lld leaves, build %.3f ms (device), frame %.3f ms (host), %.1f Mrays/s incl. build\n",
               f, (long long)info.n, (long long)info.nu, info.last_build_ms, t1 - t0, (double)w * h * spp / ((t1 - t0) * 1e3));
    }

    uint32_t* fb = (uint32_t*)malloc((size_t)w * h * 4);
    CHECK(bihrt_framebuffer_read(ctx, fb));
    FILE* o = fopen(out, "wb");
    if (o) {
        fprintf(o, "P6\n%d %d\n255\n", w, h);
        for (int j = h - 1; j >= 0; j--)                       /* row 0 = bottom */
            for (int i = 0; i < w; i++) {
                uint32_t p = fb[(size_t)j * w + i];
                unsigned char rgb[3] = { (unsigned char)(p & 255), (unsigned char)((p >> 8) & 255), (unsigned char)((p >> 16) & 255) };
                fwrite(rgb, 1, 3, o);
            }
        fclose(o);
        printf("wrote %s\n", out);
    }
    free(fb);
    for (int g = 1; g < gpus && !use_group; g++) bihrt_destroy(helper[g]);
    DESTROY_ALL();
    return 0;
}
