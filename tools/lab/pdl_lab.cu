// How much does a kernel boundary cost inside a CUDA graph on this GPU, and how much of it does programmatic dependent launch hide?
// A chain of N kernels (245 x 256 threads, each spins ~T us on the clock after an optional griddepcontrol.wait).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void k_work(unsigned* p, int spin_cycles, int pdl) {
    if (pdl) asm volatile("griddepcontrol.launch_dependents;");
    __shared__ unsigned s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = 0;      // prologue that does not depend on the previous kernel
    __syncthreads();
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    long long t0 = clock64();
    while (clock64() - t0 < spin_cycles) { }
    if (threadIdx.x == 0) atomicAdd(p, s[blockIdx.x & 2047] + 1);
}
static float run(int n, int blocks, int spin, int pdl, cudaStream_t st, unsigned* d) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; i++) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = a; cfg.numAttrs = pdl ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, k_work, d, spin, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 20; r++) {
        cudaEventRecord(e0, st); CK(cudaGraphLaunch(ge, st)); cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    return best * 1e3f;
}
int main() {
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    unsigned* d; CK(cudaMalloc(&d, 4)); CK(cudaMemset(d, 0, 4));
    for (int blocks : {245, 1184, 3917}) for (int spin : {0, 10000, 30000}) {
        float a1 = run(1, blocks, spin, 0, st, d), a14 = run(15, blocks, spin, 0, st, d);
        float b1 = run(1, blocks, spin, 1, st, d), b14 = run(15, blocks, spin, 1, st, d);
        printf("blocks %4d spin %5d cyc: plain 1 kernel %.2f us, +%.2f us per extra kernel | PDL 1 kernel %.2f us, +%.2f us per extra kernel\n",
               blocks, spin, a1, (a14 - a1) / 14, b1, (b14 - b1) / 14);
    }
    return 0;
}
