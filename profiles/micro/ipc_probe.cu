// Peer-memory probe for the sharded trainer's device-initiated exchange: CUDA IPC between the one-process-
// per-GPU ranks of a torchrun job, and the round-trip latency of a flag written into a peer's memory
// over NVLink.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC
//   -o profiles/micro/libipc_probe.so profiles/micro/ipc_probe.cu ; run: profiles/micro/ipc_probe.py
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void pingpong_kernel(volatile unsigned *mine, volatile unsigned *peer, int rank, unsigned iters,
                                unsigned long long *ns_out) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (unsigned i = 1; i <= iters; i++) {
        if (rank == 0) {
            *peer = i;
            __threadfence_system();
            while (*mine < i) { }
        } else {
            while (*mine < i) { }
            *peer = i;
            __threadfence_system();
        }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    *ns_out = t1 - t0;
}

// bulk push: every thread stores 16 bytes into the peer buffer, then one flag
__global__ void push_kernel(uint4 *peer, size_t n16, volatile unsigned *peer_flag, unsigned tag) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        peer[i] = make_uint4(tag, (unsigned)i, 0, 0);
    __threadfence_system();
}

extern "C" {
int probe_alloc(size_t bytes, void **p) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
    return (int)e;
}
int probe_export(void *p, void *handle64) { return (int)cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, p); }
int probe_open(const void *handle64, void **p) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess);
}
int probe_pingpong(void *mine, void *peer, int rank, unsigned iters, double *us_per_roundtrip) {
    unsigned long long *d_ns, ns = 0;
    cudaMalloc(&d_ns, 8);
    pingpong_kernel<<<1, 1>>>((volatile unsigned *)mine, (volatile unsigned *)peer, rank, iters, d_ns);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&ns, d_ns, 8, cudaMemcpyDeviceToHost);
    cudaFree(d_ns);
    *us_per_roundtrip = ns * 1e-3 / iters;
    return (int)e;
}
int probe_push(void *peer, size_t bytes, int reps, double *gbs) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    push_kernel<<<148 * 4, 256>>>((uint4 *)peer, bytes / 16, nullptr, 1);
    cudaEventRecord(a);
    for (int r = 0; r < reps; r++) push_kernel<<<148 * 4, 256>>>((uint4 *)peer, bytes / 16, nullptr, 2 + r);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    *gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
    return (int)e;
}
const char *probe_err(int e) { return cudaGetErrorString((cudaError_t)e); }
}
