// Microbenchmark: cost of a cooperative-groups grid barrier on B200 for several grid shapes,
// and of an L2-resident table scan.  nvcc -arch=sm_100a -O3 -o gridsync gridsync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void sync_loop(int iters, unsigned long long *sink) {
    cg::grid_group grid = cg::this_grid();
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) {
        acc += i;
        grid.sync();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

__global__ void scan_loop(const unsigned *keys, const unsigned long long *cnt, size_t n, int iters, unsigned long long *sink) {
    cg::grid_group grid = cg::this_grid();
    unsigned long long best = 0;
    for (int it = 0; it < iters; it++) {
        for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
            if (keys[s] == 0xFFFFFFFFu) continue;
            unsigned long long c = cnt[s];
            best = c > best ? c : best;
        }
        grid.sync();
    }
    if (best == 12345) *sink = best;
}

int main() {
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000;
    for (int tpb : {256, 1024}) {
        for (int per_sm : {1, 2, 4}) {
            if (tpb == 1024 && per_sm > 2) continue;
            int grid = 148 * per_sm;
            void *args[] = {&iters, &sink};
            cudaLaunchCooperativeKernel((void *)sync_loop, dim3(grid), dim3(tpb), args, 0, 0);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void *)sync_loop, dim3(grid), dim3(tpb), args, 0, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("grid.sync  grid=%4d x %4d threads : %.2f us per barrier (%s)\n", grid, tpb, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int log2 : {18, 20, 22}) {
        size_t n = (size_t)1 << log2;
        unsigned *keys; unsigned long long *cnt;
        cudaMalloc(&keys, n * 4); cudaMalloc(&cnt, n * 8);
        cudaMemset(keys, 0x11, n * 4); cudaMemset(cnt, 0, n * 8);
        int it2 = 500, grid = 592;
        void *args[] = {&keys, &cnt, &n, &it2, &sink};
        cudaLaunchCooperativeKernel((void *)scan_loop, dim3(grid), dim3(256), args, 0, 0);
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((void *)scan_loop, dim3(grid), dim3(256), args, 0, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("table scan 2^%d slots (%zu MB) + barrier: %.2f us per pass\n", log2, n * 12 >> 20, ms * 1e3 / it2);
        cudaFree(keys); cudaFree(cnt);
    }
    return 0;
}
