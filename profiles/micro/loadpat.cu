// Microbenchmark: how fast can one-thread-per-record walkers stream their OWN record (240 KB apart from
// the neighbour lane's) into the SM?  The fused encoder's refill has exactly this access pattern, and the
// round-1 profile charged most of the l1tex pipe to it (32 distinct lines per LDG.128 request).
//   A  4 x LDG.128 per 64 B per lane                      (round-1 refill)
//   B  8 x LDG.128 per 128 B line per lane
//   C  4 x 256-bit ld.global.v8.f32 per 128 B line per lane
//   D  cp.async.bulk (TMA, UBLKCP) 128 B per lane into a padded slab + mbarrier per lane, then LDS.128
//   E  the same with 256 B per lane
//   F  warp-cooperative: the warp reads lane l's line with one coalesced LDG.32 request, l = 0..31
//   G  cp.async (LDGSTS) 16 B x 8 per lane into the slab
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loadpat loadpat.cu ; ./loadpat [records_per_thread]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRec = 60000;  // floats per record

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float sum4(uint4 v) {
    return __uint_as_float(v.x) + __uint_as_float(v.y) + __uint_as_float(v.z) + __uint_as_float(v.w);
}

template <int MODE>
__global__ void __launch_bounds__(768, 1) stream_kernel(const float *__restrict__ in, size_t n_rec, float *sink, int slab_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t r_lo = n_rec * blockIdx.x / gridDim.x, r_hi = n_rec * (blockIdx.x + 1) / gridDim.x;
    float acc = 0.f;
    if (MODE == 0 || MODE == 1 || MODE == 2) {
        for (size_t r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
            const float *p = in + r * kRec;
            if (MODE == 0) {
                for (int g = 0; g < kRec; g += 16) {
#pragma unroll
                    for (int j = 0; j < 4; j++) acc += sum4(__ldg(reinterpret_cast<const uint4 *>(p + g) + j));
                }
            } else if (MODE == 1) {
                for (int g = 0; g < kRec; g += 32) {
#pragma unroll
                    for (int j = 0; j < 8; j++) acc += sum4(__ldg(reinterpret_cast<const uint4 *>(p + g) + j));
                }
            } else {
                for (int g = 0; g < kRec; g += 32) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float a0, a1, a2, a3, a4, a5, a6, a7;
                        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)
                                     : "l"(p + g + 8 * j));
                        acc += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
                    }
                }
            }
        }
    } else if (MODE == 3 || MODE == 4) {
        // per-lane slab (slab_bytes + 16 pad) and mbarrier
        const int stride = slab_bytes + 16;
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
        uint8_t *slabs = smem + ((blockDim.x * 8 + 127) & ~127);
        const uint32_t bar = smem_u32(&bars[threadIdx.x]);
        const uint32_t slab = smem_u32(slabs + (size_t)threadIdx.x * stride);
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        uint32_t parity = 0;
        const int per = slab_bytes / 4;  // floats per slab
        for (size_t r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
            const float *p = in + r * kRec;
            // prologue
            asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(bar), "r"(slab_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(slab), "l"(p), "r"(slab_bytes), "r"(bar) : "memory");
            for (int g = 0; g < kRec; g += per) {
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                }
                parity ^= 1;
                for (int j = 0; j < slab_bytes / 16; j++) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(slab + j * 16));
                    acc += sum4(v);
                }
                if (g + per < kRec) {
                    const int bytes = min(slab_bytes, (kRec - g - per) * 4);
                    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(slab), "l"(p + g + per), "r"(bytes), "r"(bar) : "memory");
                }
            }
        }
    } else if (MODE == 5) {
        // warp-cooperative: 32 walkers of the warp, one line at a time, coalesced
        for (size_t r0 = r_lo + (size_t)warp * 32; r0 < r_hi; r0 += blockDim.x) {
            for (int g = 0; g < kRec; g += 32) {
#pragma unroll 8
                for (int l = 0; l < 32; l++) {
                    const size_t r = r0 + l;
                    if (r < r_hi) acc += __ldg(in + r * kRec + g + lane);
                }
            }
        }
    } else if (MODE == 6) {
        const int stride = 128 + 16;
        uint8_t *slabs = smem;
        const uint32_t slab = smem_u32(slabs + (size_t)threadIdx.x * stride);
        for (size_t r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
            const float *p = in + r * kRec;
            for (int g = 0; g < kRec; g += 32) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slab + j * 16), "l"(p + g + 4 * j) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                for (int j = 0; j < 8; j++) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(slab + j * 16));
                    acc += sum4(v);
                }
            }
        }
    }
    if (acc == 1234.5678f) *sink = acc;
}

template <int MODE>
static void run(const char *name, const float *d_in, size_t n_rec, float *d_sink, int threads, int slab_bytes) {
    size_t smem = 0;
    if (MODE == 3 || MODE == 4) smem = ((threads * 8 + 127) & ~127) + (size_t)threads * (slab_bytes + 16);
    if (MODE == 6) smem = (size_t)threads * 144;
    cudaFuncSetAttribute(stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
        cudaEventRecord(e0);
        stream_kernel<MODE><<<148, threads, smem>>>(d_in, n_rec, d_sink, slab_bytes);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("%-34s threads %4d: %s\n", name, threads, cudaGetErrorString(e)); return; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double gb = (double)n_rec * kRec * 4 / 1e9;
    printf("%-34s threads %4d: %8.3f ms  %8.1f GB/s\n", name, threads, best, gb / (best * 1e-3));
}

int main(int argc, char **argv) {
    const size_t n_rec = argc > 1 ? (size_t)atol(argv[1]) : 40000;  // 9.6 GB
    float *d_in, *d_sink;
    cudaMalloc(&d_in, n_rec * kRec * 4 + 1024);
    cudaMalloc(&d_sink, 4);
    cudaMemset(d_in, 0, n_rec * kRec * 4 + 1024);
    for (int threads : {256, 384, 512, 768}) {
        run<0>("A 4xLDG.128 / 64 B", d_in, n_rec, d_sink, threads, 0);
        run<1>("B 8xLDG.128 / 128 B", d_in, n_rec, d_sink, threads, 0);
        run<2>("C 4xLDG.256 / 128 B", d_in, n_rec, d_sink, threads, 0);
        run<3>("D TMA bulk 128 B + LDS.128", d_in, n_rec, d_sink, threads, 128);
        if (threads <= 512) run<4>("E TMA bulk 256 B + LDS.128", d_in, n_rec, d_sink, threads, 256);
        run<5>("F warp-cooperative LDG.32 lines", d_in, n_rec, d_sink, threads, 0);
        run<6>("G LDGSTS 8x16 B + LDS.128", d_in, n_rec, d_sink, threads, 0);
    }
    return 0;
}
