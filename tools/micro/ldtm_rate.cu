// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput on B200 by shape and by number of reading warps.
// Prints bytes per clock per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 ldtm_rate.cu -o ldtm_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(SHAPE, v, addr)                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 "                                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),      \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),            \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),          \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),          \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                               \
                 : "r"(addr))

// MODE 0: 32x32b.x32 (one lane = one row, 32 columns)   4 KiB per warp instruction
// MODE 1: 16x256b.x4 (16 lanes x 32 columns x ... )      32 registers as well
// MODE 2: 16x128b.x8
// MODE 3: 32x32b.x32 with two loads in flight before the wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cycles, int iters, int warps_active) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < warps_active) {
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t v[32], w[32];
        for (int it = 0; it < iters; it++) {
            const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 128) & 255);
            if (MODE == 0) { LD32("32x32b.x32", v, taddr + col); }
            if (MODE == 1) { LD32("16x256b.x8", v, taddr + col); }
            if (MODE == 2) { LD32("16x128b.x16", v, taddr + col); }
            if (MODE == 3) { LD32("32x32b.x32", v, taddr + col); LD32("32x32b.x32", w, taddr + ((col + 32) & 255)); }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i++) acc ^= v[i];
            if (MODE == 3) {
#pragma unroll
                for (int i = 0; i < 32; i++) acc ^= w[i];
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int warps_active, int bytes_per_iter) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, 512>>>(out, cyc, 100, warps_active);
    k<MODE><<<148, 512>>>(out, cyc, iters, warps_active);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    printf("%-28s warps %2d  %s  %.1f B/clk/SM  (%.0f clk per warp-instruction)\n", name, warps_active, cudaGetErrorString(e),
           (double)bytes_per_iter * warps_active * iters / avg, avg / iters / (MODE == 3 ? 2 : 1));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8, 16}) run<0>("32x32b.x32", w, 4096);
    for (int w : {4, 8}) run<3>("32x32b.x32 two in flight", w, 8192);
    for (int w : {4, 8}) run<1>("16x256b.x8", w, 4096);
    for (int w : {4, 8}) run<2>("16x128b.x16", w, 4096);
    return 0;
}
