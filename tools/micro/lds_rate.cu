// Micro-benchmark: shared-memory load rate on B200 for the access patterns of the PQ ADC scan.
//   mode 0: contiguous row (addr = lane*4 + imm)           -- the textbook conflict-free LDS.32
//   mode 1: scattered rows, bank == lane (one 256-B line per lane chosen by a register byte): the ADC pattern
//   mode 2: mode 1 with LDS.64 (8-byte entries, half-warp conflict-free)
// Prints LDS warp-instructions per clock per SM.   nvcc -arch=sm_100a -O3 lds_rate.cu -o lds_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cycles, int iters, uint32_t seed) {
    extern __shared__ __align__(1024) unsigned char lut[];
    for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) ((uint32_t*)lut)[i] = i * 2654435761u >> 20;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t cb = (uint32_t)(lane << 2) | (1u << 8) | (2u << 16);
    uint32_t w = seed * (threadIdx.x + 1) * 2654435761u;
    uint32_t acc = 0, acc2 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (MODE == 0) {
                    acc += *(const uint32_t*)(lut + (lane << 2) + ((w >> 28) << 8) * 0 + (b * 8 + u) * 256);
                } else if (MODE == 1) {
                    const uint32_t addr = prmt(w, cb, 0x7004u | (((u % 3 == 0) ? 7u : (u % 3 == 1) ? 5u : 6u) << 8) | ((uint32_t)b << 4));
                    acc += *(const uint32_t*)(lut + addr + u * 4);
                } else {
                    const uint32_t off = (uint32_t)((lane + u) & 31) << 3;
                    const uint32_t addr = prmt(w, off, 0x7704u | ((uint32_t)b << 4));
                    const uint2 e = *(const uint2*)(lut + addr + (u % 3) * 65536);
                    acc += e.x; acc2 += e.y;
                }
            }
        }
        w = w * 1664525u + 1013904223u + acc * (MODE == 0 ? 0u : 0u);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc2;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
    const int iters = 2000;
    for (int r = 0; r < 2; r++) k<MODE><<<148, 1024, 196608>>>(out, cyc, iters, 12345u);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    const double lds = (double)iters * 32 * 32;  // warp-level LDS instructions per SM
    printf("%-44s %8.0f cycles, %.3f LDS/clk/SM (%s)\n", name, avg, lds / avg, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0>("contiguous LDS.32");
    run<1>("scattered, bank==lane, PRMT address, LDS.32");
    run<2>("scattered, half-warp conflict-free, LDS.64");
    return 0;
}
