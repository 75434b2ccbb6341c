// kernels_filter.cu -- predicate -> dense allow-bitmap.
// Replaces the 1-byte-per-row masks of internal/simd/simd.go:572-761 (MatchInt64 / MatchFloat32),
// the AND of internal/simd/simd.go:119-126 and the row -> bitset scatter of
// internal/query/filter_evaluator.go:700-758 with one pass that emits 1 bit per row.
#include "kernels.cuh"

namespace lb {

template <typename T>
__device__ __forceinline__ bool cmp(T v, T val, int op) {
    switch (op) {
        case 0: return v == val;
        case 1: return v != val;
        case 2: return v > val;
        case 3: return v >= val;
        case 4: return v < val;
        default: return v <= val;
    }
}

// one warp -> one 32-bit bitmap word via ballot; coalesced column reads
template <typename T>
__global__ void filter_kernel(const T* __restrict__ col, int64_t n, int op, T val, int and_into,
                              uint32_t* __restrict__ bitmap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool m = (i < n) && cmp<T>(col[i < n ? i : 0], val, op);
    uint32_t w = __ballot_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && (i < n)) {
        int64_t word = i >> 5;
        bitmap[word] = and_into ? (bitmap[word] & w) : w;
    }
}

cudaError_t launch_filter_i64(const int64_t* col, int64_t n, int op, int64_t val, int and_into, uint32_t* bitmap,
                              cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    filter_kernel<int64_t><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(col, n, op, val, and_into, bitmap);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_filter_f32(const float* col, int64_t n, int op, float val, int and_into, uint32_t* bitmap,
                              cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    filter_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(col, n, op, val, and_into, bitmap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace lb
