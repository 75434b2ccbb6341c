// pq_scan.cu -- the PQ asymmetric-distance (ADC) code scan, round 2: coarse integer pass -> certify -> exact.
//
// Reference semantics (never changed): results[i] = float32(sqrt(float64(sum_j table[j*256 + codes[i*M+j]]))),
// the sum accumulated sequentially in j with plain fp32 adds (internal/simd/simd.go:345-355; every lane of the
// AVX2 gather kernel pq_amd64.s:14-168 agrees).  That order is what the EXACT stage below reproduces -- but only
// for the few candidates that survive a coarse pass whose job is to touch every code byte once, at HBM speed.
//
// Why the round-1 kernel sat at 0.29 of the HBM roofline: one fp32 LUT in shared memory, 32 lanes = 32 rows
// looking up the SAME sub-quantiser j with random codes => random banks, 3.1 wavefronts per load
// (profiles/r1_ncu_summary.md).  The fix is a layout in which the bank of every look-up is decided by the lane,
// not by the code:
//
//   * lane l owns row 32*T + l of tile T and, at step t of group g, looks up sub-quantiser j = 32 g + (l + t) % 32
//     -- 32 lanes, 32 DIFFERENT sub-quantisers;
//   * the table of a group is laid out [code][slot]: one line of 32 slots per code value, slot s holding the
//     entry of sub-quantiser 32 g + s % 32 for that code.  The word a lane reads sits in bank (l + t) % 32
//     whatever the code is: every warp-wide look-up is exactly one conflict-free wavefront;
//   * integer sums are associative, so the rotated visiting order costs nothing -- the table is quantised to
//     integers (25 bits per entry for one query, 12 bits when four queries share a 64-bit entry) and the coarse
//     key is an exact integer sum;
//   * the code bytes are stored in HBM in the order the lanes consume them (lb_pq owns its mirror): 32-row tiles
//     of 16-byte chunks, chunk i of row l at tile + (32 i + l) * 16 -- one LDG.128 per lane reads 512 contiguous
//     bytes per warp -- with the bytes of every 32-byte group rotated left by (row % 32), so the byte a lane needs
//     at step t is byte t of its registers: no shuffles, no dynamic register indexing;
//   * the address of a look-up is assembled by ONE byte-permute: byte 1 = the code (x 256 B per table line),
//     byte 0 = the lane/step offset, group in the immediate.  Single-query tables store every line twice
//     (64 slots) so that (l + t) needs no modulo and t lives in the load's immediate offset.
//
// Coarse keys order rows up to the quantisation step; the exact stage recomputes the reference's fp32 sum for the
// kc best, sorts by (distance, id), and CERTIFIES that no row outside the candidate set can reach the k'-th exact
// distance (bound: M/2 + 1 quantisation steps + the fp32 summation error).  Queries that fail are flagged and
// re-done by the exhaustive fp32 kernel (kernels_pq.cu).
//
// Roofline: HBM for one query per pass (N * Mp bytes; C3: 960 MB), shared-memory bandwidth when four queries
// share a pass (8 B per (row, sub-quantiser) look-up for 4 queries).
#include "kernels.cuh"

namespace lb {

constexpr int PQS_THREADS = 512;
constexpr int PQS_WARPS = PQS_THREADS / 32;
constexpr int PQS_GROUP_BYTES = 65536;  // one group's table: 256 codes x 256 B

// ---------------------------------------------------------------------------------------------
// flat [n][M] codes -> tiled + rotated mirror (appending at row0).  One thread per (row, 16-byte chunk).
// ---------------------------------------------------------------------------------------------
__global__ void pq_tile_codes_kernel(const uint8_t* __restrict__ flat, int64_t n, int M, int Mp, int64_t row0,
                                     uint8_t* __restrict__ tiled) {
    const int chunks = Mp / 16;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * chunks) return;
    const int64_t i = t / chunks;
    const int ch = (int)(t % chunks);
    const int64_t r = row0 + i;
    const int rot = (int)(r & 31);
    const uint8_t* src = flat + (size_t)i * M;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const int pos = ch * 16 + b;            // byte position in the rotated row
        const int g = pos >> 5, tt = pos & 31;
        const int j = g * 32 + ((tt + rot) & 31);  // the sub-quantiser consumed at step tt by this row's lane
        const uint32_t v = (j < M) ? src[j] : 0u;
        w[b >> 2] |= v << (8 * (b & 3));
    }
    uint4* dst = reinterpret_cast<uint4*>(tiled + (size_t)(r >> 5) * 32 * Mp) + ((size_t)ch * 32 + (r & 31));
    *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

cudaError_t launch_pq_tile_codes(const uint8_t* flat, int64_t n, int M, int Mp, int64_t row0, uint8_t* tiled,
                                 cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (Mp / 16);
    pq_tile_codes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(flat, n, M, Mp, row0, tiled);
    count_launch();
    return cudaGetLastError();
}

// code byte of (row, sub-quantiser j) in the tiled mirror
__device__ __forceinline__ uint32_t tiled_code(const uint8_t* __restrict__ tiled, int Mp, uint32_t row, int j) {
    const int g = j >> 5;
    const int tt = ((j & 31) - (int)(row & 31)) & 31;
    const int pos = g * 32 + tt;
    return tiled[(size_t)(row >> 5) * 32 * Mp + ((size_t)(pos >> 4) * 32 + (row & 31)) * 16 + (pos & 15)];
}

// ---------------------------------------------------------------------------------------------
// LUT quantisation.  One block per query GROUP (NQ queries).  Per query: min_j over each sub-quantiser, the
// widest range R, scale = SMAX / R; entry = rint((lut - min_j) * scale) in double.  Reconstruction:
//   sum_j lut[j][c_j] = base + key / scale + e,  |e| <= (M/2 + 1) / scale     (base = sum_j min_j)
// NQ = 1: u32 entries, SMAX = (2^32 - 1) / M, table lines stored twice (64 slots).
// NQ = 4: four u16 (12-bit: SMAX = 4095) entries packed into 8 bytes, 32 slots.
// ---------------------------------------------------------------------------------------------
struct PqQParams {
    double base;       // sum of the per-sub-quantiser minima
    double inv_scale;  // real units per key unit
    double smax_sum;   // upper bound of any real sum (for the fp32 summation error bound)
};

template <int NQ>
__global__ void __launch_bounds__(256)
adc_quantise_kernel(const float* __restrict__ luts, int M, int G, int nq, uint8_t* __restrict__ lutq,
                    PqQParams* __restrict__ params) {
    __shared__ float s_min[4][256];   // [q][j] (M <= 96 here; sized for safety to 256)
    __shared__ double s_rng[4][256];
    __shared__ double s_scale[4];
    const int qg = blockIdx.x, tid = threadIdx.x;
    for (int qi = 0; qi < NQ; qi++) {
        const int q = qg * NQ + qi;
        for (int j = tid; j < M; j += blockDim.x) {
            float mn = INFINITY, mx = -INFINITY;
            if (q < nq) {
                const float* t = luts + ((size_t)q * M + j) * 256;
                for (int c = 0; c < 256; c++) { const float v = t[c]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
            } else { mn = 0.f; mx = 0.f; }
            s_min[qi][j] = mn;
            s_rng[qi][j] = (double)mx - (double)mn;
        }
    }
    __syncthreads();
    if (tid < NQ) {
        const int q = qg * NQ + tid;
        double base = 0, R = 0, top = 0;
        for (int j = 0; j < M; j++) {
            base += (double)s_min[tid][j];
            R = fmax(R, s_rng[tid][j]);
            top += (double)s_min[tid][j] + s_rng[tid][j];
        }
        const double smax = (NQ == 1) ? floor(4294960000.0 / (double)M) : 4095.0;  // sums stay below 2^32 - 1
        const double scale = (R > 0 && isfinite(R)) ? smax / R : 0.0;
        s_scale[tid] = scale;
        if (q < nq) {
            params[q].base = base;
            params[q].inv_scale = scale > 0 ? 1.0 / scale : 0.0;
            params[q].smax_sum = top;
        }
    }
    __syncthreads();
    uint8_t* out = lutq + (size_t)qg * G * PQS_GROUP_BYTES;
    if (NQ == 1) {
        // [g][code][64 slots] u32, slot s <- sub-quantiser 32 g + s % 32
        const int q = qg;
        const double scale = s_scale[0];
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
        const int total = G * 256 * 64;
        for (int e = tid; e < total; e += blockDim.x) {
            const int s = e & 63, c = (e >> 6) & 255, g = e >> 14;
            const int j = g * 32 + (s & 31);
            uint32_t v = 0;
            if (j < M && q < nq) {
                const double x = ((double)luts[((size_t)q * M + j) * 256 + c] - (double)s_min[0][j]) * scale;
                v = (uint32_t)llrint(x);
            }
            o32[e] = v;
        }
    } else {
        // [g][code][32 slots] of 4 x u16
        uint2* o64 = reinterpret_cast<uint2*>(out);
        const int total = G * 256 * 32;
        for (int e = tid; e < total; e += blockDim.x) {
            const int s = e & 31, c = (e >> 5) & 255, g = e >> 13;
            const int j = g * 32 + s;
            uint32_t v[4] = {0, 0, 0, 0};
            if (j < M) {
#pragma unroll
                for (int qi = 0; qi < 4; qi++) {
                    const int q = qg * 4 + qi;
                    if (q < nq) {
                        const double x = ((double)luts[((size_t)q * M + j) * 256 + c] - (double)s_min[qi][j]) * s_scale[qi];
                        v[qi] = (uint32_t)llrint(x);
                    }
                }
            }
            o64[e] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The coarse scan.  grid = (query groups, parts); 512 threads; one CTA per SM (table: G x 64 KB of shared memory).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

struct PqCoarseArgs {
    const uint4* tiles;       // tiled + rotated codes
    uint32_t n_rows;
    uint32_t tiles_per_part;
    const uint8_t* lutq;      // [qgroups][G * 64 KB]
    const uint32_t* tomb;
    uint32_t tomb_bits;
    const uint32_t* allow;
    int kc, cap, nq;
    // output: per query one compact list every CTA appends its survivors to (one atomic per CTA and query)
    uint64_t* compact;        // [nq][stride]  (key << 32 | row)
    uint32_t* out_cnt;        // [nq], zeroed by the caller
    size_t stride;            // >= parts * kc
    // shared progressive threshold: the smallest kc-th-best key any CTA of the query has published.  Some CTA
    // holds kc live rows at or below it, so no row above it can be among the global kc best.
    uint32_t* g_tau;          // [nq], initialised to 0xffffffff by the caller
};

template <int NQ, int G>
__global__ void __launch_bounds__(PQS_THREADS, 1)
adc_coarse_kernel(const PqCoarseArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* lut = smem_raw;                                             // G * 64 KB
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw + G * PQS_GROUP_BYTES);  // [NQ][cap]
    __shared__ int s_cnt[NQ];
    __shared__ uint32_t s_tau[NQ];
    constexpr int CH = 2 * G;  // 16-byte chunks per row
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qg = blockIdx.x, part = blockIdx.y;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.lutq + (size_t)qg * G * PQS_GROUP_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(lut);
        for (int i = tid; i < G * (PQS_GROUP_BYTES / 16); i += PQS_THREADS) dst[i] = __ldg(src + i);
    }
    const int nvalid = min(NQ, a.nq - qg * NQ);  // queries of this group that exist (the last group may be short)
    if (tid < NQ) { s_cnt[tid] = 0; s_tau[tid] = (tid < nvalid) ? __ldcg(a.g_tau + qg * NQ + tid) : 0u; }
    __syncthreads();

    const uint32_t n_tiles = (a.n_rows + 31) >> 5;
    const uint32_t tile_begin = (uint32_t)part * a.tiles_per_part;
    const uint32_t tile_end = min(n_tiles, tile_begin + a.tiles_per_part);
    const uint32_t iters = (tile_end > tile_begin) ? (tile_end - tile_begin + PQS_WARPS - 1) / PQS_WARPS : 0;

    uint4 cur[CH], nxt[CH];
    {
        const uint32_t t0 = tile_begin + warp;
#pragma unroll
        for (int i = 0; i < CH; i++) {
            cur[i] = make_uint4(0, 0, 0, 0);
            if (t0 < tile_end) cur[i] = __ldcs(a.tiles + ((size_t)t0 * CH + i) * 32 + lane);
        }
    }
    // lane constant for the byte permute: byte0 = 4 * lane (NQ = 1), bytes 1..3 = 1, 2, 0 (group selectors)
    const uint32_t cb = (uint32_t)(lane << 2) | (1u << 8) | (2u << 16);

    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t tile = tile_begin + it * PQS_WARPS + warp;
        const uint32_t tn = tile + PQS_WARPS;
#pragma unroll
        for (int i = 0; i < CH; i++) {
            nxt[i] = make_uint4(0, 0, 0, 0);
            if (tn < tile_end) nxt[i] = __ldcs(a.tiles + ((size_t)tn * CH + i) * 32 + lane);
        }
        uint32_t key[NQ];
        if (NQ == 1) {
            uint32_t acc = 0;
#pragma unroll
            for (int g = 0; g < G; g++) {
                // byte2 of the address = g: selector index 7 -> 0, 5 -> 1, 6 -> 2 (bytes of cb)
                const uint32_t gsel = (g == 0) ? 7u : (g == 1) ? 5u : 6u;
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const uint4 v = cur[g * 2 + c];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int wi = 0; wi < 4; wi++)
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int t = c * 16 + wi * 4 + b;
                            const uint32_t addr = prmt(w[wi], cb, 0x7004u | (gsel << 8) | ((uint32_t)b << 4));
                            acc += *reinterpret_cast<const uint32_t*>(lut + addr + t * 4);
                        }
                }
            }
            key[0] = acc;
        } else {
            uint32_t tot[4] = {0, 0, 0, 0};
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const uint4 v = cur[g * 2 + c];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    uint32_t p01 = 0, p23 = 0;  // 2 x u16 partial sums over 16 steps (entries <= 4095)
#pragma unroll
                    for (int wi = 0; wi < 4; wi++)
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int t = c * 16 + wi * 4 + b;
                            const uint32_t off = (uint32_t)((lane + t) & 31) << 3;
                            const uint32_t addr = prmt(w[wi], off, 0x7704u | ((uint32_t)b << 4));
                            const uint2 e = *reinterpret_cast<const uint2*>(lut + addr + g * PQS_GROUP_BYTES);
                            p01 += e.x;
                            p23 += e.y;
                        }
                    tot[0] += p01 & 0xffffu; tot[1] += p01 >> 16;
                    tot[2] += p23 & 0xffffu; tot[3] += p23 >> 16;
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; q++) key[q] = tot[q];
        }
        const uint32_t row = (tile << 5) + lane;
        if (tile < tile_end && row < a.n_rows) {
            bool any = false;
#pragma unroll
            for (int q = 0; q < NQ; q++) any = any || (q < nvalid && key[q] <= s_tau[q]);
            if (any) {
                bool ok = true;
                if (a.tomb != nullptr && row < a.tomb_bits && bit_set(a.tomb, row)) ok = false;
                if (ok && a.allow != nullptr && !bit_set(a.allow, row)) ok = false;
                if (ok) {
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        if (q < nvalid && key[q] <= s_tau[q]) {
                            const int pos = atomicAdd(&s_cnt[q], 1);
                            if (pos < a.cap) cand[(size_t)q * a.cap + pos] = ((uint64_t)key[q] << 32) | row;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CH; i++) cur[i] = nxt[i];
        __syncthreads();
        // compaction of any list that could overflow during the next trip (<= 512 appends per trip)
#pragma unroll 1
        for (int q = 0; q < NQ; q++) {
            const int c = min(s_cnt[q], a.cap);
            if (c > a.cap - PQS_THREADS) {  // block-uniform
                uint64_t* buf = cand + (size_t)q * a.cap;
                const int n2 = next_pow2(c);
                for (int t = c + tid; t < n2; t += PQS_THREADS) buf[t] = kInvalid;
                __syncthreads();
                block_bitonic_sort(buf, n2);
                if (tid == 0) {
                    s_cnt[q] = min(c, a.kc);
                    if (c >= a.kc) {
                        const uint32_t lt = (uint32_t)(buf[a.kc - 1] >> 32);
                        atomicMin(a.g_tau + qg * NQ + q, lt);
                        s_tau[q] = min(s_tau[q], lt);
                    }
                }
                __syncthreads();
            }
        }
        // fold in what the other CTAs of the query have published (stale values are safe: it only shrinks)
        if (tid < nvalid) s_tau[tid] = min(s_tau[tid], __ldcg(a.g_tau + qg * NQ + tid));
        __syncthreads();
    }
    __syncthreads();
    // emit: this CTA's best kc per query, restricted to entries at or below the query's shared bound, appended to
    // the query's compact global list (one global atomic per CTA and query; order is irrelevant, the merge sorts)
    __shared__ uint32_t s_pos, s_keep, s_w, s_gt;
#pragma unroll 1
    for (int q = 0; q < nvalid; q++) {
        const int gq = qg * NQ + q;
        int c = min(s_cnt[q], a.cap);
        uint64_t* buf = cand + (size_t)q * a.cap;
        if (c > a.kc) {  // only then does the order decide what is kept
            const int n2 = next_pow2(c);
            for (int t = c + tid; t < n2; t += PQS_THREADS) buf[t] = kInvalid;
            __syncthreads();
            block_bitonic_sort(buf, n2);
            c = a.kc;
            if (tid == 0) atomicMin(a.g_tau + gq, (uint32_t)(buf[a.kc - 1] >> 32));
        }
        if (tid == 0) { s_keep = 0; s_w = 0; s_gt = __ldcg(a.g_tau + gq); }  // ONE read: count and copy must agree
        __syncthreads();
        const uint32_t gt = s_gt;
        uint32_t keep = 0;
        for (int t = tid; t < c; t += PQS_THREADS) keep += ((uint32_t)(buf[t] >> 32) <= gt) ? 1u : 0u;
        if (keep) atomicAdd(&s_keep, keep);
        __syncthreads();
        if (tid == 0) s_pos = s_keep ? atomicAdd(a.out_cnt + gq, s_keep) : 0u;
        __syncthreads();
        uint64_t* out = a.compact + (size_t)gq * a.stride + s_pos;
        for (int t = tid; t < c; t += PQS_THREADS) {
            const uint64_t e = buf[t];
            if ((uint32_t)(e >> 32) <= gt) out[atomicAdd(&s_w, 1u)] = e;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Exact stage: the reference's sequential fp32 sum + sqrt for the kc coarse candidates of each query, sort by
// (distance, id), certification, first k_out packed (distance, row) entries out (ascending; kInvalid padding).
// grid = nq, block = 256.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adc_exact_kernel(const uint8_t* __restrict__ tiled, int M, int Mp, const float* __restrict__ luts,
                 const uint64_t* __restrict__ coarse, int kc, int k_out, const PqQParams* __restrict__ params,
                 uint64_t* __restrict__ out, uint32_t* __restrict__ cert_flags, uint32_t* __restrict__ cert_count) {
    __shared__ uint64_t keys[2048];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n2 = next_pow2(max(kc, 2));
    const float* lut = luts + (size_t)q * M * 256;
    for (int ci = tid; ci < n2; ci += blockDim.x) {
        uint64_t mine = kInvalid;
        if (ci < kc) {
            const uint64_t p = coarse[(size_t)q * kc + ci];
            if (p != kInvalid) {
                const uint32_t row = (uint32_t)p;
                float sum = 0.f;
                for (int j = 0; j < M; j++) sum = __fadd_rn(sum, __ldg(lut + j * 256 + tiled_code(tiled, Mp, row, j)));
                mine = pack_key(__fsqrt_rn(sum), row);
            }
        }
        keys[ci] = mine;
    }
    __syncthreads();
    block_bitonic_sort(keys, n2);
    if (tid == 0 && cert_flags != nullptr) {
        // every row outside the candidate set has an integer key >= the largest candidate key (the list holds the
        // kc smallest by (key, row)); its real sum is >= base + (key - M/2 - 1) / scale, its fp32 sequential sum
        // >= that * (1 - M 2^-24) (non-negative terms), and sqrt is monotone.  Certified when even that lower bound
        // lies strictly above the k_out-th exact distance.
        bool cert = true;
        const uint64_t last = coarse[(size_t)q * kc + kc - 1];  // merge output is sorted: the largest key, if full
        const uint64_t kth = (k_out - 1 < n2) ? keys[k_out - 1] : kInvalid;
        if (last != kInvalid && kth != kInvalid) {
            const PqQParams pr = params[q];
            const double klast = (double)(uint32_t)(last >> 32);
            double lb = pr.base + (klast - 0.5 * M - 1.0) * pr.inv_scale;
            lb -= fabs(pr.smax_sum) * (double)M * 6.0e-8 * 1.01;
            const double dk = (double)key_of(kth);
            cert = pr.inv_scale > 0.0 && lb > 0.0 && sqrt(lb) * (1.0 - 2.0e-7) > dk;
        }
        cert_flags[q] = cert ? 0u : 1u;
        if (!cert && cert_count != nullptr) atomicAdd(cert_count, 1u);
    }
    for (int t = tid; t < k_out; t += blockDim.x) out[(size_t)q * k_out + t] = (t < n2) ? keys[t] : kInvalid;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool adc_coarse_eligible(int M, int kc, int nq_per_pass) {
    if (M < 1 || M > 96) return false;
    if (nq_per_pass == 4) return kc <= 512;
    return kc <= 1024;
}

size_t adc_lutq_bytes(int M, int nq, int nq_per_pass) {
    const int G = (M + 31) / 32;
    const int groups = (nq + nq_per_pass - 1) / nq_per_pass;
    return (size_t)groups * G * PQS_GROUP_BYTES;
}

template <int NQ, int G>
static cudaError_t launch_coarse_t(const PqCoarseArgs& a, int qgroups, int parts, cudaStream_t st) {
    auto kern = adc_coarse_kernel<NQ, G>;
    LB_SMEM_OPTIN(kern);
    const size_t smem = (size_t)G * PQS_GROUP_BYTES + (size_t)NQ * a.cap * 8;
    dim3 grid(qgroups, parts);
    kern<<<grid, PQS_THREADS, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

// luts: fp32 [nq][M*256] (adc_lut_kernel).  Scratch supplied by the caller: lutq (adc_lutq_bytes), params
// (adc_params_bytes), compact [nq][stride >= parts * kc] u64, out_cnt [nq] zeroed, g_tau [nq] set to 0xff bytes.
cudaError_t launch_adc_coarse(const uint8_t* tiled, uint32_t n_rows, int M, const float* luts, int nq, int nq_per_pass,
                              const uint32_t* tomb, uint32_t tomb_bits, const uint32_t* allow, int kc, int parts,
                              uint32_t tiles_per_part, uint8_t* lutq, void* params, uint64_t* compact,
                              uint32_t* out_cnt, size_t stride, uint32_t* g_tau, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const int G = (M + 31) / 32;
    const int qgroups = (nq + nq_per_pass - 1) / nq_per_pass;
    if (nq_per_pass == 1) adc_quantise_kernel<1><<<qgroups, 256, 0, st>>>(luts, M, G, nq, lutq, (PqQParams*)params);
    else adc_quantise_kernel<4><<<qgroups, 256, 0, st>>>(luts, M, G, nq, lutq, (PqQParams*)params);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    PqCoarseArgs a;
    a.tiles = reinterpret_cast<const uint4*>(tiled); a.n_rows = n_rows; a.tiles_per_part = tiles_per_part;
    a.lutq = lutq; a.tomb = tomb; a.tomb_bits = tomb_bits; a.allow = allow;
    a.kc = kc; a.nq = nq; a.compact = compact; a.out_cnt = out_cnt; a.stride = stride; a.g_tau = g_tau;
    a.cap = (nq_per_pass == 1) ? 2048 : 1024;
    if (a.cap < next_pow2(kc + PQS_THREADS)) a.cap = next_pow2(kc + PQS_THREADS);
#define LB_PQC(NQ_, G_) return launch_coarse_t<NQ_, G_>(a, qgroups, parts, st)
    if (nq_per_pass == 1) {
        if (G == 1) LB_PQC(1, 1);
        if (G == 2) LB_PQC(1, 2);
        LB_PQC(1, 3);
    }
    if (G == 1) LB_PQC(4, 1);
    if (G == 2) LB_PQC(4, 2);
    LB_PQC(4, 3);
#undef LB_PQC
}

cudaError_t launch_adc_exact(const uint8_t* tiled, int M, const float* luts, const uint64_t* coarse, int nq, int kc,
                             int k_out, const void* params, uint64_t* out, uint32_t* cert_flags, uint32_t* cert_count,
                             cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    if (kc > 2048) return cudaErrorInvalidValue;
    const int Mp = ((M + 31) / 32) * 32;
    adc_exact_kernel<<<nq, 256, 0, st>>>(tiled, M, Mp, luts, coarse, kc, k_out, (const PqQParams*)params, out,
                                         cert_flags, cert_count);
    count_launch();
    return cudaGetLastError();
}

size_t adc_params_bytes(int nq) { return (size_t)nq * sizeof(PqQParams); }

}  // namespace lb
