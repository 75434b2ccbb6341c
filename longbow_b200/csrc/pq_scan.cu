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
#include <cstdio>

#include "kernels.cuh"

namespace lb {

bool g_pq_ring = false;  // lb_set_option("pq_ring"): single-query passes through the bulk-copy ring kernel.
                         // Measured 0.267 ms against 0.227 ms for the register kernel at C3: kept selectable, off by default
int g_pq_ahead = 2;  // lb_set_option("pq_ahead"): measured flat for 1..6 trips, worse at 12 (L2 thrash)
constexpr int PQS_THREADS = 640;   // 20 warps, one CTA per SM, <= 102 registers: room for two tiles of codes per lane
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

// grid = (query groups, QSPLIT): every block recomputes the (cheap) per-query statistics and fills its slice of
// the group's table, so a single query's table build is spread over QSPLIT SMs.
constexpr int PQS_QSPLIT = 16;

template <int NQ>
__global__ void __launch_bounds__(256)
adc_quantise_kernel(const float* __restrict__ luts, int M, int G, int nq, uint8_t* __restrict__ lutq,
                    PqQParams* __restrict__ params) {
    __shared__ float s_min[NQ][96];
    __shared__ double s_rng[NQ][96];
    __shared__ double s_scale[NQ];
    const int qg = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // statistics: one warp per (query, sub-quantiser), lanes stride the 256 codes (coalesced)
    for (int w = warp; w < NQ * M; w += 8) {
        const int qi = w / M, j = w - qi * M;
        const int q = qg * NQ + qi;
        float mn = INFINITY, mx = -INFINITY;
        if (q < nq) {
            const float* t = luts + ((size_t)q * M + j) * 256;
#pragma unroll
            for (int c = 0; c < 8; c++) { const float v = __ldg(t + c * 32 + lane); mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
        } else { mn = 0.f; mx = 0.f; }
        if (lane == 0) { s_min[qi][j] = mn; s_rng[qi][j] = (double)mx - (double)mn; }
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
        if (q < nq && blockIdx.y == 0) {
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
        for (int e = blockIdx.y * 256 + tid; e < total; e += gridDim.y * 256) {
            const int s = e & 63, c = (e >> 6) & 255, g = e >> 14;
            const int j = g * 32 + (s & 31);
            uint32_t v = 0;
            if (j < M && q < nq) {
                const double x = ((double)__ldg(luts + ((size_t)q * M + j) * 256 + c) - (double)s_min[0][j]) * scale;
                v = (uint32_t)llrint(x);
            }
            o32[e] = v;
        }
    } else {
        // [g][code][32 slots] of 4 x u16
        uint2* o64 = reinterpret_cast<uint2*>(out);
        const int total = G * 256 * 32;
        for (int e = blockIdx.y * 256 + tid; e < total; e += gridDim.y * 256) {
            const int s = e & 31, c = (e >> 5) & 255, g = e >> 13;
            const int j = g * 32 + s;
            uint32_t v[4] = {0, 0, 0, 0};
            if (j < M) {
#pragma unroll
                for (int qi = 0; qi < NQ; qi++) {
                    const int q = qg * NQ + qi;
                    if (q < nq) {
                        const double x = ((double)__ldg(luts + ((size_t)q * M + j) * 256 + c) - (double)s_min[qi][j]) * s_scale[qi];
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
    // Shared progressive threshold, two sources (both only ever shrink; stale reads are safe):
    //  g_tau [nq]: the smallest kc-th-best key any CTA of the query has published after compacting its list --
    //    that CTA holds kc live rows at or below it;
    //  g_min [nq][nmin]: the smallest live key every WARP of every CTA has seen so far (nmin = parts * 16 disjoint
    //    row sets).  The kc-th smallest of those minima is an upper bound of the global kc-th best key: kc
    //    different live rows lie at or below it.  With thousands of small row sets this bound sits within a few
    //    percent of the true kc-th best after a fraction of the scan, with no sorting at all.
    uint32_t* g_tau;          // initialised to 0xffffffff by the caller
    uint32_t* g_min;          // initialised to 0xffffffff by the caller
    int nmin;
    // [nq], zeroed by the caller: set when a candidate had to be dropped because a CTA's list was full between two
    // overflow checks (only degenerate tables can do that); the exact stage then reports the query uncertified
    uint32_t* overflow;
    int ahead;                // trips the L2 prefetch runs ahead of the register loads (single-query passes)
};

// kc-th smallest (rounded UP to a histogram bin edge: still a valid upper bound) of the set entries of v[0, n).
// Returns 0xffffffff when fewer than kc entries are set.  All threads of the block call; s_hist: int[258].
__device__ __forceinline__ uint32_t block_kth_upper_bound(const uint32_t* __restrict__ v, int n, int kc, int* s_hist,
                                                          int tid) {
    constexpr int E = (4096 + PQS_THREADS - 1) / PQS_THREADS;
    uint32_t mine[E];
    uint32_t lo = 0xffffffffu, hi = 0u;
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int i = tid + e * PQS_THREADS;
        mine[e] = (i < n) ? __ldcg(v + i) : 0xffffffffu;
        if (mine[e] != 0xffffffffu) { lo = min(lo, mine[e]); hi = max(hi, mine[e]); cnt++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (tid < 256) s_hist[tid] = 0;
    __shared__ uint32_t s_lo[PQS_WARPS], s_hi[PQS_WARPS];
    __shared__ int s_c[PQS_WARPS];
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; s_c[tid >> 5] = cnt; }
    __syncthreads();
    lo = 0xffffffffu; hi = 0u; cnt = 0;
#pragma unroll
    for (int w = 0; w < PQS_WARPS; w++) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); cnt += s_c[w]; }
    uint32_t result = 0xffffffffu;
    if (cnt >= kc) {  // block-uniform
        int shift = 0;
        while (((hi - lo) >> shift) >= 256u) shift++;
#pragma unroll
        for (int e = 0; e < E; e++)
            if (mine[e] != 0xffffffffu) atomicAdd(&s_hist[(mine[e] - lo) >> shift], 1);
        __syncthreads();
        if (tid == 0) {
            int cum = 0, b = 0;
            for (; b < 256; b++) { cum += s_hist[b]; if (cum >= kc) break; }
            const uint64_t edge = (uint64_t)lo + (((uint64_t)b + 1) << shift) - 1;
            s_hist[256] = (int)(uint32_t)min(edge, (uint64_t)hi);
        }
        __syncthreads();
        result = (uint32_t)s_hist[256];
    }
    __syncthreads();
    return result;
}

template <int NQ, int G>
__device__ __forceinline__ void adc_tile_keys(const uint4 (&cur)[2 * G], const unsigned char* lut, int lane, uint32_t cb,
                                              uint32_t (&key)[NQ]) {
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
}

template <int NQ, int G>
__global__ void __launch_bounds__(PQS_THREADS, 1)
adc_coarse_kernel(const PqCoarseArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* lut = smem_raw;                                             // G * 64 KB
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw + G * PQS_GROUP_BYTES);  // [NQ][cap]
    __shared__ int s_cnt[NQ];
    __shared__ uint32_t s_tau[NQ];
    __shared__ int s_hist[258];
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

    // lane constant for the byte permute: byte0 = 4 * lane (NQ = 1), bytes 1..3 = 1, 2, 0 (group selectors)
    const uint32_t cb = (uint32_t)(lane << 2) | (1u << 8) | (2u << 16);
    uint32_t lmin[NQ];  // smallest LIVE key this lane has seen, per query
#pragma unroll
    for (int q = 0; q < NQ; q++) lmin[q] = 0xffffffffu;

    // The hot loop is issue-bound (ncu: ~49 % issue-active at the pass's plateau, every other pipe lower), so
    // everything outside the 2.5 instructions per look-up is kept off it: ONE compare per row and query against an
    // exclusive bound thr1 = max(tau + 1, lmin) held in a register (key < thr1  <=>  key <= tau || key < lmin);
    // row index, range checks, bitmaps, the shared-memory threshold and the list append live in the rare path.
    uint32_t thr1[NQ];
    auto refresh_thr = [&]() {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const uint32_t t = s_tau[q];
            thr1[q] = (q < nvalid) ? max(t == 0xffffffffu ? t : t + 1u, lmin[q]) : 0u;
        }
    };
    refresh_thr();
    auto rare = [&](uint32_t tile, const uint32_t (&key)[NQ]) {
        const uint32_t row = (tile << 5) + lane;
        if (tile >= tile_end || row >= a.n_rows) return;
        bool ok = true;
        if (a.tomb != nullptr && row < a.tomb_bits && bit_set(a.tomb, row)) ok = false;
        if (ok && a.allow != nullptr && !bit_set(a.allow, row)) ok = false;
        if (!ok) return;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            if (q >= nvalid || key[q] >= thr1[q]) continue;
            lmin[q] = min(lmin[q], key[q]);
            if (key[q] <= s_tau[q]) {
                const int pos = atomicAdd(&s_cnt[q], 1);
                if (pos < a.cap) cand[(size_t)q * a.cap + pos] = ((uint64_t)key[q] << 32) | row;
                else a.overflow[qg * NQ + q] = 1u;
            }
        }
        refresh_thr();
    };
    auto consume = [&](const uint4 (&cur)[CH], uint32_t tile) {
        uint32_t key[NQ];
        adc_tile_keys<NQ, G>(cur, lut, lane, cb, key);
        bool any = false;
#pragma unroll
        for (int q = 0; q < NQ; q++) any = any || (key[q] < thr1[q]);
        if (any) rare(tile, key);
    };
    // tiles past the end are clamped to the last one (their keys never pass the range check in the rare path): no
    // predicates, no zero fill
    const uint32_t tile_last = tile_end - 1;  // only used when iters > 0
    auto prefetch = [&](uint4 (&dst)[CH], uint32_t tile) {
        const uint4* p = a.tiles + (size_t)min(tile, tile_last) * (CH * 32) + lane;
#pragma unroll
        for (int i = 0; i < CH; i++) dst[i] = __ldcs(p + i * 32);
    };
    auto housekeeping = [&](uint32_t it) {
        // Block barriers only on a doubling schedule (after trips 1, 2, 4, 8, ...): list compaction and threshold
        // refresh from the shared bounds, which improve with the logarithm of the rows scanned.  Between them the
        // warps run free -- a barrier every trip locks their load and look-up phases together, and then HBM latency
        // and shared-memory look-ups add up instead of overlapping (measured: 203k + 278k cycles per SM).  A list
        // that fills up between two barriers sets the overflow flag (degenerate tables only).
        // Every 8th trip is an overflow check only (compaction if a list is more than half full).
        const bool pow2 = ((it + 1) & it) == 0;
        const bool chk = ((it + 1) & 7) == 0;
        if (!pow2 && !chk) return;  // block-uniform
        __syncthreads();
#pragma unroll 1
        for (int q = 0; q < nvalid; q++) {
            const int c = min(s_cnt[q], a.cap);
            // scheduled refresh: keep the local kc-th best current (the sort is short, the list holds about kc entries
            // plus what arrived since the last refresh); overflow check: only a list that is half full
            if (pow2 ? (c > a.kc) : (c > a.cap / 2)) {  // block-uniform
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
        if (pow2) {
            // publish this warp's minima, then bound the query's kc-th best by the kc-th smallest published minimum
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                uint32_t m = lmin[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0 && q < nvalid && m != 0xffffffffu)
                    atomicMin(a.g_min + (size_t)(qg * NQ + q) * a.nmin + part * PQS_WARPS + warp, m);
            }
            __threadfence();
            __syncthreads();
#pragma unroll 1
            for (int q = 0; q < nvalid; q++) {
                const uint32_t b = block_kth_upper_bound(a.g_min + (size_t)(qg * NQ + q) * a.nmin, a.nmin, a.kc, s_hist, tid);
                if (tid == 0) s_tau[q] = min(min(s_tau[q], b), __ldcg(a.g_tau + qg * NQ + q));
            }
        }
        __syncthreads();
        refresh_thr();
    };

    // two trips per loop iteration with ping-pong register buffers: the next tile's codes are in flight while the
    // current tile's look-ups run
    // The register double buffer keeps one tile (3 KB) per warp in flight: 60 KB per SM, which at the loaded HBM
    // latency (~2.5 us) sustains only ~3.7 TB/s.  The rest of the latency is taken off the critical path by an L2
    // prefetch PQS_AHEAD trips ahead: one lane issues a bulk prefetch of the warp's future tile (no registers, no
    // shared memory), so the later LDG hits L2.
    auto l2_prefetch = [&](uint32_t tile) {
        if (NQ == 1 && lane == 0) {  // batched passes re-read the codes from L2 anyway
            const void* p = a.tiles + (size_t)min(tile, tile_last) * (CH * 32);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(CH * 32 * 16)) : "memory");
        }
    };
    const uint32_t PQS_AHEAD = (uint32_t)a.ahead;
    uint4 buf0[CH], buf1[CH];
    uint32_t it = 0;
    if (iters > 0) {
        for (uint32_t p = 1; p < PQS_AHEAD; p++) l2_prefetch(tile_begin + p * PQS_WARPS + warp);
        prefetch(buf0, tile_begin + warp);
    }
    // housekeeping after trips 0, 1, 3, 7 and then every 8th (one compare per trip on the hot path)
    uint32_t next_hk = 0;
    auto maybe_hk = [&](uint32_t t) {
        if (t == next_hk) {
            housekeeping(t);
            next_hk = (t < 7) ? 2 * t + 1 : t + 8;
        }
    };
#pragma unroll 1
    for (; it + 1 < iters; it += 2) {
        const uint32_t tile = tile_begin + it * PQS_WARPS + warp;
        l2_prefetch(tile + PQS_AHEAD * PQS_WARPS);
        l2_prefetch(tile + (PQS_AHEAD + 1) * PQS_WARPS);
        prefetch(buf1, tile + PQS_WARPS);
        consume(buf0, tile);
        maybe_hk(it);
        prefetch(buf0, tile + 2 * PQS_WARPS);
        consume(buf1, tile + PQS_WARPS);
        maybe_hk(it + 1);
    }
    if (it < iters) consume(buf0, tile_begin + it * PQS_WARPS + warp);
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
        if (tid == 0) { s_keep = 0; s_w = 0; s_gt = min(s_tau[q], __ldcg(a.g_tau + gq)); }  // ONE value for count and copy
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
// grid = nq, block = 256.  Candidates are processed 64 at a time: all threads first fetch the 64 x M table values
// (code byte, then LUT entry: two dependent loads, M independent pairs per candidate in flight), then thread c adds
// candidate c's values in j order out of shared memory (row pitch M + 1: conflict free).
// ---------------------------------------------------------------------------------------------
constexpr int PQX_CHUNK = 64;

__global__ void __launch_bounds__(256)
adc_exact_kernel(const uint8_t* __restrict__ tiled, int M, int Mp, const float* __restrict__ luts,
                 const uint64_t* __restrict__ coarse, int kc, int k_out, const PqQParams* __restrict__ params,
                 const uint32_t* __restrict__ overflow, uint64_t* __restrict__ out, uint32_t* __restrict__ cert_flags,
                 uint32_t* __restrict__ cert_count, const PqGemmCert g) {
    __shared__ uint64_t keys[1024];
    __shared__ float vals[PQX_CHUNK * 97];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n2 = next_pow2(max(kc, 2));
    const float* lut = luts + (size_t)q * M * 256;
    const int pitch = M + 1;
    for (int c0 = 0; c0 < n2; c0 += PQX_CHUNK) {
        for (int e = tid; e < PQX_CHUNK * M; e += blockDim.x) {
            const int c = e / M, j = e - c * M;
            const int ci = c0 + c;
            float v = 0.f;
            if (ci < kc) {
                const uint64_t p = coarse[(size_t)q * kc + ci];
                if (p != kInvalid) v = __ldg(lut + j * 256 + tiled_code(tiled, Mp, (uint32_t)p, j));
            }
            vals[c * pitch + j] = v;
        }
        __syncthreads();
        if (tid < PQX_CHUNK && c0 + tid < n2) {
            const int ci = c0 + tid;
            uint64_t mine = kInvalid;
            if (ci < kc) {
                const uint64_t p = coarse[(size_t)q * kc + ci];
                if (p != kInvalid) {
                    float sum = 0.f;
                    for (int j = 0; j < M; j++) sum = __fadd_rn(sum, vals[tid * pitch + j]);
                    mine = pack_key(__fsqrt_rn(sum), (uint32_t)p);
                }
            }
            keys[ci] = mine;
        }
        __syncthreads();
    }
    block_bitonic_sort(keys, n2);
    if (tid == 0 && cert_flags != nullptr) {
        // every row outside the candidate set has an integer key >= the largest candidate key (the list holds the
        // kc smallest by (key, row)); its real sum is >= base + (key - M/2 - 1) / scale, its fp32 sequential sum
        // >= that * (1 - M 2^-24) (non-negative terms), and sqrt is monotone.  Certified when even that lower bound
        // lies strictly above the k_out-th exact distance.
        bool cert = true;
        const uint64_t last = coarse[(size_t)q * kc + kc - 1];  // merge output is sorted: the largest key, if full
        const uint64_t kth = (k_out - 1 < n2) ? keys[k_out - 1] : kInvalid;
        if (g.qn != nullptr) {
            // Coarse stage = tensor-core L2 keys over fp16-decoded rows (pq_gemm.cu).  The candidate list is the kc
            // smallest (key, row) in no particular order; if it is full, every other live row has a key >= its
            // largest one.  For such a row: |q_h - x_h|^2 >= |q_h|^2 + key - 2 beta |q_h||x_h| (truncating fp32
            // accumulation of the dot product), the unrounded distance is at least the rounded one minus
            // 2^-11 (|q| + |x|) + sqrt(dims) 2^-24, and the reference's fp32 table sum is within 1e-5 of it.
            uint32_t kmax = 0;
            bool full = true;
            for (int i = 0; i < kc; i++) {
                const uint64_t p = coarse[(size_t)q * kc + i];
                if (p == kInvalid) full = false;
                else kmax = max(kmax, (uint32_t)(p >> 32));
            }
            if (full && kth != kInvalid) {
                const float2 qq = g.qn[q];
                const double X = sqrt((double)__uint_as_float(*g.xmax2)) * 1.001;
                const double d2 = (double)qq.x + (double)ordered_to_float(kmax) - 2.0 * g.beta * sqrt((double)qq.x) * X;
                double lb = d2 > 0.0 ? sqrt(d2) : 0.0;
                lb -= 4.8828125e-4 * ((double)qq.y + X) + sqrt((double)g.dims) * 1.2e-7;
                cert = lb > 0.0 && lb * (1.0 - 1.0e-5) > (double)key_of(kth);
            }
        } else
        if (last != kInvalid && kth != kInvalid) {
            const PqQParams pr = params[q];
            const double klast = (double)(uint32_t)(last >> 32);
            double lb = pr.base + (klast - 0.5 * M - 1.0) * pr.inv_scale;
            lb -= fabs(pr.smax_sum) * (double)M * 6.0e-8 * 1.01;
            const double dk = (double)key_of(kth);
            cert = pr.inv_scale > 0.0 && lb > 0.0 && sqrt(lb) * (1.0 - 2.0e-7) > dk;
        }
#ifdef LB_PQ_DEBUG
        if (!cert || (overflow != nullptr && overflow[q])) {
            const PqQParams pr = params[q];
            printf("q %d uncert: ovf %u last %llx kth %llx klast %u base %f inv %g smax %f dk %f n2 %d\n", q, overflow ? overflow[q] : 0u,
                   (unsigned long long)last, (unsigned long long)kth, (uint32_t)(last >> 32), pr.base, pr.inv_scale, pr.smax_sum,
                   (double)key_of(kth), n2);
        }
#endif
        if (overflow != nullptr && overflow[q]) cert = false;
        cert_flags[q] = cert ? 0u : 1u;
        if (!cert && cert_count != nullptr) atomicAdd(cert_count, 1u);
    }
    for (int t = tid; t < k_out; t += blockDim.x) out[(size_t)q * k_out + t] = (t < n2) ? keys[t] : kInvalid;
}

// ---------------------------------------------------------------------------------------------
// Single-query coarse scan with the code tiles staged through a shared-memory RING fed by bulk copies
// (cp.async.bulk + mbarrier), warp-specialised: one producer warp keeps up to NSLOT tiles (3 KB each) in flight,
// sixteen consumer warps take them round-robin.  The register double buffer of adc_coarse_kernel keeps one tile per
// warp in flight -- 60 KB per SM, which the loaded HBM latency turns into ~3.8 TB/s; the ring decouples the bytes
// in flight (~80 KB per SM, no registers) from the look-up work.  To make room the table is stored ONCE
// (adc_coarse_kernel stores every line twice): lines of 256 bytes hold two groups of 32 sub-quantisers side by side,
// and the slot (lane + step) mod 32 costs one extra integer instruction per look-up (there is issue headroom).
// Per tile the shared-memory pipe moves 96 look-up wavefronts + 24 (codes out of the ring) + 24 (the bulk copy's
// write) = 144, against 135 cycles of HBM time per tile and SM: the pass is balanced at ~0.16 ms.
// ---------------------------------------------------------------------------------------------
constexpr int RING_CONSUMERS = 16;
constexpr int RING_PRODUCERS = 4;  // producer warps (one issuing lane each)
constexpr int RING_THREADS = (RING_CONSUMERS + RING_PRODUCERS) * 32;
constexpr int RING_CT = RING_CONSUMERS * 32;  // consumer threads (named barrier 1)

__device__ __forceinline__ uint32_t ring_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_csync() { asm volatile("bar.sync 1, %0;" ::"n"(RING_CT) : "memory"); }
__device__ __forceinline__ void ring_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ring_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a lost arrival traps after ~2 s instead of hanging the GPU
__device__ __forceinline__ void ring_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; !ok; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(100000u) : "memory");
        if (!ok && (spins & 63u) == 63u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void ring_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// consumer-only bitonic sort (named barrier, RING_CT threads)
__device__ __forceinline__ void ring_bitonic_sort(uint64_t* a, int n, int tid) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n >> 1); t += RING_CT) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            ring_csync();
        }
    }
}

// kc-th smallest (rounded up to a histogram bin edge) of the set entries of v[0, n), consumer threads only
__device__ __forceinline__ uint32_t ring_kth_upper_bound(const uint32_t* __restrict__ v, int n, int kc, int* s_hist,
                                                         uint32_t* s_lo, uint32_t* s_hi, int* s_c, int tid) {
    constexpr int E = (4096 + RING_CT - 1) / RING_CT;
    uint32_t mine[E];
    uint32_t lo = 0xffffffffu, hi = 0u;
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int i = tid + e * RING_CT;
        mine[e] = (i < n) ? __ldcg(v + i) : 0xffffffffu;
        if (mine[e] != 0xffffffffu) { lo = min(lo, mine[e]); hi = max(hi, mine[e]); cnt++; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (tid < 256) s_hist[tid] = 0;
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; s_c[tid >> 5] = cnt; }
    ring_csync();
    lo = 0xffffffffu; hi = 0u; cnt = 0;
#pragma unroll
    for (int w = 0; w < RING_CONSUMERS; w++) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); cnt += s_c[w]; }
    uint32_t result = 0xffffffffu;
    if (cnt >= kc) {  // uniform
        int shift = 0;
        while (((hi - lo) >> shift) >= 256u) shift++;
#pragma unroll
        for (int e = 0; e < E; e++)
            if (mine[e] != 0xffffffffu) atomicAdd(&s_hist[(mine[e] - lo) >> shift], 1);
        ring_csync();
        if (tid == 0) {
            int cum = 0, b = 0;
            for (; b < 256; b++) { cum += s_hist[b]; if (cum >= kc) break; }
            const uint64_t edge = (uint64_t)lo + (((uint64_t)b + 1) << shift) - 1;
            s_hist[256] = (int)(uint32_t)min(edge, (uint64_t)hi);
        }
        ring_csync();
        result = (uint32_t)s_hist[256];
    }
    ring_csync();
    return result;
}

template <int G>
__global__ void __launch_bounds__(RING_THREADS, 1)
adc_coarse_ring_kernel(const PqCoarseArgs a, int nslot) {
    constexpr int CH = 2 * G;
    constexpr int NT = (G + 1) / 2;                 // 64 KB tables of two groups each
    constexpr uint32_t TILE_BYTES = CH * 32 * 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* lut = smem_raw;                                                   // NT * 64 KB
    unsigned char* ring = smem_raw + NT * PQS_GROUP_BYTES;                           // nslot * TILE_BYTES
    uint64_t* cand = reinterpret_cast<uint64_t*>(ring + (size_t)nslot * TILE_BYTES);  // [cap]
    uint64_t* bars = cand + a.cap;                                                   // full[nslot], empty[nslot]
    __shared__ int s_cnt;
    __shared__ uint32_t s_tau;
    __shared__ int s_hist[258];
    __shared__ uint32_t s_lo[RING_CONSUMERS], s_hi[RING_CONSUMERS];
    __shared__ int s_c[RING_CONSUMERS];
    __shared__ uint32_t s_pos, s_keep, s_w, s_gt;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = blockIdx.x, part = blockIdx.y;
    const uint32_t bar0 = ring_smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (nslot + s); };
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.lutq + (size_t)q * NT * PQS_GROUP_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(lut);
        for (int i = tid; i < NT * (PQS_GROUP_BYTES / 16); i += RING_THREADS) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        s_cnt = 0;
        s_tau = __ldcg(a.g_tau + q);
        for (int s = 0; s < nslot; s++) { ring_mbar_init(full_bar(s), 1); ring_mbar_init(empty_bar(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only full-block barrier: producer and consumers part ways here

    const uint32_t n_tiles = (a.n_rows + 31) >> 5;
    const uint32_t tile_begin = (uint32_t)part * a.tiles_per_part;
    const uint32_t tile_end = min(n_tiles, tile_begin + a.tiles_per_part);
    const uint32_t n_my = tile_end > tile_begin ? tile_end - tile_begin : 0;

    if (warp >= RING_CONSUMERS) {
        // ================================ producers ================================
        // Four producer warps, one issuing lane each, take the tiles seq = p, p + 4, ...: a single thread cannot
        // issue a 3 KB copy every 70 ns (measured: one issuing lane halves the pass rate), and lanes of ONE warp
        // waiting on different barriers stall each other.  Two tiles that share a slot are a whole round apart and
        // ordered by the slot's empty barrier, whichever warp issues them.
        if (lane == 0) {
            for (uint32_t seq = warp - RING_CONSUMERS; seq < n_my; seq += RING_PRODUCERS) {
                const int slot = (int)(seq % (uint32_t)nslot);
                const uint32_t round = seq / (uint32_t)nslot;
                ring_mbar_wait(empty_bar(slot), (round & 1u) ^ 1u);
                ring_mbar_expect_tx(full_bar(slot), TILE_BYTES);
                ring_bulk_load(ring_smem_u32(ring + (size_t)slot * TILE_BYTES),
                               a.tiles + (size_t)(tile_begin + seq) * (CH * 32), TILE_BYTES, full_bar(slot));
            }
        }
        return;
    }
    // ================================ consumers ================================
    const uint32_t iters = (n_my + RING_CONSUMERS - 1) / RING_CONSUMERS;
    uint32_t lmin = 0xffffffffu;  // smallest LIVE key this lane has seen

    auto housekeeping = [&](uint32_t it) {
        const bool pow2 = ((it + 1) & it) == 0;
        const bool chk = ((it + 1) & 7) == 0;
        if (!pow2 && !chk) return;
        ring_csync();
        {
            const int c = min(s_cnt, a.cap);
            if (pow2 ? (c > a.kc) : (c > a.cap / 2)) {
                const int n2 = next_pow2(c);
                for (int t = c + tid; t < n2; t += RING_CT) cand[t] = kInvalid;
                ring_csync();
                ring_bitonic_sort(cand, n2, tid);
                if (tid == 0) {
                    s_cnt = min(c, a.kc);
                    if (c >= a.kc) {
                        const uint32_t lt = (uint32_t)(cand[a.kc - 1] >> 32);
                        atomicMin(a.g_tau + q, lt);
                        s_tau = min(s_tau, lt);
                    }
                }
                ring_csync();
            }
        }
        if (pow2) {
            uint32_t m = lmin;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0 && m != 0xffffffffu) atomicMin(a.g_min + (size_t)q * a.nmin + part * RING_CONSUMERS + warp, m);
            __threadfence();
            ring_csync();
            const uint32_t b = ring_kth_upper_bound(a.g_min + (size_t)q * a.nmin, a.nmin, a.kc, s_hist, s_lo, s_hi, s_c, tid);
            if (tid == 0) s_tau = min(min(s_tau, b), __ldcg(a.g_tau + q));
        }
        ring_csync();
    };

#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
        const uint32_t seq = it * RING_CONSUMERS + warp;
        if (seq < n_my) {
            const int slot = (int)(seq % (uint32_t)nslot);
            const uint32_t round = seq / (uint32_t)nslot;
            ring_mbar_wait(full_bar(slot), round & 1u);
            const uint4* tp = reinterpret_cast<const uint4*>(ring + (size_t)slot * TILE_BYTES) + lane;
            uint4 cur[CH];
#pragma unroll
            for (int i = 0; i < CH; i++) cur[i] = tp[i * 32];
            __syncwarp();
            if (lane == 0) ring_mbar_arrive(empty_bar(slot));  // the tile is in registers: the slot may be refilled
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int wi = 0; wi < 4; wi++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int t = c * 16 + wi * 4 + b;
                        const uint32_t off0 = (uint32_t)((lane + t) & 31) << 2;   // slot (lane + t) mod 32 of the line
                        const uint32_t off1 = off0 | 128u;                          // ... of the line's second group
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            const uint4 v = cur[g * 2 + c];
                            const uint32_t w = (wi == 0) ? v.x : (wi == 1) ? v.y : (wi == 2) ? v.z : v.w;
                            const uint32_t addr = prmt(w, (g & 1) ? off1 : off0, 0x7704u | ((uint32_t)b << 4));
                            acc += *reinterpret_cast<const uint32_t*>(lut + addr + (g >> 1) * PQS_GROUP_BYTES);
                        }
                    }
            const uint32_t tile = tile_begin + seq;
            const uint32_t row = (tile << 5) + lane;
            if (row < a.n_rows && (acc <= s_tau || acc < lmin)) {
                bool ok = true;
                if (a.tomb != nullptr && row < a.tomb_bits && bit_set(a.tomb, row)) ok = false;
                if (ok && a.allow != nullptr && !bit_set(a.allow, row)) ok = false;
                if (ok) {
                    lmin = min(lmin, acc);
                    if (acc <= s_tau) {
                        const int pos = atomicAdd(&s_cnt, 1);
                        if (pos < a.cap) cand[pos] = ((uint64_t)acc << 32) | row;
                        else a.overflow[q] = 1u;
                    }
                }
            }
        }
        housekeeping(it);
    }
    ring_csync();
    // emit (as adc_coarse_kernel)
    {
        int c = min(s_cnt, a.cap);
        if (c > a.kc) {
            const int n2 = next_pow2(c);
            for (int t = c + tid; t < n2; t += RING_CT) cand[t] = kInvalid;
            ring_csync();
            ring_bitonic_sort(cand, n2, tid);
            c = a.kc;
            if (tid == 0) atomicMin(a.g_tau + q, (uint32_t)(cand[a.kc - 1] >> 32));
        }
        if (tid == 0) { s_keep = 0; s_w = 0; s_gt = min(s_tau, __ldcg(a.g_tau + q)); }
        ring_csync();
        const uint32_t gt = s_gt;
        uint32_t keep = 0;
        for (int t = tid; t < c; t += RING_CT) keep += ((uint32_t)(cand[t] >> 32) <= gt) ? 1u : 0u;
        if (keep) atomicAdd(&s_keep, keep);
        ring_csync();
        if (tid == 0) s_pos = s_keep ? atomicAdd(a.out_cnt + q, s_keep) : 0u;
        ring_csync();
        uint64_t* out = a.compact + (size_t)q * a.stride + s_pos;
        for (int t = tid; t < c; t += RING_CT) {
            const uint64_t e = cand[t];
            if ((uint32_t)(e >> 32) <= gt) out[atomicAdd(&s_w, 1u)] = e;
        }
    }
}

// table layout of adc_coarse_ring_kernel: [table = group / 2][code][(group & 1) * 32 + slot] u32
__global__ void __launch_bounds__(256)
adc_quantise_paired_kernel(const float* __restrict__ luts, int M, int G, int nq, uint8_t* __restrict__ lutq,
                           PqQParams* __restrict__ params) {
    __shared__ float s_min[96];
    __shared__ double s_rng[96];
    __shared__ double s_scale;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = warp; j < M; j += 8) {
        float mn = INFINITY, mx = -INFINITY;
        const float* t = luts + ((size_t)q * M + j) * 256;
#pragma unroll
        for (int c = 0; c < 8; c++) { const float v = __ldg(t + c * 32 + lane); mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { s_min[j] = mn; s_rng[j] = (double)mx - (double)mn; }
    }
    __syncthreads();
    if (tid == 0) {
        double base = 0, R = 0, top = 0;
        for (int j = 0; j < M; j++) { base += (double)s_min[j]; R = fmax(R, s_rng[j]); top += (double)s_min[j] + s_rng[j]; }
        const double smax = floor(4294960000.0 / (double)M);
        const double scale = (R > 0 && isfinite(R)) ? smax / R : 0.0;
        s_scale = scale;
        if (blockIdx.y == 0) {
            params[q].base = base;
            params[q].inv_scale = scale > 0 ? 1.0 / scale : 0.0;
            params[q].smax_sum = top;
        }
    }
    __syncthreads();
    const int NT = (G + 1) / 2;
    uint32_t* o32 = reinterpret_cast<uint32_t*>(lutq + (size_t)q * NT * PQS_GROUP_BYTES);
    const int total = NT * 256 * 64;
    const double scale = s_scale;
    for (int e = blockIdx.y * 256 + tid; e < total; e += gridDim.y * 256) {
        const int slot = e & 63, c = (e >> 6) & 255, T = e >> 14;
        const int g = 2 * T + (slot >> 5);
        const int j = g * 32 + (slot & 31);
        uint32_t v = 0;
        if (g < G && j < M) {
            const double x = ((double)__ldg(luts + ((size_t)q * M + j) * 256 + c) - (double)s_min[j]) * scale;
            v = (uint32_t)llrint(x);
        }
        o32[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool adc_coarse_eligible(int M, int kc, int nq_per_pass) {
    if (M < 1 || M > 96) return false;
    if (nq_per_pass == 4) return kc <= 256;
    return kc <= 768;
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
// (adc_params_bytes), compact [nq][stride >= parts * kc] u64, out_cnt [nq] + overflow [nq] zeroed, g_tau [nq] and
// g_min [nq][adc_min_slots(parts)] set to 0xff bytes.
cudaError_t launch_adc_coarse(const uint8_t* tiled, uint32_t n_rows, int M, const float* luts, int nq, int nq_per_pass,
                              const uint32_t* tomb, uint32_t tomb_bits, const uint32_t* allow, int kc, int parts,
                              uint32_t tiles_per_part, uint8_t* lutq, void* params, uint64_t* compact,
                              uint32_t* out_cnt, size_t stride, uint32_t* g_tau, uint32_t* g_min, uint32_t* overflow,
                              cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const int G = (M + 31) / 32;
    const int qgroups = (nq + nq_per_pass - 1) / nq_per_pass;
    if (nq_per_pass == 1 && g_pq_ring) {
        // ring kernel: paired table layout, sixteen consumer warps
        const int NT = (G + 1) / 2;
        const dim3 qg(nq, nq >= 64 ? 4 : PQS_QSPLIT);
        adc_quantise_paired_kernel<<<qg, 256, 0, st>>>(luts, M, G, nq, lutq, (PqQParams*)params);
        count_launch();
        cudaError_t e0 = cudaGetLastError();
        if (e0 != cudaSuccess) return e0;
        PqCoarseArgs a;
        a.tiles = reinterpret_cast<const uint4*>(tiled); a.n_rows = n_rows; a.tiles_per_part = tiles_per_part;
        a.lutq = lutq; a.tomb = tomb; a.tomb_bits = tomb_bits; a.allow = allow;
        a.kc = kc; a.nq = nq; a.compact = compact; a.out_cnt = out_cnt; a.stride = stride; a.g_tau = g_tau;
        a.g_min = g_min; a.nmin = parts * RING_CONSUMERS; a.overflow = overflow; a.ahead = 0;
        a.cap = 2048;
        const size_t tile_bytes = (size_t)G * 1024;
        const size_t fixed = (size_t)NT * PQS_GROUP_BYTES + (size_t)a.cap * 8;
        int nslot = (int)((232448 - 4096 - fixed) / (tile_bytes + 16));
        if (nslot > 48) nslot = 48;
        if (nslot < 4) return cudaErrorInvalidValue;
        const size_t smem = fixed + (size_t)nslot * (tile_bytes + 16);
        dim3 grid(nq, parts);
#define LB_RING(G_)                                                                   \
        {                                                                             \
            auto kern = adc_coarse_ring_kernel<G_>;                                   \
            LB_SMEM_OPTIN(kern);                                                      \
            kern<<<grid, RING_THREADS, smem, st>>>(a, nslot);                         \
        }
        if (G == 1) LB_RING(1) else if (G == 2) LB_RING(2) else LB_RING(3)
#undef LB_RING
        count_launch();
        return cudaGetLastError();
    }
    const dim3 qgrid(qgroups, qgroups >= 64 ? 4 : PQS_QSPLIT);
    if (nq_per_pass == 1) adc_quantise_kernel<1><<<qgrid, 256, 0, st>>>(luts, M, G, nq, lutq, (PqQParams*)params);
    else adc_quantise_kernel<4><<<qgrid, 256, 0, st>>>(luts, M, G, nq, lutq, (PqQParams*)params);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    PqCoarseArgs a;
    a.tiles = reinterpret_cast<const uint4*>(tiled); a.n_rows = n_rows; a.tiles_per_part = tiles_per_part;
    a.lutq = lutq; a.tomb = tomb; a.tomb_bits = tomb_bits; a.allow = allow;
    a.kc = kc; a.nq = nq; a.compact = compact; a.out_cnt = out_cnt; a.stride = stride; a.g_tau = g_tau;
    a.g_min = g_min; a.nmin = parts * PQS_WARPS; a.overflow = overflow; a.ahead = g_pq_ahead;
    a.cap = (nq_per_pass == 1) ? 2048 : 1024;
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
                             int k_out, const void* params, const uint32_t* overflow, uint64_t* out,
                             uint32_t* cert_flags, uint32_t* cert_count, cudaStream_t st, const PqGemmCert* gemm) {
    if (nq <= 0) return cudaSuccess;
    if (kc > 1024 || M > 96) return cudaErrorInvalidValue;
    const int Mp = ((M + 31) / 32) * 32;
    PqGemmCert g = {};
    if (gemm) g = *gemm;
    adc_exact_kernel<<<nq, 256, 0, st>>>(tiled, M, Mp, luts, coarse, kc, k_out, (const PqQParams*)params, overflow, out,
                                         cert_flags, cert_count, g);
    count_launch();
    return cudaGetLastError();
}

size_t adc_params_bytes(int nq) { return (size_t)nq * sizeof(PqQParams); }
int adc_min_slots(int parts) { return parts * (PQS_WARPS > 16 ? PQS_WARPS : 16); }  // one published minimum per warp of every CTA

}  // namespace lb
