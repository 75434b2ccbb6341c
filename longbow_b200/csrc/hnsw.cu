// hnsw.cu -- batched HNSW layer search on the GPU: ArrowHNSW.searchLayer
// (internal/store/arrow_hnsw.go:1108-1385) for thousands of concurrent queries over the adjacency arrays of
// GraphData (internal/store/types/graph_data.go:605-670), followed by the in-kernel-bitmap re-rank
// (lb_index_rerank) -- BASELINE config 5 end to end on the device.
//
// One warp per query, the reference's algorithm step for step, so the frontier it returns is identical to the CPU
// walk's (oracle/lb_oracle.c lbo_hnsw_search_layer), not merely "as good":
//   * candidate min-heap and result max-heap live in shared memory as packed (ordered distance << 32 | id) words
//     -- one u64 compare is the (distance, id) order -- and are operated by lane 0;
//   * an expansion loads the popped node's neighbour list with one coalesced 128-byte read per 32 neighbours,
//     every lane claims its neighbour in the query's visited set (open-addressing hash table in global memory,
//     atomicCAS: insert == "was not visited"), lanes holding a new node compute its distance with the exact
//     reference arithmetic (common.cuh ExactAcc: 4 lanes, no FMA, sqrt via double), each lane streaming its own
//     row with eight 16-byte loads in flight;
//   * lane 0 then replays the neighbours IN LIST ORDER through the reference's acceptance test
//     (len < ef || d < worst, strict) -- the order matters because every accepted neighbour can move the worst.
// The graph itself (construction, upper-layer descent to the entry point) stays on the host.
#include <algorithm>
#include <cstdio>
#include <new>

#pragma GCC visibility push(default)
#include "../../include/longbow_b200.h"
#pragma GCC visibility pop
#include "kernels.cuh"

namespace lb {

int api_fail(int code, const char* what);
int api_fail_cuda(cudaError_t e, const char* where);
int api_use_device(int device);
int api_index_view(const lb_index* idx, IndexView* out);  // api.cu
int api_rerank_device(lb_index* idx, const void* d_q, int64_t nq, const uint32_t* d_ids, int c, int k,
                      const uint64_t* d_allow, float* d_dist, int64_t* d_lab, cudaStream_t st);

bool g_hnsw_coop = true;  // lb_set_option("hnsw_coop")
constexpr int HN_WARPS_MAX = 8;        // queries (warps) per CTA: 4..8, chosen at launch for the most resident warps
constexpr uint32_t HN_EMPTY = 0xffffffffu;
constexpr int HN_SR = 48;              // staging row-buffers per warp (trips of 16 rows, 3 chunks in flight)
constexpr int HN_SR_QUAD = 24;         // fp32 rows, quad form: trips of 8 rows, 3 chunks in flight
template <typename T> __host__ __device__ constexpr int hn_stage_rows() { return sizeof(T) == 4 ? HN_SR_QUAD : HN_SR; }

struct HnswArgs {
    const void* db;
    uint32_t n_rows;
    int dim;
    const uint32_t* neighbors;  // [n][max_degree]
    const int32_t* counts;      // [n] or null
    int max_degree;
    const void* queries;        // [nq][dim]
    int nq;
    const uint32_t* entries;    // [nq]
    int ef, cand_cap;
    uint32_t* visited;          // [nq][ht_size], HN_EMPTY-initialised
    uint32_t ht_mask;           // ht_size - 1
    uint32_t* out_ids;          // [nq][ef]
    float* out_d;               // [nq][ef]
    uint32_t* out_visited;      // [nq] or null
    uint32_t* fail_count;       // [1]: queries that hit a table / heap limit (results invalid for those)
    uint32_t* fail_flags;       // [nq] or null
    int coop;                   // rows are 16-byte aligned multiples of 16 bytes: cooperative gather
};

// ---- binary heaps of packed u64 in shared memory, lane 0 only
__device__ __forceinline__ void heap_push(uint64_t* h, int& n, uint64_t v, bool maxheap) {
    int i = n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        const uint64_t hp = h[p];
        if (maxheap ? (hp >= v) : (hp <= v)) break;
        h[i] = hp;
        i = p;
    }
    h[i] = v;
}
__device__ __forceinline__ uint64_t heap_pop(uint64_t* h, int& n, bool maxheap) {
    const uint64_t top = h[0];
    const uint64_t v = h[--n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        uint64_t hc = h[c];
        if (c + 1 < n) {
            const uint64_t hr = h[c + 1];
            if (maxheap ? (hr > hc) : (hr < hc)) { c++; hc = hr; }
        }
        if (maxheap ? (hc <= v) : (hc >= v)) break;
        h[i] = hc;
        i = c;
    }
    if (n > 0) h[i] = v;
    return top;
}

template <typename T, int METRIC>
__global__ void __launch_bounds__(HN_WARPS_MAX * 32)
hnsw_search_layer_kernel(const HnswArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * (int)(blockDim.x >> 5) + warp;
    // per-warp regions: query (fp32) | result heap [ef + 1] | candidate heap [cand_cap]
    const size_t q_bytes = ((size_t)a.dim * 4 + 15) & ~(size_t)15;
    const size_t heap_bytes = (((size_t)(a.ef + 1 + a.cand_cap) * 8) + 15) & ~(size_t)15;  // cp.async needs 16-byte slots
    const size_t per_warp = q_bytes + heap_bytes + (a.coop ? (size_t)hn_stage_rows<T>() * RC_PITCH : 0);
    unsigned char* base = smem_raw + (size_t)warp * per_warp;
    unsigned char* stage = base + q_bytes + heap_bytes;  // [hn_stage_rows][RC_PITCH] when coop
    float* qf = reinterpret_cast<float*>(base);
    uint64_t* res = reinterpret_cast<uint64_t*>(base + q_bytes);
    uint64_t* cand = res + (a.ef + 1);
    if (q >= a.nq) return;
    const T* qrow = reinterpret_cast<const T*>(a.queries) + (size_t)q * a.dim;
    for (int i = lane; i < a.dim; i += 32) qf[i] = Elem<T>::widen(qrow[i]);
    __syncwarp();
    const T* db = reinterpret_cast<const T*>(a.db);
    const bool vec_ok = (a.dim % Elem<T>::kVec == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    uint32_t* ht = a.visited + (size_t)q * (a.ht_mask + 1);
    auto dist_of = [&](uint32_t id) -> float {
        float d = exact_pair<T, METRIC>(qf, db + (size_t)id * a.dim, a.dim, vec_ok);
        if (METRIC == METRIC_DOT) d = -d;
        return __fadd_rn(d, 0.f);  // -0 -> +0: equal distances must compare equal in the packed order
    };
    // claim `id` in the visited set; true = it was not there
    auto visit = [&](uint32_t id) -> int {
        uint32_t slot = (id * 2654435761u) & a.ht_mask;
        for (uint32_t probe = 0; probe <= a.ht_mask; probe++) {
            const uint32_t old = atomicCAS(ht + slot, HN_EMPTY, id);
            if (old == HN_EMPTY) return 1;
            if (old == id) return 0;
            slot = (slot + 1) & a.ht_mask;
        }
        return -1;  // table full
    };

    int nr = 0, nc = 0;  // heap sizes (lane 0's copies are authoritative; kept uniform via shuffles below)
    uint32_t n_vis = 0;
    bool failed = false;
    const uint32_t entry = a.entries[q];
    if (entry < a.n_rows) {
        float ed = 0.f;
        if (lane == 0) {
            visit(entry);
            ed = dist_of(entry);
            const uint64_t p = pack_key(ed, entry);
            heap_push(cand, nc, p, false);
            heap_push(res, nr, p, true);
        }
        n_vis = 1;
    }
    nr = __shfl_sync(0xffffffffu, nr, 0);
    nc = __shfl_sync(0xffffffffu, nc, 0);
    const uint32_t vis_limit = (a.ht_mask + 1) - ((a.ht_mask + 1) >> 2);  // 75 % load
    while (nc > 0) {
        uint64_t cur = 0;
        int stop = 0;
        if (lane == 0) {
            cur = heap_pop(cand, nc, false);
            if (nr > 0 && nr >= a.ef && key_of(cur) > key_of(res[0])) stop = 1;  // arrow_hnsw.go:1322-1329
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        if (stop) break;
        cur = __shfl_sync(0xffffffffu, cur, 0);
        const uint32_t cid = id_of(cur);
        int cnt = a.counts ? a.counts[cid] : a.max_degree;
        cnt = min(max(cnt, 0), a.max_degree);
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const int i = b0 + lane;
            uint32_t nb = HN_EMPTY;
            if (i < cnt) nb = __ldg(a.neighbors + (size_t)cid * a.max_degree + i);
            // a duplicate inside one list is "already visited" for every occurrence but the first (the reference
            // marks sequentially): a lane defers to any lower lane holding the same id
            const bool dup = (__match_any_sync(0xffffffffu, nb) & ((1u << lane) - 1u)) != 0u;
            int fresh = 0;
            if (nb < a.n_rows && !dup) fresh = visit(nb);
            if (fresh < 0) { failed = true; fresh = 0; }
            const unsigned fm = __ballot_sync(0xffffffffu, fresh != 0);
            const int nfresh = __popc(fm);
            n_vis += nfresh;
            // Distances of the new nodes.  Aligned rows: the new ids are compacted to lanes 0..nfresh-1 (list order
            // kept) and the warp gathers their rows together through shared memory -- 128-byte chunks copied with
            // cp.async, several chunks in flight, lane i accumulating node i in the reference order (common.cuh
            // rc_gather, the re-rank's gather): 3-5 memory round trips per expansion instead of the 12 a lane needs
            // to stream a 1.5 KB row with eight 16-byte loads in flight.  Other rows: one lane per node.
            float d = 0.f;
            uint32_t cnb = nb;  // node held by this lane in the order the replay walks
            if (a.coop) {
                const unsigned src = __fns(fm, 0, lane + 1);      // lane of the (lane + 1)-th new node
                cnb = __shfl_sync(0xffffffffu, nb, src < 32 ? src : 0);
                const size_t row_bytes = (size_t)a.dim * sizeof(T);
                const int n_chunks = (int)((row_bytes + RC_CHUNK - 1) / RC_CHUNK);
                if constexpr (sizeof(T) == 4) {
                    // fp32 rows: trips of 8, four lanes per row (rc_gather_quad)
                    for (int base_i = 0; base_i < nfresh; base_i += 8) {
                        const unsigned char* src_row[8];
                        bool src_ok[8];
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = 4 * i + (lane >> 3);
                            const int sl = base_i + r;
                            const uint32_t rid = __shfl_sync(0xffffffffu, cnb, sl < 32 ? sl : 0);
                            src_ok[i] = sl < nfresh;
                            src_row[i] = reinterpret_cast<const unsigned char*>(db) + (size_t)rid * row_bytes + (lane & 7) * 16;
                        }
#pragma unroll
                        for (int i = 2; i < 8; i++) { src_ok[i] = false; src_row[i] = nullptr; }
                        QuadPart<METRIC> part;
                        part.init();
                        const bool okq = base_i + (lane >> 2) < nfresh;   // this lane's quad holds a row
                        rc_gather_quad<METRIC, HN_SR_QUAD / 8>(stage, src_row, src_ok, row_bytes, n_chunks, a.dim, qf, okq, lane, part);
                        ExactAcc<METRIC> acc;
                        part.collect(acc, lane);
                        float dd = acc.finish();
                        if (METRIC == METRIC_DOT) dd = -dd;
                        dd = __fadd_rn(dd, 0.f);
                        // the replay wants row j of the trip in lane base_i + j
                        const int from = 4 * (lane - base_i);
                        const float got = __shfl_sync(0xffffffffu, dd, (from >= 0 && from < 32) ? from : 0);
                        if (lane >= base_i && lane < base_i + 8 && lane < nfresh) d = got;
                    }
                } else
                for (int base_i = 0; base_i < nfresh; base_i += 16) {   // trips of 16 rows (HN_SR = 48 row buffers)
                    const int mine_i = lane - base_i;                   // row slot of this lane in the trip
                    const bool okc = lane >= base_i && lane < nfresh && mine_i < 16;
                    const unsigned char* src_row[8];
                    bool src_ok[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = 4 * i + (lane >> 3);              // trip row copied by copy instruction i
                        const int sl = base_i + r;                      // lane that owns it
                        const uint32_t rid = __shfl_sync(0xffffffffu, cnb, sl < 32 ? sl : 0);
                        src_ok[i] = r < 16 && sl < nfresh;
                        src_row[i] = reinterpret_cast<const unsigned char*>(db) + (size_t)rid * row_bytes + (lane & 7) * 16;
                    }
                    // lane (base_i + j) must accumulate trip row j: rc_gather reads row `lane` of the staging space,
                    // so the trip is issued from a rotated view: lanes base_i.. act as rows 0..
                    const int rows_here = min(16, nfresh - base_i);
                    ExactAcc<METRIC> acc;
                    acc.init();
                    const int vlane = okc ? mine_i : 31;                // staging row this lane reads (31: unused)
                    if (rows_here <= 8) rc_gather<T, METRIC, HN_SR / 8, 8>(stage, src_row, src_ok, row_bytes, n_chunks, a.dim, qf, okc, lane, acc, vlane);
                    else rc_gather<T, METRIC, HN_SR / 16, 16>(stage, src_row, src_ok, row_bytes, n_chunks, a.dim, qf, okc, lane, acc, vlane);
                    if (okc) {
                        float dd = acc.finish();
                        if (METRIC == METRIC_DOT) dd = -dd;
                        d = __fadd_rn(dd, 0.f);
                    }
                }
            } else if (fresh) {
                d = dist_of(nb);
            }
            // replay in list order through the acceptance test (arrow_hnsw.go:1349-1370)
            unsigned m = a.coop ? (nfresh >= 32 ? 0xffffffffu : ((1u << nfresh) - 1u)) : fm;
            {
                // Once the result set is full, a node at or beyond the CURRENT worst result can never be accepted by
                // the replay (the worst only improves while the list is walked): drop those lanes up front.
                float worst0 = 3.402823466e+38f;
                int full0 = 0;
                if (lane == 0 && nr >= a.ef) { worst0 = key_of(res[0]); full0 = 1; }
                worst0 = __shfl_sync(0xffffffffu, worst0, 0);
                full0 = __shfl_sync(0xffffffffu, full0, 0);
                if (full0) m &= __ballot_sync(0xffffffffu, d < worst0);
            }
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const float dl = __shfl_sync(0xffffffffu, d, l);
                const uint32_t nl = __shfl_sync(0xffffffffu, cnb, l);
                int accept = 0;
                if (lane == 0) accept = (nr < a.ef || dl < key_of(res[0])) ? 1 : 0;
                accept = __shfl_sync(0xffffffffu, accept, 0);
                if (!accept) continue;
                int ncu = __shfl_sync(0xffffffffu, nc, 0);
                if (ncu >= a.cand_cap) {
                    // Candidate heap full (only possible once the result set is full).  Candidates strictly worse
                    // than the worst result can never be expanded -- the worst only improves, and popping one of
                    // them is exactly the stop test -- so the whole warp prunes them: compact the survivors to the
                    // front and sort them ascending (a sorted array is a min-heap).
                    const int nru = __shfl_sync(0xffffffffu, nr, 0);
                    const uint64_t worst = res[0];
                    int kept = 0;
                    if (nru >= a.ef) {
                        for (int c0 = 0; c0 < ncu; c0 += 32) {
                            const int t = c0 + lane;
                            const uint64_t v = (t < ncu) ? cand[t] : kInvalid;
                            const bool keep = (t < ncu) && !(key_of(v) > key_of(worst));
                            const unsigned km = __ballot_sync(0xffffffffu, keep);
                            __syncwarp();
                            if (keep) cand[kept + __popc(km & ((1u << lane) - 1u))] = v;
                            kept += __popc(km);
                            __syncwarp();
                        }
                        const int s2 = next_pow2(max(kept, 2));
                        for (int t = kept + lane; t < s2; t += 32) cand[t] = kInvalid;
                        __syncwarp();
                        warp_bitonic_sort(cand, s2, lane);
                    } else {
                        kept = ncu;
                    }
                    if (kept >= a.cand_cap) failed = true;  // more exact ties with the worst than the heap can hold
                    nc = kept;
                }
                if (lane == 0 && !failed) {
                    const uint64_t p = pack_key(dl, nl);
                    heap_push(cand, nc, p, false);
                    heap_push(res, nr, p, true);
                    if (nr > a.ef) heap_pop(res, nr, true);
                }
                __syncwarp();
            }
            __syncwarp();
        }
        nr = __shfl_sync(0xffffffffu, nr, 0);
        nc = __shfl_sync(0xffffffffu, nc, 0);
        failed = __any_sync(0xffffffffu, failed);
        if (failed || n_vis > vis_limit) { failed = true; break; }
    }
    nr = __shfl_sync(0xffffffffu, nr, 0);
    __syncwarp();
    // ascending output: sort the result heap's array in place (power-of-two padded)
    const int n2 = next_pow2(max(a.ef + 1, 2));
    // res has ef + 1 slots; the sort needs n2: use the candidate heap's space as the tail (it is dead now)
    for (int t = nr + lane; t < n2; t += 32) res[t] = kInvalid;  // res and cand are contiguous: cand_cap >= n2 - ef - 1
    __syncwarp();
    warp_bitonic_sort(res, n2, lane);
    for (int t = lane; t < a.ef; t += 32) {
        const uint64_t p = (t < nr && !failed) ? res[t] : kInvalid;
        a.out_ids[(size_t)q * a.ef + t] = (p == kInvalid) ? HN_EMPTY : id_of(p);
        a.out_d[(size_t)q * a.ef + t] = (p == kInvalid) ? 3.402823466e+38f : key_of(p);
    }
    if (lane == 0) {
        if (a.out_visited) a.out_visited[q] = n_vis;
        if (a.fail_flags) a.fail_flags[q] = failed ? 1u : 0u;
        if (failed) atomicAdd(a.fail_count, 1u);
    }
}

template <typename T>
static cudaError_t launch_hnsw_t(const HnswArgs& a, int metric, cudaStream_t st) {
    const size_t q_bytes = ((size_t)a.dim * 4 + 15) & ~(size_t)15;
    const size_t heap_bytes = (((size_t)(a.ef + 1 + a.cand_cap) * 8) + 15) & ~(size_t)15;
    const size_t per_warp = q_bytes + heap_bytes + (a.coop ? (size_t)hn_stage_rows<T>() * RC_PITCH : 0);
    // Warps per CTA: the walk is a chain of dependent memory round trips, so what matters is how many queries are
    // resident.  Pick the CTA size that packs the most warps into an SM's shared memory (1 KiB reserved per CTA,
    // 64 registers per thread): 4096 queries of C5w fit one wave of 148 x 28.
    int wpc = 4, best = 0;
    for (int w = 4; w <= HN_WARPS_MAX; w++) {
        const size_t cta = (size_t)w * per_warp + 1024;
        int ctas = (int)(233472 / cta);
        if (ctas > 32 / w) ctas = 32 / w;          // register file: 64 regs x 32 lanes x w warps per CTA
        if (ctas * w > best) { best = ctas * w; wpc = w; }
    }
    const size_t smem = (size_t)wpc * per_warp;
    const int grid = (a.nq + wpc - 1) / wpc;
#define LB_HN(M)                                                              \
    {                                                                         \
        auto kern = hnsw_search_layer_kernel<T, M>;                           \
        LB_SMEM_OPTIN(kern);                                                  \
        static std::atomic<int> carve_{0};                                    \
        if (!carve_.exchange(1))                                              \
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        kern<<<grid, wpc * 32, smem, st>>>(a);                                \
    }
    switch (metric) {
        case METRIC_L2: LB_HN(METRIC_L2) break;
        case METRIC_COSINE: LB_HN(METRIC_COSINE) break;
        default: LB_HN(METRIC_DOT) break;
    }
#undef LB_HN
    count_launch();
    return cudaGetLastError();
}

}  // namespace lb

using namespace lb;

struct lb_graph {
    lb_index* idx = nullptr;
    int device = 0, max_degree = 0;
    uint32_t* neighbors = nullptr;  // [n][max_degree]
    int32_t* counts = nullptr;      // [n] or null
    int64_t n = 0;
};

#define GCK(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return api_fail_cuda(e__, #call);  \
    } while (0)

extern "C" {

int lb_graph_create(lb_index* idx, int max_degree, lb_graph** out) {
    if (!out) return api_fail(LB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!idx || max_degree <= 0 || max_degree > 1024) return api_fail(LB_ERR_INVALID, "bad index / max_degree");
    IndexView v;
    int rc = api_index_view(idx, &v);
    if (rc) return rc;
    rc = api_use_device(v.device);
    if (rc) return rc;
    lb_graph* g = new (std::nothrow) lb_graph();
    if (!g) return api_fail(LB_ERR_OOM, "host allocation failed");
    g->idx = idx; g->device = v.device; g->max_degree = max_degree;
    *out = g;
    return LB_OK;
}

void lb_graph_free(lb_graph* g) {
    if (!g) return;
    if (cudaSetDevice(g->device) == cudaSuccess) {
        cudaDeviceSynchronize();
        if (g->neighbors) cudaFree(g->neighbors);
        if (g->counts) cudaFree(g->counts);
    }
    cudaGetLastError();
    delete g;
}

static int graph_set(lb_graph* g, const uint32_t* neighbors, const int32_t* counts, int64_t n, bool on_device,
                     cudaStream_t st) {
    if (!g || !neighbors || n <= 0) return api_fail(LB_ERR_INVALID, "bad argument");
    int rc = api_use_device(g->device);
    if (rc) return rc;
    GCK(cudaDeviceSynchronize());
    if (g->neighbors) { cudaFree(g->neighbors); g->neighbors = nullptr; }
    if (g->counts) { cudaFree(g->counts); g->counts = nullptr; }
    g->n = 0;
    GCK(cudaMalloc((void**)&g->neighbors, (size_t)n * g->max_degree * 4));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    GCK(cudaMemcpyAsync(g->neighbors, neighbors, (size_t)n * g->max_degree * 4, kind, st));
    if (counts) {
        GCK(cudaMalloc((void**)&g->counts, (size_t)n * 4));
        GCK(cudaMemcpyAsync(g->counts, counts, (size_t)n * 4, kind, st));
    }
    if (!on_device) GCK(cudaStreamSynchronize(st));
    g->n = n;
    return LB_OK;
}

int lb_graph_set_layer(lb_graph* g, const uint32_t* neighbors, const int32_t* counts, int64_t n) {
    return graph_set(g, neighbors, counts, n, false, cudaStreamPerThread);
}
int lb_graph_set_layer_device(lb_graph* g, const uint32_t* d_neighbors, const int32_t* d_counts, int64_t n, void* stream) {
    return graph_set(g, d_neighbors, d_counts, n, true, (cudaStream_t)stream);
}

// the walk on `st`; all pointers on the device.  *h_failed (optional) = queries that hit a limit (synchronises).
static int graph_walk(lb_graph* g, const void* d_q, int64_t nq, const uint32_t* d_entries, int ef, uint32_t* d_ids,
                      float* d_dist, uint32_t* d_visited, int ht_scale, uint32_t* d_fail_count, cudaStream_t st) {
    IndexView v;
    int rc = api_index_view(g->idx, &v);
    if (rc) return rc;
    if (g->n == 0) return api_fail(LB_ERR_STATE, "graph has no adjacency");
    if (g->n > v.size) return api_fail(LB_ERR_STATE, "graph names more nodes than the index holds rows");
    if (ef <= 0 || ef > 1024) return api_fail(LB_ERR_UNSUPPORTED, "ef must be in 1..1024");
    if (v.dtype == DT_I8 && v.metric == METRIC_COSINE) return api_fail(LB_ERR_UNSUPPORTED, "no int8 cosine kernel");
    HnswArgs a;
    a.db = v.rows; a.n_rows = (uint32_t)g->n; a.dim = v.dim;
    a.neighbors = g->neighbors; a.counts = g->counts; a.max_degree = g->max_degree;
    a.queries = d_q; a.entries = d_entries; a.ef = ef;
    const int n2 = next_pow2(ef + 1 > 2 ? ef + 1 : 2);
    // power of two (the prune step sorts it in place).  2 ef is enough: a prune keeps only candidates not worse than
    // the worst result, and every such candidate is itself in the result set -- at most ef of them (plus exact ties)
    a.cand_cap = next_pow2(2 * ef);
    if (a.cand_cap < n2) a.cand_cap = n2;
    uint32_t ht = 4096;
    while (ht < (uint32_t)ef * 96u) ht <<= 1;
    ht <<= ht_scale;
    a.ht_mask = ht - 1;
    a.out_ids = d_ids; a.out_d = d_dist; a.out_visited = d_visited; a.fail_count = d_fail_count; a.fail_flags = nullptr;
    {
        const size_t rb = (size_t)v.dim * (v.dtype == DT_F32 ? 4 : v.dtype == DT_F16 ? 2 : 1);
        a.coop = (rb % 16 == 0) && ((reinterpret_cast<uintptr_t>(v.rows) & 15) == 0) && g_hnsw_coop ? 1 : 0;
    }
    // visited tables are allocated per chunk of queries so that scratch stays bounded (<= ~1 GB)
    const int64_t chunk_max = std::max<int64_t>(64, (int64_t)(1ull << 30) / ((int64_t)ht * 4));
    for (int64_t qo = 0; qo < nq; qo += chunk_max) {
        const int64_t cq = std::min<int64_t>(chunk_max, nq - qo);
        uint32_t* table = nullptr;
        GCK(cudaMallocAsync((void**)&table, (size_t)cq * ht * 4, st));
        GCK(cudaMemsetAsync(table, 0xff, (size_t)cq * ht * 4, st));
        HnswArgs b = a;
        b.nq = (int)cq; b.visited = table;
        b.queries = (const char*)d_q + (size_t)qo * v.dim * (v.dtype == DT_F32 ? 4 : v.dtype == DT_F16 ? 2 : 1);
        b.entries = d_entries + qo;
        b.out_ids = d_ids + (size_t)qo * ef; b.out_d = d_dist + (size_t)qo * ef;
        b.out_visited = d_visited ? d_visited + qo : nullptr;
        cudaError_t e;
        switch (v.dtype) {
            case DT_F32: e = launch_hnsw_t<float>(b, v.metric, st); break;
            case DT_F16: e = launch_hnsw_t<__half>(b, v.metric, st); break;
            case DT_I8: e = launch_hnsw_t<int8_t>(b, v.metric, st); break;
            default: e = cudaErrorInvalidValue;
        }
        cudaFreeAsync(table, st);
        if (e != cudaSuccess) return api_fail_cuda(e, "hnsw_search_layer_kernel");
    }
    return LB_OK;
}

int lb_graph_search_layer_device(lb_graph* g, const void* d_queries, int64_t nq, const uint32_t* d_entry_points, int ef,
                                 uint32_t* d_ids, float* d_distances, uint32_t* d_visited, uint32_t* d_fail_count,
                                 void* stream) {
    if (!g || nq < 0) return api_fail(LB_ERR_INVALID, "bad argument");
    if (nq == 0) return LB_OK;
    if (!d_queries || !d_entry_points || !d_ids || !d_distances || !d_fail_count)
        return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(g->device);
    if (rc) return rc;
    return graph_walk(g, d_queries, nq, d_entry_points, ef, d_ids, d_distances, d_visited, 0, d_fail_count,
                      (cudaStream_t)stream);
}

// host buffers; k > 0: walk + re-rank (tombstones of the index and `allow` applied in-kernel) -> [nq][k];
// k == 0: the raw frontier -> ids [nq][ef] (as int64 labels, -1 padded) and distances [nq][ef]
static int graph_search_host(lb_graph* g, const void* queries, int64_t nq, const uint32_t* entry_points, int ef, int k,
                             const uint64_t* allow, float* distances, int64_t* labels, uint32_t* frontier_ids,
                             float* frontier_d, uint32_t* visited) {
    if (!g || nq < 0 || k < 0) return api_fail(LB_ERR_INVALID, "bad argument");
    if (nq == 0) return LB_OK;
    if (!queries || !entry_points) return api_fail(LB_ERR_INVALID, "NULL buffer");
    IndexView v;
    int rc = api_index_view(g->idx, &v);
    if (rc) return rc;
    rc = api_use_device(g->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    const size_t es = v.dtype == DT_F32 ? 4 : v.dtype == DT_F16 ? 2 : 1;
    void* d_q = nullptr; uint32_t *d_e = nullptr, *d_ids = nullptr, *d_vis = nullptr, *d_fail = nullptr;
    float *d_fd = nullptr, *d_od = nullptr; int64_t* d_ol = nullptr; uint64_t* d_allow = nullptr;
    auto cleanup = [&]() {
        for (void* p : {(void*)d_q, (void*)d_e, (void*)d_ids, (void*)d_vis, (void*)d_fail, (void*)d_fd, (void*)d_od,
                        (void*)d_ol, (void*)d_allow})
            if (p) cudaFreeAsync(p, st);
    };
#define HCK(call)                                                                        \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) { cleanup(); return api_fail_cuda(e__, #call); }         \
    } while (0)
    HCK(cudaMallocAsync(&d_q, (size_t)nq * v.dim * es, st));
    HCK(cudaMallocAsync((void**)&d_e, (size_t)nq * 4, st));
    HCK(cudaMallocAsync((void**)&d_ids, (size_t)nq * ef * 4, st));
    HCK(cudaMallocAsync((void**)&d_fd, (size_t)nq * ef * 4, st));
    HCK(cudaMallocAsync((void**)&d_vis, (size_t)nq * 4, st));
    HCK(cudaMallocAsync((void**)&d_fail, 4, st));
    HCK(cudaMemcpyAsync(d_q, queries, (size_t)nq * v.dim * es, cudaMemcpyHostToDevice, st));
    HCK(cudaMemcpyAsync(d_e, entry_points, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
    uint32_t n_fail = 0;
    for (int scale = 0; scale <= 4; scale += 2) {  // a walk that outgrows its visited table is re-run with 4x, 16x
        HCK(cudaMemsetAsync(d_fail, 0, 4, st));
        rc = graph_walk(g, d_q, nq, d_e, ef, d_ids, d_fd, d_vis, scale, d_fail, st);
        if (rc) { cleanup(); return rc; }
        HCK(cudaMemcpyAsync(&n_fail, d_fail, 4, cudaMemcpyDeviceToHost, st));
        HCK(cudaStreamSynchronize(st));
        if (n_fail == 0) break;
    }
    if (n_fail) { cleanup(); return api_fail(LB_ERR_UNSUPPORTED, "graph walk exceeded its visited-set / candidate limits"); }
    if (frontier_ids) HCK(cudaMemcpyAsync(frontier_ids, d_ids, (size_t)nq * ef * 4, cudaMemcpyDeviceToHost, st));
    if (frontier_d) HCK(cudaMemcpyAsync(frontier_d, d_fd, (size_t)nq * ef * 4, cudaMemcpyDeviceToHost, st));
    if (visited) HCK(cudaMemcpyAsync(visited, d_vis, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (k > 0) {
        if (!distances || !labels) { cleanup(); return api_fail(LB_ERR_INVALID, "NULL output buffer"); }
        HCK(cudaMallocAsync((void**)&d_od, (size_t)nq * k * 4, st));
        HCK(cudaMallocAsync((void**)&d_ol, (size_t)nq * k * 8, st));
        if (allow) {
            const size_t words = (size_t)((v.size + 63) / 64);
            HCK(cudaMallocAsync((void**)&d_allow, words * 8, st));
            HCK(cudaMemcpyAsync(d_allow, allow, words * 8, cudaMemcpyHostToDevice, st));
        }
        rc = api_rerank_device(g->idx, d_q, nq, d_ids, ef, k, d_allow, d_od, d_ol, st);
        if (rc) { cudaStreamSynchronize(st); cleanup(); return rc; }
        HCK(cudaMemcpyAsync(distances, d_od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
        HCK(cudaMemcpyAsync(labels, d_ol, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    }
    HCK(cudaStreamSynchronize(st));
    cleanup();
#undef HCK
    return LB_OK;
}

int lb_graph_search_layer(lb_graph* g, const void* queries, int64_t nq, const uint32_t* entry_points, int ef,
                          uint32_t* ids, float* distances, uint32_t* visited) {
    if (!ids || !distances) return api_fail(LB_ERR_INVALID, "NULL output buffer");
    return graph_search_host(g, queries, nq, entry_points, ef, 0, nullptr, nullptr, nullptr, ids, distances, visited);
}

int lb_graph_search(lb_graph* g, const void* queries, int64_t nq, const uint32_t* entry_points, int ef, int k,
                    const uint64_t* allow, float* distances, int64_t* labels) {
    if (k <= 0) return api_fail(LB_ERR_INVALID, "k must be positive");
    return graph_search_host(g, queries, nq, entry_points, ef, k, allow, distances, labels, nullptr, nullptr, nullptr);
}

int lb_graph_search_device(lb_graph* g, const void* d_queries, int64_t nq, const uint32_t* d_entry_points, int ef, int k,
                           const uint64_t* d_allow, float* d_distances, int64_t* d_labels, uint32_t* d_fail_count,
                           void* stream) {
    if (!g || nq < 0 || k <= 0) return api_fail(LB_ERR_INVALID, "bad argument");
    if (nq == 0) return LB_OK;
    if (!d_queries || !d_entry_points || !d_distances || !d_labels || !d_fail_count)
        return api_fail(LB_ERR_INVALID, "NULL buffer");
    int rc = api_use_device(g->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* d_ids = nullptr; float* d_fd = nullptr;
    GCK(cudaMallocAsync((void**)&d_ids, (size_t)nq * ef * 4, st));
    cudaError_t e = cudaMallocAsync((void**)&d_fd, (size_t)nq * ef * 4, st);
    if (e != cudaSuccess) { cudaFreeAsync(d_ids, st); return api_fail_cuda(e, "cudaMallocAsync"); }
    rc = graph_walk(g, d_queries, nq, d_entry_points, ef, d_ids, d_fd, nullptr, 0, d_fail_count, st);
    if (rc == LB_OK) rc = api_rerank_device(g->idx, d_queries, nq, d_ids, ef, k, d_allow, d_distances, d_labels, st);
    cudaFreeAsync(d_ids, st);
    cudaFreeAsync(d_fd, st);
    return rc;
}

}  // extern "C"
