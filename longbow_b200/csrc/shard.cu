// shard.cu -- row-sharded multi-GPU index behind the C ABI, for a host that drives all GPUs from ONE process
// (the Go server: internal/store/sharded_hnsw.go:378-503 is the CPU pattern it replaces).
//
// Rows are split into contiguous, 64-row-aligned ranges (one per device, so a global bitmap is sliced by words);
// queries are replicated; every device answers its range with the ordinary coarse scan -> exact re-score chain,
// and the re-score kernel's epilogue stores the shard's [nq,k] (distance, global label) record STRAIGHT INTO THE
// ROOT GPU's gather buffer through its NVLink peer mapping -- the "all-gather" of the exchange is the last store of
// the compute kernel, not a separate collective.  The root stream then waits on one CUDA event per shard and runs
// the (distance, label) merge kernel.  Merging exact per-shard top-k lists is exact.
// (Ranks in SEPARATE processes use csrc/exchange.cu, which needs flags instead of events.)
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/longbow_b200.h"
#pragma GCC visibility pop
#include "kernels.cuh"

namespace lb {
int api_fail(int code, const char* what);
int api_fail_cuda(cudaError_t e, const char* where);
int api_use_device(int device);
int api_search_core(lb_index* idx, const void* d_q, int64_t nq, int k, const uint64_t* d_allow, float* d_dist,
                    int64_t* d_lab, cudaStream_t st, uint32_t* d_flags, uint32_t* d_count);
int api_exact_search_one(lb_index* idx, const void* d_q1, int k, const uint64_t* d_allow, float* d_out_d,
                         int64_t* d_out_l, cudaStream_t st);
}  // namespace lb
using namespace lb;

struct lb_shard {
    int n = 0, dim = 0, dtype = 0, metric = 0, root = 0;
    std::vector<int> devices;
    std::vector<lb_index*> idx;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> done;
    int64_t total_rows = 0, rows_per_shard = 0, added = 0;
    char* gather = nullptr;  // on the root device: [n][slot_bytes]
    size_t slot_bytes = 0;
    std::vector<std::vector<uint64_t>> tomb;  // host copies are not kept; per-shard device bitmaps live in the lb_index
    int64_t last_uncertified = 0;
    // one search owns the per-device streams and the root's gather buffer: calls on one handle are serialised here
    // (a search already occupies every GPU of the set; the Go side may hold only a read lock, faiss_gpu.go:108)
    std::mutex mu;
};

#define SCK(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return api_fail_cuda(e__, #call);  \
    } while (0)

static size_t elem_bytes(int dt) { return dt == DT_F32 ? 4 : dt == DT_F16 ? 2 : 1; }
static size_t loff_of(int64_t nq, int k) { return (((size_t)nq * k * 4) + 15) & ~(size_t)15; }

extern "C" {

int lb_shard_create(const int* devices, int n_devices, int dim, int dtype, int metric, int64_t total_rows,
                    lb_shard** out) {
    if (!out) return api_fail(LB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 16 || total_rows <= 0) return api_fail(LB_ERR_INVALID, "bad devices / total_rows");
    lb_shard* s = new (std::nothrow) lb_shard();
    if (!s) return api_fail(LB_ERR_OOM, "host allocation failed");
    s->n = n_devices; s->dim = dim; s->dtype = dtype; s->metric = metric; s->root = 0;
    s->total_rows = total_rows;
    const int64_t per = (total_rows + n_devices - 1) / n_devices;
    s->rows_per_shard = ((per + 63) / 64) * 64;  // 64-aligned: a global bitmap is sliced by whole words
    for (int g = 0; g < n_devices; g++) {
        lb_index* ix = nullptr;
        int rc = lb_index_create(devices[g], dim, dtype, metric, &ix);
        if (rc) { lb_shard_free(s); return rc; }
        s->devices.push_back(devices[g]);
        s->idx.push_back(ix);
        lb_index_set_id_base(ix, (int64_t)g * s->rows_per_shard);
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        s->streams.push_back(st);
        s->done.push_back(ev);
        if (e != cudaSuccess) { lb_shard_free(s); return api_fail_cuda(e, "cudaStreamCreate"); }
        const int64_t lo = (int64_t)g * s->rows_per_shard;
        const int64_t cnt = std::max<int64_t>(0, std::min(total_rows, lo + s->rows_per_shard) - lo);
        if (cnt > 0) { rc = lb_index_reserve(ix, cnt); if (rc) { lb_shard_free(s); return rc; } }
    }
    // every non-root device writes its record into the root's gather buffer: peer access device -> root
    for (int g = 1; g < n_devices; g++) {
        if (s->devices[g] == s->devices[s->root]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, s->devices[g], s->devices[s->root]);
        if (!can) { lb_shard_free(s); return api_fail(LB_ERR_UNSUPPORTED, "no peer access to the root device"); }
        cudaSetDevice(s->devices[g]);
        cudaError_t e = cudaDeviceEnablePeerAccess(s->devices[s->root], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { lb_shard_free(s); return api_fail_cuda(e, "cudaDeviceEnablePeerAccess"); }
        cudaGetLastError();
    }
    *out = s;
    return LB_OK;
}

void lb_shard_free(lb_shard* s) {
    if (!s) return;
    for (size_t g = 0; g < s->idx.size(); g++) {
        if (cudaSetDevice(s->devices[g]) == cudaSuccess) {
            if (g < s->streams.size() && s->streams[g]) { cudaStreamSynchronize(s->streams[g]); cudaStreamDestroy(s->streams[g]); }
            if (g < s->done.size() && s->done[g]) cudaEventDestroy(s->done[g]);
        }
        lb_index_free(s->idx[g]);
    }
    if (s->gather && !s->devices.empty() && cudaSetDevice(s->devices[s->root]) == cudaSuccess) cudaFree(s->gather);
    cudaGetLastError();
    delete s;
}

int64_t lb_shard_size(const lb_shard* s) { return s ? s->added : -1; }
int lb_shard_count(const lb_shard* s) { return s ? s->n : -1; }
int64_t lb_shard_rows_per_shard(const lb_shard* s) { return s ? s->rows_per_shard : -1; }
int64_t lb_shard_last_uncertified(const lb_shard* s) { return s ? s->last_uncertified : -1; }

// rows are appended in global order; a block that crosses a range boundary is split
int lb_shard_add(lb_shard* s, const void* rows, int64_t n) {
    if (!s || n < 0) return api_fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return LB_OK;
    if (!rows) return api_fail(LB_ERR_INVALID, "rows is NULL");
    std::lock_guard<std::mutex> lock(s->mu);
    if (s->added + n > s->total_rows) return api_fail(LB_ERR_INVALID, "more rows than the shard set was created for");
    const size_t rb = (size_t)s->dim * elem_bytes(s->dtype);
    int64_t done = 0;
    while (done < n) {
        const int64_t gpos = s->added + done;
        const int g = (int)(gpos / s->rows_per_shard);
        const int64_t room = (int64_t)(g + 1) * s->rows_per_shard - gpos;
        const int64_t take = std::min(room, n - done);
        int rc = lb_index_add(s->idx[g], (const char*)rows + (size_t)done * rb, take);
        if (rc) return rc;
        done += take;
    }
    s->added += n;
    return LB_OK;
}

// global tombstone bitmap (bit i <-> global row i), sliced by whole words per shard
int lb_shard_set_tombstones(lb_shard* s, const uint64_t* bitmap, int64_t nbits) {
    if (!s) return api_fail(LB_ERR_INVALID, "shard set is NULL");
    std::lock_guard<std::mutex> lock(s->mu);
    for (int g = 0; g < s->n; g++) {
        const int64_t lo = (int64_t)g * s->rows_per_shard;
        int rc;
        if (!bitmap || nbits <= lo) rc = lb_index_set_tombstones(s->idx[g], nullptr, 0);
        else rc = lb_index_set_tombstones(s->idx[g], bitmap + lo / 64, std::min(nbits - lo, s->rows_per_shard));
        if (rc) return rc;
    }
    return LB_OK;
}

int lb_shard_search(lb_shard* s, const void* queries, int64_t nq, int k, const uint64_t* allow, float* distances,
                    int64_t* labels) {
    if (!s) return api_fail(LB_ERR_INVALID, "shard set is NULL");
    if (k <= 0 || nq < 0) return api_fail(LB_ERR_INVALID, "k must be positive and nq non-negative");
    if (nq == 0) return LB_OK;
    if (!queries || !distances || !labels) return api_fail(LB_ERR_INVALID, "NULL buffer");
    if ((int64_t)s->n * k > 16384) return api_fail(LB_ERR_UNSUPPORTED, "shards * k > 16384");
    std::lock_guard<std::mutex> lock(s->mu);
    const size_t qb = (size_t)nq * s->dim * elem_bytes(s->dtype);
    const size_t loff = loff_of(nq, k);
    const size_t rec = (loff + (size_t)nq * k * 8 + 255) & ~(size_t)255;
    int rc = api_use_device(s->devices[s->root]);
    if (rc) return rc;
    if (rec > s->slot_bytes) {
        SCK(cudaDeviceSynchronize());
        if (s->gather) cudaFree(s->gather);
        s->gather = nullptr; s->slot_bytes = 0;
        SCK(cudaMalloc((void**)&s->gather, rec * s->n));
        s->slot_bytes = rec;
    }
    const int n = s->n;
    std::vector<void*> d_q(n, nullptr);
    std::vector<uint64_t*> d_allow(n, nullptr);
    std::vector<uint32_t*> d_flags(n, nullptr), d_count(n, nullptr);
    std::vector<uint32_t> h_count(n, 0);
    auto cleanup = [&]() {
        for (int g = 0; g < n; g++) {
            if (cudaSetDevice(s->devices[g]) != cudaSuccess) continue;
            for (void* p : {(void*)d_q[g], (void*)d_allow[g], (void*)d_flags[g], (void*)d_count[g]})
                if (p) cudaFreeAsync(p, s->streams[g]);
        }
    };
#define XCK(call)                                                                  \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) { cleanup(); return api_fail_cuda(e__, #call); }   \
    } while (0)
    // 1. every shard: queries (+ its slice of the predicate bitmap) up, search, record straight into the root's slot
    for (int g = 0; g < n; g++) {
        rc = api_use_device(s->devices[g]);
        if (rc) { cleanup(); return rc; }
        cudaStream_t st = s->streams[g];
        XCK(cudaMallocAsync(&d_q[g], qb, st));
        XCK(cudaMemcpyAsync(d_q[g], queries, qb, cudaMemcpyHostToDevice, st));
        const int64_t lo = (int64_t)g * s->rows_per_shard;
        const int64_t cnt = lb_index_size(s->idx[g]);
        if (allow && cnt > 0) {
            const size_t words = (size_t)((cnt + 63) / 64);
            XCK(cudaMallocAsync((void**)&d_allow[g], words * 8, st));
            XCK(cudaMemcpyAsync(d_allow[g], allow + lo / 64, words * 8, cudaMemcpyHostToDevice, st));
        }
        XCK(cudaMallocAsync((void**)&d_flags[g], (size_t)nq * 4, st));
        XCK(cudaMallocAsync((void**)&d_count[g], 4, st));
        XCK(cudaMemsetAsync(d_count[g], 0, 4, st));
        XCK(cudaMemsetAsync(d_flags[g], 0, (size_t)nq * 4, st));
        char* slot = s->gather + (size_t)g * s->slot_bytes;  // root memory: a peer store for g != root
        rc = api_search_core(s->idx[g], d_q[g], nq, k, d_allow[g], reinterpret_cast<float*>(slot),
                             reinterpret_cast<int64_t*>(slot + loff), st, d_flags[g], d_count[g]);
        if (rc) { cleanup(); return rc; }
        XCK(cudaMemcpyAsync(&h_count[g], d_count[g], 4, cudaMemcpyDeviceToHost, st));
        XCK(cudaEventRecord(s->done[g], st));
    }
    // 2. root: wait for every shard's record, merge, results down
    rc = api_use_device(s->devices[s->root]);
    if (rc) { cleanup(); return rc; }
    cudaStream_t rs = s->streams[s->root];
    for (int g = 0; g < n; g++)
        if (g != s->root) XCK(cudaStreamWaitEvent(rs, s->done[g], 0));
    float* d_od = nullptr; int64_t* d_ol = nullptr;
    XCK(cudaMallocAsync((void**)&d_od, (size_t)nq * k * 4, rs));
    XCK(cudaMallocAsync((void**)&d_ol, (size_t)nq * k * 8, rs));
    cudaError_t e = launch_merge_topk_strided(s->gather, s->slot_bytes, s->gather + loff, s->slot_bytes, n, (int)nq, k, k,
                                              d_od, d_ol, rs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(distances, d_od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, rs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(labels, d_ol, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, rs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rs);
    for (int g = 0; g < n && e == cudaSuccess; g++) {
        if (cudaSetDevice(s->devices[g]) == cudaSuccess) e = cudaStreamSynchronize(s->streams[g]);
    }
    if (e != cudaSuccess) { cudaFreeAsync(d_od, rs); cudaFreeAsync(d_ol, rs); cleanup(); return api_fail_cuda(e, "shard merge"); }
    // 3. certification: queries any shard flagged are re-done exhaustively on every shard and merged on the host
    int64_t uncert = 0;
    for (int g = 0; g < n; g++) uncert += h_count[g];
    s->last_uncertified = uncert;
    if (uncert > 0) {
        std::vector<uint32_t> any((size_t)nq, 0), fl((size_t)nq);
        for (int g = 0; g < n; g++) {
            if (!h_count[g]) continue;
            cudaSetDevice(s->devices[g]);
            XCK(cudaMemcpy(fl.data(), d_flags[g], (size_t)nq * 4, cudaMemcpyDeviceToHost));
            for (int64_t q = 0; q < nq; q++) any[q] |= fl[q];
        }
        const size_t qstride = (size_t)s->dim * elem_bytes(s->dtype);
        std::vector<float> hd((size_t)n * k);
        std::vector<int64_t> hl((size_t)n * k);
        for (int64_t q = 0; q < nq; q++) {
            if (!any[q]) continue;
            for (int g = 0; g < n; g++) {
                rc = api_use_device(s->devices[g]);
                if (rc) { cleanup(); return rc; }
                cudaStream_t st = s->streams[g];
                float* td; int64_t* tl;
                XCK(cudaMallocAsync((void**)&td, (size_t)k * 4, st));
                XCK(cudaMallocAsync((void**)&tl, (size_t)k * 8, st));
                if (lb_index_size(s->idx[g]) > 0) {
                    rc = api_exact_search_one(s->idx[g], (const char*)d_q[g] + (size_t)q * qstride, k, d_allow[g], td, tl, st);
                    if (rc) { cleanup(); return rc; }
                    XCK(cudaMemcpyAsync(hd.data() + (size_t)g * k, td, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
                    XCK(cudaMemcpyAsync(hl.data() + (size_t)g * k, tl, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
                    XCK(cudaStreamSynchronize(st));
                } else {
                    for (int j = 0; j < k; j++) { hd[(size_t)g * k + j] = 3.402823466e+38f; hl[(size_t)g * k + j] = -1; }
                }
                cudaFreeAsync(td, st);
                cudaFreeAsync(tl, st);
            }
            std::vector<int> order;
            for (int t = 0; t < n * k; t++) if (hl[t] >= 0) order.push_back(t);
            std::sort(order.begin(), order.end(), [&](int a, int b) {
                return hd[a] < hd[b] || (hd[a] == hd[b] && hl[a] < hl[b]);
            });
            for (int j = 0; j < k; j++) {
                const bool ok = j < (int)order.size();
                distances[(size_t)q * k + j] = ok ? hd[order[j]] : 3.402823466e+38f;
                labels[(size_t)q * k + j] = ok ? hl[order[j]] : -1;
            }
        }
    }
    cudaSetDevice(s->devices[s->root]);
    cudaFreeAsync(d_od, rs);
    cudaFreeAsync(d_ol, rs);
    cleanup();
#undef XCK
    return LB_OK;
}

}  // extern "C"
