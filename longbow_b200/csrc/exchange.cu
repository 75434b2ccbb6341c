// exchange.cu -- the multi-GPU exchange step of the row-sharded search, written as our own
// kernels over NVLink peer memory instead of a library collective.
//
// What it replaces: the concat + sort tail of ShardedHNSW.SearchVectors (internal/store/sharded_hnsw.go:432-503)
// and MergeSortedStreams (internal/store/result_merger.go:34-100).  Every rank (one GPU each -- separate
// processes connected through CUDA IPC, or the devices of one process connected through peer access) owns a
// receive buffer of 2 (parity) x world record slots.  A batch's exchange is
//   push   : the local [nq,k] (distance, label) record is written straight into slot[parity][rank] of EVERY
//            peer with 16-byte stores over NVLink (all-gather semantics: every rank ends up with every record);
//   signal : one release-store of the batch's sequence number into flags[rank] on every peer;
//   wait   : one warp polls the flags of all sources (acquire loads, bounded by a timeout) until they reach the
//            sequence number;
//   merge  : the (distance, label) merge kernel merges the world records of the batch.
// No rank ever waits before pushing, so there is no cycle; a slot of parity p is only overwritten by batch
// s+2 after the pusher has merged batch s+1, which needed the receiver's signal s+1, which the receiver issues
// (stream order) after its own merge of batch s -- the last reader of that slot.
#include <cstdio>
#include <cstring>
#include <new>

#pragma GCC visibility push(default)
#include "../../include/longbow_b200.h"
#pragma GCC visibility pop
#include "kernels.cuh"

namespace lb {

int api_fail(int code, const char* what);            // api.cu
int api_fail_cuda(cudaError_t e, const char* where);  // api.cu
int api_use_device(int device);                       // api.cu

constexpr int EX_MAXW = 16;

struct PeerPtrs {
    char* recv[EX_MAXW];
    uint32_t* flags[EX_MAXW];
};

__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint4* __restrict__ src, size_t n16, PeerPtrs pp, size_t slot_off, int rank, int world) {
    // grid.y enumerates the peers (skipping self), grid.x stripes the record
    int peer = blockIdx.y;
    if (peer >= rank) peer++;
    if (peer >= world) return;
    uint4* dst = reinterpret_cast<uint4*>(pp.recv[peer] + slot_off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

__global__ void exchange_signal_kernel(PeerPtrs pp, int rank, int world, uint32_t seq) {
    const int p = threadIdx.x;
    if (p >= world) return;
    // the pushes were issued by the previous kernel on this stream; the fence + release store orders them
    // before the flag for any observer that acquires it at system scope
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[p] + rank), "r"(seq) : "memory");
}

// One warp waits for the batch's records: lane p polls source p's flag (acquire at system scope) until it reaches
// the sequence number, at most ~10 s.  The MERGE kernel must not do this waiting itself: a thousand resident merge
// CTAs spinning on a slow peer hold every SM's shared memory and lock the next batch's persistent scan out, and the
// ranks then convoy (measured: 8 GPUs slower than 4).  A single warp costs nothing while it waits.
__global__ void exchange_wait_kernel(const uint32_t* flags, uint32_t seq, int world, uint32_t* err) {
    const int p = threadIdx.x;
    if (p >= world) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
        if ((int32_t)(v - seq) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { atomicExch(err, 1u); break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

}  // namespace lb

using namespace lb;

struct lb_exchange {
    int device = 0, rank = 0, world = 1;
    size_t slot_bytes = 0;
    char* recv = nullptr;       // [2][world][slot_bytes] | flags[EX_MAXW] | err
    uint32_t* flags = nullptr;  // inside the same allocation: visible to peers through the one mapping
    uint32_t* err = nullptr;
    PeerPtrs peers;
    bool ipc_opened[EX_MAXW];
    bool connected = false;
    uint32_t seq = 0;           // sequence number of the batch in flight (0 = none yet)
};

extern "C" {

int lb_exchange_create(int device, int rank, int world, size_t max_record_bytes, lb_exchange** out) {
    if (!out) return api_fail(LB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > EX_MAXW || rank < 0 || rank >= world || max_record_bytes == 0)
        return api_fail(LB_ERR_INVALID, "bad rank / world / record size");
    int rc = api_use_device(device);
    if (rc) return rc;
    lb_exchange* ex = new (std::nothrow) lb_exchange();
    if (!ex) return api_fail(LB_ERR_OOM, "host allocation failed");
    ex->device = device; ex->rank = rank; ex->world = world;
    ex->slot_bytes = (max_record_bytes + 255) & ~(size_t)255;
    const size_t data = 2 * (size_t)world * ex->slot_bytes;
    cudaError_t e = cudaMalloc((void**)&ex->recv, data + 256);  // plain cudaMalloc: IPC-exportable
    if (e != cudaSuccess) { delete ex; return api_fail_cuda(e, "cudaMalloc(exchange)"); }
    e = cudaMemset(ex->recv + data, 0, 256);
    if (e != cudaSuccess) { cudaFree(ex->recv); delete ex; return api_fail_cuda(e, "cudaMemset(exchange)"); }
    ex->flags = reinterpret_cast<uint32_t*>(ex->recv + data);
    ex->err = ex->flags + EX_MAXW;
    memset(&ex->peers, 0, sizeof ex->peers);
    memset(ex->ipc_opened, 0, sizeof ex->ipc_opened);
    ex->peers.recv[rank] = ex->recv;
    ex->peers.flags[rank] = ex->flags;
    ex->connected = (world == 1);
    *out = ex;
    return LB_OK;
}

void lb_exchange_free(lb_exchange* ex) {
    if (!ex) return;
    if (cudaSetDevice(ex->device) == cudaSuccess) {
        cudaDeviceSynchronize();
        for (int p = 0; p < ex->world; p++)
            if (ex->ipc_opened[p]) cudaIpcCloseMemHandle(ex->peers.recv[p]);
        if (ex->recv) cudaFree(ex->recv);
    }
    cudaGetLastError();
    delete ex;
}

int lb_exchange_handle(lb_exchange* ex, void* handle64) {
    if (!ex || !handle64) return api_fail(LB_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    int rc = api_use_device(ex->device);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ex->recv);
    if (e != cudaSuccess) return api_fail_cuda(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, 64);
    return LB_OK;
}

int lb_exchange_connect_ipc(lb_exchange* ex, const void* handles) {
    if (!ex || !handles) return api_fail(LB_ERR_INVALID, "NULL argument");
    int rc = api_use_device(ex->device);
    if (rc) return rc;
    const size_t data = 2 * (size_t)ex->world * ex->slot_bytes;
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank || ex->peers.recv[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)p * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return api_fail_cuda(e, "cudaIpcOpenMemHandle (peer receive buffer)");
        ex->peers.recv[p] = (char*)ptr;
        ex->peers.flags[p] = reinterpret_cast<uint32_t*>((char*)ptr + data);
        ex->ipc_opened[p] = true;
    }
    ex->connected = true;
    return LB_OK;
}

int lb_exchange_connect_local(lb_exchange* ex, lb_exchange* const* all) {
    if (!ex || !all) return api_fail(LB_ERR_INVALID, "NULL argument");
    int rc = api_use_device(ex->device);
    if (rc) return rc;
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank) continue;
        lb_exchange* o = all[p];
        if (!o || o->world != ex->world || o->rank != p || o->slot_bytes != ex->slot_bytes)
            return api_fail(LB_ERR_INVALID, "peer exchange handles do not match");
        if (o->device != ex->device) {
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, ex->device, o->device);
            if (e != cudaSuccess || !can) { cudaGetLastError(); return api_fail(LB_ERR_UNSUPPORTED, "no peer access between the devices"); }
            e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return api_fail_cuda(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        ex->peers.recv[p] = o->recv;
        ex->peers.flags[p] = o->flags;
    }
    ex->connected = true;
    return LB_OK;
}

static size_t label_offset_of(int64_t nq, int k) { return (((size_t)nq * k * 4) + 15) & ~(size_t)15; }

int lb_exchange_slot(lb_exchange* ex, int64_t nq, int k, float** d_dist, int64_t** d_labels) {
    if (!ex || !d_dist || !d_labels || nq <= 0 || k <= 0) return api_fail(LB_ERR_INVALID, "bad argument");
    const size_t need = label_offset_of(nq, k) + (size_t)nq * k * 8;
    if (need > ex->slot_bytes) return api_fail(LB_ERR_INVALID, "record larger than the exchange was created for");
    ex->seq++;
    char* slot = ex->recv + ((size_t)(ex->seq & 1) * ex->world + ex->rank) * ex->slot_bytes;
    *d_dist = reinterpret_cast<float*>(slot);
    *d_labels = reinterpret_cast<int64_t*>(slot + label_offset_of(nq, k));
    return LB_OK;
}

int lb_exchange_all_gather_merge(lb_exchange* ex, int64_t nq, int k_in, int k, float* d_out_d, int64_t* d_out_l,
                                 void* stream) {
    if (!ex || nq <= 0 || k_in <= 0 || k <= 0 || !d_out_d || !d_out_l) return api_fail(LB_ERR_INVALID, "bad argument");
    if (!ex->connected) return api_fail(LB_ERR_STATE, "exchange is not connected to its peers");
    if (ex->seq == 0) return api_fail(LB_ERR_STATE, "lb_exchange_slot has not been called for this batch");
    if ((int64_t)ex->world * k_in > 16384) return api_fail(LB_ERR_UNSUPPORTED, "world * k_in > 16384");
    int rc = api_use_device(ex->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t loff = label_offset_of(nq, k_in);
    const size_t rec = loff + (size_t)nq * k_in * 8;
    const size_t parity_off = (size_t)(ex->seq & 1) * ex->world * ex->slot_bytes;
    const size_t slot_off = parity_off + (size_t)ex->rank * ex->slot_bytes;
    if (ex->world > 1) {
        const size_t n16 = (rec + 15) / 16;
        int bx = (int)((n16 + 255) / 256);
        if (bx > 32) bx = 32;
        dim3 grid(bx, ex->world - 1);
        exchange_push_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(ex->recv + slot_off), n16, ex->peers,
                                                   slot_off, ex->rank, ex->world);
        exchange_signal_kernel<<<1, 32, 0, st>>>(ex->peers, ex->rank, ex->world, ex->seq);
        count_launch(); count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return api_fail_cuda(e, "exchange push");
    }
    if (ex->world > 1) {
        exchange_wait_kernel<<<1, 32, 0, st>>>(ex->flags, ex->seq, ex->world, ex->err);
        count_launch();
    }
    cudaError_t e = launch_merge_topk_strided(ex->recv + parity_off, ex->slot_bytes, ex->recv + parity_off + loff,
                                              ex->slot_bytes, ex->world, (int)nq, k_in, k, d_out_d, d_out_l, st);
    if (e != cudaSuccess) return api_fail_cuda(e, "exchange merge");
    return LB_OK;
}

int lb_exchange_error(lb_exchange* ex) {
    if (!ex) return api_fail(LB_ERR_INVALID, "exchange is NULL");
    int rc = api_use_device(ex->device);
    if (rc) return rc;
    uint32_t v = 0;
    cudaError_t e = cudaMemcpy(&v, ex->err, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return api_fail_cuda(e, "cudaMemcpy(exchange error word)");
    if (v) return api_fail(LB_ERR_STATE, "exchange: a peer's record did not arrive within the timeout");
    return LB_OK;
}

}  // extern "C"
