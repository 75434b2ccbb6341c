// api.cu -- the C ABI of liblongbow_b200.so (include/longbow_b200.h): handles, HBM mirrors,
// scratch management, launch planning.  No CPU fallback anywhere: without a CUDA device every
// entry point fails with a code.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/longbow_b200.h"
#pragma GCC visibility pop
#include "kernels.cuh"

namespace lb {

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local std::string t_err;
static int fail(int code, const char* what) {
    t_err = what ? what : "";
    return code;
}
static int fail_cuda(cudaError_t e, const char* where) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    t_err = buf;
    cudaGetLastError();  // clear sticky-less errors
    return e == cudaErrorMemoryAllocation ? LB_ERR_OOM : LB_ERR_CUDA;
}
#define CK(call)                                             \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

static size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : dt == DT_F16 ? 2 : 1; }

struct DeviceInfo {
    int sm_count = 0;
    bool ok = false;
};
static DeviceInfo g_dev[64];
static std::mutex g_dev_mu;

static int use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(LB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= n || device >= 64) return fail(LB_ERR_NO_DEVICE, "device id out of range");
    CK(cudaSetDevice(device));
    std::lock_guard<std::mutex> g(g_dev_mu);
    if (!g_dev[device].ok) {
        cudaDeviceProp p;
        CK(cudaGetDeviceProperties(&p, device));
        if (p.major < 10) return fail(LB_ERR_NO_DEVICE, "device is not sm_100 class (built for sm_100a only)");
        g_dev[device].sm_count = p.multiProcessorCount;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ull;  // keep freed scratch cached in the pool
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        g_dev[device].ok = true;
    }
    return LB_OK;
}

// the same helpers for the other translation units with C-ABI entry points (exchange.cu, shard.cu, pq_scan.cu ...)
int api_fail(int code, const char* what) { return fail(code, what); }
int api_fail_cuda(cudaError_t e, const char* where) { return fail_cuda(e, where); }
int api_use_device(int device) { return use_device(device); }
int api_sm_count(int device) { return (device >= 0 && device < 64) ? g_dev[device].sm_count : 0; }

// stream-ordered scratch
struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    cudaError_t get(void** p, size_t bytes) {
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMallocAsync(p, bytes, st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
};

// ---- options (lb_set_option)
static std::atomic<int> g_opt_dense_scan{0};
static std::atomic<int> g_opt_tc_debug{0};
static std::atomic<int> g_opt_tc_boot_tiles{0};  // 0 = auto
static std::atomic<int> g_opt_f32_tc{1};         // fp32 indexes: 3xTF32 tensor-core scan (0: SIMT scan)
static std::atomic<int> g_opt_pq_scan{0};  // 0 auto, 1 exhaustive fp32 kernel, 2 coarse (1 query / pass), 3 coarse (4 / pass), 4 decode + tensor-core scan
static std::atomic<int> g_opt_pq_gemm{1};   // auto policy may use the decode + tensor-core path for batches (0: never)
static std::atomic<int> g_opt_exhaustive_k{256};  // single-query searches with k >= this take the exhaustive exact chain (0: only beyond the fused selector)
static std::atomic<int> g_opt_certify{1};        // host searches: certify the coarse stage, redo flagged queries exactly
static std::atomic<int> g_opt_tc_boot{1};     // bootstrap threshold scan on/off (A/B timing)    // timing probes of the tensor-core scan (results invalid when != 0)  // 0 auto, 1 force SIMT, 2 force tensor-core (error if ineligible)

// ---- dominant-kernel timing (lb_prof_*)
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
// channel 0: the dominant scan kernel (units = queries x rows); channel 1: auxiliary streaming kernels worth their own
// roofline line (the PQ decode: units = bytes moved)
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events[2];
static double g_prof_units[2] = {0, 0};
struct ProfScope {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    double units;
    int ch;
    ProfScope(cudaStream_t s, double u, int channel = 0) : st(s), units(u), ch(channel) {
        if (g_prof_on.load(std::memory_order_relaxed)) {
            if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, st);
            else e0 = e1 = nullptr;
        }
    }
    ~ProfScope() {
        if (e0 && e1) {
            cudaEventRecord(e1, st);
            std::lock_guard<std::mutex> g(g_prof_mu);
            g_prof_events[ch].emplace_back(e0, e1);
            g_prof_units[ch] += units;
        }
    }
};

}  // namespace lb

using namespace lb;

// ---------------------------------------------------------------------------------------------
// Growable device buffer on the CUDA virtual-memory API: one large address reservation, physical chunks mapped
// behind the used part as it grows.  Growing never copies and never synchronises the device (round 1's grow()
// allocated 1.5x, device-synchronised and copied: 2.5x transient HBM on a 100 M-row mirror), the base pointer never
// changes, and a 180 GB part can be filled to the brim.
// ---------------------------------------------------------------------------------------------
// The driver API is reached through cudaGetDriverEntryPoint, not by linking libcuda: the library must load (and export
// its symbols) on a machine without a driver, and fail loudly only when a GPU call is made.
namespace drv {
struct Api {
    CUresult (*GetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    bool ok = false;
};
static const Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult qr;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) == cudaSuccess &&
                   qr == cudaDriverEntryPointSuccess && *fn != nullptr;
        };
        a.ok = get("cuMemGetAllocationGranularity", (void**)&a.GetAllocationGranularity) &&
               get("cuMemAddressReserve", (void**)&a.AddressReserve) && get("cuMemAddressFree", (void**)&a.AddressFree) &&
               get("cuMemCreate", (void**)&a.Create) && get("cuMemRelease", (void**)&a.Release) &&
               get("cuMemMap", (void**)&a.Map) && get("cuMemUnmap", (void**)&a.Unmap) &&
               get("cuMemSetAccess", (void**)&a.SetAccess);
    });
    return a;
}
}  // namespace drv

struct VBuf {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0;
    int device = 0;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;

    void* ptr() const { return reinterpret_cast<void*>(base); }
    int reserve_va(int dev, size_t max_bytes) {
        device = dev;
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = dev;
        const drv::Api& d = drv::api();
        if (!d.ok) return fail(LB_ERR_CUDA, "CUDA driver virtual-memory entry points unavailable");
        if (d.GetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
            return fail(LB_ERR_CUDA, "cuMemGetAllocationGranularity failed");
        reserved = ((max_bytes + gran - 1) / gran) * gran;
        if (d.AddressReserve(&base, reserved, 0, 0, 0) != CUDA_SUCCESS) {
            base = 0; reserved = 0;
            return fail(LB_ERR_OOM, "cuMemAddressReserve failed");
        }
        return LB_OK;
    }
    // make [0, bytes) usable
    int ensure(size_t bytes) {
        if (bytes <= mapped) return LB_OK;
        if (bytes > reserved) return fail(LB_ERR_OOM, "index grew beyond its address reservation");
        // grow in steps of at least 1/8 of what is mapped (bounded number of chunks), rounded to the granularity
        size_t want = bytes - mapped;
        const size_t step = mapped / 8;
        if (want < step) want = step;
        want = ((want + gran - 1) / gran) * gran;
        if (mapped + want > reserved) want = reserved - mapped;
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        const drv::Api& d = drv::api();
        CUmemGenericAllocationHandle h;
        CUresult r = d.Create(&h, want, &prop, 0);
        if (r != CUDA_SUCCESS && want > ((bytes - mapped + gran - 1) / gran) * gran) {  // retry with the exact need
            want = ((bytes - mapped + gran - 1) / gran) * gran;
            r = d.Create(&h, want, &prop, 0);
        }
        if (r != CUDA_SUCCESS) return fail(LB_ERR_OOM, "cuMemCreate failed (device memory exhausted)");
        if (d.Map(base + mapped, want, 0, h, 0) != CUDA_SUCCESS) { d.Release(h); return fail(LB_ERR_CUDA, "cuMemMap failed"); }
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (d.SetAccess(base + mapped, want, &acc, 1) != CUDA_SUCCESS) {
            d.Unmap(base + mapped, want); d.Release(h);
            return fail(LB_ERR_CUDA, "cuMemSetAccess failed");
        }
        chunks.emplace_back(h, want);
        mapped += want;
        return LB_OK;
    }
    void release() {
        if (base) {
            size_t off = 0;
            const drv::Api& d = drv::api();
            for (auto& c : chunks) { d.Unmap(base + off, c.second); d.Release(c.first); off += c.second; }
            d.AddressFree(base, reserved);
        }
        chunks.clear();
        base = 0; reserved = mapped = 0;
    }
};

// ---------------------------------------------------------------------------------------------
struct lb_index {
    int device, dim, dtype, metric;
    void* rows = nullptr;  // [capacity][dim]
    float* aux = nullptr;  // [capacity + 256] coarse-key auxiliaries
    float* nrm = nullptr;  // [capacity] cosine only: exact |x|^2 in reference lane order
    float* max_norm2 = nullptr;  // device scalar: max |x|^2 over the rows (certification bound)
    float* lo = nullptr;   // fp32 only, built on first tensor-core search: x - tf32(x) for the 3xTF32 scan
    int64_t lo_rows = 0, lo_cap = 0;
    std::mutex lo_mu;
    int64_t size = 0, capacity = 0;
    VBuf v_rows, v_aux, v_nrm, v_lo;  // backing store of rows / aux / nrm / lo (virtual-memory growth)
    int64_t max_rows = 0;             // rows the address reservations cover
    uint32_t* tomb = nullptr;
    int64_t tomb_bits = 0, tomb_cap_words = 0;
    std::vector<void*> retired;       // replaced bitmap buffers: freed lazily (kernels in flight may still read them)
    int64_t id_base = 0;
    int sm_count = 0;
    std::atomic<int64_t> last_uncertified{0};
};

// rows [0, need) usable: maps more physical memory behind the mirrors; never copies, never synchronises
static int index_ensure_rows(lb_index* idx, int64_t need) {
    if (need <= idx->capacity) return LB_OK;
    if (need > idx->max_rows) return fail(LB_ERR_OOM, "more rows than this device can hold");
    const size_t rb = (size_t)idx->dim * (idx->dtype == DT_F32 ? 4 : idx->dtype == DT_F16 ? 2 : 1);
    int rc = idx->v_rows.ensure((size_t)need * rb);
    if (rc) return rc;
    rc = idx->v_aux.ensure(((size_t)need + 256) * 4);  // +256: tile-tail reads
    if (rc) return rc;
    if (idx->metric == METRIC_COSINE) { rc = idx->v_nrm.ensure((size_t)need * 4); if (rc) return rc; }
    // capacity = what all mirrors cover
    int64_t cap = (int64_t)(idx->v_rows.mapped / rb);
    cap = std::min<int64_t>(cap, (int64_t)(idx->v_aux.mapped / 4) - 256);
    if (idx->metric == METRIC_COSINE) cap = std::min<int64_t>(cap, (int64_t)(idx->v_nrm.mapped / 4));
    idx->capacity = std::min<int64_t>(cap, idx->max_rows);
    idx->rows = idx->v_rows.ptr();
    idx->aux = reinterpret_cast<float*>(idx->v_aux.ptr());
    idx->nrm = idx->metric == METRIC_COSINE ? reinterpret_cast<float*>(idx->v_nrm.ptr()) : nullptr;
    return LB_OK;
}

struct lb_pq {
    int device, dims, M, K, sub;
    float* codebooks = nullptr;  // [M][K][sub]
    void* codebook16 = nullptr;  // the same, fp16: decode source of the batched tensor-core coarse stage (pq_gemm.cu)
    float* cnorm2 = nullptr;     // [M][256] |fp16 centroid|^2
    float* xn2 = nullptr;        // [xn2_cap + 256] per row: |decoded fp16 vector|^2 (the dense scan's L2 auxiliary)
    int64_t xn2_cap = 0;
    uint32_t* xmax2 = nullptr;   // device word: largest xn2 so far (float bits)
    uint8_t* codes = nullptr;    // [capacity][M] row-major (flatCodes, adc_table.go:57): exhaustive fp32 scan, fallback
    uint8_t* tiled = nullptr;    // the same codes as 32-row tiles of rotated 16-byte chunks (pq_scan.cu): coarse scan
    int64_t tiled_cap = 0;       // rows (multiple of 32)
    int64_t size = 0, capacity = 0;
    std::atomic<int64_t> last_uncertified{0};
    lb_index* raw = nullptr;
    uint32_t* tomb = nullptr;
    int64_t tomb_bits = 0;
    int sm_count = 0;
};

struct faiss_res {
    int device;
};

static int grow(void** buf, int64_t* cap, int64_t need, size_t row_bytes, int64_t used, float** aux,
                float** nrm = nullptr) {
    if (need <= *cap) return LB_OK;
    int64_t ncap = *cap + (*cap >> 1);
    if (ncap < need) ncap = need;
    if (ncap < 1024) ncap = 1024;
    void* nb = nullptr;
    cudaError_t e = cudaMalloc(&nb, (size_t)ncap * row_bytes);
    if (e != cudaSuccess && ncap > need) {  // retry with the exact size
        cudaGetLastError();
        ncap = need;
        e = cudaMalloc(&nb, (size_t)ncap * row_bytes);
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(rows)");
    float* na = nullptr;
    if (aux) {
        e = cudaMalloc((void**)&na, ((size_t)ncap + 256) * sizeof(float));  // +256: tile-tail reads
        if (e != cudaSuccess) { cudaFree(nb); return fail_cuda(e, "cudaMalloc(aux)"); }
    }
    float* nn = nullptr;
    if (nrm) {
        e = cudaMalloc((void**)&nn, (size_t)ncap * sizeof(float));
        if (e != cudaSuccess) { cudaFree(nb); if (na) cudaFree(na); return fail_cuda(e, "cudaMalloc(nrm)"); }
    }
    CK(cudaDeviceSynchronize());
    if (used > 0) {
        CK(cudaMemcpy(nb, *buf, (size_t)used * row_bytes, cudaMemcpyDeviceToDevice));
        if (aux) CK(cudaMemcpy(na, *aux, (size_t)used * sizeof(float), cudaMemcpyDeviceToDevice));
        if (nrm) CK(cudaMemcpy(nn, *nrm, (size_t)used * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    if (*buf) cudaFree(*buf);
    *buf = nb;
    if (aux) { if (*aux) cudaFree(*aux); *aux = na; }
    if (nrm) { if (*nrm) cudaFree(*nrm); *nrm = nn; }
    *cap = ncap;
    return LB_OK;
}

namespace lb {
int api_index_view(const lb_index* idx, IndexView* out) {
    if (!idx || !out) return fail(LB_ERR_INVALID, "index is NULL");
    out->rows = idx->rows; out->size = idx->size; out->dim = idx->dim; out->dtype = idx->dtype;
    out->metric = idx->metric; out->device = idx->device;
    return LB_OK;
}
}  // namespace lb

static bool supported(int dtype, int metric) {
    if (metric < 0 || metric > 2) return false;
    if (dtype == DT_F32 || dtype == DT_F16) return true;
    if (dtype == DT_I8) return metric != METRIC_COSINE;  // dispatch.go:241-242: no int8 cosine kernel
    return false;
}

extern "C" {

const char* lb_last_error(void) { return t_err.c_str(); }

int lb_set_option(const char* name, int value) {
    if (!name) return fail(LB_ERR_INVALID, "name is NULL");
    if (strcmp(name, "dense_scan") == 0) {
        if (value < 0 || value > 3) return fail(LB_ERR_INVALID, "dense_scan: 0 auto, 1 simt, 2 tensor-core, 3 streaming");
        g_opt_dense_scan.store(value);
        return LB_OK;
    }
    if (strcmp(name, "tc_boot") == 0) {
        g_opt_tc_boot.store(value ? 1 : 0);
        return LB_OK;
    }
    if (strcmp(name, "tc_boot_tiles") == 0) {
        g_opt_tc_boot_tiles.store(value < 0 ? 0 : value);
        return LB_OK;
    }
    if (strcmp(name, "certify") == 0) {
        g_opt_certify.store(value ? 1 : 0);
        return LB_OK;
    }
    if (strcmp(name, "exhaustive_k") == 0) {
        g_opt_exhaustive_k.store(value < 0 ? 0 : value);
        return LB_OK;
    }
    if (strcmp(name, "f32_tc") == 0) {
        g_opt_f32_tc.store(value ? 1 : 0);
        return LB_OK;
    }
    if (strcmp(name, "ssel_warp") == 0) {
        g_ssel_warp = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "tc_reserve_sms") == 0) {
        if (value < 0 || value > 64) return fail(LB_ERR_INVALID, "tc_reserve_sms: 0..64");
        g_tc_reserve_sms = value;
        return LB_OK;
    }
    if (strcmp(name, "tc_pair") == 0) {
        g_tc_pair = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "rescore_legacy") == 0) {
        g_rescore_legacy = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "hnsw_coop") == 0) {
        g_hnsw_coop = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "pq_ring") == 0) {
        g_pq_ring = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "pq_ahead") == 0) {
        g_pq_ahead = value < 1 ? 1 : (value > 64 ? 64 : value);
        return LB_OK;
    }
    if (strcmp(name, "rescore_block") == 0) {
        g_rescore_block = value != 0;
        return LB_OK;
    }
    if (strcmp(name, "tc_debug") == 0) {
        g_opt_tc_debug.store(value);
        return LB_OK;
    }
    if (strcmp(name, "pq_gemm") == 0) {
        g_opt_pq_gemm.store(value ? 1 : 0);
        return LB_OK;
    }
    if (strcmp(name, "pq_scan") == 0) {
        if (value < 0 || value > 4)
            return fail(LB_ERR_INVALID, "pq_scan: 0 auto, 1 exhaustive fp32, 2 coarse x1, 3 coarse x4, 4 decode + tensor cores");
        g_opt_pq_scan.store(value);
        return LB_OK;
    }
    return fail(LB_ERR_INVALID, "unknown option");
}

int lb_prof_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    return LB_OK;
}
static int prof_read_channel(int ch, double* total_ms, int64_t* launches, double* units, int reset) {
    std::lock_guard<std::mutex> g(g_prof_mu);
    double tot = 0;
    int64_t n = 0;
    for (auto& pr : g_prof_events[ch]) {
        float ms = 0;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            tot += ms;
            n++;
        }
    }
    cudaGetLastError();
    if (units) *units = g_prof_units[ch];
    if (reset) {
        for (auto& pr : g_prof_events[ch]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        g_prof_events[ch].clear();
        g_prof_units[ch] = 0;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return LB_OK;
}
int lb_prof_read(double* total_ms, int64_t* launches, double* units, int reset) {
    return prof_read_channel(0, total_ms, launches, units, reset);
}
int lb_prof_read_aux(double* total_ms, int64_t* launches, double* units, int reset) {
    return prof_read_channel(1, total_ms, launches, units, reset);
}
int64_t lb_kernel_launch_count(void) { return g_launches.load(); }

int lb_device_info(int device, int* sm_count, size_t* total_mem, char* name, size_t name_len) {
    int rc = use_device(device);
    if (rc) return rc;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (name && name_len) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    return LB_OK;
}

// ------------------------------------------------------------------------------- dense index
int lb_index_create(int device, int dim, int dtype, int metric, lb_index** out) {
    if (!out) return fail(LB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return fail(LB_ERR_INVALID, "dimension must be positive");
    if (!supported(dtype, metric)) return fail(LB_ERR_UNSUPPORTED, "no kernel for this (metric, dtype)");
    int rc = use_device(device);
    if (rc) return rc;
    lb_index* idx = new (std::nothrow) lb_index();
    if (!idx) return fail(LB_ERR_OOM, "host allocation failed");
    idx->device = device; idx->dim = dim; idx->dtype = dtype; idx->metric = metric;
    idx->sm_count = g_dev[device].sm_count;
    {
        // address reservations sized for the whole device: rows that fit its memory (at most 2^32 - 2^20 ids)
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t rb = (size_t)dim * dtype_size(dtype);
        int64_t mr = (int64_t)(total_b / rb);
        if (mr > 0xfff00000ll) mr = 0xfff00000ll;
        if (mr < 1024) mr = 1024;
        idx->max_rows = mr;
        int rc2 = idx->v_rows.reserve_va(device, (size_t)mr * rb);
        if (!rc2) rc2 = idx->v_aux.reserve_va(device, ((size_t)mr + 256) * 4);
        if (!rc2 && metric == METRIC_COSINE) rc2 = idx->v_nrm.reserve_va(device, (size_t)mr * 4);
        if (!rc2 && dtype == DT_F32) rc2 = idx->v_lo.reserve_va(device, (size_t)mr * rb);
        if (rc2) { idx->v_rows.release(); idx->v_aux.release(); idx->v_nrm.release(); idx->v_lo.release(); delete idx; return rc2; }
    }
    {
        cudaError_t e = cudaMalloc((void**)&idx->max_norm2, 4);
        if (e == cudaSuccess) e = cudaMemset(idx->max_norm2, 0, 4);
        if (e != cudaSuccess) { if (idx->max_norm2) cudaFree(idx->max_norm2); delete idx; return fail_cuda(e, "cudaMalloc(stats)"); }
    }
    *out = idx;
    return LB_OK;
}

void lb_index_free(lb_index* idx) {
    if (!idx) return;
    if (cudaSetDevice(idx->device) == cudaSuccess) {
        cudaDeviceSynchronize();
        idx->v_rows.release(); idx->v_aux.release(); idx->v_nrm.release(); idx->v_lo.release();
        if (idx->max_norm2) cudaFree(idx->max_norm2);
        if (idx->tomb) cudaFree(idx->tomb);
        for (void* p : idx->retired) cudaFree(p);
    }
    cudaGetLastError();
    delete idx;
}

int lb_index_reserve(lb_index* idx, int64_t n_rows) {
    if (!idx || n_rows < 0) return fail(LB_ERR_INVALID, "bad argument");
    int rc = use_device(idx->device);
    if (rc) return rc;
    return index_ensure_rows(idx, n_rows);  // maps physical memory behind the reservation: no copy, no sync
}

static int index_add_common(lb_index* idx, const void* src, int64_t n, bool src_on_device, cudaStream_t st) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (n < 0) return fail(LB_ERR_INVALID, "n < 0");
    if (n == 0) return LB_OK;
    if (!src) return fail(LB_ERR_INVALID, "rows is NULL");
    int rc = use_device(idx->device);
    if (rc) return rc;
    if (idx->size + n > 0xfff00000ll) return fail(LB_ERR_INVALID, "more than 2^32 rows per device handle");
    size_t rb = (size_t)idx->dim * dtype_size(idx->dtype);
    rc = index_ensure_rows(idx, idx->size + n);
    if (rc) return rc;
    char* dst = (char*)idx->rows + (size_t)idx->size * rb;
    CK(cudaMemcpyAsync(dst, src, (size_t)n * rb, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    int64_t row0 = idx->size;
    idx->size += n;
    if (idx->metric != METRIC_DOT)
        CK(launch_row_aux(idx->dtype, idx->rows, idx->size, idx->dim, idx->metric, idx->aux, row0, st));
    if (idx->metric == METRIC_COSINE)
        CK(launch_row_norm_exact(idx->dtype, idx->rows, idx->size, idx->dim, idx->nrm, row0, st));
    if (idx->max_norm2) CK(launch_row_maxnorm(idx->dtype, idx->rows, idx->size, idx->dim, row0, idx->max_norm2, st));
    if (!src_on_device) CK(cudaStreamSynchronize(st));  // cgo: the Go slice may move after return
    return LB_OK;
}

int lb_index_add(lb_index* idx, const void* rows, int64_t n) {
    return index_add_common(idx, rows, n, false, cudaStreamPerThread);
}
int lb_index_add_device(lb_index* idx, const void* d_rows, int64_t n, void* stream) {
    return index_add_common(idx, d_rows, n, true, (cudaStream_t)stream);
}
// Arrow ingest (internal/store/arrow_utils.go:112-171): rows of a FixedSizeList<T, dim> column live in the child
// values buffer, row r at element (list_offset + r) * dim -- unless the buffer was flattened by Arrow IPC and is
// only long enough for the relative offset, in which case row 0 is the buffer's first element (the reference's
// "truncated buffer heuristic", :146-160).  With pin != 0 the caller's buffer is page-locked for the duration of the
// call (cudaHostRegister) and uploaded in chunks on two streams, so the DMA of chunk i+1 overlaps the row-statistics
// kernels of chunk i and runs at full PCIe rate instead of the pageable-copy rate.
int lb_index_add_arrow(lb_index* idx, const void* values, size_t values_len_bytes, int64_t list_offset, int64_t n_rows,
                       int pin) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (n_rows < 0 || list_offset < 0) return fail(LB_ERR_INVALID, "negative size");
    if (n_rows == 0) return LB_OK;
    if (!values) return fail(LB_ERR_INVALID, "record is nil");
    const size_t rb = (size_t)idx->dim * dtype_size(idx->dtype);
    size_t start = (size_t)list_offset * rb;
    if (values_len_bytes < start + (size_t)n_rows * rb) {
        if (values_len_bytes >= (size_t)n_rows * rb) start = 0;  // truncated buffer: index 0 is logical row list_offset
        else return fail(LB_ERR_INVALID, "ExtractVector: buffer out of bounds (too small even for relative access)");
    }
    const char* src = (const char*)values + start;
    if (!pin) return index_add_common(idx, src, n_rows, false, cudaStreamPerThread);
    int rc = use_device(idx->device);
    if (rc) return rc;
    const size_t bytes = (size_t)n_rows * rb;
    bool registered = cudaHostRegister((void*)src, bytes, cudaHostRegisterDefault) == cudaSuccess;
    if (!registered) {
        cudaGetLastError();
        registered = cudaHostRegister((void*)src, bytes, cudaHostRegisterReadOnly) == cudaSuccess;  // mmap'ed IPC files
        if (!registered) cudaGetLastError();
    }
    if (idx->size + n_rows > 0xfff00000ll) { if (registered) cudaHostUnregister((void*)src); return fail(LB_ERR_INVALID, "more than 2^32 rows per device handle"); }
    rc = index_ensure_rows(idx, idx->size + n_rows);
    if (rc) { if (registered) cudaHostUnregister((void*)src); return rc; }
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaError_t e = cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking);
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)((32u << 20) / rb));  // 32 MB per copy
    int which = 0;
    for (int64_t r0 = 0; r0 < n_rows && e == cudaSuccess; r0 += chunk_rows, which ^= 1) {
        const int64_t cr = std::min(chunk_rows, n_rows - r0);
        const int64_t row0 = idx->size + r0;
        e = cudaMemcpyAsync((char*)idx->rows + (size_t)row0 * rb, src + (size_t)r0 * rb, (size_t)cr * rb,
                            cudaMemcpyHostToDevice, st[which]);
        if (e == cudaSuccess && idx->metric != METRIC_DOT)
            e = launch_row_aux(idx->dtype, idx->rows, row0 + cr, idx->dim, idx->metric, idx->aux, row0, st[which]);
        if (e == cudaSuccess && idx->metric == METRIC_COSINE)
            e = launch_row_norm_exact(idx->dtype, idx->rows, row0 + cr, idx->dim, idx->nrm, row0, st[which]);
        if (e == cudaSuccess && idx->max_norm2)
            e = launch_row_maxnorm(idx->dtype, idx->rows, row0 + cr, idx->dim, row0, idx->max_norm2, st[which]);
    }
    for (int i = 0; i < 2; i++)
        if (st[i]) { cudaError_t e2 = cudaStreamSynchronize(st[i]); if (e == cudaSuccess) e = e2; cudaStreamDestroy(st[i]); }
    if (registered) cudaHostUnregister((void*)src);
    if (e != cudaSuccess) return fail_cuda(e, "lb_index_add_arrow");
    idx->size += n_rows;
    return LB_OK;
}

int64_t lb_index_size(const lb_index* idx) { return idx ? idx->size : -1; }
int lb_index_dim(const lb_index* idx) { return idx ? idx->dim : -1; }
int lb_index_set_id_base(lb_index* idx, int64_t b) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    idx->id_base = b;
    return LB_OK;
}
int64_t lb_index_last_uncertified(const lb_index* idx) { return idx ? idx->last_uncertified.load() : -1; }

// Replaces a device bitmap without stalling the device: the new copy is uploaded on `st`, the old buffer is
// RETIRED (searches enqueued earlier through the *_device entry points may still read it) and freed once more than
// a few have piled up -- one device synchronisation per eight updates instead of two per update.
static int set_bitmap(int device, uint32_t** dst, int64_t* dst_bits, const uint64_t* src, int64_t nbits,
                      bool on_device, cudaStream_t st, std::vector<void*>* retired) {
    int rc = use_device(device);
    if (rc) return rc;
    auto retire = [&](void* p) {
        if (!p) return;
        if (!retired) { cudaDeviceSynchronize(); cudaFree(p); return; }
        retired->push_back(p);
        if (retired->size() > 8) {
            cudaDeviceSynchronize();
            for (void* r : *retired) cudaFree(r);
            retired->clear();
        }
    };
    if (!src || nbits <= 0) {
        retire(*dst);
        *dst = nullptr; *dst_bits = 0;
        return LB_OK;
    }
    size_t words = (size_t)((nbits + 63) / 64);
    uint32_t* nb = nullptr;
    CK(cudaMalloc((void**)&nb, words * 8));
    CK(cudaMemcpyAsync(nb, src, words * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!on_device) CK(cudaStreamSynchronize(st));  // the host buffer may move after return (cgo)
    else {
        // later searches run on other streams: make the upload visible to them by finishing it here (it is small)
        CK(cudaStreamSynchronize(st));
    }
    retire(*dst);
    *dst = nb; *dst_bits = (int64_t)words * 64;
    return LB_OK;
}

int lb_index_set_tombstones(lb_index* idx, const uint64_t* bitmap, int64_t nbits) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    return set_bitmap(idx->device, &idx->tomb, &idx->tomb_bits, bitmap, nbits, false, cudaStreamPerThread, &idx->retired);
}
int lb_index_set_tombstones_device(lb_index* idx, const uint64_t* d_bitmap, int64_t nbits, void* stream) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    return set_bitmap(idx->device, &idx->tomb, &idx->tomb_bits, d_bitmap, nbits, true, (cudaStream_t)stream, &idx->retired);
}

// coarse candidates per query: margin over k absorbs coarse-key rounding; the re-score stage
// restores the reference's exact values and order.
static int coarse_k(int k) {
    int m = k / 4;
    if (m < 16) m = 16;
    int kc = k + m;
    return ((kc + 31) / 32) * 32;
}

// Which searches take the exhaustive exact chain (every row's exact distance, then the two-level select_k): k beyond
// the fused selector's candidate capacity, and single-query calls with k >= 256.  Measured on the C2 database
// (1 M x 768 fp16, tools/large_k_probe.py, profiles/r2_large_k_probe.txt): one query through the padded tensor-core
// block costs 0.65 / 0.69 / 3.3 / 11.1 ms at k = 288 / 352 / 448 / 704 (the 1024-entry selector from a 512-entry coarse
// list up), the exhaustive chain 0.59 / 0.62 / 0.66 / 0.79 ms; below k = 224 the streaming scan wins (0.34 against
// 0.57 ms at k = 160).  Both plans return the same (distance, id) lists.
static bool exhaustive_plan(int64_t nq, int k) {
    if (coarse_k(k) > 896) return true;
    const int xk = g_opt_exhaustive_k.load(std::memory_order_relaxed);
    return xk > 0 && nq == 1 && k >= xk;
}

// fp32 indexes: the low parts of the rows for the 3xTF32 tensor-core scan, built (or extended after an add) by
// the first search that needs them.  Searches may run concurrently (read lock on the Go side), so the build is
// serialised here; rows only ever grow, and add() needs exclusive access anyway.
static int ensure_lo(lb_index* idx, cudaStream_t st) {
    std::lock_guard<std::mutex> g(idx->lo_mu);
    if (idx->lo_rows == idx->size && idx->lo != nullptr) return LB_OK;
    int rc = idx->v_lo.ensure((size_t)idx->size * idx->dim * 4);
    if (rc) return rc;
    idx->lo = reinterpret_cast<float*>(idx->v_lo.ptr());
    idx->lo_cap = idx->size;
    const size_t off = (size_t)idx->lo_rows * idx->dim;
    CK(launch_split_lo((const float*)idx->rows + off, idx->lo + off, (size_t)(idx->size - idx->lo_rows) * idx->dim, st));
    CK(cudaStreamSynchronize(st));  // other threads' streams may use it as soon as the lock drops
    idx->lo_rows = idx->size;
    return LB_OK;
}

static int exact_search_one(lb_index* idx, const void* d_q1, int k, const uint64_t* d_allow, float* d_out_d,
                            int64_t* d_out_l, cudaStream_t st);

// d_cert_flags [nq] / d_cert_count [1] (optional, device): certification of the coarse stage, see RescoreArgs
static int search_core(lb_index* idx, const void* d_q, int64_t nq, int k, const uint64_t* d_allow, float* d_dist,
                       int64_t* d_lab, cudaStream_t st, uint32_t* d_cert_flags = nullptr,
                       uint32_t* d_cert_count = nullptr) {
    if (nq == 0) return LB_OK;
    const int kc = coarse_k(k);
    if (exhaustive_plan(nq, k)) {
        // k beyond the fused selector (k > 704), or one query with a large k: exhaustive exact kernel, one query at a time
        if (k > 2048) return fail(LB_ERR_UNSUPPORTED, "k > 2048");
        if (idx->size == 0) {
            Scratch scr0(st);
            uint64_t* merged0;
            CK(scr0.get((void**)&merged0, 8));
            CK(launch_unpack_topk(merged0, (int)nq, 0, k, 0, d_dist, d_lab, st));
            return LB_OK;
        }
        const size_t qstride = (size_t)idx->dim * dtype_size(idx->dtype);
        for (int64_t q = 0; q < nq; q++) {
            int rc = exact_search_one(idx, (const char*)d_q + (size_t)q * qstride, k, d_allow, d_dist + (size_t)q * k,
                                      d_lab + (size_t)q * k, st);
            if (rc) return rc;
        }
        return LB_OK;
    }
    Scratch scr(st);
    if (idx->size == 0) {
        uint64_t* merged;
        CK(scr.get((void**)&merged, 8));
        CK(launch_unpack_topk(merged, (int)nq, 0, k, 0, d_dist, d_lab, st));
        return LB_OK;
    }
    // queries are processed in chunks so scratch stays bounded
    const int64_t qchunk = 4096;
    for (int64_t qo = 0; qo < nq; qo += qchunk) {
        const int cq = (int)((nq - qo) < qchunk ? (nq - qo) : qchunk);
        ScanArgs a;
        a.dtype = idx->dtype; a.metric = idx->metric;
        a.db = idx->rows; a.aux = idx->aux; a.n_rows = (uint32_t)idx->size; a.dim = idx->dim;
        a.queries = (const char*)d_q + (size_t)qo * idx->dim * dtype_size(idx->dtype);
        a.nq = cq;
        a.tomb = idx->tomb; a.tomb_bits = (uint32_t)(idx->tomb_bits > 0xffffffffll ? 0xffffffffll : idx->tomb_bits);
        a.allow = (const uint32_t*)d_allow;
        a.kc = kc;
        const int mode = g_opt_dense_scan.load(std::memory_order_relaxed);
        const bool tc_ok = dense_tc_eligible(idx->dtype, idx->dim, idx->rows, a.queries, kc) &&
                           (idx->dtype != DT_F32 || g_opt_f32_tc.load(std::memory_order_relaxed));
        if (mode == 2 && !tc_ok) return fail(LB_ERR_UNSUPPORTED, "tensor-core scan not eligible for this index");
        const bool stream_ok = dense_stream_eligible(idx->dtype, idx->dim, idx->rows, cq, kc);
        if (mode == 3 && !stream_ok) return fail(LB_ERR_UNSUPPORTED, "streaming scan not eligible for this search");
        const bool use_stream = stream_ok && (mode == 3 || (mode == 0 && cq == 1));  // single query: HBM-bound streaming kernel
        // (measured: from 2 queries up the tensor-core scan with one padded query block is already faster)
        const bool use_tc = !use_stream && tc_ok && mode != 1 && mode != 3;
        uint64_t *partial, *merged;
        int parts;
        bool merged_done = false;
        bool simt_keys = false;  // the SIMT scan ranks L2 by |q - x|^2, the other scans by |x|^2 - 2 q.x
        if (use_stream) {
            a.debug = 0;
            // bootstrap sample ~1% of the rows (2048..8192), skipped for small indexes
            int64_t S = (idx->size / 128 + 255) / 256 * 256;
            if (S < 2048) S = 2048;
            if (S > 8192) S = 8192;
            if (!g_opt_tc_boot.load(std::memory_order_relaxed) || S * 4 > idx->size) S = 0;
            int grid = dense_stream_grid(idx->sm_count, cq);
            const int64_t max_grid = (idx->size - S + 255) / 256;
            if (grid > max_grid) grid = (int)max_grid;
            if (grid < 1) grid = 1;
            parts = 2;  // (unused by the compact merge below)
            // per query: the sample's kc best (head) + one compact list every CTA appends its survivors to
            const size_t stride = (size_t)grid * kc;
            uint64_t *head = nullptr, *compact;
            uint32_t* out_cnt;
            CK(scr.get((void**)&compact, (size_t)cq * stride * 8));
            CK(scr.get((void**)&out_cnt, (size_t)cq * 4));
            CK(cudaMemsetAsync(out_cnt, 0, (size_t)cq * 4, st));
            if (S) {
                float *keys, *tau, *edges;
                uint32_t* edge_cnt;
                int* sel_done;
                CK(scr.get((void**)&head, (size_t)cq * kc * 8));
                CK(scr.get((void**)&sel_done, (size_t)cq * 4));
                CK(scr.get((void**)&keys, (size_t)cq * S * 4));
                CK(scr.get((void**)&tau, (size_t)cq * 4));
                CK(scr.get((void**)&edges, (size_t)cq * LB_NEDGE * 4));
                CK(scr.get((void**)&edge_cnt, (size_t)cq * LB_NEDGE * 4));
                ScanArgs b = a;
                b.keys_out = keys; b.keys_ld = (int)S; b.partial = nullptr;
                int gb = (int)((S + 255) / 256);
                CK(launch_dense_scan_stream(b, gb, 0, (uint32_t)S, nullptr, 0, st));
                CK(launch_sample_select(keys, (int)S, (int)S, a.n_rows, a.tomb, a.tomb_bits, a.allow, cq, kc, head, tau,
                                        edges, edge_cnt, sel_done, st));
                a.edges = edges; a.edge_cnt = edge_cnt; a.tau_init = tau;
            }
            a.partial = compact;
            {
                ProfScope prof(st, (double)cq * (double)(idx->size - S));
                CK(launch_dense_scan_stream(a, grid, (uint32_t)S, 0, out_cnt, stride, st));
            }
            CK(scr.get((void**)&merged, (size_t)cq * kc * 8));
            CK(launch_merge_select_compact(head, compact, out_cnt, stride, cq, kc, merged, a.edges, a.edge_cnt, st));
            merged_done = true;
        } else if (use_tc) {
            if (idx->dtype == DT_F32) {
                int rc = ensure_lo(idx, st);
                if (rc) return rc;
                float* qlo;
                CK(scr.get((void**)&qlo, (size_t)cq * idx->dim * 4));
                CK(launch_split_lo((const float*)a.queries, qlo, (size_t)cq * idx->dim, st));
                a.db_lo = idx->lo; a.queries_lo = qlo;
            }
            const int n_tiles = (int)((idx->size + 255) / 256);
            a.tq = 128; a.cap = 0; a.rows_per_part = 0;
            a.debug = g_opt_tc_debug.load(std::memory_order_relaxed);
            // Bootstrap: the first boot_tiles row tiles are scanned in "dump keys" mode into a small
            // [nq][S] matrix; sample_select gives each query its kc best sample rows, their kc-th key (a
            // valid upper bound of the global kc-th best: the scan's starting threshold) and the ladder of
            // sample keys that seeds the shared progressive threshold.  Small indexes skip it.
            int boot_tiles = 0;
            size_t cand_bytes;
            int gm;
            dense_scan_tc_plan(cq, n_tiles, idx->sm_count, kc, &gm, &cand_bytes);
            if (g_opt_tc_boot.load(std::memory_order_relaxed)) {
                int bt = g_opt_tc_boot_tiles.load(std::memory_order_relaxed);
                if (bt <= 0) {  // auto: ~1% of the index, 8..32 tiles (measured optimum at C2: 32 of 3907)
                    bt = n_tiles / 128;
                    if (bt < 8) bt = 8;
                    if (bt > 32) bt = 32;
                }
                if (bt > 128) bt = 128;
                if (bt * 4 <= n_tiles) boot_tiles = bt;
            }
            dense_scan_tc_plan(cq, n_tiles - boot_tiles, idx->sm_count, kc, &gm, &cand_bytes);
            uint64_t* cand;
            CK(scr.get((void**)&cand, cand_bytes));
            const int extra = boot_tiles ? 1 : 0;
            parts = 2 * gm + extra;
            CK(scr.get((void**)&partial, (size_t)parts * cq * kc * 8));
            a.partial = partial;
            if (boot_tiles) {
                const int S = boot_tiles * 256;
                float *keys, *tau, *edges;
                uint32_t* edge_cnt;
                int* sel_done;
                CK(scr.get((void**)&sel_done, (size_t)cq * 4));
                CK(scr.get((void**)&keys, (size_t)cq * S * 4));
                CK(scr.get((void**)&tau, (size_t)cq * 4));
                CK(scr.get((void**)&edges, (size_t)cq * LB_NEDGE * 4));
                CK(scr.get((void**)&edge_cnt, (size_t)cq * LB_NEDGE * 4));
                ScanArgs b = a;
                b.parts = 0; b.partial = nullptr; b.tile_begin = 0; b.tile_end = boot_tiles; b.part_offset = 0;
                b.keys_out = keys; b.keys_ld = S;
                CK(launch_dense_scan_tc(b, idx->sm_count, cand, st));
                CK(launch_sample_select(keys, S, S, a.n_rows, a.tomb, a.tomb_bits, a.allow, cq, kc, partial, tau, edges,
                                        edge_cnt, sel_done, st));
                a.edges = edges; a.edge_cnt = edge_cnt;
                a.tau_init = tau;
            }
            a.parts = 2 * gm; a.tile_begin = boot_tiles; a.tile_end = n_tiles; a.part_offset = extra;
            ProfScope prof(st, (double)cq * (double)(idx->size - (int64_t)boot_tiles * 256));  // main scan only
            CK(launch_dense_scan_tc(a, idx->sm_count, cand, st));
        } else {
            a.tq = (kc <= 128 && cq > 32) ? 64 : (kc <= 384 && cq > 16) ? 32 : 16;
            if (kc > 384) a.tq = 16; else if (kc > 128 && a.tq == 64) a.tq = 32;
            a.cap = next_pow2(kc + 128);
            const int qblocks = (cq + a.tq - 1) / a.tq;
            int target = 2 * idx->sm_count;
            parts = (target + qblocks - 1) / qblocks;
            int64_t max_parts = (idx->size + 511) / 512;
            if (parts > max_parts) parts = (int)max_parts;
            if (parts < 1) parts = 1;
            int64_t rpp = (idx->size + parts - 1) / parts;
            rpp = ((rpp + 127) / 128) * 128;
            parts = (int)((idx->size + rpp - 1) / rpp);
            a.parts = parts; a.rows_per_part = (uint32_t)rpp;
            CK(scr.get((void**)&partial, (size_t)parts * cq * kc * 8));
            a.partial = partial;
            ProfScope prof(st, (double)cq * (double)idx->size);
            CK(launch_dense_scan_simt(a, st));
            simt_keys = true;
        }
        if (merged_done) {
        } else if (parts > 1) {
            CK(scr.get((void**)&merged, (size_t)cq * kc * 8));
            // the re-score stage sorts its kc exact distances, so an unordered top-kc is enough
            CK(launch_merge_select(partial, parts, cq, kc, merged, nullptr, a.edges, a.edge_cnt, st));
        } else {
            merged = partial;
        }
        RescoreArgs r;
        r.nrm = idx->nrm;
        r.dtype = idx->dtype; r.metric = idx->metric; r.db = idx->rows; r.n_rows = (uint32_t)idx->size;
        r.dim = idx->dim; r.queries = a.queries; r.nq = cq; r.packed = merged; r.ids32 = nullptr;
        r.c = kc; r.k = k; r.tomb = nullptr; r.tomb_bits = 0; r.allow = nullptr; r.id_base = idx->id_base;
        r.out_d = d_dist + (size_t)qo * k; r.out_l = d_lab + (size_t)qo * k; r.negate_dot = 1;
        // int8 coarse keys are integers carried in fp32: exact (nothing to certify, ties are ordered by the packed
        // (key, id) compare) while |x|^2 + 2|q.x| < 2^24, i.e. dim * 127^2 * 3 < 2^24; longer int8 rows round and
        // are certified like the float types
        const bool keys_exact = idx->dtype == DT_I8 && (int64_t)idx->dim * 16129 * 3 < (1 << 24);
        if ((d_cert_flags != nullptr || d_cert_count != nullptr) && keys_exact) {
            if (d_cert_flags) CK(cudaMemsetAsync(d_cert_flags + qo, 0, (size_t)cq * 4, st));
        } else if (d_cert_flags != nullptr && idx->max_norm2 != nullptr) {
            r.cert_flags = d_cert_flags + qo; r.cert_count = d_cert_count; r.max_norm2 = idx->max_norm2;
            r.key_space = (simt_keys && idx->metric == METRIC_L2) ? 1 : 0;
            // coarse-key error bound relative to |q||x|: truncating fp32 accumulation over dim terms (measured
            // ~4e-6 at dim 256, test_coarse_keys_accuracy), plus the 3xTF32 operand residue for fp32 rows
            float beta = (float)idx->dim * 1.2e-7f;
            if (beta < 8e-6f) beta = 8e-6f;
            if (idx->dtype == DT_F32) beta += 1e-6f;
            if (idx->dtype == DT_I8) beta = 1.2e-7f;  // integer products, one fp32 rounding of the total
            r.beta = beta;
        }
        CK(launch_rescore(r, st));
    }
    return LB_OK;
}

// Exhaustive exact search of ONE query (device pointers in and out): the reference's arithmetic for every row,
// bitmaps applied, k smallest by (distance, row).  Slow (one thread per row over the whole index); used for the
// queries the certification flags and for k beyond the fused selector (e.g. SearchHybrid's k * 10 candidates,
// internal/store/hnsw_gpu.go:85).
static int exact_search_one(lb_index* idx, const void* d_q1, int k, const uint64_t* d_allow, float* d_out_d,
                            int64_t* d_out_l, cudaStream_t st) {
    Scratch scr(st);
    const int64_t n = idx->size;
    if (k > 2048) return fail(LB_ERR_UNSUPPORTED, "k > 2048");
    float* d_o;
    uint64_t *p, *m;
    CK(scr.get((void**)&d_o, (size_t)n * 4));
    CK(scr.get((void**)&p, select_k_scratch_entries(n, k) * 8));
    CK(scr.get((void**)&m, (size_t)k * 8));
    CK(launch_batch_flat(idx->metric, idx->dtype, idx->rows, n, idx->dim, d_q1, d_o, 1, st));
    CK(launch_mask_rows(d_o, n, idx->tomb, (uint32_t)(idx->tomb_bits > 0xffffffffll ? 0xffffffffll : idx->tomb_bits),
                        (const uint32_t*)d_allow, st));
    CK(launch_select_k(d_o, n, k, p, m, d_out_l, d_out_d, idx->id_base, st));
    return LB_OK;
}

}  // extern "C"
namespace lb {
int api_search_core(lb_index* idx, const void* d_q, int64_t nq, int k, const uint64_t* d_allow, float* d_dist,
                    int64_t* d_lab, cudaStream_t st, uint32_t* d_flags, uint32_t* d_count) {
    if (exhaustive_plan(nq, k) || idx->size == 0) d_flags = d_count = nullptr;  // exhaustive / empty: nothing to certify
    return search_core(idx, d_q, nq, k, d_allow, d_dist, d_lab, st, d_flags, d_count);
}
int api_exact_search_one(lb_index* idx, const void* d_q1, int k, const uint64_t* d_allow, float* d_out_d,
                         int64_t* d_out_l, cudaStream_t st) {
    return exact_search_one(idx, d_q1, k, d_allow, d_out_d, d_out_l, st);
}
}  // namespace lb
extern "C" {

int lb_index_search_device(lb_index* idx, const void* d_queries, int64_t nq, int k, const uint64_t* d_allow,
                           float* d_distances, int64_t* d_labels, void* stream) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive and nq non-negative");
    if (nq > 0 && (!d_queries || !d_distances || !d_labels)) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(idx->device);
    if (rc) return rc;
    return search_core(idx, d_queries, nq, k, d_allow, d_distances, d_labels, (cudaStream_t)stream);
}

int lb_index_search_device_cert(lb_index* idx, const void* d_queries, int64_t nq, int k, const uint64_t* d_allow,
                                float* d_distances, int64_t* d_labels, uint32_t* d_uncert_flags,
                                uint32_t* d_uncert_count, void* stream) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive and nq non-negative");
    if (nq > 0 && (!d_queries || !d_distances || !d_labels)) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(idx->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (nq == 0) return LB_OK;
    if (d_uncert_flags == nullptr && d_uncert_count == nullptr)
        return search_core(idx, d_queries, nq, k, d_allow, d_distances, d_labels, st);
    if (exhaustive_plan(nq, k) || idx->size == 0) {  // exhaustive exact path: nothing to certify
        if (d_uncert_flags) CK(cudaMemsetAsync(d_uncert_flags, 0, (size_t)nq * 4, st));
        return search_core(idx, d_queries, nq, k, d_allow, d_distances, d_labels, st);
    }
    // the re-score kernel needs both outputs; a caller that wants only one gets the other from scratch
    Scratch scr(st);
    uint32_t* flags = d_uncert_flags;
    uint32_t* count = d_uncert_count;
    if (!flags) CK(scr.get((void**)&flags, (size_t)nq * 4));
    if (!count) { CK(scr.get((void**)&count, 4)); CK(cudaMemsetAsync(count, 0, 4, st)); }
    return search_core(idx, d_queries, nq, k, d_allow, d_distances, d_labels, st, flags, count);
}

int lb_index_search_exact_device(lb_index* idx, const void* d_queries, int64_t nq, int k, const uint64_t* d_allow,
                                 const uint32_t* h_flags, float* d_distances, int64_t* d_labels, void* stream) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive and nq non-negative");
    if (k > 2048) return fail(LB_ERR_UNSUPPORTED, "k > 2048");
    if (nq > 0 && (!d_queries || !d_distances || !d_labels)) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(idx->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t qstride = (size_t)idx->dim * dtype_size(idx->dtype);
    for (int64_t q = 0; q < nq; q++) {
        if (h_flags && !h_flags[q]) continue;
        if (idx->size == 0) {
            Scratch scr0(st);
            uint64_t* m0;
            CK(scr0.get((void**)&m0, 8));
            CK(launch_unpack_topk(m0, 1, 0, k, 0, d_distances + (size_t)q * k, d_labels + (size_t)q * k, st));
            continue;
        }
        rc = exact_search_one(idx, (const char*)d_queries + (size_t)q * qstride, k, d_allow, d_distances + (size_t)q * k,
                              d_labels + (size_t)q * k, st);
        if (rc) return rc;
    }
    return LB_OK;
}

int lb_index_search(lb_index* idx, const void* queries, int64_t nq, int k, const uint64_t* allow, float* distances,
                    int64_t* labels) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive and nq non-negative");
    if (nq == 0) return LB_OK;
    if (!queries || !distances || !labels) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(idx->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t qb = (size_t)nq * idx->dim * dtype_size(idx->dtype);
    void* d_q; float* d_d; int64_t* d_l; uint64_t* d_allow = nullptr;
    CK(scr.get(&d_q, qb));
    CK(scr.get((void**)&d_d, (size_t)nq * k * 4));
    CK(scr.get((void**)&d_l, (size_t)nq * k * 8));
    CK(cudaMemcpyAsync(d_q, queries, qb, cudaMemcpyHostToDevice, st));
    if (allow) {
        size_t words = (size_t)((idx->size + 63) / 64);
        CK(scr.get((void**)&d_allow, words * 8));
        CK(cudaMemcpyAsync(d_allow, allow, words * 8, cudaMemcpyHostToDevice, st));
    }
    // certification: queries whose candidate margin does not cover the coarse error bound are re-done exactly
    uint32_t *d_flags = nullptr, *d_count = nullptr;
    if (idx->max_norm2 != nullptr && g_opt_certify.load(std::memory_order_relaxed)) {
        CK(scr.get((void**)&d_flags, (size_t)nq * 4));
        CK(scr.get((void**)&d_count, 4));
        CK(cudaMemsetAsync(d_count, 0, 4, st));  // (every flag is written by its re-score block)
    }
    rc = search_core(idx, d_q, nq, k, d_allow, d_d, d_l, st, d_flags, d_count);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    // the 4-byte count lands in a per-thread pinned word (a copy into pageable memory would stall every stream)
    static thread_local uint32_t* t_pinned = nullptr;
    if (d_count && t_pinned == nullptr && cudaHostAlloc((void**)&t_pinned, 64, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        t_pinned = nullptr;
        d_count = nullptr;  // cannot report: skip the fallback rather than slow every call down
    }
    CK(cudaMemcpyAsync(distances, d_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(labels, d_l, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    if (d_count) CK(cudaMemcpyAsync(t_pinned, d_count, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_uncert = d_count ? *t_pinned : 0u;
    idx->last_uncertified.store((int64_t)n_uncert);
    if (n_uncert > 0) {
        std::vector<uint32_t> flags((size_t)nq);
        CK(cudaMemcpy(flags.data(), d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        const size_t qstride = (size_t)idx->dim * dtype_size(idx->dtype);
        for (int64_t q = 0; q < nq; q++) {
            if (!flags[(size_t)q]) continue;
            rc = exact_search_one(idx, (const char*)d_q + (size_t)q * qstride, k, d_allow, d_d + (size_t)q * k,
                                  d_l + (size_t)q * k, st);
            if (rc) return rc;
            CK(cudaMemcpyAsync(distances + (size_t)q * k, d_d + (size_t)q * k, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(labels + (size_t)q * k, d_l + (size_t)q * k, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
    }
    return LB_OK;
}

static int rerank_core(lb_index* idx, const void* d_q, int64_t nq, const uint32_t* d_ids, int c, int k,
                       const uint64_t* d_allow, float* d_dist, int64_t* d_lab, cudaStream_t st) {
    if (nq == 0) return LB_OK;
    if (c > 1024) return fail(LB_ERR_UNSUPPORTED, "more than 1024 candidates per query");
    RescoreArgs r;
    r.nrm = idx->nrm;
    r.dtype = idx->dtype; r.metric = idx->metric; r.db = idx->rows; r.n_rows = (uint32_t)idx->size;
    r.dim = idx->dim; r.queries = d_q; r.nq = (int)nq; r.packed = nullptr; r.ids32 = d_ids;
    r.c = c; r.k = k; r.tomb = idx->tomb;
    r.tomb_bits = (uint32_t)(idx->tomb_bits > 0xffffffffll ? 0xffffffffll : idx->tomb_bits);
    r.allow = (const uint32_t*)d_allow; r.id_base = idx->id_base; r.out_d = d_dist; r.out_l = d_lab;
    r.negate_dot = 1;
    CK(launch_rescore(r, st));
    return LB_OK;
}

}  // extern "C"
namespace lb {
int api_rerank_device(lb_index* idx, const void* d_q, int64_t nq, const uint32_t* d_ids, int c, int k,
                      const uint64_t* d_allow, float* d_dist, int64_t* d_lab, cudaStream_t st) {
    return rerank_core(idx, d_q, nq, d_ids, c, k, d_allow, d_dist, d_lab, st);
}
}  // namespace lb
extern "C" {

int lb_index_rerank_device(lb_index* idx, const void* d_queries, int64_t nq, const uint32_t* d_cand_ids, int c,
                           int k, const uint64_t* d_allow, float* d_distances, int64_t* d_labels, void* stream) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0 || c <= 0) return fail(LB_ERR_INVALID, "k, c must be positive");
    int rc = use_device(idx->device);
    if (rc) return rc;
    return rerank_core(idx, d_queries, nq, d_cand_ids, c, k, d_allow, d_distances, d_labels, (cudaStream_t)stream);
}

int lb_index_rerank(lb_index* idx, const void* queries, int64_t nq, const uint32_t* cand_ids, int c, int k,
                    const uint64_t* allow, float* distances, int64_t* labels) {
    if (!idx) return fail(LB_ERR_INVALID, "index is NULL");
    if (k <= 0 || nq < 0 || c <= 0) return fail(LB_ERR_INVALID, "k, c must be positive");
    if (nq == 0) return LB_OK;
    if (!queries || !cand_ids || !distances || !labels) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(idx->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t qb = (size_t)nq * idx->dim * dtype_size(idx->dtype);
    void* d_q; uint32_t* d_ids; float* d_d; int64_t* d_l; uint64_t* d_allow = nullptr;
    CK(scr.get(&d_q, qb));
    CK(scr.get((void**)&d_ids, (size_t)nq * c * 4));
    CK(scr.get((void**)&d_d, (size_t)nq * k * 4));
    CK(scr.get((void**)&d_l, (size_t)nq * k * 8));
    CK(cudaMemcpyAsync(d_q, queries, qb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ids, cand_ids, (size_t)nq * c * 4, cudaMemcpyHostToDevice, st));
    if (allow) {
        size_t words = (size_t)((idx->size + 63) / 64);
        CK(scr.get((void**)&d_allow, words * 8));
        CK(cudaMemcpyAsync(d_allow, allow, words * 8, cudaMemcpyHostToDevice, st));
    }
    rc = rerank_core(idx, d_q, nq, d_ids, c, k, d_allow, d_d, d_l, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(distances, d_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(labels, d_l, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_index_distances(lb_index* idx, const void* query, float* out) {
    if (!idx || !query || !out) return fail(LB_ERR_INVALID, "NULL argument");
    int rc = use_device(idx->device);
    if (rc) return rc;
    if (idx->size == 0) return LB_OK;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t qb = (size_t)idx->dim * dtype_size(idx->dtype);
    void* d_q; float* d_o;
    CK(scr.get(&d_q, qb));
    CK(scr.get((void**)&d_o, (size_t)idx->size * 4));
    CK(cudaMemcpyAsync(d_q, query, qb, cudaMemcpyHostToDevice, st));
    CK(launch_batch_flat(idx->metric, idx->dtype, idx->rows, idx->size, idx->dim, d_q, d_o, 1, st));
    CK(cudaMemcpyAsync(out, d_o, (size_t)idx->size * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

// Diagnostics: the COARSE ranking keys the tensor-core scan computes for rows [0, n_rows) -- |x|^2 - 2 q.x (L2),
// -q.x / |x| (cosine), -q.x (dot) -- so tests can bound the coarse error against float64 (these keys only
// rank candidates; every returned distance comes from the exact re-score).
int lb_index_coarse_keys(lb_index* idx, const void* queries, int64_t nq, int64_t n_rows, float* out) {
    if (!idx || !queries || !out) return fail(LB_ERR_INVALID, "NULL argument");
    if (nq <= 0 || n_rows <= 0 || n_rows > idx->size) return fail(LB_ERR_INVALID, "bad size");
    int rc = use_device(idx->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    const size_t qb = (size_t)nq * idx->dim * dtype_size(idx->dtype);
    void* d_q;
    CK(scr.get(&d_q, qb));
    CK(cudaMemcpyAsync(d_q, queries, qb, cudaMemcpyHostToDevice, st));
    if (!dense_tc_eligible(idx->dtype, idx->dim, idx->rows, d_q, 32))
        return fail(LB_ERR_UNSUPPORTED, "tensor-core scan not eligible for this index");
    const int tiles = (int)((n_rows + 255) / 256);
    const int S = tiles * 256;
    ScanArgs a;
    a.dtype = idx->dtype; a.metric = idx->metric; a.db = idx->rows; a.aux = idx->aux;
    a.n_rows = (uint32_t)idx->size; a.dim = idx->dim; a.queries = d_q; a.nq = (int)nq;
    a.tomb = nullptr; a.tomb_bits = 0; a.allow = nullptr; a.kc = 32; a.cap = 0; a.tq = 128; a.rows_per_part = 0;
    if (idx->dtype == DT_F32) {
        rc = ensure_lo(idx, st);
        if (rc) return rc;
        float* qlo;
        CK(scr.get((void**)&qlo, (size_t)nq * idx->dim * 4));
        CK(launch_split_lo((const float*)d_q, qlo, (size_t)nq * idx->dim, st));
        a.db_lo = idx->lo; a.queries_lo = qlo;
    }
    float* keys;
    uint64_t* cand;
    size_t cand_bytes;
    int gm;
    dense_scan_tc_plan((int)nq, tiles, idx->sm_count, a.kc, &gm, &cand_bytes);
    CK(scr.get((void**)&cand, cand_bytes));
    CK(scr.get((void**)&keys, (size_t)nq * S * 4));
    a.parts = 0; a.partial = nullptr; a.tile_begin = 0; a.tile_end = tiles; a.part_offset = 0;
    a.keys_out = keys; a.keys_ld = S;
    CK(launch_dense_scan_tc(a, idx->sm_count, cand, st));
    CK(cudaMemcpy2DAsync(out, (size_t)n_rows * 4, keys, (size_t)S * 4, (size_t)n_rows * 4, (size_t)nq,
                         cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

// ------------------------------------------------------------------ faiss_gpu_* (faiss_gpu.go:16-21)
void* faiss_gpu_resources_new(int device) {
    if (use_device(device) != LB_OK) return nullptr;
    faiss_res* r = new (std::nothrow) faiss_res();
    if (r) r->device = device;
    return r;
}
void faiss_gpu_resources_free(void* res) { delete (faiss_res*)res; }
void* faiss_gpu_index_flat_l2_new(void* res, int dim) {
    if (!res) { fail(LB_ERR_INVALID, "resources is NULL"); return nullptr; }
    lb_index* idx = nullptr;
    if (lb_index_create(((faiss_res*)res)->device, dim, LB_F32, LB_METRIC_L2, &idx) != LB_OK) return nullptr;
    return idx;
}
void faiss_gpu_index_flat_l2_free(void* idx) { lb_index_free((lb_index*)idx); }
int faiss_gpu_index_add(void* idx, int64_t n, float* vectors) { return lb_index_add((lb_index*)idx, vectors, n); }
int faiss_gpu_index_search(void* idx, int64_t n, float* queries, int k, float* distances, int64_t* labels) {
    return lb_index_search((lb_index*)idx, queries, n, k, nullptr, distances, labels);
}

// ------------------------------------------------------------------------------ stateless simd
int lb_simd_distance_batch_flat(int device, int metric, int dtype, const void* query, const void* flat, int64_t n,
                                int dim, float* results) {
    if (n < 0 || dim <= 0) return fail(LB_ERR_INVALID, "bad size");
    if (n == 0) return LB_OK;  // batch_operations.go:65-67
    if (!query || !flat || !results) return fail(LB_ERR_INVALID, "NULL buffer");
    if (!(dtype == DT_U8 ? metric == METRIC_L2 : supported(dtype, metric)))
        return fail(LB_ERR_UNSUPPORTED, "no kernel for this (metric, dtype)");
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t es = dtype_size(dtype);
    void *d_q, *d_f; float* d_o;
    CK(scr.get(&d_q, (size_t)dim * es));
    CK(scr.get(&d_f, (size_t)n * dim * es));
    CK(scr.get((void**)&d_o, (size_t)n * 4));
    CK(cudaMemcpyAsync(d_q, query, (size_t)dim * es, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_f, flat, (size_t)n * dim * es, cudaMemcpyHostToDevice, st));
    CK(launch_batch_flat(metric, dtype, d_f, n, dim, d_q, d_o, 0, st));
    CK(cudaMemcpyAsync(results, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_simd_adc_distance_batch(int device, const float* table, const uint8_t* flat_codes, int m, int64_t n,
                               float* results) {
    if (!table || !flat_codes) return fail(LB_ERR_INVALID, "simd: empty table or codes");  // batch_operations.go:120
    if (m <= 0) return fail(LB_ERR_INVALID, "simd: invalid m parameter");                   // :123
    if (m > 200) return fail(LB_ERR_UNSUPPORTED, "m too large for a shared-memory LUT");
    if (n <= 0) return LB_OK;
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    float *d_t, *d_o; uint8_t* d_c;
    CK(scr.get((void**)&d_t, (size_t)m * 1024));
    CK(scr.get((void**)&d_c, (size_t)n * m));
    CK(scr.get((void**)&d_o, (size_t)n * 4));
    CK(cudaMemcpyAsync(d_t, table, (size_t)m * 1024, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_c, flat_codes, (size_t)n * m, cudaMemcpyHostToDevice, st));
    CK(launch_adc_batch(d_t, d_c, m, n, d_o, st));
    CK(cudaMemcpyAsync(results, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_select_k(int device, const float* distances, int64_t n, int k, int64_t* out_indices, float* out_distances) {
    if (k <= 0 || n < 0 || !out_indices) return fail(LB_ERR_INVALID, "bad argument");
    if (n > 0 && !distances) return fail(LB_ERR_INVALID, "NULL buffer");
    if (k > 2048) return fail(LB_ERR_UNSUPPORTED, "k > 2048");
    if (n > 0xfff00000ll) return fail(LB_ERR_UNSUPPORTED, "n too large");
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    float *d_d, *d_od; int64_t* d_oi; uint64_t *p, *m;
    CK(scr.get((void**)&d_d, (size_t)n * 4));
    CK(scr.get((void**)&d_od, (size_t)k * 4));
    CK(scr.get((void**)&d_oi, (size_t)k * 8));
    CK(scr.get((void**)&p, select_k_scratch_entries(n, k) * 8));
    CK(scr.get((void**)&m, (size_t)k * 8));
    if (n > 0) CK(cudaMemcpyAsync(d_d, distances, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(launch_select_k(d_d, n, k, p, m, d_oi, d_od, 0, st));
    CK(cudaMemcpyAsync(out_indices, d_oi, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
    if (out_distances) CK(cudaMemcpyAsync(out_distances, d_od, (size_t)k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_merge_topk_device(int device, const float* d_distances, const int64_t* d_labels, int parts, int64_t nq,
                         int k_in, int k, float* d_out_distances, int64_t* d_out_labels, void* stream) {
    if (parts <= 0 || k_in <= 0 || k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "bad argument");
    int rc = use_device(device);
    if (rc) return rc;
    if ((int64_t)parts * k_in > 16384) return fail(LB_ERR_UNSUPPORTED, "parts * k_in > 16384");
    CK(launch_merge_topk(d_distances, d_labels, parts, (int)nq, k_in, k, d_out_distances, d_out_labels,
                         (cudaStream_t)stream));
    return LB_OK;
}

int lb_merge_topk_packed_device(int device, const void* d_records, size_t part_stride, size_t label_offset, int parts,
                                int64_t nq, int k_in, int k, float* d_out_distances, int64_t* d_out_labels,
                                void* stream) {
    if (parts <= 0 || k_in <= 0 || k <= 0 || nq < 0 || !d_records) return fail(LB_ERR_INVALID, "bad argument");
    if ((part_stride & 7) || (label_offset & 7) || label_offset < (size_t)nq * k_in * 4 ||
        part_stride < label_offset + (size_t)nq * k_in * 8)
        return fail(LB_ERR_INVALID, "record layout: [nq*k_in f32 | pad to 8 | nq*k_in i64] per part");
    int rc = use_device(device);
    if (rc) return rc;
    if ((int64_t)parts * k_in > 16384) return fail(LB_ERR_UNSUPPORTED, "parts * k_in > 16384");
    CK(launch_merge_topk_strided(d_records, part_stride, (const char*)d_records + label_offset, part_stride, parts,
                                 (int)nq, k_in, k, d_out_distances, d_out_labels, (cudaStream_t)stream));
    return LB_OK;
}

int lb_merge_topk(int device, const float* distances, const int64_t* labels, int parts, int64_t nq, int k_in, int k,
                  float* out_distances, int64_t* out_labels) {
    if (parts <= 0 || k_in <= 0 || k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "bad argument");
    if (nq == 0) return LB_OK;
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t cnt = (size_t)parts * nq * k_in;
    float *d_d, *d_od; int64_t *d_l, *d_ol;
    CK(scr.get((void**)&d_d, cnt * 4));
    CK(scr.get((void**)&d_l, cnt * 8));
    CK(scr.get((void**)&d_od, (size_t)nq * k * 4));
    CK(scr.get((void**)&d_ol, (size_t)nq * k * 8));
    CK(cudaMemcpyAsync(d_d, distances, cnt * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_l, labels, cnt * 8, cudaMemcpyHostToDevice, st));
    rc = lb_merge_topk_device(device, d_d, d_l, parts, nq, k_in, k, d_od, d_ol, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(out_distances, d_od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_labels, d_ol, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

// --------------------------------------------------------------------------------------- PQ
int lb_pq_create(int device, const void* blob, size_t blob_len, lb_pq** out) {
    if (!out) return fail(LB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!blob || blob_len < 12) return fail(LB_ERR_INVALID, "invalid PQ data: too short");  // persistence.go:40
    uint32_t hdr[3];
    memcpy(hdr, blob, 12);
    int dims = (int)hdr[0], M = (int)hdr[1], K = (int)hdr[2];
    if (M == 0 || dims % M != 0) return fail(LB_ERR_INVALID, "invalid PQ parameters in serialized data");  // :48
    int sub = dims / M;
    size_t expect = 12 + (size_t)M * K * sub * 4;
    if (blob_len != expect) return fail(LB_ERR_INVALID, "invalid PQ data: size mismatch");  // :54
    if (K != 256) return fail(LB_ERR_UNSUPPORTED, "only K = 256 is coherent with simd.ADCDistanceBatch (simd.go:350)");
    if (M > 200) return fail(LB_ERR_UNSUPPORTED, "M too large for a shared-memory LUT (M <= 200)");
    int rc = use_device(device);
    if (rc) return rc;
    lb_pq* pq = new (std::nothrow) lb_pq();
    if (!pq) return fail(LB_ERR_OOM, "host allocation failed");
    pq->device = device; pq->dims = dims; pq->M = M; pq->K = K; pq->sub = sub;
    pq->sm_count = g_dev[device].sm_count;
    cudaError_t e = cudaMalloc((void**)&pq->codebooks, (size_t)M * K * sub * 4);
    if (e != cudaSuccess) { delete pq; return fail_cuda(e, "cudaMalloc(codebooks)"); }
    e = cudaMemcpy(pq->codebooks, (const char*)blob + 12, (size_t)M * K * sub * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(pq->codebooks); delete pq; return fail_cuda(e, "cudaMemcpy(codebooks)"); }
    // fp16 copy for the batched tensor-core coarse stage -- only when every centroid component survives the trip
    // (finite and far from the fp16 range limit; the certification bound assumes relative 2^-11 rounding)
    bool fp16_ok = pq_gemm_eligible(dims, M, sub);
    if (fp16_ok) {
        const float* cbh = reinterpret_cast<const float*>((const char*)blob + 12);
        for (size_t i = 0; i < (size_t)M * K * sub; i++) {
            float v;
            memcpy(&v, cbh + i, 4);
            v = v < 0 ? -v : v;
            if (!(v <= 1.0e4f)) { fp16_ok = false; break; }   // also catches NaN / inf
        }
    }
    if (fp16_ok) {
        e = cudaMalloc(&pq->codebook16, (size_t)M * K * sub * 2);
        if (e == cudaSuccess) e = launch_pq_codebook16(pq->codebooks, pq->codebook16, (size_t)M * K * sub, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaMalloc((void**)&pq->cnorm2, (size_t)M * 256 * 4);
        if (e == cudaSuccess) e = launch_pq_centroid_norms(pq->codebook16, M, sub, pq->cnorm2, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaMalloc((void**)&pq->xmax2, 4);
        if (e == cudaSuccess) e = cudaMemsetAsync(pq->xmax2, 0, 4, cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
        if (e != cudaSuccess) { lb_pq_free(pq); return fail_cuda(e, "fp16 codebooks"); }
    }
    *out = pq;
    return LB_OK;
}

void lb_pq_free(lb_pq* pq) {
    if (!pq) return;
    if (cudaSetDevice(pq->device) == cudaSuccess) {
        cudaDeviceSynchronize();
        if (pq->codebooks) cudaFree(pq->codebooks);
        if (pq->codebook16) cudaFree(pq->codebook16);
        if (pq->cnorm2) cudaFree(pq->cnorm2);
        if (pq->xn2) cudaFree(pq->xn2);
        if (pq->xmax2) cudaFree(pq->xmax2);
        if (pq->codes) cudaFree(pq->codes);
        if (pq->tiled) cudaFree(pq->tiled);
        if (pq->tomb) cudaFree(pq->tomb);
    }
    cudaGetLastError();
    delete pq;
}

int lb_pq_params(const lb_pq* pq, int* dims, int* m, int* k, int* sub_dim) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (dims) *dims = pq->dims;
    if (m) *m = pq->M;
    if (k) *k = pq->K;
    if (sub_dim) *sub_dim = pq->sub;
    return LB_OK;
}

static int pq_add_common(lb_pq* pq, const uint8_t* src, int64_t n, bool on_device, cudaStream_t st) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (n < 0) return fail(LB_ERR_INVALID, "n < 0");
    if (n == 0) return LB_OK;
    if (!src) return fail(LB_ERR_INVALID, "codes is NULL");
    int rc = use_device(pq->device);
    if (rc) return rc;
    if (pq->size + n > 0xfff00000ll) return fail(LB_ERR_INVALID, "more than 2^32 rows per device handle");
    rc = grow((void**)&pq->codes, &pq->capacity, pq->size + n, (size_t)pq->M, pq->size, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(pq->codes + (size_t)pq->size * pq->M, src, (size_t)n * pq->M,
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (pq->M <= 96) {  // the coarse scan's mirror (pq_scan.cu)
        const int Mp = ((pq->M + 31) / 32) * 32;
        const int64_t need = ((pq->capacity + 31) / 32) * 32;
        if (need > pq->tiled_cap) {
            uint8_t* nt = nullptr;
            CK(cudaMalloc((void**)&nt, (size_t)need * Mp));
            CK(cudaDeviceSynchronize());
            if (pq->tiled && pq->size > 0)
                CK(cudaMemcpy(nt, pq->tiled, (size_t)(((pq->size + 31) / 32) * 32) * Mp, cudaMemcpyDeviceToDevice));
            if (pq->tiled) cudaFree(pq->tiled);
            pq->tiled = nt;
            pq->tiled_cap = need;
        }
        CK(launch_pq_tile_codes(pq->codes + (size_t)pq->size * pq->M, n, pq->M, Mp, pq->size, pq->tiled, st));
    }
    if (pq->codebook16 != nullptr) {  // row norms of the decoded vectors (pq_gemm.cu)
        if (pq->capacity > pq->xn2_cap) {
            float* nx = nullptr;
            CK(cudaMalloc((void**)&nx, (size_t)(pq->capacity + 256) * 4));
            CK(cudaDeviceSynchronize());
            CK(cudaMemset(nx, 0, (size_t)(pq->capacity + 256) * 4));
            if (pq->xn2 && pq->size > 0) CK(cudaMemcpy(nx, pq->xn2, (size_t)pq->size * 4, cudaMemcpyDeviceToDevice));
            if (pq->xn2) cudaFree(pq->xn2);
            pq->xn2 = nx;
            pq->xn2_cap = pq->capacity;
        }
        CK(launch_pq_row_norms(pq->codes, pq->cnorm2, pq->M, pq->size, n, pq->xn2, pq->xmax2, st));
    }
    pq->size += n;
    if (!on_device) CK(cudaStreamSynchronize(st));
    return LB_OK;
}
int lb_pq_add_codes(lb_pq* pq, const uint8_t* codes, int64_t n) {
    return pq_add_common(pq, codes, n, false, cudaStreamPerThread);
}
int lb_pq_add_codes_device(lb_pq* pq, const uint8_t* d_codes, int64_t n, void* stream) {
    return pq_add_common(pq, d_codes, n, true, (cudaStream_t)stream);
}
int64_t lb_pq_size(const lb_pq* pq) { return pq ? pq->size : -1; }

int lb_pq_attach_raw(lb_pq* pq, lb_index* raw) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (raw) {
        if (raw->dtype != DT_F32 || raw->metric != METRIC_L2 || raw->dim != pq->dims || raw->device != pq->device)
            return fail(LB_ERR_INVALID, "raw index must be fp32 / L2 / same dims / same device");
    }
    pq->raw = raw;
    return LB_OK;
}

int lb_pq_set_tombstones(lb_pq* pq, const uint64_t* bitmap, int64_t nbits) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    return set_bitmap(pq->device, &pq->tomb, &pq->tomb_bits, bitmap, nbits, false, cudaStreamPerThread, nullptr);
}

int lb_pq_build_adc_table(lb_pq* pq, const float* query, float* table) {
    if (!pq || !query || !table) return fail(LB_ERR_INVALID, "NULL argument");
    int rc = use_device(pq->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    float *d_q, *d_t;
    CK(scr.get((void**)&d_q, (size_t)pq->dims * 4));
    CK(scr.get((void**)&d_t, (size_t)pq->M * pq->K * 4));
    CK(cudaMemcpyAsync(d_q, query, (size_t)pq->dims * 4, cudaMemcpyHostToDevice, st));
    CK(launch_adc_lut(pq->codebooks, pq->M, pq->K, pq->sub, d_q, 1, d_t, st));
    CK(cudaMemcpyAsync(table, d_t, (size_t)pq->M * pq->K * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int lb_pq_encode(lb_pq* pq, const float* vectors, int64_t n, uint8_t* codes) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (n < 0) return fail(LB_ERR_INVALID, "n < 0");
    if (n == 0) return LB_OK;
    if (!vectors || !codes) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(pq->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    float* d_v; uint8_t* d_c;
    CK(scr.get((void**)&d_v, (size_t)n * pq->dims * 4));
    CK(scr.get((void**)&d_c, (size_t)n * pq->M));
    CK(cudaMemcpyAsync(d_v, vectors, (size_t)n * pq->dims * 4, cudaMemcpyHostToDevice, st));
    CK(launch_pq_encode(pq->codebooks, pq->M, pq->K, pq->sub, d_v, n, d_c, st));
    CK(cudaMemcpyAsync(codes, d_c, (size_t)n * pq->M, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

}  // extern "C"

// memset-style fill of int32 words on the device
static __global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" {

int lb_pq_train(int device, const float* vectors, int64_t n, int dims, int m, int k, int max_iter,
                const int32_t* init_idx, float* codebooks, int32_t* iters_run) {
    if (!vectors || !init_idx || !codebooks) return fail(LB_ERR_INVALID, "NULL buffer");
    if (n <= 0) return fail(LB_ERR_INVALID, "empty training data");                         // encoder.go:41-43
    if (m <= 0 || dims <= 0 || dims % m != 0) return fail(LB_ERR_INVALID, "dimension must be divisible by M");
    if (k <= 0 || n < k) return fail(LB_ERR_INVALID, "insufficient data for k-means: n < k");  // kmeans.go:65-67
    if (k > 256) return fail(LB_ERR_UNSUPPORTED, "K > 256 (codes are bytes)");
    if (dims / m > 128) return fail(LB_ERR_UNSUPPORTED, "sub-vector dimension > 128");
    if (max_iter <= 0) max_iter = 20;                                                       // encoder.go:62
    for (int64_t i = 0; i < (int64_t)m * k; i++)
        if (init_idx[i] < 0 || init_idx[i] >= n) return fail(LB_ERR_INVALID, "init_idx out of range");
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    const int sub = dims / m;
    float *d_data, *d_cent;
    int32_t *d_init, *d_assign, *d_active, *d_iters;
    uint32_t* d_changed;
    CK(scr.get((void**)&d_data, (size_t)n * dims * 4));
    CK(scr.get((void**)&d_cent, (size_t)m * k * sub * 4));
    CK(scr.get((void**)&d_init, (size_t)m * k * 4));
    CK(scr.get((void**)&d_assign, (size_t)m * n * 4));
    CK(scr.get((void**)&d_active, (size_t)m * 4));
    CK(scr.get((void**)&d_iters, (size_t)m * 4));
    CK(scr.get((void**)&d_changed, (size_t)m * 4));
    CK(cudaMemcpyAsync(d_data, vectors, (size_t)n * dims * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_init, init_idx, (size_t)m * k * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_changed, 0, (size_t)m * 4, st));
    CK(cudaMemsetAsync(d_iters, 0, (size_t)m * 4, st));
    const int64_t na = (int64_t)m * n;
    fill_i32_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(d_assign, na, -1);
    fill_i32_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_active, m, 1);
    count_launch(); count_launch();
    CK(launch_pq_train(d_data, n, dims, m, k, max_iter, d_init, d_cent, d_assign, d_changed, d_active, d_iters, st));
    CK(cudaMemcpyAsync(codebooks, d_cent, (size_t)m * k * sub * 4, cudaMemcpyDeviceToHost, st));
    if (iters_run) CK(cudaMemcpyAsync(iters_run, d_iters, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}


static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// The round-1 path: every code row gets the reference's sequential fp32 sum (adc_scan_kernel).  Exact keys, no
// certification needed.  Still the path for M > 96, very large k', and the repair of uncertified queries.
static int pq_search_exhaustive(lb_pq* pq, const float* q, int cq, const float* luts, int k, int kc, bool rerank,
                                const uint64_t* d_allow, float* d_dist, int64_t* d_lab, Scratch& scr, cudaStream_t st) {
    PqScanArgs a;
    a.codes = pq->codes; a.n_rows = (uint32_t)pq->size; a.M = pq->M; a.luts = luts; a.nq = cq;
    a.tomb = pq->tomb; a.tomb_bits = (uint32_t)(pq->tomb_bits > 0xffffffffll ? 0xffffffffll : pq->tomb_bits);
    a.allow = (const uint32_t*)d_allow;
    a.kc = kc; a.cap = next_pow2(kc + 256);
    int target = 4 * pq->sm_count;
    int parts = (target + cq - 1) / cq;
    int64_t max_parts = (pq->size + 2047) / 2048;
    if (parts > max_parts) parts = (int)max_parts;
    if (parts < 1) parts = 1;
    int64_t rpp = (pq->size + parts - 1) / parts;
    rpp = ((rpp + 255) / 256) * 256;
    parts = (int)((pq->size + rpp - 1) / rpp);
    a.parts = parts; a.rows_per_part = (uint32_t)rpp;
    uint64_t *partial, *merged;
    CK(scr.get((void**)&partial, (size_t)parts * cq * kc * 8));
    a.partial = partial;
    {
        ProfScope prof(st, (double)cq * (double)pq->size);
        CK(launch_adc_scan(a, st));
    }
    if (parts > 1) {
        CK(scr.get((void**)&merged, (size_t)cq * kc * 8));
        CK(launch_merge_partials(partial, parts, cq, kc, merged, st));
    } else {
        merged = partial;
    }
    if (rerank) {
        RescoreArgs r;
        r.dtype = DT_F32; r.metric = METRIC_L2; r.db = pq->raw->rows; r.n_rows = (uint32_t)pq->raw->size;
        r.dim = pq->dims; r.queries = q; r.nq = cq; r.packed = merged; r.ids32 = nullptr; r.c = kc; r.k = k;
        r.tomb = nullptr; r.tomb_bits = 0; r.allow = nullptr; r.id_base = 0;
        r.out_d = d_dist; r.out_l = d_lab; r.negate_dot = 1;
        CK(launch_rescore(r, st));
    } else {
        CK(launch_unpack_topk(merged, cq, kc, k, 0, d_dist, d_lab, st));
    }
    return LB_OK;
}

// d_flags [nq] / d_count [1] (device, optional): certification of the coarse pass (pq_scan.cu)

// Coarse candidates of a dense fp16 / L2 view through the tensor-core scan: bootstrap sample -> shared-threshold
// scan -> merge-select.  `a` names the view (db, aux, n_rows, dim), the queries and the bitmaps; *merged_out receives
// [nq][kc] packed (key, row), the kc smallest by (key, row) in no particular order.  The same chain search_core runs
// for a dense index (kept separate: that one also handles the streaming / SIMT / 3xTF32 cases).
static int tc_coarse_candidates(ScanArgs a, int kc, int sm_count, Scratch& scr, cudaStream_t st, uint64_t** merged_out) {
    const int cq = a.nq;
    const int n_tiles = (int)(((int64_t)a.n_rows + 255) / 256);
    a.kc = kc; a.tq = 128; a.cap = 0; a.rows_per_part = 0;
    a.debug = g_opt_tc_debug.load(std::memory_order_relaxed);
    int boot_tiles = 0, gm;
    size_t cand_bytes;
    if (g_opt_tc_boot.load(std::memory_order_relaxed)) {
        int bt = g_opt_tc_boot_tiles.load(std::memory_order_relaxed);
        if (bt <= 0) {
            bt = n_tiles / 128;
            if (bt < 8) bt = 8;
            if (bt > 32) bt = 32;
        }
        if (bt > 128) bt = 128;
        if (bt * 4 <= n_tiles) boot_tiles = bt;
    }
    dense_scan_tc_plan(cq, n_tiles - boot_tiles, sm_count, kc, &gm, &cand_bytes);
    uint64_t *cand, *partial, *merged;
    CK(scr.get((void**)&cand, cand_bytes));
    const int extra = boot_tiles ? 1 : 0;
    const int parts = 2 * gm + extra;
    CK(scr.get((void**)&partial, (size_t)parts * cq * kc * 8));
    a.partial = partial;
    if (boot_tiles) {
        const int S = boot_tiles * 256;
        float *keys, *tau, *edges;
        uint32_t* edge_cnt;
        int* sel_done;
        CK(scr.get((void**)&sel_done, (size_t)cq * 4));
        CK(scr.get((void**)&keys, (size_t)cq * S * 4));
        CK(scr.get((void**)&tau, (size_t)cq * 4));
        CK(scr.get((void**)&edges, (size_t)cq * LB_NEDGE * 4));
        CK(scr.get((void**)&edge_cnt, (size_t)cq * LB_NEDGE * 4));
        ScanArgs b = a;
        b.parts = 0; b.partial = nullptr; b.tile_begin = 0; b.tile_end = boot_tiles; b.part_offset = 0;
        b.keys_out = keys; b.keys_ld = S;
        CK(launch_dense_scan_tc(b, sm_count, cand, st));
        CK(launch_sample_select(keys, S, S, a.n_rows, a.tomb, a.tomb_bits, a.allow, cq, kc, partial, tau, edges, edge_cnt,
                                sel_done, st));
        a.edges = edges; a.edge_cnt = edge_cnt; a.tau_init = tau;
    }
    a.parts = 2 * gm; a.tile_begin = boot_tiles; a.tile_end = n_tiles; a.part_offset = extra;
    {
        ProfScope prof(st, (double)cq * (double)((int64_t)a.n_rows - (int64_t)boot_tiles * 256));
        CK(launch_dense_scan_tc(a, sm_count, cand, st));
    }
    if (parts > 1) {
        CK(scr.get((void**)&merged, (size_t)cq * kc * 8));
        CK(launch_merge_select(partial, parts, cq, kc, merged, nullptr, a.edges, a.edge_cnt, st));
    } else {
        merged = partial;
    }
    *merged_out = merged;
    return LB_OK;
}

static int pq_search_core(lb_pq* pq, const float* d_q, int64_t nq, int k, int kprime, const uint64_t* d_allow,
                          float* d_dist, int64_t* d_lab, cudaStream_t st, uint32_t* d_flags = nullptr,
                          uint32_t* d_count = nullptr) {
    if (nq == 0) return LB_OK;
    const bool rerank = pq->raw != nullptr;
    const int kout = rerank ? (kprime > k ? kprime : k) : k;  // exact ADC top-kout, then (optionally) the fp32 re-rank
    if (kout > 1024) return fail(LB_ERR_UNSUPPORTED, "k' > 1024");
    if (rerank && pq->raw->size < pq->size) return fail(LB_ERR_STATE, "raw index has fewer rows than codes");
    Scratch scr(st);
    if (pq->size == 0) {
        uint64_t* merged;
        CK(scr.get((void**)&merged, 8));
        CK(launch_unpack_topk(merged, (int)nq, 0, k, 0, d_dist, d_lab, st));
        if (d_flags) CK(cudaMemsetAsync(d_flags, 0, (size_t)nq * 4, st));
        return LB_OK;
    }
    const int mode = g_opt_pq_scan.load(std::memory_order_relaxed);
    const int kc = coarse_k(kout);  // coarse candidates: margin over kout absorbs the quantisation step
    // (the tensor-core path decodes the codes once per query chunk: larger chunks amortise it)
    const int64_t qchunk = (mode == 4 || (mode == 0 && nq >= 64 && pq->codebook16 != nullptr)) ? 1024 : 512;
    for (int64_t qo = 0; qo < nq; qo += qchunk) {
        const int cq = (int)((nq - qo) < qchunk ? (nq - qo) : qchunk);
        const float* q = d_q + (size_t)qo * pq->dims;
        float* luts;
        CK(scr.get((void**)&luts, (size_t)cq * pq->M * 1024));
        CK(launch_adc_lut(pq->codebooks, pq->M, pq->K, pq->sub, q, cq, luts, st));
        // Batches: coarse stage on the tensor cores over a decoded fp16 slab (pq_gemm.cu).  From ~64 queries up the
        // look-up scan is bound by the shared-memory pipe (N*M look-ups per query) while the decode is paid once per
        // slab and the dense scan runs at its tensor rate.
        const bool gemm_ok = pq->codebook16 != nullptr && pq->xn2 != nullptr && pq->tiled != nullptr && pq->M <= 96 && pq->size >= 4096 &&
                             dense_tc_eligible(DT_F16, pq->dims, pq->codebook16, pq->codebook16, 2 * kout + 64);
        if (mode == 4 && (!gemm_ok || kout > 704))
            return fail(LB_ERR_UNSUPPORTED, "decode + tensor-core PQ scan not eligible");
        if (mode == 4 || (mode == 0 && gemm_ok && cq >= 64 && kout <= 352 && g_opt_pq_gemm.load(std::memory_order_relaxed))) {  // auto: only with the full 2 k' margin
            int kg = 2 * kout;  // wider candidate margin than the look-up path: the fp16 rounding bound is looser
            if (kg < kc) kg = kc;
            if (kg > 704) kg = 704;
            void* q16; float* qn; void* slab; uint64_t *all, *merged, *exact;
            const int64_t slab_rows_max = 4 << 20;   // 4 Mi rows per slab (6 GiB at 768 dims)
            const int64_t slab_rows = pq->size < slab_rows_max ? ((pq->size + 255) / 256) * 256 : slab_rows_max;
            const int n_slabs = (int)((pq->size + slab_rows - 1) / slab_rows);
            CK(scr.get(&q16, (size_t)cq * pq->dims * 2));
            CK(scr.get((void**)&qn, (size_t)cq * 8));
            CK(scr.get(&slab, (size_t)slab_rows * pq->dims * 2));
            CK(scr.get((void**)&all, (size_t)n_slabs * cq * kg * 8));
            CK(scr.get((void**)&exact, (size_t)cq * kout * 8));
            CK(launch_pq_q16(q, cq, pq->dims, q16, qn, st));
            const int64_t tomb_bits = pq->tomb ? pq->tomb_bits : 0;
            for (int sl = 0; sl < n_slabs; sl++) {
                const int64_t r0 = (int64_t)sl * slab_rows;
                const int64_t rn = (pq->size - r0) < slab_rows ? (pq->size - r0) : slab_rows;
                {
                    ProfScope prof(st, (double)rn * ((double)pq->dims * 2 + pq->M), 1);  // bytes written + code bytes read
                    CK(launch_pq_decode(pq->codes, pq->codebook16, pq->M, pq->sub, (uint32_t)r0, (uint32_t)rn, slab,
                                        pq->sm_count, st));
                }
                ScanArgs a;
                a.dtype = DT_F16; a.metric = METRIC_L2; a.db = slab; a.aux = pq->xn2 + r0; a.n_rows = (uint32_t)rn;
                a.dim = pq->dims; a.queries = q16; a.nq = cq;
                // bitmaps are indexed by global row: r0 is a multiple of 256, so the slab's view is a word offset
                const int64_t tb = tomb_bits - r0;
                a.tomb = (pq->tomb && tb > 0) ? pq->tomb + r0 / 32 : nullptr;
                a.tomb_bits = (uint32_t)(tb > 0 ? (tb > 0xffffffffll ? 0xffffffffll : tb) : 0);
                a.allow = d_allow ? (const uint32_t*)d_allow + r0 / 32 : nullptr;
                uint64_t* mg;
                int rc = tc_coarse_candidates(a, kg, pq->sm_count, scr, st, &mg);
                if (rc) return rc;
                uint64_t* dst = all + (size_t)sl * cq * kg;
                CK(cudaMemcpyAsync(dst, mg, (size_t)cq * kg * 8, cudaMemcpyDeviceToDevice, st));
                CK(launch_pq_offset_rows(dst, (size_t)cq * kg, (uint32_t)r0, st));
            }
            if (n_slabs > 1) {
                CK(scr.get((void**)&merged, (size_t)cq * kg * 8));
                CK(launch_merge_select(all, n_slabs, cq, kg, merged, nullptr, nullptr, nullptr, st));
            } else {
                merged = all;
            }
            uint32_t* flags = d_flags ? d_flags + qo : nullptr;
            if (!flags && d_count) CK(scr.get((void**)&flags, (size_t)cq * 4));
            PqGemmCert gc;
            gc.qn = (const float2*)qn; gc.xmax2 = pq->xmax2; gc.dims = pq->dims;
            gc.beta = (float)pq->dims * 1.2e-7f;
            if (gc.beta < 8e-6f) gc.beta = 8e-6f;
            CK(launch_adc_exact(pq->tiled, pq->M, luts, merged, cq, kg, kout, nullptr, nullptr, exact, flags, d_count, st, &gc));
            if (rerank) {
                RescoreArgs r;
                r.dtype = DT_F32; r.metric = METRIC_L2; r.db = pq->raw->rows; r.n_rows = (uint32_t)pq->raw->size;
                r.dim = pq->dims; r.queries = q; r.nq = cq; r.packed = exact; r.ids32 = nullptr; r.c = kout; r.k = k;
                r.tomb = nullptr; r.tomb_bits = 0; r.allow = nullptr; r.id_base = 0;
                r.out_d = d_dist + (size_t)qo * k; r.out_l = d_lab + (size_t)qo * k; r.negate_dot = 1;
                CK(launch_rescore(r, st));
            } else {
                CK(launch_unpack_topk(exact, cq, kout, k, 0, d_dist + (size_t)qo * k, d_lab + (size_t)qo * k, st));
            }
            continue;
        }
        int nqpp = (mode == 2) ? 1 : (mode == 3) ? 4 : (cq >= 2 ? 4 : 1);
        if (nqpp == 4 && !adc_coarse_eligible(pq->M, kc, 4)) nqpp = 1;
        const bool coarse = mode != 1 && pq->tiled != nullptr && adc_coarse_eligible(pq->M, kc, nqpp) &&
                            pq->size >= 4096;
        if ((mode == 2 || mode == 3) && !coarse) return fail(LB_ERR_UNSUPPORTED, "coarse ADC scan not eligible");
        if (!coarse) {
            int rc = pq_search_exhaustive(pq, q, cq, luts, k, kout, rerank, d_allow, d_dist + (size_t)qo * k,
                                          d_lab + (size_t)qo * k, scr, st);
            if (rc) return rc;
            if (d_flags) CK(cudaMemsetAsync(d_flags + qo, 0, (size_t)cq * 4, st));
            continue;
        }
        const int qgroups = (cq + nqpp - 1) / nqpp;
        const int64_t n_tiles = (pq->size + 31) / 32;
        int unit = pq->sm_count / gcd_i(qgroups, pq->sm_count);
        int parts = unit;
        const int64_t max_parts = (n_tiles + 63) / 64;  // at least 2048 rows per part
        if (parts > max_parts) parts = (int)max_parts;
        if (parts < 1) parts = 1;
        const uint32_t tpp = (uint32_t)((n_tiles + parts - 1) / parts);
        parts = (int)((n_tiles + tpp - 1) / tpp);
        const size_t stride = (size_t)parts * kc;
        uint8_t* lutq; void* params; uint64_t *compact, *merged, *exact; uint32_t *out_cnt, *g_tau, *g_min, *ovf;
        CK(scr.get((void**)&lutq, adc_lutq_bytes(pq->M, cq, nqpp)));
        CK(scr.get(&params, adc_params_bytes(cq)));
        CK(scr.get((void**)&compact, (size_t)cq * stride * 8));
        CK(scr.get((void**)&out_cnt, (size_t)cq * 4));
        CK(scr.get((void**)&g_tau, (size_t)cq * 4));
        CK(scr.get((void**)&merged, (size_t)cq * kc * 8));
        CK(scr.get((void**)&exact, (size_t)cq * kout * 8));
        // one allocation: [out_cnt | overflow] zeroed, [g_tau | g_min] set to 0xffffffff
        const size_t nmin = (size_t)adc_min_slots(parts);
        CK(scr.get((void**)&ovf, (size_t)cq * 4));
        CK(scr.get((void**)&g_min, (size_t)cq * nmin * 4));
        CK(cudaMemsetAsync(out_cnt, 0, (size_t)cq * 4, st));
        CK(cudaMemsetAsync(ovf, 0, (size_t)cq * 4, st));
        CK(cudaMemsetAsync(g_tau, 0xff, (size_t)cq * 4, st));
        CK(cudaMemsetAsync(g_min, 0xff, (size_t)cq * nmin * 4, st));
        {
            ProfScope prof(st, (double)cq * (double)pq->size);
            CK(launch_adc_coarse(pq->tiled, (uint32_t)pq->size, pq->M, luts, cq, nqpp, pq->tomb,
                                 (uint32_t)(pq->tomb_bits > 0xffffffffll ? 0xffffffffll : pq->tomb_bits),
                                 (const uint32_t*)d_allow, kc, parts, tpp, lutq, params, compact, out_cnt, stride, g_tau,
                                 g_min, ovf, st));
        }
        CK(launch_merge_select_compact(nullptr, compact, out_cnt, stride, cq, kc, merged, nullptr, nullptr, st));
        uint32_t* flags = d_flags ? d_flags + qo : nullptr;
        if (!flags && d_count) CK(scr.get((void**)&flags, (size_t)cq * 4));
        CK(launch_adc_exact(pq->tiled, pq->M, luts, merged, cq, kc, kout, params, ovf, exact, flags, d_count, st));
        if (rerank) {
            RescoreArgs r;
            r.dtype = DT_F32; r.metric = METRIC_L2; r.db = pq->raw->rows; r.n_rows = (uint32_t)pq->raw->size;
            r.dim = pq->dims; r.queries = q; r.nq = cq; r.packed = exact; r.ids32 = nullptr; r.c = kout; r.k = k;
            r.tomb = nullptr; r.tomb_bits = 0; r.allow = nullptr; r.id_base = 0;
            r.out_d = d_dist + (size_t)qo * k; r.out_l = d_lab + (size_t)qo * k; r.negate_dot = 1;
            CK(launch_rescore(r, st));
        } else {
            CK(launch_unpack_topk(exact, cq, kout, k, 0, d_dist + (size_t)qo * k, d_lab + (size_t)qo * k, st));
        }
    }
    return LB_OK;
}

// exhaustive repair of the queries whose host flag is set (device pointers in / out)
static int pq_repair(lb_pq* pq, const float* d_q, int64_t nq, int k, int kprime, const uint64_t* d_allow,
                     const uint32_t* h_flags, float* d_dist, int64_t* d_lab, cudaStream_t st) {
    const bool rerank = pq->raw != nullptr;
    const int kout = rerank ? (kprime > k ? kprime : k) : k;
    Scratch scr(st);
    for (int64_t qi = 0; qi < nq; qi++) {
        if (!h_flags[qi]) continue;
        const float* q = d_q + (size_t)qi * pq->dims;
        float* luts;
        CK(scr.get((void**)&luts, (size_t)pq->M * 1024));
        CK(launch_adc_lut(pq->codebooks, pq->M, pq->K, pq->sub, q, 1, luts, st));
        int rc = pq_search_exhaustive(pq, q, 1, luts, k, kout, rerank, d_allow, d_dist + (size_t)qi * k,
                                      d_lab + (size_t)qi * k, scr, st);
        if (rc) return rc;
    }
    return LB_OK;
}

int lb_pq_search_device(lb_pq* pq, const float* d_queries, int64_t nq, int k, int kprime, const uint64_t* d_allow,
                        float* d_distances, int64_t* d_labels, void* stream) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive");
    int rc = use_device(pq->device);
    if (rc) return rc;
    return pq_search_core(pq, d_queries, nq, k, kprime, d_allow, d_distances, d_labels, (cudaStream_t)stream);
}

int lb_pq_search(lb_pq* pq, const float* queries, int64_t nq, int k, int kprime, const uint64_t* allow,
                 float* distances, int64_t* labels) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive");
    if (nq == 0) return LB_OK;
    if (!queries || !distances || !labels) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(pq->device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    float *d_q, *d_d; int64_t* d_l; uint64_t* d_allow = nullptr;
    CK(scr.get((void**)&d_q, (size_t)nq * pq->dims * 4));
    CK(scr.get((void**)&d_d, (size_t)nq * k * 4));
    CK(scr.get((void**)&d_l, (size_t)nq * k * 8));
    CK(cudaMemcpyAsync(d_q, queries, (size_t)nq * pq->dims * 4, cudaMemcpyHostToDevice, st));
    if (allow) {
        size_t words = (size_t)((pq->size + 63) / 64);
        CK(scr.get((void**)&d_allow, words * 8));
        CK(cudaMemcpyAsync(d_allow, allow, words * 8, cudaMemcpyHostToDevice, st));
    }
    uint32_t *d_flags, *d_count;
    CK(scr.get((void**)&d_flags, (size_t)nq * 4));
    CK(scr.get((void**)&d_count, 4));
    CK(cudaMemsetAsync(d_count, 0, 4, st));
    rc = pq_search_core(pq, d_q, nq, k, kprime, d_allow, d_d, d_l, st, d_flags, d_count);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    uint32_t n_uncert = 0;
    CK(cudaMemcpyAsync(&n_uncert, d_count, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    pq->last_uncertified.store((int64_t)n_uncert);
    if (n_uncert > 0) {  // coarse margin did not cover the quantisation bound: redo those queries exhaustively
        std::vector<uint32_t> flags((size_t)nq);
        CK(cudaMemcpy(flags.data(), d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost));
        rc = pq_repair(pq, d_q, nq, k, kprime, d_allow, flags.data(), d_d, d_l, st);
        if (rc) { cudaStreamSynchronize(st); return rc; }
    }
    CK(cudaMemcpyAsync(distances, d_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(labels, d_l, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

int64_t lb_pq_last_uncertified(const lb_pq* pq) { return pq ? pq->last_uncertified.load() : -1; }

int lb_pq_search_device_cert(lb_pq* pq, const float* d_queries, int64_t nq, int k, int kprime, const uint64_t* d_allow,
                             float* d_distances, int64_t* d_labels, uint32_t* d_uncert_flags,
                             uint32_t* d_uncert_count, void* stream) {
    if (!pq) return fail(LB_ERR_INVALID, "pq is NULL");
    if (k <= 0 || nq < 0) return fail(LB_ERR_INVALID, "k must be positive");
    int rc = use_device(pq->device);
    if (rc) return rc;
    return pq_search_core(pq, d_queries, nq, k, kprime, d_allow, d_distances, d_labels, (cudaStream_t)stream,
                          d_uncert_flags, d_uncert_count);
}

// ----------------------------------------------------------------------------------- predicates
}  // extern "C"

template <typename T, typename F>
static int filter_common(int device, const T* column, int64_t n, int op, int and_into, uint64_t* bitmap, F launch) {
    if (n < 0 || op < 0 || op > 5) return fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return LB_OK;
    if (!column || !bitmap) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(device);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    Scratch scr(st);
    size_t words = (size_t)((n + 63) / 64);
    T* d_c; uint32_t* d_b;
    CK(scr.get((void**)&d_c, (size_t)n * sizeof(T)));
    CK(scr.get((void**)&d_b, words * 8));
    CK(cudaMemcpyAsync(d_c, column, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
    if (and_into) CK(cudaMemcpyAsync(d_b, bitmap, words * 8, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(d_b, 0, words * 8, st));
    CK(launch(d_c, d_b, st));
    CK(cudaMemcpyAsync(bitmap, d_b, words * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return LB_OK;
}

extern "C" {

int lb_filter_i64(int device, const int64_t* column, int64_t n, int op, int64_t value, int and_into, uint64_t* bitmap) {
    return filter_common<int64_t>(device, column, n, op, and_into, bitmap,
                                  [&](const int64_t* c, uint32_t* b, cudaStream_t st) {
                                      return launch_filter_i64(c, n, op, value, and_into, b, st);
                                  });
}
int lb_filter_i64_device(int device, const int64_t* d_column, int64_t n, int op, int64_t value, int and_into,
                         uint64_t* d_bitmap, void* stream) {
    if (n < 0 || op < 0 || op > 5) return fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return LB_OK;
    if (!d_column || !d_bitmap) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(device);
    if (rc) return rc;
    CK(launch_filter_i64(d_column, n, op, value, and_into, (uint32_t*)d_bitmap, (cudaStream_t)stream));
    return LB_OK;
}
int lb_filter_f32_device(int device, const float* d_column, int64_t n, int op, float value, int and_into,
                         uint64_t* d_bitmap, void* stream) {
    if (n < 0 || op < 0 || op > 5) return fail(LB_ERR_INVALID, "bad argument");
    if (n == 0) return LB_OK;
    if (!d_column || !d_bitmap) return fail(LB_ERR_INVALID, "NULL buffer");
    int rc = use_device(device);
    if (rc) return rc;
    CK(launch_filter_f32(d_column, n, op, value, and_into, (uint32_t*)d_bitmap, (cudaStream_t)stream));
    return LB_OK;
}
int lb_filter_f32(int device, const float* column, int64_t n, int op, float value, int and_into, uint64_t* bitmap) {
    return filter_common<float>(device, column, n, op, and_into, bitmap,
                                [&](const float* c, uint32_t* b, cudaStream_t st) {
                                    return launch_filter_f32(c, n, op, value, and_into, b, st);
                                });
}

}  // extern "C"
