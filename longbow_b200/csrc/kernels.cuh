// kernels.cuh -- host-side launch interface between api.cu and the kernel translation units.
#pragma once
#include <atomic>

#include "common.cuh"

namespace lb {

// Dynamic shared-memory opt-in.  cudaFuncAttributeMaxDynamicSharedMemorySize is per-function, per-device GLOBAL
// state and searches run concurrently (Go read lock, faiss_gpu.go:108), so it is raised to the device maximum
// ONCE per (kernel, device) and never written with a per-call size (two callers with different k could
// otherwise interleave "set small" / "launch large").  `done` is a bit per device.
inline cudaError_t smem_optin(const void* func, std::atomic<uint64_t>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, func);
    if (e != cudaSuccess) return e;
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return cudaSuccess;
}
#define LB_SMEM_OPTIN(kern)                                                     \
    do {                                                                        \
        static std::atomic<uint64_t> lb_done_{0};                               \
        cudaError_t lb_e_ = ::lb::smem_optin((const void*)(kern), lb_done_);    \
        if (lb_e_ != cudaSuccess) return lb_e_;                                 \
    } while (0)

extern bool g_rescore_legacy;
extern bool g_rescore_block;
extern int g_pq_ahead;
extern bool g_pq_ring;
extern bool g_hnsw_coop;
extern bool g_tc_pair;
extern int g_tc_reserve_sms;
extern bool g_ssel_warp;
void count_launch();  // api.cu: process-wide launch counter (bench evidence)

constexpr int LB_NEDGE = 16;  // rungs of the shared threshold ladder (dense_tc.cu)

struct ScanArgs {
    int dtype, metric;
    const void* db;          // [n_rows][dim] row-major, element type = dtype
    const float* aux;        // per-row |x|^2 (L2) or 1/|x| (cosine); may be null for dot
    uint32_t n_rows;
    int dim;
    const void* queries;     // [nq][dim], same dtype
    int nq;
    const uint32_t* tomb;    // dense bitmaps viewed as 32-bit words (little endian == u64 layout)
    uint32_t tomb_bits;
    const uint32_t* allow;
    int kc;                  // candidates kept per (part, query)
    int cap;                 // per-query candidate buffer capacity (power of two >= kc + TN)
    int tq;                  // queries per CTA: 64, 32 or 16
    int parts;
    uint32_t rows_per_part;  // multiple of 128
    uint64_t* partial;       // [parts][nq][kc]
    int debug = 0;           // tensor-core scan timing probes (lb_set_option "tc_debug")
    // tensor-core scan only: sub-range of 256-row tiles, first partial[] slot, bootstrap thresholds
    int tile_begin = 0, tile_end = -1, part_offset = 0;
    const float* tau_init = nullptr;
    float* keys_out = nullptr;  // bootstrap sample mode: dump raw keys [nq][keys_ld]
    int keys_ld = 0;
    // shared progressive threshold (tensor-core scan): per query a ladder of LB_NEDGE keys taken from the
    // bootstrap sample and the number of live rows seen so far at or below each
    const void* db_lo = nullptr;       // fp32 tensor-core scan (3xTF32): x - tf32(x) of the rows / queries
    const void* queries_lo = nullptr;
    const float* edges = nullptr;  // [nq][LB_NEDGE] ascending, already nextafter()'d
    uint32_t* edge_cnt = nullptr;  // [nq][LB_NEDGE]
};

// Ascending 1-based sample ranks of the threshold ladder.  The first n_spare slots have no rank of their own (the
// ladder reached rank 1 before it ran out of slots): their edges are extrapolated below the sample, `doublings`
// halvings of the tail probability past the lowest rank spread evenly over them (any value is a valid edge -- an
// edge only becomes the threshold once kc live rows were counted below it).
struct EdgeRanks { int r[LB_NEDGE]; int n_spare; float doublings; };

struct RescoreArgs {
    int dtype, metric;
    const void* db;
    uint32_t n_rows;
    int dim;
    const void* queries;
    int nq;
    const uint64_t* packed;  // [nq][c] packed candidates, or null
    const uint32_t* ids32;   // [nq][c] raw VectorIDs, or null
    int c, k;
    const uint32_t* tomb;
    uint32_t tomb_bits;
    const uint32_t* allow;
    int64_t id_base;
    float* out_d;            // [nq][k]
    int64_t* out_l;          // [nq][k]
    int negate_dot;          // 1: dot returned as a distance (negated)
    const float* nrm = nullptr;  // cosine only: exact per-row |x|^2 (reference lane order), or null
    // certification of the coarse stage (packed mode only; null = off).  cert_flags[q] = 1 when the margin
    // between the kc-th coarse key and the k-th exact result does not cover the coarse error bound, i.e. a row
    // outside the candidate set could belong to the true top-k; cert_count accumulates such queries.
    uint32_t* cert_flags = nullptr;
    uint32_t* cert_count = nullptr;
    const float* max_norm2 = nullptr;  // device scalar: max |x|^2 over the index (float bits, non-negative)
    float beta = 0.f;                  // relative coarse-key error bound (of |q||x|)
    int key_space = 0;                 // see CertArgs
};

struct CertArgs {
    uint32_t* flags;
    uint32_t* count;
    const float* max_norm2;
    float beta;
    int key_space;  // 0: expanded keys (|x|^2 - 2 q.x, tensor-core / streaming scans); 1: difference form |q - x|^2 (SIMT L2)
};

cudaError_t launch_row_maxnorm(int dtype, const void* db, int64_t n, int dim, int64_t row0, float* max_norm2,
                               cudaStream_t st);
cudaError_t launch_mask_rows(float* dist, int64_t n, const uint32_t* tomb, uint32_t tomb_bits, const uint32_t* allow,
                             cudaStream_t st);

size_t dense_scan_simt_smem(int tq, int cap);
cudaError_t launch_dense_scan_simt(const ScanArgs& a, cudaStream_t st);
cudaError_t launch_row_aux(int dtype, const void* db, int64_t n, int dim, int metric, float* aux, int64_t row0,
                           cudaStream_t st);
cudaError_t launch_merge_partials(const uint64_t* partial, int parts, int nq, int kc, uint64_t* merged,
                                  cudaStream_t st);
cudaError_t launch_merge_select(const uint64_t* partial, int parts, int nq, int kc, uint64_t* merged, uint64_t* kth,
                                const float* edges, const uint32_t* edge_cnt, cudaStream_t st);
// compact form: per query, `head` kc entries at head[q*kc] (may be null) plus counts[q] entries at
// compact[q*stride]
cudaError_t launch_merge_select_compact(const uint64_t* head, const uint64_t* compact, const uint32_t* counts,
                                        size_t stride, int nq, int kc, uint64_t* merged, const float* edges,
                                        const uint32_t* edge_cnt, cudaStream_t st);
cudaError_t launch_row_norm_exact(int dtype, const void* db, int64_t n, int dim, float* nrm, int64_t row0,
                                  cudaStream_t st);
cudaError_t launch_rescore(const RescoreArgs& a, cudaStream_t st);
cudaError_t launch_batch_flat(int metric, int dtype, const void* db, int64_t n, int dim, const void* query,
                              float* out, int negate_dot, cudaStream_t st);
size_t select_k_scratch_entries(int64_t n, int k);  // 8-byte entries launch_select_k needs in scratch_partial
cudaError_t launch_select_k(const float* d, int64_t n, int k, uint64_t* scratch_partial, uint64_t* scratch_merged,
                            int64_t* out_idx, float* out_d, int64_t id_base, cudaStream_t st);
cudaError_t launch_merge_topk_strided(const void* in_d_base, size_t stride_d, const void* in_l_base, size_t stride_l,
                                      int parts, int nq, int k_in, int k, float* out_d, int64_t* out_l,
                                      cudaStream_t st);
cudaError_t launch_merge_topk_wait(const void* in_d_base, size_t stride_d, const void* in_l_base, size_t stride_l,
                                   int parts, int nq, int k_in, int k, float* out_d, int64_t* out_l,
                                   const uint32_t* wait_flags, uint32_t wait_seq, int rank, uint32_t* err,
                                   cudaStream_t st);
cudaError_t launch_merge_topk(const float* in_d, const int64_t* in_l, int parts, int nq, int k_in, int k,
                              float* out_d, int64_t* out_l, cudaStream_t st);

// ---- tensor-core scan (dense_tc.cu)
bool dense_tc_eligible(int dtype, int dim, const void* db, const void* queries, int kc);
void dense_scan_tc_plan(int nq, int n_row_tiles, int sm_count, int kc, int* groups_out, size_t* cand_bytes);
cudaError_t launch_sample_select(const float* keys, int ld, int S, uint32_t n_rows, const uint32_t* tomb,
                                 uint32_t tomb_bits, const uint32_t* allow, int nq, int kc, uint64_t* out,
                                 float* tau, float* edges, uint32_t* edge_cnt, int* done, cudaStream_t st);
cudaError_t launch_dense_scan_tc(const ScanArgs& s, int sm_count, uint64_t* cand, cudaStream_t st);
cudaError_t launch_split_lo(const float* x, float* lo, size_t n, cudaStream_t st);

// ---- streaming scan for small query batches (dense_stream.cu)
bool dense_stream_eligible(int dtype, int dim, const void* db, int nq, int kc);
int dense_stream_grid(int sm_count, int nq);
cudaError_t launch_dense_scan_stream(const ScanArgs& s, int grid, uint32_t row_begin, uint32_t dump_rows,
                                     uint32_t* out_cnt, size_t out_stride, cudaStream_t st);

// ---- PQ (kernels_pq.cu)
struct PqScanArgs {
    const uint8_t* codes;    // [n][M] row-major
    uint32_t n_rows;
    int M;
    const float* luts;       // [nq][M*256]
    int nq;
    const uint32_t* tomb;
    uint32_t tomb_bits;
    const uint32_t* allow;
    int kc, cap;
    int parts;
    uint32_t rows_per_part;
    uint64_t* partial;       // [parts][nq][kc]
};
cudaError_t launch_adc_lut(const float* codebooks, int M, int K, int sub, const float* queries, int nq, float* luts,
                           cudaStream_t st);
cudaError_t launch_adc_scan(const PqScanArgs& a, cudaStream_t st);
cudaError_t launch_adc_batch(const float* table, const uint8_t* codes, int M, int64_t n, float* out, cudaStream_t st);
cudaError_t launch_pq_encode(const float* codebooks, int M, int K, int sub, const float* vecs, int64_t n,
                             uint8_t* codes, cudaStream_t st);
cudaError_t launch_pq_train(const float* d_data, int64_t n, int dims, int M, int K, int max_iter,
                            const int32_t* d_init_idx, float* d_cent, int32_t* d_assign, uint32_t* d_changed,
                            int32_t* d_active, int32_t* d_iters, cudaStream_t st);
// ---- PQ coarse -> certify -> exact scan over the tiled code mirror (pq_scan.cu)
cudaError_t launch_pq_tile_codes(const uint8_t* flat, int64_t n, int M, int Mp, int64_t row0, uint8_t* tiled,
                                 cudaStream_t st);
bool adc_coarse_eligible(int M, int kc, int nq_per_pass);
size_t adc_lutq_bytes(int M, int nq, int nq_per_pass);
size_t adc_params_bytes(int nq);
int adc_min_slots(int parts);
cudaError_t launch_adc_coarse(const uint8_t* tiled, uint32_t n_rows, int M, const float* luts, int nq, int nq_per_pass,
                              const uint32_t* tomb, uint32_t tomb_bits, const uint32_t* allow, int kc, int parts,
                              uint32_t tiles_per_part, uint8_t* lutq, void* params, uint64_t* compact,
                              uint32_t* out_cnt, size_t stride, uint32_t* g_tau, uint32_t* g_min, uint32_t* overflow,
                              cudaStream_t st);
// certification inputs when the coarse candidates come from the tensor-core scan over decoded rows (pq_gemm.cu)
struct PqGemmCert {
    const float2* qn;        // per query { |q_h|^2, |q| }; null = look-up path (integer keys, `params`)
    const uint32_t* xmax2;   // largest |x_h|^2 over the decoded rows (bits of a non-negative float)
    float beta;              // dot-product error bound of the scan, relative to |q_h||x_h|
    int dims;
};
cudaError_t launch_adc_exact(const uint8_t* tiled, int M, const float* luts, const uint64_t* coarse, int nq, int kc,
                             int k_out, const void* params, const uint32_t* overflow, uint64_t* out,
                             uint32_t* cert_flags, uint32_t* cert_count, cudaStream_t st,
                             const PqGemmCert* gemm = nullptr);
// pq_gemm.cu: batched PQ coarse stage through the dense tensor-core scan
bool pq_gemm_eligible(int dims, int M, int sub);
cudaError_t launch_pq_codebook16(const float* cb, void* out, size_t n, cudaStream_t st);
cudaError_t launch_pq_q16(const float* q, int nq, int dims, void* q16, float* qn, cudaStream_t st);
cudaError_t launch_pq_centroid_norms(const void* cb16, int M, int sub, float* n2, cudaStream_t st);
cudaError_t launch_pq_row_norms(const uint8_t* codes, const float* n2, int M, int64_t r0, int64_t n, float* xn2,
                                uint32_t* xmax2, cudaStream_t st);
cudaError_t launch_pq_decode(const uint8_t* codes, const void* cb16, int M, int sub, uint32_t r0, uint32_t n, void* x,
                             int sm_count, cudaStream_t st);
cudaError_t launch_pq_offset_rows(uint64_t* p, size_t n, uint32_t r0, cudaStream_t st);
cudaError_t launch_unpack_topk(const uint64_t* merged, int nq, int kc, int k, int64_t id_base, float* out_d,
                               int64_t* out_l, cudaStream_t st);

// what the translation units outside api.cu may know about an lb_index handle
struct IndexView { const void* rows; int64_t size; int dim, dtype, metric, device; };

// ---- predicates (kernels_filter.cu)
cudaError_t launch_filter_i64(const int64_t* col, int64_t n, int op, int64_t val, int and_into, uint32_t* bitmap,
                              cudaStream_t st);
cudaError_t launch_filter_f32(const float* col, int64_t n, int op, float val, int and_into, uint32_t* bitmap,
                              cudaStream_t st);

}  // namespace lb
