/*
 * longbow_b200.h -- C ABI of liblongbow_b200.so, the B200-native (sm_100a) replacement for
 * Longbow's vector-distance / k-NN hot path.
 *
 * Plain pointers and sizes only; no C++/torch types.  This is exactly what the reference's
 * cgo layer binds (see INTEGRATION.md for the Go stubs).  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - Return value: 0 = success (internal/gpu/faiss_gpu.go:99-101,134-137 treats non-zero as
 *     "failed with code %d").  Non-zero codes are the LB_ERR_* values below and are stable.
 *     Constructors that return a pointer return NULL on failure (faiss_gpu.go:57-66).
 *   - No CPU fallback: if there is no usable CUDA device the call fails with a code.
 *   - Host entry points are synchronous: inputs are fully consumed (copied to the device)
 *     and outputs fully written before they return, so the caller (Go) may free or move its
 *     slices immediately (cgo pointer rules).  They are thread-safe: any number of threads may
 *     search one handle concurrently (faiss_gpu.go:40,108 takes a read lock); add / free /
 *     set_* need exclusive access, as the reference's write lock provides.
 *   - *_device entry points take device pointers plus a cudaStream_t (as void*), enqueue work
 *     on that stream and return without synchronising.
 *   - Labels are 0-based insertion positions (the ids passed to gpu.Index.Add never reach C,
 *     faiss_gpu.go:93-97) plus the handle's id_base, as int64.  Missing results: label -1,
 *     distance FLT_MAX.
 *   - Distances are the reference's simd values: sqrt'd Euclidean
 *     (internal/simd/distance_functions.go:17-31), 1 - cosine similarity (:47-55), and the
 *     NEGATED dot product (internal/store/distance_resolvers.go:11-16; docs/distance_metrics.md:42-55).
 *     Results are ordered by (distance, label) ascending.
 *   - Bitmaps are dense, little-endian 64-bit words, bit i <-> row (VectorID) i.
 *       tombstones: bit set = deleted (internal/store/arrow_hnsw.go:147,468-472)
 *       allow:      bit set = passes the predicate (internal/query/bitmap.go:13-120)
 */
#ifndef LONGBOW_B200_H
#define LONGBOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_OK 0
#define LB_ERR_INVALID 1     /* bad argument (NULL handle, k <= 0, size mismatch ...) */
#define LB_ERR_CUDA 2        /* CUDA runtime/driver error; text via lb_last_error() */
#define LB_ERR_OOM 3         /* device or host allocation failed */
#define LB_ERR_UNSUPPORTED 4 /* no kernel for this (metric, dtype): registry miss, dispatch.go:275-278 */
#define LB_ERR_STATE 5       /* handle not in the required state (e.g. PQ without codes) */
#define LB_ERR_NO_DEVICE 6   /* no CUDA device / device id out of range */

/* internal/simd/registry.go:8-14 */
enum lb_metric { LB_METRIC_L2 = 0, LB_METRIC_COSINE = 1, LB_METRIC_DOT = 2 };
/* internal/simd/registry.go:31-47 (first four) */
enum lb_dtype { LB_F32 = 0, LB_F16 = 1, LB_I8 = 2, LB_U8 = 3 };

/* Thread-local text of the last failure on the calling thread ("" if none). */
const char *lb_last_error(void);
/* Library / device info: SM count, device name; returns LB_ERR_NO_DEVICE without a GPU. */
int lb_device_info(int device, int *sm_count, size_t *total_mem, char *name, size_t name_len);

/* ------------------------------------------------------------------------------------------
 * 1. The six symbols internal/gpu/faiss_gpu.go:16-21 declares and links today.
 *    fp32, Euclidean (sqrt'd, matching simd.EuclideanDistance and the Metal backends,
 *    internal/gpu/metal_gpu.go:104), labels = insertion positions.
 * ---------------------------------------------------------------------------------------- */
void *faiss_gpu_resources_new(int device);                    /* faiss_gpu.go:16 */
void faiss_gpu_resources_free(void *res);                     /* faiss_gpu.go:17 */
void *faiss_gpu_index_flat_l2_new(void *res, int dim);        /* faiss_gpu.go:18 */
void faiss_gpu_index_flat_l2_free(void *idx);                 /* faiss_gpu.go:19 */
int faiss_gpu_index_add(void *idx, int64_t n, float *vectors);/* faiss_gpu.go:20 */
int faiss_gpu_index_search(void *idx, int64_t n, float *queries, int k, float *distances,
                           int64_t *labels);                  /* faiss_gpu.go:21 */

/* ------------------------------------------------------------------------------------------
 * 2. Dense index: rows of one Arrow FixedSizeList<T, dim> column mirrored in HBM.
 *    Replaces BruteForceIndex.SearchVectors (internal/store/adaptive_index.go:161-225), the
 *    re-rank stage (internal/store/parallel_search.go:147-365, hnsw_batch.go:104-245) and the
 *    simd batch functions (internal/simd/batch_operations.go:17-157) for resident data.
 * ---------------------------------------------------------------------------------------- */
typedef struct lb_index lb_index;

int lb_index_create(int device, int dim, int dtype, int metric, lb_index **out);
void lb_index_free(lb_index *idx);
/* Pre-size the HBM mirror (rows).  Optional.  The mirror lives in one virtual-address reservation per index and
 * grows by mapping more physical memory behind it: growth never copies rows and never synchronises the device. */
int lb_index_reserve(lb_index *idx, int64_t n_rows);
/* Append n rows from a host buffer laid out as the Arrow child values buffer: row-major,
 * contiguous, n*dim elements (internal/store/arrow_utils.go:112-171).  gpu.Index.Add. */
int lb_index_add(lb_index *idx, const void *rows, int64_t n);
int lb_index_add_device(lb_index *idx, const void *d_rows, int64_t n, void *stream);
/* Append the rows of one RecordBatch's vector column straight from its Arrow buffer: `values` is the child values
 * buffer of the FixedSizeList<T, dim> column (values_len_bytes long), list_offset the list array's Offset().  Row r
 * is read at element (list_offset + r) * dim, with the reference's truncated-buffer rule for IPC-flattened buffers
 * (internal/store/arrow_utils.go:112-171).  pin != 0: the buffer is page-locked for the call and uploaded in
 * chunks on two streams (DMA of one chunk under the row-statistics kernels of the previous one). */
int lb_index_add_arrow(lb_index *idx, const void *values, size_t values_len_bytes, int64_t list_offset,
                       int64_t n_rows, int pin);
int64_t lb_index_size(const lb_index *idx);
int lb_index_dim(const lb_index *idx);
/* Labels returned = local row + id_base (row-sharded multi-GPU: base of this shard). */
int lb_index_set_id_base(lb_index *idx, int64_t id_base);
/* Tombstones (index state).  words = ceil(nbits/64); NULL clears.  nbits may be < size
 * (missing bits = live). */
int lb_index_set_tombstones(lb_index *idx, const uint64_t *bitmap, int64_t nbits);
int lb_index_set_tombstones_device(lb_index *idx, const uint64_t *d_bitmap, int64_t nbits, void *stream);

/* Batched exact k-NN: nq queries (row-major [nq*dim], same dtype as the index) -> [nq*k]
 * distances and labels.  allow: optional predicate bitmap over local rows shared by the
 * batch (one Filters set per VectorSearchRequest, internal/query/requests.go:4-20), nbits
 * must cover the index.  Replaces the per-query loop at
 * internal/store/vector_search_action.go:73.
 * k <= 704 uses the fused scan + selector (one query: HBM-bound streaming scan; more: tensor-core scan);
 * 704 < k <= 2048 (e.g. SearchHybrid's k*10 candidates, internal/store/hnsw_gpu.go:85) takes an
 * exhaustive exact kernel, one query at a time; larger k is LB_ERR_UNSUPPORTED. */
int lb_index_search(lb_index *idx, const void *queries, int64_t nq, int k, const uint64_t *allow,
                    float *distances, int64_t *labels);
int lb_index_search_device(lb_index *idx, const void *d_queries, int64_t nq, int k,
                           const uint64_t *d_allow, float *d_distances, int64_t *d_labels,
                           void *stream);

/* lb_index_search_device + certification outputs, still without any host synchronisation:
 * d_uncert_flags [nq] receives 1 for every query whose coarse-stage margin did not cover the coarse error bound
 * (a row outside its candidate set might belong to the top-k) and 0 otherwise; *d_uncert_count is INCREMENTED
 * by the number of such queries (the caller zeroes it, and may let it accumulate over many batches).  Either
 * may be NULL.  Nothing is re-done here: the caller reads the count when convenient and repairs flagged
 * queries with lb_index_search_exact_device.  (lb_index_search, the host entry point, does both itself.) */
int lb_index_search_device_cert(lb_index *idx, const void *d_queries, int64_t nq, int k, const uint64_t *d_allow,
                                float *d_distances, int64_t *d_labels, uint32_t *d_uncert_flags,
                                uint32_t *d_uncert_count, void *stream);
/* Exhaustive exact search -- the reference's arithmetic for every live row, no coarse stage -- of the queries
 * whose HOST flag h_flags[q] is non-zero (h_flags NULL = all), written into the same [nq*k] layout; the other
 * rows of the outputs are left untouched.  Queries and outputs are device pointers. */
int lb_index_search_exact_device(lb_index *idx, const void *d_queries, int64_t nq, int k, const uint64_t *d_allow,
                                 const uint32_t *h_flags, float *d_distances, int64_t *d_labels, void *stream);

/* Re-rank: per query a list of c candidate VectorIDs (uint32, internal/core/types.go:7);
 * ids that are out of range, tombstoned or fail `allow` are dropped; exact distances of the
 * rest; ascending; first k.  ArrowHNSW.RerankBatch / processChunkInternal. */
int lb_index_rerank(lb_index *idx, const void *queries, int64_t nq, const uint32_t *cand_ids, int c,
                    int k, const uint64_t *allow, float *distances, int64_t *labels);
int lb_index_rerank_device(lb_index *idx, const void *d_queries, int64_t nq, const uint32_t *d_cand_ids,
                           int c, int k, const uint64_t *d_allow, float *d_distances, int64_t *d_labels,
                           void *stream);

/* One query against every resident row -> size() distances in row order
 * (simd.EuclideanDistanceBatchFlat with the flat buffer already in HBM). */
int lb_index_distances(lb_index *idx, const void *query, float *out);

/* Number of queries of the last lb_index_search (host) call on this handle whose coarse-stage result could not
 * be CERTIFIED -- the margin between the kc-th coarse key and the k-th exact distance did not cover the coarse
 * error bound, so a row outside the candidate set might have belonged to the top-k -- and which were therefore
 * recomputed by an exhaustive exact scan before returning (DESIGN.md 4.1 "certification").
 * lb_index_search_device does not certify; lb_index_search_device_cert reports the flags on the device. */
int64_t lb_index_last_uncertified(const lb_index *idx);
/* Diagnostics: the COARSE ranking keys the tensor-core scan computes for rows [0, n_rows) of nq queries
 * (out: [nq][n_rows]): |x|^2 - 2 q.x (L2), -q.x/|x| (cosine), -q.x (dot).  They only rank candidates -- every
 * returned distance comes from the exact re-score -- and exist so tests can bound the coarse error. */
int lb_index_coarse_keys(lb_index *idx, const void *queries, int64_t nq, int64_t n_rows, float *out);

/* ------------------------------------------------------------------------------------------
 * 3. Stateless simd surface (host buffers in, host buffers out; data uploaded per call).
 *    Mirrors internal/simd/batch_operations.go so internal/store can route through cgo.
 * ---------------------------------------------------------------------------------------- */
/* simd.{Euclidean,Cosine,DotProduct}DistanceBatch / EuclideanDistanceBatchFlat /
 * EuclideanDistanceF16Batch / EuclideanDistanceSQ8Batch: one query x n rows in a flat buffer.
 * Dot returns the RAW dot product here (simd.DotProductBatch), not negated.  LB_U8 + L2 returns
 * the squared distance as float32(int32) (internal/simd/sq8.go:37-66). */
int lb_simd_distance_batch_flat(int device, int metric, int dtype, const void *query, const void *flat,
                                int64_t n, int dim, float *results);
/* simd.ADCDistanceBatch (batch_operations.go:119-127): table [m*256] fp32, codes [n*m]. */
int lb_simd_adc_distance_batch(int device, const float *table, const uint8_t *flat_codes, int m,
                               int64_t n, float *results);
/* Scalar quantisation (internal/simd/sq8.go:70-104): QuantizeSQ8 -- scale = 255/(max-min) (0 when equal),
 * (v-min)*scale clamped to [0,255], truncated to a byte -- ComputeBounds, and the de-quantisation the HNSW
 * distance computer applies inline (internal/store/arrow_hnsw.go:1176-1186): min + code * (max-min)/255. */
int lb_simd_quantize_sq8(int device, const float *src, int64_t n, float min_val, float max_val, uint8_t *dst);
int lb_simd_dequantize_sq8(int device, const uint8_t *src, int64_t n, float min_val, float max_val, float *dst);
int lb_simd_compute_bounds(int device, const float *vec, int64_t n, float *min_val, float *max_val);
/* One fp32 query against n SQ8 rows, de-quantised inline, single sequential fp32 accumulator, sqrt through
 * double (arrow_hnsw.go:1176-1186: the distance an SQ8-enabled ArrowHNSW ranks by). */
int lb_simd_sq8_dequant_distance_batch(int device, const float *query, const uint8_t *rows, int64_t n, int dim,
                                       float min_val, float max_val, float *results);
/* simd.FindNearestCentroid (internal/simd/simd.go:278-326): k <= 8 compares squared distances, larger k the
 * sqrt'd batch results; strict '<' keeps the first minimum. */
int lb_simd_find_nearest_centroid(int device, const float *query, const float *centroids, int sub_dim, int k,
                                  int *out_index, float *out_distance);
/* Arrow compute "select_k_neighbors" (internal/store/arrow_kernels.go:230-345): indices of the
 * k smallest distances, (distance, index) ascending. */
int lb_select_k(int device, const float *distances, int64_t n, int k, int64_t *out_indices,
                float *out_distances);
/* Shard merge (internal/store/sharded_hnsw.go:432-503, result_merger.go:34-100):
 * [parts][nq][k_in] sorted lists -> [nq][k] by (distance, label); label -1 = padding. */
int lb_merge_topk(int device, const float *distances, const int64_t *labels, int parts, int64_t nq,
                  int k_in, int k, float *out_distances, int64_t *out_labels);
int lb_merge_topk_device(int device, const float *d_distances, const int64_t *d_labels, int parts,
                         int64_t nq, int k_in, int k, float *d_out_distances, int64_t *d_out_labels,
                         void *stream);

/* Same merge over PACKED per-part records, the layout of the multi-GPU exchange (one all-gather or one peer
 * push per batch): part p's record starts at d_records + p * part_stride and holds [nq*k_in] f32 distances at
 * offset 0 and [nq*k_in] i64 labels at label_offset (both multiples of 8 bytes). */
int lb_merge_topk_packed_device(int device, const void *d_records, size_t part_stride, size_t label_offset,
                                int parts, int64_t nq, int k_in, int k, float *d_out_distances,
                                int64_t *d_out_labels, void *stream);

/* ------------------------------------------------------------------------------------------
 * 3b. Multi-GPU exchange: all-gather of the per-GPU top-k records + merge, as our own kernels over
 *     NVLink peer memory (replaces the concat + sort tail of ShardedHNSW.SearchVectors,
 *     internal/store/sharded_hnsw.go:432-503, and MergeSortedStreams, result_merger.go:34-100).
 *     One lb_exchange per rank (= GPU).  Ranks in different processes connect through CUDA IPC handles
 *     shipped over any host channel (lb_exchange_handle / lb_exchange_connect_ipc); ranks in one process
 *     connect directly (lb_exchange_connect_local).  Per batch, on every rank and in the same order:
 *       lb_exchange_slot            -> where the local search writes its [nq*k] distances / labels
 *       (local search on `stream`)
 *       lb_exchange_all_gather_merge -> push to all peers, signal, wait for all records, merge -> outputs
 *     All of it is enqueued on the caller's stream; nothing synchronises the host.
 * ---------------------------------------------------------------------------------------- */
typedef struct lb_exchange lb_exchange;
int lb_exchange_create(int device, int rank, int world, size_t max_record_bytes, lb_exchange **out);
void lb_exchange_free(lb_exchange *ex);
/* 64-byte CUDA IPC handle of this rank's receive buffer. */
int lb_exchange_handle(lb_exchange *ex, void *handle64);
/* handles: [world][64] bytes, entry r = rank r's handle (own entry ignored). */
int lb_exchange_connect_ipc(lb_exchange *ex, const void *handles);
/* all: [world] exchange handles of this process, entry r = rank r (enables peer access as needed). */
int lb_exchange_connect_local(lb_exchange *ex, lb_exchange *const *all);
/* Starts the next batch: device pointers of the local record (inside this rank's own receive slot).
 * Record size = round16(nq*k*4) + nq*k*8 bytes <= max_record_bytes. */
int lb_exchange_slot(lb_exchange *ex, int64_t nq, int k, float **d_distances, int64_t **d_labels);
int lb_exchange_all_gather_merge(lb_exchange *ex, int64_t nq, int k_in, int k, float *d_out_distances,
                                 int64_t *d_out_labels, void *stream);
/* LB_OK unless a merge gave up waiting for a peer (synchronises the device). */
int lb_exchange_error(lb_exchange *ex);

/* ------------------------------------------------------------------------------------------
 * 3c. Row-sharded multi-GPU index driven from ONE process (SURVEY.md 8b "multi-GPU handle (devices[],
 *     row-sharded)", 8e; replaces ShardedHNSW's fan-out + merge, internal/store/sharded_hnsw.go:378-503).
 *     Global rows are split into contiguous ranges of lb_shard_rows_per_shard() rows (a multiple of 64, so
 *     global bitmaps are sliced by whole words); labels are global row positions.  Every device runs the
 *     ordinary search chain and its re-score kernel stores the shard's record directly into the root GPU's
 *     gather buffer over NVLink peer memory; the root merges by (distance, label).  Same certification
 *     contract as lb_index_search (uncertified queries are re-done exhaustively on every shard).
 *     Thread safety: calls on one handle are serialised inside the library (a search occupies every GPU of the
 *     set and owns the root's gather buffer), so callers may hold only a read lock as faiss_gpu.go:108 does.
 * ---------------------------------------------------------------------------------------- */
typedef struct lb_shard lb_shard;
int lb_shard_create(const int *devices, int n_devices, int dim, int dtype, int metric, int64_t total_rows,
                    lb_shard **out);
void lb_shard_free(lb_shard *s);
int lb_shard_add(lb_shard *s, const void *rows, int64_t n); /* appended in global row order */
int64_t lb_shard_size(const lb_shard *s);
int lb_shard_count(const lb_shard *s);
int64_t lb_shard_rows_per_shard(const lb_shard *s);
int lb_shard_set_tombstones(lb_shard *s, const uint64_t *bitmap, int64_t nbits); /* global bitmap */
int lb_shard_search(lb_shard *s, const void *queries, int64_t nq, int k, const uint64_t *allow, float *distances,
                    int64_t *labels);
int64_t lb_shard_last_uncertified(const lb_shard *s);

/* ------------------------------------------------------------------------------------------
 * 4. Product quantisation (internal/pq): ADC LUT build, ADC code scan + top-k', fp32 re-rank.
 * ---------------------------------------------------------------------------------------- */
typedef struct lb_pq lb_pq;

/* From the serialised encoder: [dims,M,K u32 LE][M*K*subDim f32 LE]
 * (internal/pq/persistence.go:15-36).  K must be 256 (simd.ADCDistanceBatch hard-codes the
 * stride, internal/simd/simd.go:350). */
int lb_pq_create(int device, const void *blob, size_t blob_len, lb_pq **out);
void lb_pq_free(lb_pq *pq);
int lb_pq_params(const lb_pq *pq, int *dims, int *m, int *k, int *sub_dim);
/* Append n PQ codes, row-major [n*M] (host / device). */
int lb_pq_add_codes(lb_pq *pq, const uint8_t *codes, int64_t n);
int lb_pq_add_codes_device(lb_pq *pq, const uint8_t *d_codes, int64_t n, void *stream);
int64_t lb_pq_size(const lb_pq *pq);
/* fp32 L2 index holding the raw vectors of the same rows, used by the re-rank stage.
 * Not owned.  NULL detaches (search then returns the ADC top-k). */
int lb_pq_attach_raw(lb_pq *pq, lb_index *raw);
int lb_pq_set_tombstones(lb_pq *pq, const uint64_t *bitmap, int64_t nbits);
/* PQEncoder.BuildADCTable (internal/pq/adc_table.go:15-51): table[m*K+k], squared L2. */
int lb_pq_build_adc_table(lb_pq *pq, const float *query, float *table);
/* PQEncoder.Encode for n vectors (internal/pq/encoder.go:76-136). */
int lb_pq_encode(lb_pq *pq, const float *vectors, int64_t n, uint8_t *codes);
/* PQEncoder.Train (internal/pq/encoder.go:39-73) = TrainKMeans per subspace (internal/pq/kmeans.go:64-151), on the
 * GPU: vectors [n*dims] fp32 (host), init_idx [m*k] = the data rows the centroids of each subspace start from (the
 * reference draws them with rand.Perm; supplying them makes the run reproducible and, given equal indices,
 * bit-identical to the reference's arithmetic), max_iter <= 0 -> 20.  Output codebooks [m][k][dims/m] in the layout
 * lb_pq_create's blob expects; iters_run [m] (optional) = iterations each subspace ran before the reference's
 * early-stop rule fired.  Empty clusters are re-seeded from row (c*7919 + iter*104729) mod n. */
int lb_pq_train(int device, const float *vectors, int64_t n, int dims, int m, int k, int max_iter,
                const int32_t *init_idx, float *codebooks, int32_t *iters_run);
/* ADC scan of all resident codes for nq fp32 queries, fused top-kprime, then (if a raw index
 * is attached) exact fp32 Euclidean re-rank of those candidates -> top-k. */
int lb_pq_search(lb_pq *pq, const float *queries, int64_t nq, int k, int kprime, const uint64_t *allow,
                 float *distances, int64_t *labels);
int lb_pq_search_device(lb_pq *pq, const float *d_queries, int64_t nq, int k, int kprime,
                        const uint64_t *d_allow, float *d_distances, int64_t *d_labels, void *stream);
/* The scan is a coarse pass over an integer-quantised, bank-conflict-free LUT followed by the reference's
 * sequential fp32 sum on the surviving candidates and a certification test.  lb_pq_search
 * re-does uncertified queries with the exhaustive fp32 kernel before returning and reports how many there were
 * here; the device entry point reports flags / an accumulating count instead (either may be NULL). */
int64_t lb_pq_last_uncertified(const lb_pq *pq);
int lb_pq_search_device_cert(lb_pq *pq, const float *d_queries, int64_t nq, int k, int kprime,
                             const uint64_t *d_allow, float *d_distances, int64_t *d_labels,
                             uint32_t *d_uncert_flags, uint32_t *d_uncert_count, void *stream);

/* ------------------------------------------------------------------------------------------
 * 4b. HNSW layer search on the GPU: ArrowHNSW.searchLayer
 *     (internal/store/arrow_hnsw.go:1108-1385) for a batch of queries, one warp per query, same algorithm
 *     step for step (candidate min-heap, result max-heap of ef, strict acceptance test, stop test), so the
 *     frontier equals the CPU walk's; heaps order on (distance, id).  The graph is built on the host; its layer
 *     arrays are mirrored as they lie in GraphData (internal/store/types/graph_data.go:605-670): per node
 *     `counts[id]` neighbours at neighbors[id * max_degree ...] (the 1024-node chunks concatenated).
 *     Node ids are rows of the lb_index the graph was created on.
 * ---------------------------------------------------------------------------------------- */
typedef struct lb_graph lb_graph;
int lb_graph_create(lb_index *idx, int max_degree, lb_graph **out);
void lb_graph_free(lb_graph *g);
/* neighbors [n][max_degree] uint32 (ids >= n are ignored: 0xffffffff padding), counts [n] int32 or NULL. */
int lb_graph_set_layer(lb_graph *g, const uint32_t *neighbors, const int32_t *counts, int64_t n);
int lb_graph_set_layer_device(lb_graph *g, const uint32_t *d_neighbors, const int32_t *d_counts, int64_t n,
                              void *stream);
/* searchLayer from one entry point per query: ids / distances [nq*ef] ascending (0xffffffff / FLT_MAX padding),
 * visited [nq] (optional) = distances computed per query. */
int lb_graph_search_layer(lb_graph *g, const void *queries, int64_t nq, const uint32_t *entry_points, int ef,
                          uint32_t *ids, float *distances, uint32_t *visited);
/* Config 5 end to end: the walk, then the re-rank of its ef candidates with the index's tombstones and the
 * `allow` predicate bitmap applied in-kernel (lb_index_rerank) -> [nq*k]. */
int lb_graph_search(lb_graph *g, const void *queries, int64_t nq, const uint32_t *entry_points, int ef, int k,
                    const uint64_t *allow, float *distances, int64_t *labels);
/* Device-pointer forms.  *d_fail_count (caller-zeroed) is incremented for every query whose walk outgrew its
 * visited-set or candidate limits (their outputs are padding); the host forms retry those with larger tables. */
int lb_graph_search_layer_device(lb_graph *g, const void *d_queries, int64_t nq, const uint32_t *d_entry_points,
                                 int ef, uint32_t *d_ids, float *d_distances, uint32_t *d_visited,
                                 uint32_t *d_fail_count, void *stream);
int lb_graph_search_device(lb_graph *g, const void *d_queries, int64_t nq, const uint32_t *d_entry_points, int ef,
                           int k, const uint64_t *d_allow, float *d_distances, int64_t *d_labels,
                           uint32_t *d_fail_count, void *stream);

/* ------------------------------------------------------------------------------------------
 * 5. Predicate -> dense bitmap (internal/simd/simd.go:572-761 compare kernels,
 *    internal/query/filter_evaluator.go:700-758).  op: 0 ==, 1 !=, 2 >, 3 >=, 4 <, 5 <=.
 *    The result is AND-ed into `bitmap` when and_into != 0, else overwrites it.
 * ---------------------------------------------------------------------------------------- */
int lb_filter_i64(int device, const int64_t *column, int64_t n, int op, int64_t value, int and_into,
                  uint64_t *bitmap);
int lb_filter_f32(int device, const float *column, int64_t n, int op, float value, int and_into,
                  uint64_t *bitmap);
/* Same with the column and the bitmap (ceil(n/64) words, zero-initialised unless and_into) resident on the device:
 * a multi-predicate filter is a chain of these on one stream, and the result feeds lb_index_search_device's
 * d_allow without ever visiting the host. */
int lb_filter_i64_device(int device, const int64_t *d_column, int64_t n, int op, int64_t value, int and_into,
                         uint64_t *d_bitmap, void *stream);
int lb_filter_f32_device(int device, const float *d_column, int64_t n, int op, float value, int and_into,
                         uint64_t *d_bitmap, void *stream);

/* The scatter step of Dataset.GenerateFilterBitset (internal/store/dataset.go:247-300): every set bit i of a
 * record batch's match bitmap (lb_filter_*_device over that batch's column) sets bit VectorID(i) of the global
 * allow-bitmap.  d_vector_ids [n_rows] gives the VectorID of each row (0xffffffff = the index has none, as when
 * GetVectorID misses); NULL means the batch was indexed contiguously: VectorID = vid_base + i.  Ids >=
 * n_vector_ids are ignored.  Call once per record batch on the same stream; the result feeds d_allow. */
int lb_filter_scatter_device(int device, const uint64_t *d_batch_bitmap, int64_t n_rows,
                             const uint32_t *d_vector_ids, uint32_t vid_base, int64_t n_vector_ids,
                             uint64_t *d_global_bitmap, void *stream);

/* Count of kernel launches issued by this library in this process (bench evidence). */
int64_t lb_kernel_launch_count(void);
/* Test / diagnostic knobs.  "dense_scan": 0 = auto (tensor-core scan when the index is eligible:
 * fp16 or int8, 16-byte row pitch), 1 = force the SIMT scan, 2 = force the tensor-core scan
 * (LB_ERR_UNSUPPORTED if not eligible).  Both scans feed the same exact re-score stage.
 * "dense_scan" 3 = force the streaming scan (1..8 queries).  "tc_pair": CTA-pair MMAs on/off.
 * "f32_tc": 3xTF32 tensor-core scan for fp32 indexes on/off.  "tc_boot_tiles": bootstrap sample size.
 * "tc_boot": 1 (default) = bootstrap-threshold pre-scan for the tensor-core path, 0 = off.
 * "tc_debug": timing probes of the tensor-core scan; results are INVALID when non-zero.
 * "pq_scan": 0 = auto (look-up passes below 64 queries per call, decode + tensor-core coarse stage from 64 up),
 *            1 = exhaustive fp32 ADC kernel, 2 = coarse look-up scan one query per pass, 3 = four per pass,
 *            4 = decode the codes to fp16 slabs and run the dense tensor-core scan (csrc/pq_gemm.cu).
 * Every mode returns the same (id, distance) pairs: the coarse stages only pick candidates for the exact stage.
 * "exhaustive_k": single-query searches with k >= this value (default 256) take the exhaustive exact chain instead of
 *            the coarse scan (cheaper there, tools/large_k_probe.py); 0 = only k beyond the fused selector (k > 704).
 * "pq_gemm": 1 (default) lets the automatic policy take path 4 for batches; 0 keeps batches on the look-up passes
 *            (path 4 borrows up to 4 Mi x dims x 2 bytes of scratch per concurrent search). */
int lb_set_option(const char *name, int value);
/* Profiling hook for bench.py's roofline: when enabled, every search brackets its dominant
 * kernel (the coarse distance scan: dense or ADC) with CUDA events on the launching stream.
 * lb_prof_read waits for the recorded events, returns their summed duration, their count and the
 * (queries x rows) pairs those launches scanned, and (reset != 0) clears them. */
int lb_prof_enable(int on);
int lb_prof_read(double *total_ms, int64_t *launches, double *units, int reset);
/* Second channel: auxiliary streaming kernels with a roofline of their own (the batched PQ path's decode; units =
 * bytes moved).  Same semantics. */
int lb_prof_read_aux(double *total_ms, int64_t *launches, double *units, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LONGBOW_B200_H */
