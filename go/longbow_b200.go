//go:build gpu && linux

// Package gpu: B200 backend for Longbow's internal/gpu, over liblongbow_b200.so.
//
// This file is meant to live next to internal/gpu/faiss_gpu.go in the reference tree.  It is
// shipped as source only: the build image of this repository has no Go toolchain, so it has not
// been compiled here (INTEGRATION.md).  It binds exactly the C ABI declared in
// include/longbow_b200.h; every call is one cgo crossing with Go-owned slices that the library
// finishes reading / writing before it returns (cgo pointer rules).
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/longbow_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/longbow_b200/lib -llongbow_b200
#include <stdlib.h>
#include "longbow_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"time"
	"unsafe"

	"github.com/23skdu/longbow/internal/metrics"
)

// Metric / dtype enums follow internal/simd/registry.go:8-14,31-47.
type Metric int

const (
	MetricEuclidean Metric = iota
	MetricCosine
	MetricDot
)

type DType int

const (
	Float32 DType = iota
	Float16
	Int8
)

func (d DType) size() int {
	switch d {
	case Float32:
		return 4
	case Float16:
		return 2
	default:
		return 1
	}
}

// call runs one C entry point and, on failure, reads the library's thread-local error text ON THE SAME OS THREAD
// (a goroutine may migrate between two cgo calls, so the pair is bracketed by LockOSThread).
func call(what string, f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return fmt.Errorf("%s failed with code %d: %s", what, int(rc), C.GoString(C.lb_last_error()))
	}
	return nil
}

// B200Index implements gpu.Index (internal/gpu/interface.go:10-19) plus the batched extensions
// internal/store type-asserts for (SearchBatch, Rerank, SetTombstones).
type B200Index struct {
	h      *C.lb_index
	dim    int
	dtype  DType
	mu     sync.RWMutex
	closed bool
}

// NewB200Index mirrors NewFaissGPUIndex (faiss_gpu.go:45-72): fp32, Euclidean.
func NewB200Index(cfg GPUConfig) (Index, error) {
	return NewB200IndexTyped(cfg, Float32, MetricEuclidean)
}

func NewB200IndexTyped(cfg GPUConfig, dt DType, m Metric) (*B200Index, error) {
	if cfg.Dimension <= 0 {
		return nil, fmt.Errorf("dimension must be positive, got %d", cfg.Dimension)
	}
	var h *C.lb_index
	if err := call("GPU index create", func() C.int {
		return C.lb_index_create(C.int(cfg.DeviceID), C.int(cfg.Dimension), C.int(dt), C.int(m), &h)
	}); err != nil {
		return nil, fmt.Errorf("failed to initialize GPU resources for device %d: %w", cfg.DeviceID, err)
	}
	idx := &B200Index{h: h, dim: cfg.Dimension, dtype: dt}
	runtime.SetFinalizer(idx, func(i *B200Index) { _ = i.Close() })
	return idx, nil
}

func (idx *B200Index) size() int64 { return int64(C.lb_index_size(idx.h)) }

// Add: ids are not passed down (labels are insertion positions, as in faiss_gpu.go:93-97).
func (idx *B200Index) Add(ids []int64, vectors []float32) error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	if idx.closed {
		return fmt.Errorf("index is closed")
	}
	if idx.dtype != Float32 {
		return fmt.Errorf("Add([]float32) on a non-fp32 index; use AddArrowBuffer")
	}
	if len(vectors)%idx.dim != 0 {
		return fmt.Errorf("vector data length %d not divisible by dimension %d", len(vectors), idx.dim)
	}
	n := len(vectors) / idx.dim
	if len(ids) != n {
		return fmt.Errorf("id count %d does not match vector count %d", len(ids), n)
	}
	if n == 0 {
		return nil
	}
	start := time.Now()
	if err := call("GPU index add", func() C.int {
		return C.lb_index_add(idx.h, unsafe.Pointer(&vectors[0]), C.int64_t(n))
	}); err != nil {
		metrics.VectorSearchGPUOperationsTotal.WithLabelValues("add", "error").Inc()
		return err
	}
	metrics.VectorSearchGPULatencySeconds.WithLabelValues("add").Observe(time.Since(start).Seconds())
	metrics.VectorSearchGPUOperationsTotal.WithLabelValues("add", "success").Inc()
	return nil
}

// AddArrowBuffer appends n rows straight from the child values buffer of a FixedSizeList column
// (internal/store/arrow_utils.go:112-171): listOffset is the list array's Offset(); the library applies the
// same truncated-buffer rule the reference does and page-locks the buffer for the upload.
func (idx *B200Index) AddArrowBuffer(buf []byte, listOffset, n int) error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	if idx.closed {
		return fmt.Errorf("index is closed")
	}
	if n == 0 {
		return nil
	}
	if len(buf) < n*idx.dim*idx.dtype.size() {
		return fmt.Errorf("values buffer holds %d bytes, %d rows need %d", len(buf), n, n*idx.dim*idx.dtype.size())
	}
	return call("GPU index add", func() C.int {
		return C.lb_index_add_arrow(idx.h, unsafe.Pointer(&buf[0]), C.size_t(len(buf)), C.int64_t(listOffset), C.int64_t(n), 1)
	})
}

func (idx *B200Index) Search(vector []float32, k int) ([]int64, []float32, error) {
	if len(vector) != idx.dim { // faiss_gpu.go:115-117
		return nil, nil, fmt.Errorf("query vector dimension %d does not match index dimension %d", len(vector), idx.dim)
	}
	return idx.SearchBatch(vector, 1, k, nil)
}

func (idx *B200Index) checkAllow(allow []uint64) error {
	if len(allow) == 0 {
		return nil
	}
	if need := (idx.size() + 63) / 64; int64(len(allow)) < need {
		return fmt.Errorf("allow bitmap has %d words, index of %d rows needs %d", len(allow), idx.size(), need)
	}
	return nil
}

// SearchBatch answers nq fp32 queries (row-major) in one call; allow is an optional dense predicate bitmap
// (bit i = VectorID i passes) -- see DenseWords below for the roaring -> dense conversion.
func (idx *B200Index) SearchBatch(queries []float32, nq, k int, allow []uint64) ([]int64, []float32, error) {
	if idx.dtype != Float32 {
		return nil, nil, fmt.Errorf("SearchBatch([]float32) on a non-fp32 index; use SearchBatchBytes")
	}
	if nq <= 0 || k <= 0 {
		return nil, nil, fmt.Errorf("nq and k must be positive, got %d and %d", nq, k)
	}
	if len(queries) != nq*idx.dim {
		return nil, nil, fmt.Errorf("query vector dimension %d does not match index dimension %d", len(queries)/nq, idx.dim)
	}
	return idx.searchRaw(unsafe.Pointer(&queries[0]), nq, k, allow)
}

// SearchBatchBytes: queries in the index's own element type (fp16 / int8 indexes), nq*dim elements.
func (idx *B200Index) SearchBatchBytes(queries []byte, nq, k int, allow []uint64) ([]int64, []float32, error) {
	if nq <= 0 || k <= 0 {
		return nil, nil, fmt.Errorf("nq and k must be positive, got %d and %d", nq, k)
	}
	if len(queries) != nq*idx.dim*idx.dtype.size() {
		return nil, nil, fmt.Errorf("query buffer holds %d bytes, %d queries of dimension %d need %d", len(queries), nq, idx.dim, nq*idx.dim*idx.dtype.size())
	}
	return idx.searchRaw(unsafe.Pointer(&queries[0]), nq, k, allow)
}

func (idx *B200Index) searchRaw(q unsafe.Pointer, nq, k int, allow []uint64) ([]int64, []float32, error) {
	idx.mu.RLock()
	defer idx.mu.RUnlock()
	if idx.closed {
		return nil, nil, fmt.Errorf("index is closed")
	}
	if err := idx.checkAllow(allow); err != nil {
		return nil, nil, err
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	start := time.Now()
	if err := call("GPU search", func() C.int {
		return C.lb_index_search(idx.h, q, C.int64_t(nq), C.int(k), ap,
			(*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		metrics.VectorSearchGPUOperationsTotal.WithLabelValues("search", "error").Inc() // faiss_gpu.go:135
		return nil, nil, err
	}
	metrics.VectorSearchGPULatencySeconds.WithLabelValues("search").Observe(time.Since(start).Seconds()) // :139-141
	metrics.VectorSearchGPUOperationsTotal.WithLabelValues("search", "success").Inc()
	return labels, distances, nil
}

// Rerank replaces processChunkInternal / RerankBatch (internal/store/parallel_search.go:147-365,
// hnsw_batch.go:206-245): candidate VectorIDs from the host graph walk, c per query (fp32 indexes).
func (idx *B200Index) Rerank(queries []float32, nq int, cand []uint32, c, k int, allow []uint64) ([]int64, []float32, error) {
	idx.mu.RLock()
	defer idx.mu.RUnlock()
	if idx.closed {
		return nil, nil, fmt.Errorf("index is closed")
	}
	if idx.dtype != Float32 {
		return nil, nil, fmt.Errorf("Rerank([]float32) on a non-fp32 index")
	}
	if nq <= 0 || k <= 0 || c <= 0 {
		return nil, nil, fmt.Errorf("nq, c and k must be positive, got %d, %d and %d", nq, c, k)
	}
	if len(queries) != nq*idx.dim {
		return nil, nil, fmt.Errorf("query vector dimension %d does not match index dimension %d", len(queries)/nq, idx.dim)
	}
	if len(cand) != nq*c {
		return nil, nil, fmt.Errorf("candidate count %d does not match %d queries x %d", len(cand), nq, c)
	}
	if err := idx.checkAllow(allow); err != nil {
		return nil, nil, err
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	if err := call("GPU rerank", func() C.int {
		return C.lb_index_rerank(idx.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), (*C.uint32_t)(unsafe.Pointer(&cand[0])),
			C.int(c), C.int(k), ap, (*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		return nil, nil, err
	}
	return labels, distances, nil
}

// SetTombstones mirrors ArrowHNSW.deleted (internal/store/arrow_hnsw.go:147,468-472) as a dense bitmap.
func (idx *B200Index) SetTombstones(words []uint64, nbits int64) error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	if idx.closed {
		return fmt.Errorf("index is closed")
	}
	if nbits > int64(len(words))*64 {
		return fmt.Errorf("tombstone bitmap has %d words for %d bits", len(words), nbits)
	}
	var p *C.uint64_t
	if len(words) > 0 {
		p = (*C.uint64_t)(unsafe.Pointer(&words[0]))
	}
	return call("GPU set tombstones", func() C.int { return C.lb_index_set_tombstones(idx.h, p, C.int64_t(nbits)) })
}

func (idx *B200Index) Close() error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	if idx.closed {
		return nil
	}
	C.lb_index_free(idx.h)
	idx.h = nil
	idx.closed = true
	return nil
}

// DenseWords converts a predicate bitset (internal/query/bitmap.go: a roaring bitmap) into the dense little-endian
// words the library consumes, through the accessor the reference has (ToUint32Array, bitmap.go:95-100).
func DenseWords(ids []uint32, nbits int64) []uint64 {
	words := make([]uint64, (nbits+63)/64)
	for _, id := range ids {
		if int64(id) < nbits {
			words[id>>6] |= 1 << (id & 63)
		}
	}
	return words
}

// ShardedB200Index: rows split over several GPUs of this process (lb_shard_*), the drop-in for ShardedHNSW's
// fan-out + merge (internal/store/sharded_hnsw.go:378-503).  Labels are global row positions.
type ShardedB200Index struct {
	h     *C.lb_shard
	dim   int
	dtype DType
	mu    sync.RWMutex
}

func NewShardedB200Index(devices []int, dim int, dt DType, m Metric, totalRows int64) (*ShardedB200Index, error) {
	if len(devices) == 0 || dim <= 0 || totalRows <= 0 {
		return nil, fmt.Errorf("devices, dimension and totalRows must be non-empty / positive")
	}
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	var h *C.lb_shard
	if err := call("GPU shard create", func() C.int {
		return C.lb_shard_create(&devs[0], C.int(len(devs)), C.int(dim), C.int(dt), C.int(m), C.int64_t(totalRows), &h)
	}); err != nil {
		return nil, err
	}
	return &ShardedB200Index{h: h, dim: dim, dtype: dt}, nil
}

func (s *ShardedB200Index) Add(vectors []float32) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	if s.h == nil {
		return fmt.Errorf("index is closed")
	}
	if s.dtype != Float32 {
		return fmt.Errorf("Add([]float32) on a non-fp32 shard set; use AddRaw")
	}
	if len(vectors) == 0 {
		return nil
	}
	if len(vectors)%s.dim != 0 {
		return fmt.Errorf("vector data length %d not divisible by dimension %d", len(vectors), s.dim)
	}
	return call("GPU shard add", func() C.int {
		return C.lb_shard_add(s.h, unsafe.Pointer(&vectors[0]), C.int64_t(len(vectors)/s.dim))
	})
}

// AddRaw appends rows of the set's own element type (fp16 / int8 / fp32) from their little-endian bytes, e.g. the
// values buffer of an Arrow FixedSizeList column.
func (s *ShardedB200Index) AddRaw(rows []byte) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	if s.h == nil {
		return fmt.Errorf("index is closed")
	}
	rb := s.dim * s.dtype.size()
	if len(rows) == 0 {
		return nil
	}
	if len(rows)%rb != 0 {
		return fmt.Errorf("row data length %d not divisible by the row size %d", len(rows), rb)
	}
	return call("GPU shard add", func() C.int {
		return C.lb_shard_add(s.h, unsafe.Pointer(&rows[0]), C.int64_t(len(rows)/rb))
	})
}

func (s *ShardedB200Index) SearchBatch(queries []float32, nq, k int) ([]int64, []float32, error) {
	s.mu.RLock()
	defer s.mu.RUnlock()
	if s.h == nil {
		return nil, nil, fmt.Errorf("index is closed")
	}
	if s.dtype != Float32 {
		return nil, nil, fmt.Errorf("SearchBatch([]float32) on a non-fp32 shard set; use SearchBatchRaw")
	}
	if nq <= 0 || k <= 0 || len(queries) != nq*s.dim {
		return nil, nil, fmt.Errorf("bad query batch: %d values for %d queries of dimension %d, k=%d", len(queries), nq, s.dim, k)
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	if err := call("GPU shard search", func() C.int {
		return C.lb_shard_search(s.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), C.int(k), nil,
			(*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		return nil, nil, err
	}
	return labels, distances, nil
}

// SearchBatchRaw: nq queries of the set's own element type as bytes (fp16 / int8 sets).
func (s *ShardedB200Index) SearchBatchRaw(queries []byte, nq, k int) ([]int64, []float32, error) {
	s.mu.RLock()
	defer s.mu.RUnlock()
	if s.h == nil {
		return nil, nil, fmt.Errorf("index is closed")
	}
	if nq <= 0 || k <= 0 || len(queries) != nq*s.dim*s.dtype.size() {
		return nil, nil, fmt.Errorf("bad query batch: %d bytes for %d queries of dimension %d, k=%d", len(queries), nq, s.dim, k)
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	if err := call("GPU shard search", func() C.int {
		return C.lb_shard_search(s.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), C.int(k), nil,
			(*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		return nil, nil, err
	}
	return labels, distances, nil
}

func (s *ShardedB200Index) Close() error {
	s.mu.Lock()
	defer s.mu.Unlock()
	if s.h != nil {
		C.lb_shard_free(s.h)
		s.h = nil
	}
	return nil
}
