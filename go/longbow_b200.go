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
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/longbow_b200/lib -llongbow_b200 -lcuda
#include <stdlib.h>
#include "longbow_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
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

func lastErr(code C.int) error {
	return fmt.Errorf("longbow_b200: code %d: %s", int(code), C.GoString(C.lb_last_error()))
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
	if rc := C.lb_index_create(C.int(cfg.DeviceID), C.int(cfg.Dimension), C.int(dt), C.int(m), &h); rc != 0 {
		return nil, fmt.Errorf("failed to initialize GPU resources for device %d: %w", cfg.DeviceID, lastErr(rc))
	}
	idx := &B200Index{h: h, dim: cfg.Dimension, dtype: dt}
	runtime.SetFinalizer(idx, (*B200Index).Close)
	return idx, nil
}

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
	if rc := C.lb_index_add(idx.h, unsafe.Pointer(&vectors[0]), C.int64_t(n)); rc != 0 {
		return fmt.Errorf("GPU index add failed: %w", lastErr(rc))
	}
	return nil
}

// AddArrowBuffer appends n rows straight from the child values buffer of a FixedSizeList column
// (internal/store/arrow_utils.go:112-171): buf must hold n*dim elements of the index dtype.
func (idx *B200Index) AddArrowBuffer(buf []byte, n int) error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	if idx.closed {
		return fmt.Errorf("index is closed")
	}
	if n == 0 {
		return nil
	}
	if rc := C.lb_index_add(idx.h, unsafe.Pointer(&buf[0]), C.int64_t(n)); rc != 0 {
		return fmt.Errorf("GPU index add failed: %w", lastErr(rc))
	}
	return nil
}

func (idx *B200Index) Search(vector []float32, k int) ([]int64, []float32, error) {
	ids, dists, err := idx.SearchBatch(vector, 1, k, nil)
	return ids, dists, err
}

// SearchBatch answers nq queries (row-major) in one call; allow is an optional dense predicate
// bitmap (bit i = VectorID i passes), e.g. query.Bitset exported with ToDenseWords().
func (idx *B200Index) SearchBatch(queries []float32, nq, k int, allow []uint64) ([]int64, []float32, error) {
	idx.mu.RLock()
	defer idx.mu.RUnlock()
	if idx.closed {
		return nil, nil, fmt.Errorf("index is closed")
	}
	if len(queries) != nq*idx.dim {
		return nil, nil, fmt.Errorf("query vector dimension %d does not match index dimension %d", len(queries)/max(nq, 1), idx.dim)
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	rc := C.lb_index_search(idx.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), C.int(k), ap,
		(*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	if rc != 0 {
		return nil, nil, fmt.Errorf("GPU search failed: %w", lastErr(rc))
	}
	return labels, distances, nil
}

// Rerank replaces processChunkInternal / RerankBatch (internal/store/parallel_search.go:147-365,
// hnsw_batch.go:206-245): candidate VectorIDs from the host graph walk, c per query.
func (idx *B200Index) Rerank(queries []float32, nq int, cand []uint32, c, k int, allow []uint64) ([]int64, []float32, error) {
	idx.mu.RLock()
	defer idx.mu.RUnlock()
	if idx.closed {
		return nil, nil, fmt.Errorf("index is closed")
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	rc := C.lb_index_rerank(idx.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), (*C.uint32_t)(unsafe.Pointer(&cand[0])),
		C.int(c), C.int(k), ap, (*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	if rc != 0 {
		return nil, nil, fmt.Errorf("GPU rerank failed: %w", lastErr(rc))
	}
	return labels, distances, nil
}

// SetTombstones mirrors ArrowHNSW.deleted (internal/store/arrow_hnsw.go:147,468-472) as a dense bitmap.
func (idx *B200Index) SetTombstones(words []uint64, nbits int64) error {
	idx.mu.Lock()
	defer idx.mu.Unlock()
	var p *C.uint64_t
	if len(words) > 0 {
		p = (*C.uint64_t)(unsafe.Pointer(&words[0]))
	}
	if rc := C.lb_index_set_tombstones(idx.h, p, C.int64_t(nbits)); rc != 0 {
		return lastErr(rc)
	}
	return nil
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
