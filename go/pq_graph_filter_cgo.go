//go:build gpu && linux

// cgo bindings for the remaining entry points of the hot path: the PQ handle (ADC scan + re-rank), the HNSW layer
// walk, predicate bitmaps and select-k.  Same rules as longbow_b200.go: one cgo crossing per call, Go-owned slices
// the library is done with when it returns, error text read on the calling OS thread.  Source only (no Go toolchain
// in the build image; tests/test_abi.py keeps the calls in step with include/longbow_b200.h).
package gpu

/*
#include "longbow_b200.h"
*/
import "C"

import (
	"fmt"
	"sync"
	"unsafe"
)

// ---------------------------------------------------------------------------------------------
// PQ: internal/pq (PQEncoder, BuildADCTable, ADCDistanceBatch) + the ADC scan -> top-k' -> fp32 re-rank -> top-k
// composition of internal/store/parallel_search.go:292-345.
// ---------------------------------------------------------------------------------------------

// B200PQ holds the code mirror of one PQ-encoded column on the device.
type B200PQ struct {
	h    *C.lb_pq
	dims int
	m    int
	k    int
	mu   sync.RWMutex
}

// NewB200PQ takes the encoder exactly as PQEncoder.Serialize writes it (internal/pq/persistence.go:15-36).
func NewB200PQ(device int, blob []byte) (*B200PQ, error) {
	if len(blob) < 12 {
		return nil, fmt.Errorf("data too short for PQ header") // persistence.go:40
	}
	var h *C.lb_pq
	if err := call("GPU PQ create", func() C.int {
		return C.lb_pq_create(C.int(device), unsafe.Pointer(&blob[0]), C.size_t(len(blob)), &h)
	}); err != nil {
		return nil, err
	}
	var dims, m, k, sub C.int
	if err := call("GPU PQ params", func() C.int { return C.lb_pq_params(h, &dims, &m, &k, &sub) }); err != nil {
		C.lb_pq_free(h)
		return nil, err
	}
	return &B200PQ{h: h, dims: int(dims), m: int(m), k: int(k)}, nil
}

// AddCodes appends n = len(codes)/M encoded rows (flatCodes layout, adc_table.go:57).
func (p *B200PQ) AddCodes(codes []byte) error {
	p.mu.Lock()
	defer p.mu.Unlock()
	if p.h == nil {
		return fmt.Errorf("PQ handle is closed")
	}
	if len(codes) == 0 {
		return nil
	}
	if len(codes)%p.m != 0 {
		return fmt.Errorf("code data length %d not divisible by M=%d", len(codes), p.m)
	}
	return call("GPU PQ add codes", func() C.int {
		return C.lb_pq_add_codes(p.h, (*C.uint8_t)(unsafe.Pointer(&codes[0])), C.int64_t(len(codes)/p.m))
	})
}

// AttachRaw names the fp32 index whose rows re-rank the ADC candidates (nil: the ADC top-k itself is returned).
func (p *B200PQ) AttachRaw(raw *B200Index) error {
	p.mu.Lock()
	defer p.mu.Unlock()
	if p.h == nil {
		return fmt.Errorf("PQ handle is closed")
	}
	var rh *C.lb_index
	if raw != nil {
		rh = raw.h
	}
	return call("GPU PQ attach", func() C.int { return C.lb_pq_attach_raw(p.h, rh) })
}

// BuildADCTable mirrors PQEncoder.BuildADCTable (internal/pq/adc_table.go:15-51): M*K squared distances.
func (p *B200PQ) BuildADCTable(query []float32) ([]float32, error) {
	p.mu.RLock()
	defer p.mu.RUnlock()
	if p.h == nil {
		return nil, fmt.Errorf("PQ handle is closed")
	}
	if len(query) != p.dims {
		return nil, fmt.Errorf("query dimension mismatch") // adc_table.go:17-19
	}
	table := make([]float32, p.m*p.k)
	if err := call("GPU PQ table", func() C.int {
		return C.lb_pq_build_adc_table(p.h, (*C.float)(unsafe.Pointer(&query[0])), (*C.float)(unsafe.Pointer(&table[0])))
	}); err != nil {
		return nil, err
	}
	return table, nil
}

// Encode mirrors PQEncoder.Encode for a batch (internal/pq/encoder.go:76-136): n rows -> n*M codes.
func (p *B200PQ) Encode(vectors []float32) ([]byte, error) {
	p.mu.RLock()
	defer p.mu.RUnlock()
	if p.h == nil {
		return nil, fmt.Errorf("PQ handle is closed")
	}
	if len(vectors) == 0 || len(vectors)%p.dims != 0 {
		return nil, fmt.Errorf("vector dimension mismatch") // encoder.go:77-79
	}
	n := len(vectors) / p.dims
	codes := make([]byte, n*p.m)
	if err := call("GPU PQ encode", func() C.int {
		return C.lb_pq_encode(p.h, (*C.float)(unsafe.Pointer(&vectors[0])), C.int64_t(n), (*C.uint8_t)(unsafe.Pointer(&codes[0])))
	}); err != nil {
		return nil, err
	}
	return codes, nil
}

// Search: ADC scan over every code, the kPrime best re-ranked with fp32 rows (when attached), k returned.
// allow is an optional dense predicate bitmap over the code rows.
func (p *B200PQ) Search(queries []float32, nq, k, kPrime int, allow []uint64) ([]int64, []float32, error) {
	p.mu.RLock()
	defer p.mu.RUnlock()
	if p.h == nil {
		return nil, nil, fmt.Errorf("PQ handle is closed")
	}
	if nq <= 0 || k <= 0 || len(queries) != nq*p.dims {
		return nil, nil, fmt.Errorf("bad query batch: %d values for %d queries of dimension %d, k=%d", len(queries), nq, p.dims, k)
	}
	if need := (int64(C.lb_pq_size(p.h)) + 63) / 64; len(allow) > 0 && int64(len(allow)) < need {
		return nil, nil, fmt.Errorf("allow bitmap has %d words, %d needed", len(allow), need)
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	if err := call("GPU PQ search", func() C.int {
		return C.lb_pq_search(p.h, (*C.float)(unsafe.Pointer(&queries[0])), C.int64_t(nq), C.int(k), C.int(kPrime), ap,
			(*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		return nil, nil, err
	}
	return labels, distances, nil
}

func (p *B200PQ) Close() error {
	p.mu.Lock()
	defer p.mu.Unlock()
	if p.h != nil {
		C.lb_pq_free(p.h)
		p.h = nil
	}
	return nil
}

// ---------------------------------------------------------------------------------------------
// HNSW layer walk: ArrowHNSW.searchLayer (internal/store/arrow_hnsw.go:1108-1385) over the adjacency layout of
// GraphData (internal/store/types/graph_data.go:605-670), many queries per call.
// ---------------------------------------------------------------------------------------------

type B200Graph struct {
	h         *C.lb_graph
	idx       *B200Index
	maxDegree int
	mu        sync.RWMutex
}

func NewB200Graph(idx *B200Index, maxDegree int) (*B200Graph, error) {
	if idx == nil || maxDegree <= 0 {
		return nil, fmt.Errorf("index and a positive maximum degree are required")
	}
	var h *C.lb_graph
	if err := call("GPU graph create", func() C.int { return C.lb_graph_create(idx.h, C.int(maxDegree), &h) }); err != nil {
		return nil, err
	}
	return &B200Graph{h: h, idx: idx, maxDegree: maxDegree}, nil
}

// SetLayer uploads layer 0: neighbors[id*maxDegree ...] and counts[id] for n = len(counts) nodes.
func (g *B200Graph) SetLayer(neighbors []uint32, counts []int32) error {
	g.mu.Lock()
	defer g.mu.Unlock()
	if g.h == nil {
		return fmt.Errorf("graph is closed")
	}
	if len(counts) == 0 || len(neighbors) != len(counts)*g.maxDegree {
		return fmt.Errorf("adjacency holds %d ids for %d nodes of degree %d", len(neighbors), len(counts), g.maxDegree)
	}
	return call("GPU graph set layer", func() C.int {
		return C.lb_graph_set_layer(g.h, (*C.uint32_t)(unsafe.Pointer(&neighbors[0])), (*C.int32_t)(unsafe.Pointer(&counts[0])),
			C.int64_t(len(counts)))
	})
}

// Search walks layer 0 from one entry point per query with the given ef, applies tombstones and the optional
// predicate bitmap in the kernel, and returns the k best of the frontier re-ranked exactly (fp32 indexes).
func (g *B200Graph) Search(queries []float32, nq int, entryPoints []uint32, ef, k int, allow []uint64) ([]int64, []float32, error) {
	g.mu.RLock()
	defer g.mu.RUnlock()
	if g.h == nil {
		return nil, nil, fmt.Errorf("graph is closed")
	}
	if g.idx.dtype != Float32 {
		return nil, nil, fmt.Errorf("Search([]float32) on a non-fp32 index")
	}
	if nq <= 0 || k <= 0 || ef <= 0 || len(entryPoints) != nq || len(queries) != nq*g.idx.dim {
		return nil, nil, fmt.Errorf("bad walk request: %d queries, %d entry points, ef=%d, k=%d", nq, len(entryPoints), ef, k)
	}
	if err := g.idx.checkAllow(allow); err != nil {
		return nil, nil, err
	}
	distances := make([]float32, nq*k)
	labels := make([]int64, nq*k)
	var ap *C.uint64_t
	if len(allow) > 0 {
		ap = (*C.uint64_t)(unsafe.Pointer(&allow[0]))
	}
	if err := call("GPU graph search", func() C.int {
		return C.lb_graph_search(g.h, unsafe.Pointer(&queries[0]), C.int64_t(nq), (*C.uint32_t)(unsafe.Pointer(&entryPoints[0])),
			C.int(ef), C.int(k), ap, (*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])))
	}); err != nil {
		return nil, nil, err
	}
	return labels, distances, nil
}

func (g *B200Graph) Close() error {
	g.mu.Lock()
	defer g.mu.Unlock()
	if g.h != nil {
		C.lb_graph_free(g.h)
		g.h = nil
	}
	return nil
}

// ---------------------------------------------------------------------------------------------
// Predicates and selection: simd.MatchInt64 / MatchFloat32 (internal/simd/simd.go:585-690) feeding
// GenerateFilterBitset (internal/query/filter_evaluator.go:700-758), and the Arrow compute function
// select_k_neighbors (internal/store/arrow_kernels.go:230-345).
// ---------------------------------------------------------------------------------------------

// CompareOp follows simd.CompareOp (internal/simd/simd.go:38-45).
type CompareOp int

const (
	CompareEq CompareOp = iota
	CompareNeq
	CompareGt
	CompareGe
	CompareLt
	CompareLe
)

// FilterInt64 evaluates column[i] <op> value into a dense bitmap (bit i = row i passes).  With andInto the result
// is AND-combined into bitmap (a conjunction of predicates), otherwise bitmap is overwritten; nil allocates one.
func FilterInt64(device int, column []int64, op CompareOp, value int64, bitmap []uint64, andInto bool) ([]uint64, error) {
	words := (len(column) + 63) / 64
	if bitmap == nil {
		bitmap = make([]uint64, words)
		andInto = false
	}
	if len(column) == 0 {
		return bitmap, nil
	}
	if len(bitmap) < words {
		return nil, fmt.Errorf("bitmap has %d words, %d rows need %d", len(bitmap), len(column), words)
	}
	and := 0
	if andInto {
		and = 1
	}
	if err := call("GPU filter", func() C.int {
		return C.lb_filter_i64(C.int(device), (*C.int64_t)(unsafe.Pointer(&column[0])), C.int64_t(len(column)), C.int(op),
			C.int64_t(value), C.int(and), (*C.uint64_t)(unsafe.Pointer(&bitmap[0])))
	}); err != nil {
		return nil, err
	}
	return bitmap, nil
}

// FilterFloat32 is FilterInt64 for float32 columns.
func FilterFloat32(device int, column []float32, op CompareOp, value float32, bitmap []uint64, andInto bool) ([]uint64, error) {
	words := (len(column) + 63) / 64
	if bitmap == nil {
		bitmap = make([]uint64, words)
		andInto = false
	}
	if len(column) == 0 {
		return bitmap, nil
	}
	if len(bitmap) < words {
		return nil, fmt.Errorf("bitmap has %d words, %d rows need %d", len(bitmap), len(column), words)
	}
	and := 0
	if andInto {
		and = 1
	}
	if err := call("GPU filter", func() C.int {
		return C.lb_filter_f32(C.int(device), (*C.float)(unsafe.Pointer(&column[0])), C.int64_t(len(column)), C.int(op),
			C.float(value), C.int(and), (*C.uint64_t)(unsafe.Pointer(&bitmap[0])))
	}); err != nil {
		return nil, err
	}
	return bitmap, nil
}

// SelectK returns the indices (and values) of the k smallest distances, ascending by (value, index).
func SelectK(device int, distances []float32, k int) ([]int64, []float32, error) {
	if k <= 0 || len(distances) == 0 {
		return nil, nil, nil
	}
	if k > len(distances) {
		k = len(distances)
	}
	indices := make([]int64, k)
	values := make([]float32, k)
	if err := call("GPU select k", func() C.int {
		return C.lb_select_k(C.int(device), (*C.float)(unsafe.Pointer(&distances[0])), C.int64_t(len(distances)), C.int(k),
			(*C.int64_t)(unsafe.Pointer(&indices[0])), (*C.float)(unsafe.Pointer(&values[0])))
	}); err != nil {
		return nil, nil, err
	}
	return indices, values, nil
}
