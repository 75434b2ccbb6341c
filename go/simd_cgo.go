//go:build gpu && linux

// Thin cgo wrappers with the signatures of internal/simd/batch_operations.go and
// internal/pq/adc_table.go, so internal/store and internal/pq can route their batch calls to
// liblongbow_b200.so.  Source only (no Go toolchain in the build image; see INTEGRATION.md).
package gpu

/*
#include "longbow_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// EuclideanDistanceBatchFlat mirrors simd.EuclideanDistanceBatchFlat (internal/simd/simd.go:203-229).
func EuclideanDistanceBatchFlat(device int, query, flat []float32, n, dims int, results []float32) error {
	if n == 0 {
		return nil
	}
	if len(flat) < n*dims || len(results) < n || len(query) != dims {
		return errors.New("simd: size mismatch")
	}
	return call("simd: batch distance", func() C.int {
		return C.lb_simd_distance_batch_flat(C.int(device), C.int(MetricEuclidean), C.int(Float32),
			unsafe.Pointer(&query[0]), unsafe.Pointer(&flat[0]), C.int64_t(n), C.int(dims), (*C.float)(unsafe.Pointer(&results[0])))
	})
}

// ADCDistanceBatch mirrors simd.ADCDistanceBatch (internal/simd/batch_operations.go:119-127).
func ADCDistanceBatch(device int, table []float32, flatCodes []byte, m int, results []float32) error {
	if len(table) == 0 || len(flatCodes) == 0 {
		return errors.New("simd: empty table or codes")
	}
	if m <= 0 {
		return errors.New("simd: invalid m parameter")
	}
	if len(results) == 0 || len(flatCodes) < len(results)*m {
		return errors.New("simd: results / codes size mismatch")
	}
	return call("simd: ADC batch distance", func() C.int {
		return C.lb_simd_adc_distance_batch(C.int(device), (*C.float)(unsafe.Pointer(&table[0])),
			(*C.uint8_t)(unsafe.Pointer(&flatCodes[0])), C.int(m), C.int64_t(len(results)), (*C.float)(unsafe.Pointer(&results[0])))
	})
}

// MergeTopK mirrors the tail of ShardedHNSW.SearchVectors (internal/store/sharded_hnsw.go:432-503):
// [parts][nq][kIn] per-shard lists with global ids -> [nq][k] by (distance, id).
func MergeTopK(device int, distances []float32, labels []int64, parts, nq, kIn, k int) ([]float32, []int64, error) {
	if parts <= 0 || nq <= 0 || kIn <= 0 || k <= 0 || len(distances) != parts*nq*kIn || len(labels) != len(distances) {
		return nil, nil, errors.New("merge: parts x nq x kIn does not match the list lengths")
	}
	od := make([]float32, nq*k)
	ol := make([]int64, nq*k)
	if err := call("GPU merge", func() C.int {
		return C.lb_merge_topk(C.int(device), (*C.float)(unsafe.Pointer(&distances[0])), (*C.int64_t)(unsafe.Pointer(&labels[0])),
			C.int(parts), C.int64_t(nq), C.int(kIn), C.int(k), (*C.float)(unsafe.Pointer(&od[0])), (*C.int64_t)(unsafe.Pointer(&ol[0])))
	}); err != nil {
		return nil, nil, err
	}
	return od, ol, nil
}
