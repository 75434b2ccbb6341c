//go:build gpu && linux

// Mirrors internal/gpu/gpu_test.go for the B200 backend (source only: no Go toolchain in the build image).
package gpu

import (
	"testing"
)

func TestB200Index_AddSearch(t *testing.T) {
	idx, err := NewB200Index(GPUConfig{DeviceID: 0, Dimension: 128})
	if err != nil {
		t.Skipf("GPU not available: %v", err)
	}
	defer idx.Close()

	vectors := make([]float32, 128*10) // 10 vectors, as internal/gpu/gpu_test.go:25-33
	for i := range vectors {
		vectors[i] = float32(i) * 0.01
	}
	ids := make([]int64, 10)
	for i := range ids {
		ids[i] = int64(i)
	}
	if err := idx.Add(ids, vectors); err != nil {
		t.Fatal(err)
	}
	resultIDs, distances, err := idx.Search(vectors[:128], 5)
	if err != nil {
		t.Fatal(err)
	}
	if len(resultIDs) != 5 || len(distances) != 5 {
		t.Fatalf("want 5 results, got %d / %d", len(resultIDs), len(distances))
	}
	if resultIDs[0] != 0 || distances[0] >= 0.01 { // gpu_test.go:45-46
		t.Fatalf("nearest should be the query itself, got id %d dist %v", resultIDs[0], distances[0])
	}
	for i := 1; i < 5; i++ { // (distance, id) ascending: our strengthening of the reference's unstable order
		if distances[i] < distances[i-1] {
			t.Fatalf("distances not ascending at %d", i)
		}
	}
}

func TestB200Index_InvalidDimension(t *testing.T) { // gpu_test.go:49-55
	if _, err := NewB200Index(GPUConfig{DeviceID: 0, Dimension: 0}); err == nil {
		t.Fatal("expected an error for dimension 0")
	}
}

func TestB200Index_BatchAndRerank(t *testing.T) {
	b, err := NewB200IndexTyped(GPUConfig{DeviceID: 0, Dimension: 4}, Float32, MetricEuclidean)
	if err != nil {
		t.Skipf("GPU not available: %v", err)
	}
	defer b.Close()
	// internal/store/arrow_kernels_test.go:56-69: rows {1,2,3,4}, {0,0,0,0}, {2,2,2,2}; query {1,2,3,4}
	rows := []float32{1, 2, 3, 4, 0, 0, 0, 0, 2, 2, 2, 2}
	if err := b.Add([]int64{0, 1, 2}, rows); err != nil {
		t.Fatal(err)
	}
	q := []float32{1, 2, 3, 4}
	ids, d, err := b.SearchBatch(q, 1, 3, nil)
	if err != nil {
		t.Fatal(err)
	}
	if ids[0] != 0 || d[0] != 0 || ids[2] != 1 { // 0, sqrt(6), sqrt(30)
		t.Fatalf("unexpected order %v %v", ids, d)
	}
	// re-rank of candidate ids with the middle row tombstoned
	if err := b.SetTombstones([]uint64{1 << 2}, 3); err != nil {
		t.Fatal(err)
	}
	ids, _, err = b.Rerank(q, 1, []uint32{2, 1, 0, 7}, 4, 3, nil) // id 7 is out of range: dropped
	if err != nil {
		t.Fatal(err)
	}
	if ids[0] != 0 || ids[1] != 1 || ids[2] != -1 {
		t.Fatalf("unexpected rerank result %v", ids)
	}
}
