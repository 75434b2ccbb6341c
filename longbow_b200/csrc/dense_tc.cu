// dense_tc.cu -- S1 coarse scan on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   D[128 queries x 256 rows] += Q[128 x K] * X[256 x K]^T      (K-major both, 128-byte swizzle)
//
// Warp-specialised persistent kernel, one CTA per SM:
//   warp 0      TMA producer   : cp.async.bulk.tensor tiles of Q and X into a 4-stage smem ring
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=256), accumulators in TMEM
//   warp 2      TMEM allocator : 512 columns = 2 accumulator stages of 256 fp32/s32 columns
//   warps 4..11 epilogue       : two groups of 4 warps drain every tile together (half the columns each);
//                                tcgen05.ld a TMEM lane (= one query) per thread, fold the negated ranking
//                                keys of 32 columns into four 8-column maxima, test those against the
//                                query's running k-th best; only flagged column groups are looked at key by
//                                key, every lane appends its own survivors (key picked out of its registers
//                                by a select tree) to the query's candidate list; warp-cooperative
//                                radix-select compaction when a list fills.
// The 128 x 256 distance tile never leaves the SM: HBM only sees the database once per batch.
//
// Work split: query block b (128 queries) x group g; CTA (b, g) streams row tiles g, g+G, g+2G ...
// so the CTAs of all query blocks touch the same row tile at about the same time and X is read
// from HBM once and from L2 by the others.  Output: partial[g][q][kc] packed (key, row), the same
// format the SIMT scan produces; merge + exact re-score follow (kernels_dense.cu).
#include <cuda.h>

#include "kernels.cuh"

namespace lb {

bool g_tc_pair = true;  // lb_set_option("tc_pair"): CTA-pair (cta_group::2) scan on / off
int g_tc_reserve_sms = 0;
bool g_ssel_warp = true;   // lb_set_option("ssel_warp"): warp-per-query sample selection for small samples

enum { KIND_F16 = 0, KIND_I8 = 1, KIND_TF32 = 2 };
static_assert(LB_NEDGE == 16, "the epilogue unpacks four uint4 of ladder counters");

template <int KIND> struct TcTraits;
template <> struct TcTraits<KIND_F16> {
    static constexpr int kElem = 2, kBlockK = 64;                      // 64 halves = 128 B per row per stage
    static constexpr uint32_t kIdescFmt = (1u << 4) | (0u << 7) | (0u << 10);  // D=f32, A=B=f16
};
template <> struct TcTraits<KIND_I8> {
    static constexpr int kElem = 1, kBlockK = 128;
    static constexpr uint32_t kIdescFmt = (2u << 4) | (1u << 7) | (1u << 10);  // D=s32, A=B=s8
};
template <> struct TcTraits<KIND_TF32> {
    static constexpr int kElem = 4, kBlockK = 32;
    static constexpr uint32_t kIdescFmt = (1u << 4) | (2u << 7) | (2u << 10);  // D=f32, A=B=tf32
};

constexpr int TC_M = 128;       // queries per CTA (TMEM lanes)
constexpr int TC_N = 256;       // rows per tile (TMEM columns per accumulator stage)
constexpr int TC_STAGES = 4;
constexpr int TC_STAGES_PAIR = 6;  // CTA-pair variant: 32 KiB stages
constexpr int TC_A_BYTES = TC_M * 128;   // 16 KiB
constexpr int TC_B_BYTES = TC_N * 128;   // 32 KiB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_THREADS = 384;   // 4 control warps + 2 epilogue groups of 4 warps

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a lost arrival traps (kernel error) after 2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    uint64_t t0 = 0;
    for (uint32_t spins = 0; !ok; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");  // sleep (<= 1 ms) instead of spinning
        if (!ok && (spins & 15u) == 15u) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on a barrier of another CTA of the cluster.  No cluster-scope release: the barriers signalled this way only
// order tensor-memory reads against the next MMA (tcgen05.fence::before_thread_sync does that), never generic-proxy
// data.  The .release.cluster form compiles to ERRBAR + CGAERRBAR, which waits for every outstanding global store
// and RED of the warp (the candidate appends) -- round 2's ncu source view had 20 % of the epilogue's stall samples
// on that fence, on the critical path of handing the accumulator stage back to the MMA issuer.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to `bar` (a shared::cluster
// address, here always the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == KIND_F16) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else if constexpr (KIND == KIND_I8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}

template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == KIND_F16) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else if constexpr (KIND == KIND_I8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return r;
}
__device__ __forceinline__ void tmem_wait_ld4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)::"memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld1(uint32_t& a) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a)::"memory");
}
// v[j] for a run-time j: registers cannot be indexed, a select tree can (16 + 8 + 4 + 2 + 1 SEL)
__device__ __forceinline__ uint32_t sel32(const uint32_t (&v)[32], int j) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = (j & 2) ? t[2 * i + 1] : t[2 * i];
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = (j & 4) ? t[2 * i + 1] : t[2 * i];
#pragma unroll
    for (int i = 0; i < 2; i++) t[i] = (j & 8) ? t[2 * i + 1] : t[2 * i];
    return (j & 16) ? t[1] : t[0];
}
// 3-input maximum (FMNMX3 on sm_100); a NaN input is ignored like fmaxf does
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// Same wait, but naming the 32 destination registers of the load it completes as read-write operands:
// no use of v[] can be scheduled above the wait even when a second tcgen05.ld is already in flight.
__device__ __forceinline__ void tmem_wait_ld32(uint32_t* v) {
    asm volatile(
        "tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);  // start address       bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset  bits [32,46): 8 rows * 128 B
    d |= (uint64_t)1 << 46;                    // descriptor version  bits [46,48) = 1 on sm_100
    d |= (uint64_t)2 << 61;                    // layout type         bits [61,64) = SWIZZLE_128B
    return d;
}

struct TcArgs {
    const float* aux;
    uint32_t n_rows;
    int nq;
    int k_blocks;       // ceil(dim / kBlockK)
    int groups;         // G: CTAs per query block
    int n_row_tiles;    // ceil(n_rows / 256)
    const uint32_t* tomb;
    uint32_t tomb_bits;
    const uint32_t* allow;
    int kc, cap;
    uint64_t* cand;     // [grid][128][cap] per-(CTA, query) candidate lists (global, L2-resident)
    uint64_t* partial;  // [G][nq][kc]
    int debug;          // timing probes only (results invalid): 1 no epilogue work, 2 reuse B tile, 4 reuse A tile
    int tile_begin, tile_end;  // row tiles [begin, end) scanned by this launch
    int part_offset;           // first partial[] slot written by this launch
    const float* tau_init;     // per-query starting threshold (bootstrap), or null
    float* keys_out;           // bootstrap sample mode: write raw keys [nq][keys_ld] instead of selecting
    int keys_ld;
    const float* edges;        // [nq][LB_NEDGE] shared threshold ladder (or null)
    uint32_t* edge_cnt;        // [nq][LB_NEDGE] live rows seen per ladder bucket, all CTAs of the query
};

// ---------------------------------------------------------------------------------------------
// Selection compaction of one candidate list (c entries in global memory, capacity 32*R, entries in
// increasing row order): keep exactly its kc smallest by (key, row) -- radix-select the kc-th key
// MSB-first over the 32-bit ordered key, break ties at that key by list position (== row order) --
// and write them back, order preserved, to dst.  Returns the new threshold (the kc-th key; +inf if
// c < kc, in which case everything is kept).  ~4x fewer instructions than sorting the list.
template <int R>
__device__ __forceinline__ float select_compact(const uint64_t* buf, uint64_t* dst, int c, int kc, int lane,
                                                int* kept_out) {
    uint64_t v[R];
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(buf + lane * R);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        const int i = lane * R + r;
        ulonglong2 t = make_ulonglong2(kInvalid, kInvalid);
        if (i < c) t = __ldcg(src + (r >> 1));
        v[r] = t.x;
        v[r + 1] = (i + 1 < c) ? t.y : kInvalid;
    }
    if (c <= kc) {  // nothing to drop: plain copy (the usual case once the bootstrap threshold is tight)
        __syncwarp();
        if (dst != buf) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (lane * R + r < c) __stcg(dst + lane * R + r, v[r]);
            }
        }
        *kept_out = c;
        return (c == kc) ? INFINITY : INFINITY;
    }
    uint32_t T = 0xffffffffu;
    if (c >= kc) {
        // bits shared by every valid key need no search: start below the common prefix
        uint32_t all_and = 0xffffffffu, all_or = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (v[r] != kInvalid) { all_and &= (uint32_t)(v[r] >> 32); all_or |= (uint32_t)(v[r] >> 32); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            all_and &= __shfl_xor_sync(0xffffffffu, all_and, o);
            all_or |= __shfl_xor_sync(0xffffffffu, all_or, o);
        }
        const uint32_t diff = all_and ^ all_or;
        const int top = diff ? (31 - __clz(diff)) : -1;  // highest bit where keys differ
        T = (top >= 31) ? 0u : (all_and & ~((2u << top) - 1u));
        if (top < 0) T = all_and;
#pragma unroll 1
        for (int bit = top; bit >= 0; bit--) {
            const uint32_t test = T | (1u << bit);
            int less = 0;
#pragma unroll
            for (int r = 0; r < R; r++) less += ((uint32_t)(v[r] >> 32) < test);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) less += __shfl_xor_sync(0xffffffffu, less, o);
            if (less < kc) T = test;  // fewer than kc keys below `test`: the kc-th key is >= test
        }
    }
    // keep key < T, plus the first (kc - #less) entries with key == T in list order
    int n_less = 0, n_tie = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t h = (uint32_t)(v[r] >> 32);
        n_less += (h < T) && (v[r] != kInvalid);
        n_tie += (h == T) && (v[r] != kInvalid);
    }
    int tot_less = n_less, tie_before = n_tie;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot_less += __shfl_xor_sync(0xffffffffu, tot_less, o);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // inclusive scan of tie counts
        const int up = __shfl_up_sync(0xffffffffu, tie_before, o);
        if (lane >= o) tie_before += up;
    }
    tie_before -= n_tie;  // exclusive
    const int tie_keep = (c >= kc) ? (kc - tot_less) : 0x7fffffff;
    // stable compaction
    int mine = 0;
    uint32_t keep_mask = 0;
    {
        int t_seen = tie_before;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t h = (uint32_t)(v[r] >> 32);
            const bool valid = v[r] != kInvalid;
            bool keep = valid && (h < T);
            if (valid && h == T) { keep = t_seen < tie_keep; t_seen++; }
            keep_mask |= (keep ? 1u : 0u) << r;
            mine += keep;
        }
    }
    int off = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += up;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    off -= mine;
    __syncwarp();  // every lane has loaded its slice before anyone overwrites the list in place
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (keep_mask & (1u << r)) __stcg(dst + off++, v[r]);
    }
    *kept_out = total;
    return (c >= kc) ? ordered_to_float(T) : INFINITY;
}

// CG = 1: one CTA per 128-query block (cta_group::1).  CG = 2: a cluster of two CTAs on one TPC works on
// two adjacent query blocks with cta_group::2 MMAs (M = 256): each CTA stages its own 128 queries and only
// HALF of every 256-row database tile, so the shared-memory fill per MMA drops from 48 KiB to 32 KiB and
// the operand reads from 12 KiB to 8 KiB per instruction -- the CG = 1 kernel is bound by exactly that
// shared-memory bandwidth.  The leader CTA (cluster rank 0) issues every MMA; accumulators
// land in each CTA's own TMEM, so the epilogue is identical.
// DUMP: the bootstrap-sample launch (raw keys written to keys_out, no selection) -- a template parameter so that
// the selecting kernels carry no per-chunk test for it.
template <int KIND, int METRIC, int CAP, int CG, bool DUMP = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
dense_scan_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
              const __grid_constant__ CUtensorMap map_qlo, const __grid_constant__ CUtensorMap map_dblo, const TcArgs a) {
    using TR = TcTraits<KIND>;
    // fp32 rows: "3xTF32".  kind::tf32 reads fp32 words and uses their top 19 bits; every operand is therefore
    // staged twice -- hi = x with the low 13 mantissa bits cleared, lo = x - hi (exact) -- and each k-step issues
    // hi*hi + hi*lo + lo*hi, which leaves an error of ~2^-21 relative: the same class as the fp16 path, so the
    // same candidate margin applies.
    constexpr bool X3 = (KIND == KIND_TF32);
    constexpr int PARTS = X3 ? 2 : 1;
    constexpr int B_ROWS = TC_N / CG;                 // database rows of a tile this CTA stages
    constexpr int B_BYTES = B_ROWS * 128;
    constexpr int STAGE_BYTES = PARTS * (TC_A_BYTES + B_BYTES);
    constexpr int STAGES = (TC_STAGES * TC_STAGE_BYTES) / STAGE_BYTES;   // 4 / 6 (pair); fp32: 2 / 3
    constexpr int B_OFF = PARTS * TC_A_BYTES;         // stage layout: A_hi [A_lo] B_hi [B_lo]
    constexpr int NBAR = STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t aux_off = STAGES * STAGE_BYTES;                      // [2 groups][2 buffers][TC_N / 2] floats
    const uint32_t bar_off = aux_off + 2u * 2u * (TC_N / 2) * 4u;
    const uint32_t bar_base = base + bar_off;
    // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2]; then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NBAR + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * NBAR + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * NBAR + 2 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (2 * NBAR + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // CTA (pair) index
    const int qb = (unit / a.groups) * CG + (int)rank, g = unit % a.groups;
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        // full: one arrive.expect_tx (the leader's producer); tmem_empty: 8 epilogue warps of every CTA of the pair
        for (int s = 0; s < NBAR; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8 * CG); }

        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // peers' barriers are initialised too
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int rt = a.tile_begin + g; rt < a.tile_end; rt += a.groups) {
                for (int kb = 0; kb < a.k_blocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    if constexpr (CG == 2) {
                        // both CTAs load; every byte is credited to the leader's full barrier
                        const uint32_t fb = map_to_cta(full_bar(stage), 0);
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                        tma_load_2d_pair(sa, &map_q, fb, kb * TR::kBlockK, qb * TC_M);
                        tma_load_2d_pair(sa + B_OFF, &map_db, fb, kb * TR::kBlockK, rt * TC_N + (int)rank * B_ROWS);
                        if constexpr (X3) {
                            tma_load_2d_pair(sa + TC_A_BYTES, &map_qlo, fb, kb * TR::kBlockK, qb * TC_M);
                            tma_load_2d_pair(sa + B_OFF + B_BYTES, &map_dblo, fb, kb * TR::kBlockK,
                                             rt * TC_N + (int)rank * B_ROWS);
                        }
                    } else {
                        const bool first = ((rt - a.tile_begin - g) / a.groups) * a.k_blocks + kb < STAGES;
                        const bool ld_a = !(a.debug & 4) || first, ld_b = !(a.debug & 2) || first;
                        mbar_arrive_expect_tx(full_bar(stage), PARTS * ((ld_a ? TC_A_BYTES : 0) + (ld_b ? B_BYTES : 0)));
                        if (ld_a) tma_load_2d(sa, &map_q, full_bar(stage), kb * TR::kBlockK, qb * TC_M);
                        if (ld_b) tma_load_2d(sa + B_OFF, &map_db, full_bar(stage), kb * TR::kBlockK, rt * TC_N);
                        if constexpr (X3) {
                            if (ld_a) tma_load_2d(sa + TC_A_BYTES, &map_qlo, full_bar(stage), kb * TR::kBlockK, qb * TC_M);
                            if (ld_b) tma_load_2d(sa + B_OFF + B_BYTES, &map_dblo, full_bar(stage), kb * TR::kBlockK, rt * TC_N);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0 && leader) {
            const uint32_t idesc = TR::kIdescFmt | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)((TC_M * CG) >> 4) << 24);
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int rt = a.tile_begin + g; rt < a.tile_end; rt += a.groups) {
                mbar_wait(tempty_bar(as), aphase ^ 1u);  // the epilogue(s) have drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * TC_N;
                for (int kb = 0; kb < a.k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + B_OFF);
                    auto mma = [&](uint64_t ad, uint64_t bd, uint32_t acc) {
                        if constexpr (CG == 2) tc_mma_pair<KIND>(d_tmem, ad, bd, idesc, acc);
                        else tc_mma<KIND>(d_tmem, ad, bd, idesc, acc);
                    };
#pragma unroll
                    for (int k = 0; k < 4; k++) {  // 4 x (UMMA_K * elem = 32 B) per 128-byte swizzle row
                        if constexpr (X3) {
                            const uint64_t dal = make_smem_desc(sa + TC_A_BYTES), dbl = make_smem_desc(sa + B_OFF + B_BYTES);
                            mma(da + 2u * k, dbl + 2u * k, (kb | k) != 0);   // hi * lo   (small terms first)
                            mma(dal + 2u * k, db + 2u * k, 1u);              // lo * hi
                            mma(da + 2u * k, db + 2u * k, 1u);               // hi * hi
                        } else {
                            mma(da + 2u * k, db + 2u * k, (kb | k) != 0);
                        }
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                    if constexpr (CG == 2) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                // accumulator complete -> epilogue(s)
                if constexpr (CG == 2) tc_commit_pair(tfull_bar(as)); else tc_commit(tfull_bar(as));
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue: fused top-k =======================
        // Two groups of four warps drain EVERY accumulator tile together: group e takes columns
        // [128e, 128e+128) of the tile (rows row0+128e ...) and keeps its own candidate lists.  A stage is
        // free again after half a tile's worth of epilogue work, so the drain hides behind the next
        // tile's MMAs even when the survivor path is busy (per-tile time = max(MMA, (MMA + drain)/2)).
        const int grp = (warp - 4) >> 2;
        const int ew = (warp - 4) & 3;   // TMEM lane quarter this warp may read (== warp % 4)
        const int tq = ew * 32 + lane;   // query (TMEM lane) of this thread
        const int q = qb * TC_M + tq;
        constexpr int cap = CAP, R = CAP / 32;
        constexpr int HALF = TC_N / 2;
        const int kc = a.kc;
        const int part = a.part_offset + g * 2 + grp;
        uint64_t* mybuf = a.cand + (((size_t)blockIdx.x * 2 + grp) * TC_M + tq) * cap;
        float* aux_s = reinterpret_cast<float*>(base_ptr + aux_off) + grp * 2 * HALF;  // two buffers per group
        int cnt = 0;
        float tau = -INFINITY;  // padding queries never select
        if (q < a.nq) tau = (a.tau_init != nullptr) ? __ldg(a.tau_init + q) : INFINITY;
        if (a.debug & 8) tau = -0.14f;  // probe: pretend a tight threshold is already known
        // Shared progressive threshold.  ed[] is this query's ladder of sample keys (ascending); gcnt[b]
        // counts the live rows every CTA of this query has met so far with a key in ladder bucket b
        // (seeded with the sample's own rows).  Whenever the counts up to some edge reach kc, at least
        // kc live rows are at or below that edge, so it is a valid threshold for everybody: the filter
        // tightens as 1/(fraction of the index scanned) instead of staying at the sample's kc-th key.
        const bool shared_tau = (a.edges != nullptr) && (q < a.nq) && !(a.debug & 16);
        float ed[LB_NEDGE];
        uint32_t* gcnt = nullptr;
        if (shared_tau) {
            gcnt = a.edge_cnt + (size_t)q * LB_NEDGE;
            const float4* ep = reinterpret_cast<const float4*>(a.edges + (size_t)q * LB_NEDGE);
#pragma unroll
            for (int i = 0; i < LB_NEDGE / 4; i++) {
                const float4 t = __ldg(ep + i);
                ed[4 * i] = t.x; ed[4 * i + 1] = t.y; ed[4 * i + 2] = t.z; ed[4 * i + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < LB_NEDGE; i++) ed[i] = INFINITY;
        }
        const uint32_t n_rows = a.n_rows, tomb_bits = a.tomb_bits;
        const uint32_t* __restrict__ tomb = a.tomb;
        const uint32_t* __restrict__ allow = a.allow;
        const bool filt = (tomb != nullptr) || (allow != nullptr);
        constexpr bool dump = DUMP;
        // timing probes (lb_set_option "tc_debug"), parked in registers the compiler cannot re-derive from the
        // parameter bank: re-reading a.debug per chunk cost the short-row epilogue 15 % of its issue slots
        uint32_t surv_mask = (a.debug & 32) ? 0u : 0xffffffffu;   // 32: filter only, survivors ignored
        uint32_t probe_bits = (uint32_t)a.debug & (1u | 128u);    // 1: no epilogue work, 128: round-1 release form
        asm volatile("mov.b32 %0, %0;" : "+r"(surv_mask));
        asm volatile("mov.b32 %0, %0;" : "+r"(probe_bits));
        int as = 0;
        uint32_t aphase = 0;
        int abuf = 0;
        // the leader issues the MMAs: both CTAs of a pair hand their accumulator stages back to ITS barriers
        const uint32_t tempty_remote[2] = {CG == 2 ? map_to_cta(tempty_bar(0), 0) : tempty_bar(0),
                                           CG == 2 ? map_to_cta(tempty_bar(1), 0) : tempty_bar(1)};
        int rt = a.tile_begin + g;
        uint32_t tile_no = 0;
        // Row auxiliaries (|x|^2 or 1/|x|) of a tile are staged in shared memory one tile ahead: the
        // global load for the NEXT tile is issued before this tile's accumulator is awaited, so its
        // L2 latency never sits between a TMEM load and the filter.
        float axn = 0.f;
        if constexpr (METRIC != METRIC_DOT) {
            if (rt < a.tile_end) axn = __ldg(a.aux + (size_t)rt * TC_N + grp * HALF + tq);
        }

        for (; rt < a.tile_end; rt += a.groups, abuf ^= 1) {
            const uint32_t row0 = (uint32_t)rt * TC_N + grp * HALF;  // first row of this group's half tile
            const float* axs = aux_s + abuf * HALF;
            if constexpr (METRIC != METRIC_DOT) {
                aux_s[abuf * HALF + tq] = axn;
                named_bar_sync(1 + grp, 128);  // the group's 4 warps; the other buffer may still be read
                if (rt + a.groups < a.tile_end)
                    axn = __ldg(a.aux + (size_t)(rt + a.groups) * TC_N + grp * HALF + tq);
            }
            // counters of the shared threshold: loaded (L2, never L1) before the wait, used after the tile.
            // Short rows make tiles cheap (one k-block = 4 MMAs): refresh only every 8th tile there.
            const bool do_refresh = shared_tau && (a.k_blocks >= 4 || (tile_no & 7) == 0);
            tile_no++;
            uint4 gc[LB_NEDGE / 4];
            if (do_refresh) {
#pragma unroll
                for (int i = 0; i < LB_NEDGE / 4; i++) gc[i] = __ldcg(reinterpret_cast<const uint4*>(gcnt) + i);
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * TC_N + grp * HALF;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();

            // one 32-column chunk of this thread's query: keys, filter, rare appends
            // one 32-column chunk of this thread's query.  Survivors are sparse (a fraction of a percent once the
            // threshold is tight), so the common path only answers "does any of my 32 keys pass?": the keys are
            // folded, negated, into four running maxima of eight columns each (one multiply / FMA and half a
            // 3-input max per key) and the maxima are tested against -tau.  Only column groups in which SOME lane
            // saw a survivor are looked at key by key.
            auto process = [&](uint32_t (&v)[32], const int c0) {
                // val = -key, bit for bit: cosine -(-dot * ax) = dot * ax, L2 -(ax - 2 dot) = fma(2, dot, -ax)
                auto val_of = [&](uint32_t raw, float axj) {
                    float dot;
                    if constexpr (KIND == KIND_I8) dot = (float)(int32_t)raw;
                    else dot = __uint_as_float(raw);
                    if constexpr (METRIC == METRIC_L2) return fmaf(2.f, dot, -axj);
                    else if constexpr (METRIC == METRIC_COSINE) return dot * axj;
                    else return dot;
                };
                if constexpr (dump) {
                    // bootstrap sample: write the keys of this chunk (row-major per query)
                    if (q < a.nq) {
                        float4* dst = reinterpret_cast<float4*>(a.keys_out + (size_t)q * a.keys_ld +
                                                                (row0 - (uint32_t)a.tile_begin * TC_N) + c0);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 ax4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if constexpr (METRIC != METRIC_DOT) ax4 = *reinterpret_cast<const float4*>(axs + c0 + j);
                            dst[j >> 2] = make_float4(-val_of(v[j], ax4.x), -val_of(v[j + 1], ax4.y),
                                                      -val_of(v[j + 2], ax4.z), -val_of(v[j + 3], ax4.w));
                        }
                    }
                    return;
                }
                constexpr bool INT_KEYS = (KIND == KIND_I8 && METRIC == METRIC_DOT);
                // integer dot, key = -dot: filter in the integer domain (key < tau <=> dot > floor(-tau)), no
                // int->float conversion per key (the conversion pipe is a quarter-rate unit)
                const int ti = __float2int_rd(-tau);  // saturates: tau = +inf admits every row, -inf none
                const float thr = -tau;
                uint32_t gm = 0;  // bit c: some column of [8c, 8c + 8) passes for this lane
                if constexpr (INT_KEYS) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        int m = __vimax3_s32((int32_t)v[8 * c], (int32_t)v[8 * c + 1], (int32_t)v[8 * c + 2]);
                        m = __vimax3_s32(m, (int32_t)v[8 * c + 3], (int32_t)v[8 * c + 4]);
                        m = __vimax3_s32(m, (int32_t)v[8 * c + 5], (int32_t)v[8 * c + 6]);
                        m = max(m, (int32_t)v[8 * c + 7]);
                        gm |= (m > ti ? 1u : 0u) << c;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                        if constexpr (METRIC != METRIC_DOT) {  // warp-uniform address: broadcast
                            a0 = *reinterpret_cast<const float4*>(axs + c0 + 8 * c);
                            a1 = *reinterpret_cast<const float4*>(axs + c0 + 8 * c + 4);
                        }
                        float m = fmax3(val_of(v[8 * c], a0.x), val_of(v[8 * c + 1], a0.y), val_of(v[8 * c + 2], a0.z));
                        m = fmax3(m, val_of(v[8 * c + 3], a0.w), val_of(v[8 * c + 4], a1.x));
                        m = fmax3(m, val_of(v[8 * c + 5], a1.y), val_of(v[8 * c + 6], a1.z));
                        m = fmaxf(m, val_of(v[8 * c + 7], a1.w));
                        gm |= (m > thr ? 1u : 0u) << c;   // NaN never passes (max ignores it, the compare is false)
                    }
                }
                const uint32_t groups = __reduce_or_sync(0xffffffffu, gm) & surv_mask;
                if (groups == 0) return;
                // per-key mask, only for the flagged groups (warp-uniform branches, static register indices)
                uint32_t hits = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (groups & (1u << c)) {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int j = 8 * c + u;
                            bool pass;
                            if constexpr (INT_KEYS) pass = (int32_t)v[j] > ti;
                            else {
                                float axj = 0.f;
                                if constexpr (METRIC != METRIC_DOT) axj = axs[c0 + j];
                                pass = val_of(v[j], axj) > thr;
                            }
                            if (pass) hits |= 1u << j;
                        }
                    }
                }
                // Every lane walks its OWN survivors, ascending columns (so every list stays in increasing row order);
                // the loop runs while some lane has one left, i.e. max-over-lanes survivors per chunk: one trip in
                // the steady state, two or three right after the bootstrap when the threshold is still loose.  The
                // key of a run-time column comes out of the registers through a five-level select tree (31 SEL).
                // (Rounds 1-2 walked the columns holding a survivor of ANY lane and re-read each from TMEM: one
                // serial tcgen05.ld round trip per column, ~60 trips per tile and warp during the first tile steps --
                // that, not the filter, was what stalled the MMA issuer: filter-only ran at the no-epilogue time.)
#pragma unroll 1
                while (__any_sync(0xffffffffu, hits != 0)) {
                    const bool mine = hits != 0;
                    const int j = mine ? (__ffs(hits) - 1) : 0;
                    hits &= hits - 1;  // 0 stays 0
                    const uint32_t raw = sel32(v, j);
                    if (mine) {
                        float axj = 0.f;
                        if constexpr (METRIC != METRIC_DOT) axj = axs[c0 + j];
                        const float key = -val_of(raw, axj);
                        const uint32_t row = row0 + c0 + j;
                        bool ok = row < n_rows;
                        if (filt && ok) {
                            if (tomb != nullptr && row < tomb_bits && bit_set(tomb, row)) ok = false;
                            if (ok && allow != nullptr && !bit_set(allow, row)) ok = false;
                        }
                        if (ok) {
                            mybuf[cnt++] = pack_key(key, row);
                            if (shared_tau) {
                                int b = 0;
#pragma unroll
                                for (int i = 0; i < LB_NEDGE; i++) b += (ed[i] <= key) ? 1 : 0;
                                if (b < LB_NEDGE) atomicAdd(gcnt + b, 1u);
                            }
                        }
                    }
                }
            };

            // hand the accumulator stage back to the MMA issuer: called once this warp's last TMEM read has landed
            // (the survivor path works from registers, so the rest of the drain needs no tensor memory)
            auto release = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) {
                        if (probe_bits & 128u)  // probe: the round-1 form (cluster-scope release)
                            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];"
                                         ::"r"(as ? tempty_remote[1] : tempty_remote[0]) : "memory");
                        else mbar_arrive_cluster(as ? tempty_remote[1] : tempty_remote[0]);
                    } else mbar_arrive(tempty_bar(as));
                }
            };
            if (!(probe_bits & 1u)) {
                // TMEM -> registers, double buffered: the load of chunk c+1 is in flight while chunk c is filtered
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
#pragma unroll 1
                for (int c0 = 0; c0 < HALF; c0 += 64) {
                    tmem_wait_ld32(va);
                    tmem_ld32(taddr + c0 + 32, vb);
                    process(va, c0);
                    tmem_wait_ld32(vb);
                    if (c0 + 64 < HALF) tmem_ld32(taddr + c0 + 64, va);
                    else release();
                    process(vb, c0 + 32);
                }
            } else {
                release();
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }

            // refresh the shared threshold from the counters fetched at the top of this tile
            if (do_refresh) {
                const uint32_t cv[LB_NEDGE] = {gc[0].x, gc[0].y, gc[0].z, gc[0].w, gc[1].x, gc[1].y, gc[1].z, gc[1].w,
                                               gc[2].x, gc[2].y, gc[2].z, gc[2].w, gc[3].x, gc[3].y, gc[3].z, gc[3].w};
                uint32_t cum = 0;
                float tg = INFINITY;
                bool found = false;
#pragma unroll
                for (int i = 0; i < LB_NEDGE; i++) {  // ascending: the lowest edge whose prefix holds kc rows
                    cum += cv[i];
                    const bool hit = cum >= (uint32_t)kc;
                    tg = (hit && !found) ? ed[i] : tg;
                    found = found || hit;
                }
                tau = fminf(tau, tg);
            }
            // warp-cooperative compaction of every list that could overflow during its next half tile
            unsigned need = __ballot_sync(0xffffffffu, cnt > cap - HALF);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int c = __shfl_sync(0xffffffffu, cnt, src);
                uint64_t* buf = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)mybuf, src));
                __syncwarp();  // order the owner's appends before the other lanes' reads
                int kept;
                const float nt = select_compact<R>(buf, buf, c, kc, lane, &kept);
                __syncwarp();
                if (lane == src) { cnt = kept; tau = nt; }
            }
        }
        // final: reduce every list to its best kc (unordered; the merge kernel sorts) and emit it
        for (int src = 0; src < 32 && !dump; src++) {
            const int c = __shfl_sync(0xffffffffu, cnt, src);
            const int qq = __shfl_sync(0xffffffffu, q, src);
            const uint64_t* buf = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)mybuf, src));
            if (qq >= a.nq) continue;  // warp-uniform
            __syncwarp();
            uint64_t* out = a.partial + ((size_t)part * a.nq + qq) * kc;
            int kept;
            select_compact<R>(buf, out, c, kc, lane, &kept);
            for (int t = kept + lane; t < kc; t += 32) out[t] = kInvalid;
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // no CTA leaves while its peer may still touch it
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Bootstrap threshold: the kc-th best key over the sample rows bounds the global kc-th best from
// above, so the main scan can start every list at it.  nextafter(+inf) turns the strict '<' of the
// filter into '<=': a row tying that key with a smaller row id must still get through.
// Bootstrap selection: per query, the kc smallest (key,row) of a dense [nq][S] sample key matrix
// (rows [0,S) of the index), honouring n_rows and the tombstone / allow bitmaps.  Block-wide
// MSB-first radix select over the packed 64-bit value; entries live in registers.
constexpr int SSEL_EMAX = 32;  // entries per thread: 8 / 16 / 32 by sample size
constexpr int SSEL_SUB = 1024;   // sub-sample whose order statistic supplies the pivot
constexpr int SSEL_COLL = 2048;  // capacity of the collected (<= pivot) set

__device__ __forceinline__ uint32_t sample_key(const float* __restrict__ keys, int ld, int S, uint32_t n_rows,
                                               const uint32_t* __restrict__ tomb, uint32_t tomb_bits,
                                               const uint32_t* __restrict__ allow, int q, uint32_t row) {
    if ((int)row >= S || row >= n_rows) return 0xffffffffu;
    if (tomb != nullptr && row < tomb_bits && bit_set(tomb, row)) return 0xffffffffu;
    if (allow != nullptr && !bit_set(allow, row)) return 0xffffffffu;
    const float k = __ldg(keys + (size_t)q * ld + row);
    return (k < INFINITY) ? float_to_ordered(k) : 0xffffffffu;  // NaN / +inf never selected
}

// The same for E rows per thread (rows e * nt + tid), for the warp-per-query kernel: ALL loads are issued first -- predicated, nothing depends on a
// loaded value yet, so they cost one memory round trip instead of E (sample_key() per row serialises: each row's
// validity branch waits for its own load) -- then the validity rules are applied.
template <int E>
__device__ __forceinline__ void sample_keys(const float* __restrict__ keys, int ld, int S, uint32_t n_rows,
                                            const uint32_t* __restrict__ tomb, uint32_t tomb_bits,
                                            const uint32_t* __restrict__ allow, int q, int tid, int nt, uint32_t (&v)[E]) {
    const float* krow = keys + (size_t)q * ld;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const uint32_t row = (uint32_t)(e * nt + tid);
        float kf = INFINITY;
        if ((int)row < S && row < n_rows) kf = __ldg(krow + row);
        v[e] = __float_as_uint(kf);
    }
#pragma unroll
    for (int e = 0; e < E; e++) {
        const float kf = __uint_as_float(v[e]);
        v[e] = (kf < INFINITY) ? float_to_ordered(kf) : 0xffffffffu;
    }
    if (tomb != nullptr || allow != nullptr) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint32_t row = (uint32_t)(e * nt + tid);
            if (v[e] != 0xffffffffu) {
                if (tomb != nullptr && row < tomb_bits && bit_set(tomb, row)) v[e] = 0xffffffffu;
                else if (allow != nullptr && !bit_set(allow, row)) v[e] = 0xffffffffu;
            }
        }
    }
}

// The ladder of one query from its sorted sample candidates (lanes 0..15 of one warp; see the comment in
// sample_select_kernel).  sorted[0, c): ascending packed (key, row), c >= kc.
__device__ __forceinline__ void emit_ladder(const uint64_t* sorted, int c, int kc, const EdgeRanks& ranks, int q,
                                            float* __restrict__ edges, uint32_t* __restrict__ edge_cnt, float* s_edge,
                                            int* s_below, int tid) {
    const float ktop = key_of(sorted[kc - 1]);
    float e;
    if (tid >= ranks.n_spare) {
        e = nextafterf(key_of(sorted[ranks.r[tid] - 1]), INFINITY);
    } else {
        const int rl = ranks.r[ranks.n_spare];
        int rm = kc / 4;
        if (rm <= rl) rm = rl + 1;
        const float kpd = (rm < kc) ? (ktop - key_of(sorted[rm - 1])) / log2f((float)kc / (float)rm) : 0.f;
        const float below = (float)(ranks.n_spare - tid) * ranks.doublings / (float)ranks.n_spare;
        e = ktop - kpd * (log2f((float)kc / (float)rl) + below);
        if (!(e == e)) e = ktop;  // inf - inf
    }
    s_edge[tid] = e;
    __syncwarp(0xffffu);
    int pos = 0;
    for (int u = 0; u < LB_NEDGE; u++) pos += (s_edge[u] < e || (s_edge[u] == e && u < tid)) ? 1 : 0;
    int lo = 0, hi = c;  // sample rows with key < e: all of them are among the c sorted ones (e <= the kc-th key, bumped)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key_of(sorted[mid]) < e) lo = mid + 1; else hi = mid;
    }
    s_below[pos] = lo;
    edges[(size_t)q * LB_NEDGE + pos] = e;
    __syncwarp(0xffffu);
    edge_cnt[(size_t)q * LB_NEDGE + tid] = (uint32_t)(s_below[tid] - (tid ? s_below[tid - 1] : 0));
}

// Fast path.  Thread t holds the ordered keys of rows t, t+nt, t+2nt ... (row ids are implicit, so 32
// registers hold 32 entries).  Pivot = the r-th smallest of a sub-sample (the threads' first entries),
// r chosen so that about 4*kc of the S keys fall at or below it; everything <= pivot is collected into
// shared memory and sorted.  If at least kc were collected, their kc smallest (key,row) are exactly the
// kc smallest of the whole sample.  Otherwise done[q] stays 0 and the exact kernel below handles q.
template <int SSEL_E>
__global__ void __launch_bounds__(1024)
sample_select_kernel(const float* __restrict__ keys, int ld, int S, uint32_t n_rows, const uint32_t* __restrict__ tomb,
                     uint32_t tomb_bits, const uint32_t* __restrict__ allow, int nq, int kc,
                     uint64_t* __restrict__ out, float* __restrict__ tau, float* __restrict__ edges,
                     uint32_t* __restrict__ edge_cnt, const EdgeRanks ranks, int* __restrict__ done) {
    __shared__ int s_out;
    __shared__ uint32_t s_sub[SSEL_SUB];
    __shared__ uint64_t s_coll[SSEL_COLL];
    const int q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    uint32_t v[SSEL_E];
#pragma unroll
    for (int e = 0; e < SSEL_E; e++)
        v[e] = sample_key(keys, ld, S, n_rows, tomb, tomb_bits, allow, q, (uint32_t)(e * nt + tid));
    if (tid == 0) { s_out = 0; done[q] = 0; }
    const int nsub = next_pow2(nt);  // nt <= 1024 == SSEL_SUB
    for (int t = tid; t < nsub; t += nt) s_sub[t] = 0xffffffffu;
    __syncthreads();
    s_sub[tid] = v[0];
    __syncthreads();
    block_bitonic_sort_t<uint32_t>(s_sub, nsub);
    int r = (int)(((int64_t)4 * kc * nt + S - 1) / S);
    if (r < 12) r = 12;
    if (r > nt) return;
    const uint32_t pivot = s_sub[r - 1];
    if (pivot == 0xffffffffu) return;
#pragma unroll
    for (int e = 0; e < SSEL_E; e++) {
        if (v[e] <= pivot) {
            const int pos = atomicAdd(&s_out, 1);
            if (pos < SSEL_COLL) s_coll[pos] = ((uint64_t)v[e] << 32) | (uint32_t)(e * nt + tid);
        }
    }
    __syncthreads();
    const int c = s_out;
    if (c < kc || c > SSEL_COLL) return;
    const int n2 = next_pow2(c);
    for (int t = c + tid; t < n2; t += nt) s_coll[t] = kInvalid;
    __syncthreads();
    block_bitonic_sort(s_coll, n2);
    for (int t = tid; t < kc; t += nt) out[(size_t)q * kc + t] = s_coll[t];
    // Threshold ladder: the sample keys at the ranks in `ranks`, bumped one ulp so that the scan's strict '<' admits
    // ties, plus -- in the slots the rank ladder left over -- edges extrapolated below the sample: the model
    // key(r) = key(kc) - kpd * log2(kc / r) continued to fractional ranks, with kpd (key units per halving of the
    // tail probability) read between ranks kc and ~kc/4, NOT off the lowest keys: a query that is itself a row of
    // the index (self-match, common in practice) puts an outlier at rank 1.  The 16 values are then sorted and every
    // bucket is seeded with the sample rows that actually fall into it, so any mix of values is a valid ladder.
    __shared__ float s_edge[LB_NEDGE];
    __shared__ int s_below[LB_NEDGE];
    if (tid < LB_NEDGE) emit_ladder(s_coll, c, kc, ranks, q, edges, edge_cnt, s_edge, s_below, tid);   // one warp
    if (tid == 0) { tau[q] = nextafterf(key_of(s_coll[kc - 1]), INFINITY); done[q] = 1; }
}

// Small samples (S <= 2048, kc <= 128): ONE WARP per query.  Lane l holds the ordered keys of rows 32 e + l in 64
// registers; the kc-th smallest key is found by an MSB-first radix search with warp votes (no sorting network, no
// block barrier), the kc smallest (key, row) are compacted in row order -- ties at the kc-th key broken by row --
// sorted (128 entries) and the ladder is read off them.  ~6 k warp instructions per query against ~54 k for the block
// kernel above: on a 125 k-row shard (8 GPUs) or C1 the sample selection was a quarter of the per-batch tail.
// Queries with fewer than kc live sample rows are left to the exact kernel (done[q] = 0).
constexpr int SSW_E = 64;
__global__ void __launch_bounds__(128)
sample_select_warp_kernel(const float* __restrict__ keys, int ld, int S, uint32_t n_rows, const uint32_t* __restrict__ tomb,
                          uint32_t tomb_bits, const uint32_t* __restrict__ allow, int nq, int kc,
                          uint64_t* __restrict__ out, float* __restrict__ tau, float* __restrict__ edges,
                          uint32_t* __restrict__ edge_cnt, const EdgeRanks ranks, int* __restrict__ done) {
    __shared__ uint64_t s_top[4][128];
    __shared__ float s_edge[4][LB_NEDGE];
    __shared__ int s_below[4][LB_NEDGE];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + w;
    if (q >= nq) return;
    uint32_t v[SSW_E];
    int live = 0;
    uint32_t a_and = 0xffffffffu, a_or = 0;
    sample_keys<SSW_E>(keys, ld, S, n_rows, tomb, tomb_bits, allow, q, lane, 32, v);
#pragma unroll
    for (int e = 0; e < SSW_E; e++) {
        if (v[e] != 0xffffffffu) { live++; a_and &= v[e]; a_or |= v[e]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        live += __shfl_xor_sync(0xffffffffu, live, o);
        a_and &= __shfl_xor_sync(0xffffffffu, a_and, o);
        a_or |= __shfl_xor_sync(0xffffffffu, a_or, o);
    }
    if (live < kc) {
        if (lane == 0) done[q] = 0;
        return;
    }
    // kc-th smallest key: the largest T with fewer than kc keys below it; bits above the highest differing bit are shared
    const uint32_t diff = a_and ^ a_or;
    const int top = diff ? (31 - __clz(diff)) : -1;
    uint32_t T = (top >= 31) ? 0u : (a_and & ~((2u << top) - 1u));
    if (top < 0) T = a_and;
#pragma unroll 1
    for (int bit = top; bit >= 0; bit--) {
        const uint32_t test = T | (1u << bit);
        int l0 = 0, l1 = 0, l2 = 0, l3 = 0;   // four chains: the per-bit step is a latency chain, not a throughput one
#pragma unroll
        for (int e = 0; e < SSW_E; e += 4) {
            l0 += (v[e] < test); l1 += (v[e + 1] < test); l2 += (v[e + 2] < test); l3 += (v[e + 3] < test);
        }
        const int less = (int)__reduce_add_sync(0xffffffffu, (unsigned)((l0 + l1) + (l2 + l3)));
        if (less < kc) T = test;
    }
    int n_less = 0;
#pragma unroll
    for (int e = 0; e < SSW_E; e++) n_less += (v[e] < T);
    n_less = (int)__reduce_add_sync(0xffffffffu, (unsigned)n_less);
    const int tie_keep = kc - n_less;   // >= 1
    // compaction in row order (e ascending, then lane): rows below T, plus the first tie_keep rows at T
    uint64_t* topl = s_top[w];
    int base = 0, ties = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int e = 0; e < SSW_E; e++) {   // (fully unrolled: v[] must stay in registers)
        const bool is_tie = v[e] == T;
        const unsigned tb = __ballot_sync(0xffffffffu, is_tie);
        const bool keep = (v[e] < T) || (is_tie && ties + __popc(tb & lt) < tie_keep);
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (keep) topl[base + __popc(kb & lt)] = ((uint64_t)v[e] << 32) | (uint32_t)(e * 32 + lane);
        base += __popc(kb);
        ties += __popc(tb);
    }
    for (int t = kc + lane; t < 128; t += 32) topl[t] = kInvalid;   // base == kc
    __syncwarp();
    warp_bitonic_sort(topl, 128, lane);
    for (int t = lane; t < kc; t += 32) out[(size_t)q * kc + t] = topl[t];
    if (lane < LB_NEDGE) emit_ladder(topl, kc, kc, ranks, q, edges, edge_cnt, s_edge[w], s_below[w], lane);
    if (lane == 0) { tau[q] = nextafterf(key_of(topl[kc - 1]), INFINITY); done[q] = 1; }
}

// Exact path for the queries the fast path left (fewer than kc sample rows at or below the pivot, too
// many, or fewer than kc valid rows at all): block-wide MSB-first search over the packed 64-bit value.
template <int SSEL_E>
__global__ void __launch_bounds__(1024)
sample_select_exact_kernel(const float* __restrict__ keys, int ld, int S, uint32_t n_rows,
                           const uint32_t* __restrict__ tomb, uint32_t tomb_bits, const uint32_t* __restrict__ allow,
                           int nq, int kc, uint64_t* __restrict__ out, float* __restrict__ tau,
                           float* __restrict__ edges, uint32_t* __restrict__ edge_cnt, const int* __restrict__ done) {
    __shared__ int s_red[64];
    __shared__ int s_out;
    const int q = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (done[q]) return;
    uint64_t v[SSEL_E];
#pragma unroll
    for (int e = 0; e < SSEL_E; e++) {
        const uint32_t row = (uint32_t)(e * nt + tid);
        const uint32_t k = sample_key(keys, ld, S, n_rows, tomb, tomb_bits, allow, q, row);
        v[e] = (k == 0xffffffffu) ? kInvalid : (((uint64_t)k << 32) | row);
    }
    if (tid == 0) s_out = 0;
    const uint64_t T = block_kth_smallest<SSEL_E>(v, kc, s_red, tid, nt / 32);
#pragma unroll
    for (int e = 0; e < SSEL_E; e++) {
        if (v[e] != kInvalid && v[e] <= T) {
            const int pos = atomicAdd(&s_out, 1);
            if (pos < kc) out[(size_t)q * kc + pos] = v[e];
        }
    }
    __syncthreads();
    for (int t = s_out + tid; t < kc; t += nt) out[(size_t)q * kc + t] = kInvalid;
    // degenerate ladder: every edge = the sample's kc-th key (or +inf when the sample holds fewer than kc
    // live rows); bucket 0 starts at the number of sample rows at or below it
    const float t0 = (T == kInvalid) ? INFINITY : nextafterf(key_of(T), INFINITY);
    const int n_sel = s_out < kc ? s_out : kc;
    if (tid < LB_NEDGE) {
        edges[(size_t)q * LB_NEDGE + tid] = t0;
        edge_cnt[(size_t)q * LB_NEDGE + tid] = (tid == 0) ? (uint32_t)n_sel : 0u;
    }
    if (tid == 0) tau[q] = t0;
}

cudaError_t launch_sample_select(const float* keys, int ld, int S, uint32_t n_rows, const uint32_t* tomb,
                                 uint32_t tomb_bits, const uint32_t* allow, int nq, int kc, uint64_t* out,
                                 float* tau, float* edges, uint32_t* edge_cnt, int* done, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    if (S > SSEL_EMAX * 1024) return cudaErrorInvalidValue;
    const int E = (S <= 8 * 256) ? 8 : (S <= 16 * 512) ? 16 : 32;
    int nt = ((S + E - 1) / E + 31) / 32 * 32;
    if (nt < 64) nt = 64;
    // ladder ranks: kc, kc/sqrt2, kc/2 ... 1 (at most LB_NEDGE of them, the largest kept), ascending;
    // unused low slots repeat the smallest rank (their buckets stay empty)
    EdgeRanks er;
    {
        int lad[64], n = 0, r = kc;
        for (;;) {
            lad[n++] = r;
            if (r == 1 || n == LB_NEDGE) break;
            int nr = (int)(r * 0.70710678f);
            r = nr < 1 ? 1 : nr;
        }
        for (int i = 0; i < LB_NEDGE; i++) {
            const int j = LB_NEDGE - 1 - i;  // position from the top
            er.r[i] = lad[j < n ? j : n - 1];
        }
        // Slots left over once the ladder has reached rank 1 continue it below the sample: the scan's kc-th best of
        // n_rows rows sits log2(n_rows / (S * kc)) tail halvings under the sample's best key (0.9: tails steepen);
        // without them the threshold stops at the sample's minimum and ~n_rows / S rows per query get through
        // (12.5 M x 128 int8, kc = 26: 1785 appended rows per query instead of ~250).
        er.n_spare = (lad[n - 1] == 1) ? LB_NEDGE - n : 0;
        const float need = 0.8f * log2f(fmaxf(1.f, (float)n_rows / ((float)S * (float)kc)));
        er.doublings = fmaxf(need, 0.5f * (float)er.n_spare);
    }
#define LB_SSEL(E_)                                                                                             \
    {                                                                                                           \
        sample_select_kernel<E_><<<nq, nt, 0, st>>>(keys, ld, S, n_rows, tomb, tomb_bits, allow, nq, kc, out, tau, \
                                                    edges, edge_cnt, er, done);                                 \
        count_launch();                                                                                         \
        sample_select_exact_kernel<E_><<<nq, nt, 0, st>>>(keys, ld, S, n_rows, tomb, tomb_bits, allow, nq, kc, out, \
                                                          tau, edges, edge_cnt, done);                          \
        count_launch();                                                                                         \
    }
    if (S <= SSW_E * 32 && kc <= 128 && g_ssel_warp) {
        sample_select_warp_kernel<<<(nq + 3) / 4, 128, 0, st>>>(keys, ld, S, n_rows, tomb, tomb_bits, allow, nq, kc, out, tau,
                                                                 edges, edge_cnt, er, done);
        count_launch();
        sample_select_exact_kernel<8><<<nq, nt, 0, st>>>(keys, ld, S, n_rows, tomb, tomb_bits, allow, nq, kc, out, tau, edges,
                                                         edge_cnt, done);
        count_launch();
    } else if (E == 8) LB_SSEL(8) else if (E == 16) LB_SSEL(16) else LB_SSEL(32)
#undef LB_SSEL
    return cudaGetLastError();
}

// fp32 operand split for the 3xTF32 scan: hi = x with the 13 low mantissa bits cleared (what kind::tf32 reads),
// lo = x - hi (exact in fp32).  In place for hi is not wanted (rows stay the reference's bytes): only lo is stored,
// the tensor core sees hi by ignoring the low bits of the original words.
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float v = x[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        lo[i] = v - hi;
    }
}

cudaError_t launch_split_lo(const float* x, float* lo, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, lo, n);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static bool make_map(CUtensorMap* m, int kind, const void* ptr, uint64_t rows, int dim, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const int elem = kind == KIND_F16 ? 2 : kind == KIND_I8 ? 1 : 4;
    const int block_k = 128 / elem;
    CUtensorMapDataType dt = kind == KIND_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                             : kind == KIND_I8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dim * elem};
    cuuint32_t box[2] = {(cuuint32_t)block_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool dense_tc_eligible(int dtype, int dim, const void* db, const void* queries, int kc) {
    const int elem = dtype == DT_F16 ? 2 : dtype == DT_I8 ? 1 : dtype == DT_F32 ? 4 : 0;
    if (elem == 0) return false;
    if (((size_t)dim * elem) % 16 != 0) return false;                    // TMA row pitch
    if ((reinterpret_cast<uintptr_t>(db) | reinterpret_cast<uintptr_t>(queries)) & 15) return false;
    if (kc + TC_N > 1024) return false;
    return get_encode() != nullptr;
}

// groups (= number of partial lists per query) and the size of the candidate scratch
void dense_scan_tc_plan(int nq, int n_row_tiles, int sm_count, int kc, int* groups_out, size_t* cand_bytes) {
    int cap = next_pow2(kc + TC_N);
    if (cap < 512) cap = 512;
    const int nqb = (nq + TC_M - 1) / TC_M;
    // lb_set_option("tc_reserve_sms"): SMs the persistent scan leaves free.  One scan CTA takes an SM's whole register
    // file, so kernels of OTHER streams (the tail of the previous batch: selection, merge, re-score, exchange) can
    // only run on SMs the scan does not occupy.
    int usable = sm_count - g_tc_reserve_sms;
    if (usable < nqb) usable = nqb;
    int groups = usable / nqb;
    if (groups < 1) groups = 1;
    if (groups > n_row_tiles) groups = n_row_tiles;
    *groups_out = groups;
    *cand_bytes = (size_t)nqb * groups * 2 * TC_M * cap * 8;  // two epilogue groups per CTA
}

cudaError_t launch_dense_scan_tc(const ScanArgs& s, int sm_count, uint64_t* cand, cudaStream_t st) {
    const int kind = s.dtype == DT_F16 ? KIND_F16 : s.dtype == DT_I8 ? KIND_I8 : KIND_TF32;
    const int elem = kind == KIND_F16 ? 2 : kind == KIND_I8 ? 1 : 4;
    const int block_k = 128 / elem;
    const int nqb_ = (s.nq + TC_M - 1) / TC_M;
    // CTA pairs (cta_group::2) whenever the query blocks pair up; lb_set_option("tc_pair", 0) disables
    const int kb_ = (s.dim + block_k - 1) / block_k;
    // (short rows are epilogue-bound: a pair only adds cross-CTA hand-shakes there -- measured slower at 128-byte rows)
    const int cg = (g_tc_pair && (nqb_ % 2 == 0) && kb_ >= 4 && !(s.debug & 6)) ? 2 : 1;
    CUtensorMap mq, mdb, mqlo, mdblo;
    if (!make_map(&mq, kind, s.queries, (uint64_t)s.nq, s.dim, TC_M)) return cudaErrorInvalidValue;
    if (!make_map(&mdb, kind, s.db, (uint64_t)s.n_rows, s.dim, TC_N / cg)) return cudaErrorInvalidValue;
    mqlo = mq; mdblo = mdb;
    if (kind == KIND_TF32) {  // 3xTF32: the low parts of both operands
        if (s.db_lo == nullptr || s.queries_lo == nullptr) return cudaErrorInvalidValue;
        if (!make_map(&mqlo, kind, s.queries_lo, (uint64_t)s.nq, s.dim, TC_M)) return cudaErrorInvalidValue;
        if (!make_map(&mdblo, kind, s.db_lo, (uint64_t)s.n_rows, s.dim, TC_N / cg)) return cudaErrorInvalidValue;
    }
    TcArgs a;
    a.aux = s.aux; a.n_rows = s.n_rows; a.nq = s.nq;
    a.k_blocks = (s.dim + block_k - 1) / block_k;
    a.n_row_tiles = (int)((s.n_rows + TC_N - 1) / TC_N);
    const int nqb = (s.nq + TC_M - 1) / TC_M;
    int groups;
    size_t cand_bytes;
    a.tile_begin = s.tile_begin;
    a.tile_end = s.tile_end < 0 ? a.n_row_tiles : s.tile_end;
    a.part_offset = s.part_offset;
    a.tau_init = s.tau_init;
    a.keys_out = s.keys_out; a.keys_ld = s.keys_ld;
    a.edges = s.edges; a.edge_cnt = s.edge_cnt;
    dense_scan_tc_plan(s.nq, a.tile_end - a.tile_begin, sm_count, s.kc, &groups, &cand_bytes);
    if (s.keys_out == nullptr && 2 * groups != s.parts) return cudaErrorInvalidValue;  // partial[] sized for s.parts lists
    a.groups = groups;
    a.tomb = s.tomb; a.tomb_bits = s.tomb_bits; a.allow = s.allow;
    a.kc = s.kc; a.cap = next_pow2(s.kc + TC_N);
    if (a.cap < 512) a.cap = 512;
    a.cand = cand; a.partial = s.partial; a.debug = s.debug;
    // the ring always occupies TC_STAGES * TC_STAGE_BYTES (192 KiB) whatever the stage size; its barrier block is
    // sized for the deepest ring (6 stages)
    const size_t smem = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + 2 * 2 * (TC_N / 2) * 4 + 8 * (2 * TC_STAGES_PAIR + 4) + 16;
    const dim3 grid(nqb * groups);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define LB_TC2(KIND_, METRIC_, CAP_, CG_)                                                                      \
    if (s.keys_out != nullptr) LB_TC3(KIND_, METRIC_, 512, CG_, true) else LB_TC3(KIND_, METRIC_, CAP_, CG_, false)
#define LB_TC3(KIND_, METRIC_, CAP_, CG_, DUMP_)                                                               \
    {                                                                                                          \
        auto kern = dense_scan_tc<KIND_, METRIC_, CAP_, CG_, DUMP_>;                                           \
        LB_SMEM_OPTIN(kern);                                                                                   \
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mq, mdb, mqlo, mdblo, a);                               \
        if (e != cudaSuccess) return e;                                                                        \
    }
#define LB_TC1(KIND_, METRIC_, CAP_)                                                                           \
    {                                                                                                          \
        if (cg == 2) LB_TC2(KIND_, METRIC_, CAP_, 2) else LB_TC2(KIND_, METRIC_, CAP_, 1)                      \
    }
#define LB_TC(KIND_, METRIC_)                                                                                  \
    {                                                                                                          \
        if (a.cap == 512) LB_TC1(KIND_, METRIC_, 512) else LB_TC1(KIND_, METRIC_, 1024)                        \
    }
    if (kind == KIND_F16) {
        if (s.metric == METRIC_L2) LB_TC(KIND_F16, METRIC_L2) else if (s.metric == METRIC_COSINE) LB_TC(KIND_F16, METRIC_COSINE) else LB_TC(KIND_F16, METRIC_DOT)
    } else if (kind == KIND_I8) {
        if (s.metric == METRIC_L2) LB_TC(KIND_I8, METRIC_L2) else LB_TC(KIND_I8, METRIC_DOT)
    } else {
        if (s.metric == METRIC_L2) LB_TC(KIND_TF32, METRIC_L2) else if (s.metric == METRIC_COSINE) LB_TC(KIND_TF32, METRIC_COSINE) else LB_TC(KIND_TF32, METRIC_DOT)
    }
#undef LB_TC
#undef LB_TC1
#undef LB_TC2
#undef LB_TC3
    count_launch();
    return cudaGetLastError();
}

}  // namespace lb
