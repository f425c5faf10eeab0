// mbc_scan_kernels.cuh -- K2/K5: fused CNF filter -> ordered compaction -> projection -> aggregates.
//
// One launch replaces the reference's per-row loop
//     iterator/ColumnarFileScan.java:156-172  while (scan.getNext) if (PredEval.Eval) Project
// together with iterator/PredEval.java:25-183 (CNF evaluation, comparison type = type of the
// lhs), iterator/TupleUtils.java:35-87 (int / float / string compare) and
// iterator/Projection.java:103-144 (copy of the selected fields), for a whole table.
//
// Shape of the kernel (sm_100a; HBM-bound integer/byte work, no tensor cores):
//   * persistent CTAs with a ring of tile slots: the 4-byte predicate columns of the next tiles are
//     brought into shared memory by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) while the
//     current tile is evaluated, so HBM latency is covered by the ring depth, not by occupancy;
//   * a warp owns 512 consecutive rows; in "unit" u lane l owns rows u*128+l*4+{0..3}, so every
//     4-byte column is read with one 128-bit access per unit (512 B per warp instruction);
//     16-byte string rows are read lane-contiguously (row = k*32+lane, 512 B per instruction) and
//     the result bits are transposed into the 4-rows-per-lane layout with warp ballots;
//   * the CNF is a small term program in the kernel parameters; the operator/type switch runs once
//     per term per 16 rows, the compares themselves are straight-line;
//   * the filter pass emits the selection bitmap (warp-shuffle assembled words) and one count per tile;
//     independent blocks turn the counts into output offsets (each sums what precedes it itself); the
//     write pass ranks the survivors of a tile with one packed warp scan (four 8-bit unit counters in
//     one register) and a CTA scan, so the output is written in ascending position order (the reference
//     emits rows in position order; bit-exact position lists need order, not atomics) and no CTA ever
//     waits for another;
//   * the write pass picks its method per GROUP of 8 tiles (32768 rows) from the group's survivor count, which the
//     offsets kernel leaves as one class byte per group:
//       - up to 12.5 %: one CTA writes the whole group (rank -> row list in shared memory, one work item -- positions, a
//         projected column, the aggregates of a column -- per warp), so sparse scans do not pay a CTA's
//         load -> scan -> gather -> store latency chain per tile and read only the survivors' sectors;
//       - up to 28 %: every CTA gathers its own tile, one thread per SURVIVOR, coalesced stores;
//       - above: write_staged_kernel reads the columns whole through a TMA ring and compacts out of shared memory with
//         warp-autonomous ranks (one bitmap word per 32-row instruction): each column is read once, nothing waits on a
//         list, and the kernel runs at the HBM roofline on the bytes it moves (measured 6.3 TB/s at 50 %);
//     in the gather methods an aggregate whose column is projected is folded from the value loaded for the projection;
//   * COUNT/SUM/MIN/MAX partials (per tile, per group or per CTA of write_staged_kernel, always in a slot of the
//     per-tile array) are reduced in slot order by a second tiny kernel, so real-valued sums are reproducible run to run.
#pragma once
#include <cstring>
#include <algorithm>

#include "mbc_internal.cuh"

namespace mbc {

// ---- kernel parameter block ---------------------------------------------------------------

struct DevOperand {
    const void* ptr;     // column base (kind 1) or unused
    int32_t kind;        // 0 literal, 1 column
    int32_t stride;      // device row stride of the column
    uint32_t bits;       // raw 32-bit literal (int or float bits)
    int32_t col;         // host bookkeeping: table column the pointer was taken from
    int32_t staged;      // index of the TMA-staged copy of this column in the tile slot, or -1
    int32_t pad;
};

struct DevTerm {
    int32_t op;          // MBC_OP_*
    int32_t cmp_type;    // MBC_ATTR_INTEGER / REAL / STRING
    int32_t end_conj;    // 1 = last term of its conjunct
    int32_t lit_words;   // string literal length in 32-bit words (zero padded)
    DevOperand lhs, rhs;
    uint32_t lit[kMaxLit / 4];   // zero padded string literal (whichever side is the literal)
};

struct DevProj {
    const void* src;
    void* dst;
    int32_t stride;      // 4, or the device string stride (multiple of 4)
    int32_t col;         // host bookkeeping
};

struct DevAgg {
    const void* src;
    int32_t kind;        // MBC_AGG_*
    int32_t type;        // MBC_ATTR_INTEGER / REAL
    int32_t col;         // host bookkeeping
    int32_t pad;
};

constexpr int kMaxStg = kMaxProj + kMaxAgg;      // columns write_staged_kernel can stage (projected fields + aggregate sources)
constexpr int kMaxStaged = 4;                    // 4-byte predicate columns staged through shared memory
constexpr int kStageColBytes = kTileRows * 4;    // one column of one tile: 16 KB

struct ScanParams {
    int64_t nrows;
    int64_t pos_base;
    int64_t out_cap;              // rows the output buffers hold (debug checks)
    // dense tiles through write_staged_kernel: every column the write pass reads is bulk-copied (TMA) 1024 rows at a time
    uint8_t proj_aggs[kMaxProj];  // aggregates folded from the values gathered for projected field c (same column, 4-byte): bit a
    uint8_t agg_group[kMaxAgg];   // other aggregates: bit mask of those sharing a's source column, on the group's first member (else 0)
    int32_t stg_min;              // groups of kGroupTiles tiles with MORE survivors than this go to write_staged_kernel (>= kSparseMax)
    int32_t stg_n;                // staged columns; 0 = the path is off and write_kernel writes the dense groups itself
    int32_t stg_stages;           // depth of the ring
    int32_t stg_bytes;            // bytes of one stage = kStgRows * sum of the staged strides
    int32_t dense_staged;         // 1 in the launch of write_kernel that leaves the dense groups to write_staged_kernel
    const void* stg_src[kMaxStg];
    int32_t stg_stride[kMaxStg];
    int32_t stg_off[kMaxStg];     // byte offset of a column's 1024 rows inside a stage
    int32_t stg_col[kMaxStg];     // host bookkeeping: table column
    int8_t proj_stg[kMaxProj];    // staged column of every projected field
    int8_t agg_stg[kMaxAgg];      // ... and of every aggregate's source (-1: COUNT)
    int32_t ntiles;
    int32_t nterms;
    int32_t nproj;
    int32_t nagg;
    int32_t tile_base;            // index of this launch's first tile in the partials arrays
    int32_t total_tiles;          // stride of the partials arrays
    int32_t nstaged;              // staged predicate columns (0..kMaxStaged)
    int32_t nstages;              // depth of the tile ring (2..kMaxStages)
    const void* staged_src[kMaxStaged];
    int32_t staged_cols[kMaxStaged];   // host bookkeeping: table column of every staged slot
    const uint32_t* sel_bitmap;   // optional precomputed selection
    const uint32_t* deleted;      // optional markedDeleted bitmap
    int64_t* out_pos;
    uint32_t* out_bitmap;
    uint32_t* tile_counts;        // qualifying rows per tile (pass 1 -> pass 2)
    unsigned long long* tile_out; // global output offset of every tile (tile_offsets_kernel)
    uint8_t* group_class;         // per group of kGroupTiles tiles (tile_offsets_kernel): kGroupEmpty / Sparse / Mid / Full
    const long long* count_in;    // running output offset before this launch (chunked scans append); NULL = 0
    long long* count_out;         // ... and after it (a different slot)
    unsigned int* work_counter;   // group tickets of the write pass (zeroed by tile_offsets_kernel)
    unsigned long long* partials; // [nagg][total_tiles]
    DevTerm terms[kMaxTerms];
    DevProj proj[kMaxProj];
    DevAgg aggs[kMaxAgg];
};

// ---- device helpers -------------------------------------------------------------------------

__device__ __forceinline__ uint4 ldg128(const void* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// result bit of one comparison given "less" and "equal"
__device__ __forceinline__ bool apply_op(int op, bool lt, bool eq) {
    switch (op) {
        case MBC_OP_EQ: return eq;
        case MBC_OP_LT: return lt;
        case MBC_OP_GT: return !(lt || eq);
        case MBC_OP_NE: return !eq;
        case MBC_OP_LE: return lt || eq;
        case MBC_OP_GE: return !lt;
        case MBC_OP_NOT: return !eq;      // PredEval.java:158-160: aopNOT behaves as NE
        default: return false;            // aopNOP / opRANGE never match
    }
}

// 16 result bits of `a op b` from per-row lt / eq bit masks
__device__ __forceinline__ uint32_t mask_op(int op, uint32_t lt, uint32_t eq) {
    switch (op) {
        case MBC_OP_EQ: return eq;
        case MBC_OP_LT: return lt;
        case MBC_OP_GT: return ~(lt | eq) & 0xFFFFu;
        case MBC_OP_NE: return ~eq & 0xFFFFu;
        case MBC_OP_LE: return lt | eq;
        case MBC_OP_GE: return ~lt & 0xFFFFu;
        case MBC_OP_NOT: return ~eq & 0xFFFFu;
        default: return 0u;
    }
}

// load the raw 32-bit values this thread owns (U units x one 128-bit load; U = 4 -> 16 rows), or broadcast a literal
template <int U = kUnits>
__device__ __forceinline__ void load_operand32(const DevOperand& o, int64_t thread_row0, const uint32_t* stage, int tile_off,
                                               uint32_t (&v)[U * kVec], const int stage_rows = kTileRows) {
    if (o.kind == 0) {
#pragma unroll
        for (int i = 0; i < U * kVec; ++i) v[i] = o.bits;
    } else if (o.staged >= 0) {
        // the tile's slice of this column was brought into shared memory by the TMA engine; lanes read
        // consecutive 16-byte chunks (conflict-free LDS.128)
        const uint32_t* base = stage + o.staged * stage_rows + tile_off;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint4 q = *reinterpret_cast<const uint4*>(base + u * kUnitRows);
            v[u * 4 + 0] = q.x;
            v[u * 4 + 1] = q.y;
            v[u * 4 + 2] = q.z;
            v[u * 4 + 3] = q.w;
        }
    } else {
        const uint32_t* base = reinterpret_cast<const uint32_t*>(o.ptr) + thread_row0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint4 q = ldg128(base + u * kUnitRows);
            v[u * 4 + 0] = q.x;
            v[u * 4 + 1] = q.y;
            v[u * 4 + 2] = q.z;
            v[u * 4 + 3] = q.w;
        }
    }
}

// U * 4 rows of one relation: bit i = rel(a[i], b[i])
#define MBC_CMP16(T, CONV, REL)                                                           \
    {                                                                                      \
        uint32_t m = 0;                                                                    \
        _Pragma("unroll") for (int i = 0; i < U * kVec; ++i) {                             \
            T x = CONV(a[i]), y = CONV(b[i]);                                              \
            m |= (uint32_t)(x REL y) << i;                                                 \
        }                                                                                  \
        return m;                                                                          \
    }

__device__ __forceinline__ int32_t as_i32(uint32_t v) { return (int32_t)v; }

// TupleUtils.java:48-57 (int) and :59-68 (float): 16 rows at once.  Bit u*4+j = row u*128+lane*4+j.
// The operator switch is outside the row loop: one compare + one bit insert per row.
template <int U = kUnits>
__device__ __forceinline__ uint32_t eval_term32(const DevTerm& t, int64_t thread_row0, const uint32_t* stage, int tile_off,
                                                const int stage_rows = kTileRows) {
    uint32_t a[U * kVec], b[U * kVec];
    load_operand32<U>(t.lhs, thread_row0, stage, tile_off, a, stage_rows);
    load_operand32<U>(t.rhs, thread_row0, stage, tile_off, b, stage_rows);
    if (t.cmp_type == MBC_ATTR_INTEGER) {
        switch (t.op) {
            case MBC_OP_EQ: MBC_CMP16(int32_t, as_i32, ==)
            case MBC_OP_LT: MBC_CMP16(int32_t, as_i32, <)
            case MBC_OP_GT: MBC_CMP16(int32_t, as_i32, >)
            case MBC_OP_NE: MBC_CMP16(int32_t, as_i32, !=)
            case MBC_OP_LE: MBC_CMP16(int32_t, as_i32, <=)
            case MBC_OP_GE: MBC_CMP16(int32_t, as_i32, >=)
            case MBC_OP_NOT: MBC_CMP16(int32_t, as_i32, !=)     // PredEval.java:158-160: aopNOT behaves as NE
            default: return 0u;                                  // aopNOP / opRANGE never match
        }
    } else {
        switch (t.op) {                                          // NaN-free by contract (TupleUtils.java:59-70)
            case MBC_OP_EQ: MBC_CMP16(float, __uint_as_float, ==)
            case MBC_OP_LT: MBC_CMP16(float, __uint_as_float, <)
            case MBC_OP_GT: MBC_CMP16(float, __uint_as_float, >)
            case MBC_OP_NE: MBC_CMP16(float, __uint_as_float, !=)
            case MBC_OP_LE: MBC_CMP16(float, __uint_as_float, <=)
            case MBC_OP_GE: MBC_CMP16(float, __uint_as_float, >=)
            case MBC_OP_NOT: MBC_CMP16(float, __uint_as_float, !=)
            default: return 0u;
        }
    }
}

// one string operand word w of row `row`: column word (zero beyond the stride) or literal word
__device__ __forceinline__ uint32_t str_word(const DevOperand& o, const uint32_t* lit, int lit_words, int64_t row, int w) {
    if (o.kind == 0) return w < lit_words ? lit[w] : 0u;
    int words = o.stride >> 2;
    if (w >= words) return 0u;
    return __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(o.ptr) + row * o.stride) + w);
}

// TupleUtils.java:70-81: String.compareTo on two zero-padded fixed-width byte strings.  For strings
// without NUL bytes (the contract) unsigned byte order over the padded width is compareTo's order.
// Rows are taken lane-contiguously (row = u*128 + k*32 + lane) and the 128 result bits of a unit
// are transposed into the 4-rows-per-lane layout through ballots.
template <int U = kUnits>
__device__ __forceinline__ uint32_t eval_term_str(const DevTerm& t, int64_t warp_row0, int lane) {
    int wl = t.lhs.kind == 0 ? t.lit_words : (t.lhs.stride >> 2);
    int wr = t.rhs.kind == 0 ? t.lit_words : (t.rhs.stride >> 2);
    int nwords = max(wl, wr);
    const bool fast16 = (t.lhs.kind != 0) != (t.rhs.kind != 0) &&           // column vs literal
                        (t.lhs.kind != 0 ? t.lhs.stride : t.rhs.stride) == 16 && t.lit_words <= 4;
    uint32_t mask = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint32_t bal[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int64_t row = warp_row0 + u * kUnitRows + k * 32 + lane;
            bool lt = false, eq = true;
            if (fast16) {
                const DevOperand& c = t.lhs.kind != 0 ? t.lhs : t.rhs;
                uint4 q = ldg128(reinterpret_cast<const char*>(c.ptr) + row * 16);
                uint32_t cw[4] = {q.x, q.y, q.z, q.w};
                bool clt = false, ceq = true;   // column vs literal
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t x = bswap32(cw[w]), y = bswap32(t.lit[w]);
                    if (ceq && x != y) { clt = x < y; ceq = false; }
                }
                if (t.lhs.kind != 0) { lt = clt; eq = ceq; }
                else { lt = !clt && !ceq; eq = ceq; }
            } else {
                for (int w = 0; w < nwords; ++w) {
                    uint32_t x = bswap32(str_word(t.lhs, t.lit, t.lit_words, row, w));
                    uint32_t y = bswap32(str_word(t.rhs, t.lit, t.lit_words, row, w));
                    if (x != y) { lt = x < y; eq = false; break; }
                }
            }
            bal[k] = __ballot_sync(0xFFFFFFFFu, apply_op(t.op, lt, eq));
        }
        // lane l needs bits (l*4 .. l*4+3) of the 128-bit unit: word l/8, shift (l%8)*4
        int sel = lane >> 3;
        uint32_t word = sel == 0 ? bal[0] : sel == 1 ? bal[1] : sel == 2 ? bal[2] : bal[3];
        mask |= ((word >> ((lane & 7) * 4)) & 0xFu) << (u * 4);
    }
    return mask;
}

// the 16 bits this thread owns out of a row bitmap (bit p = word p/32, bit p%32)
template <int U = kUnits>
__device__ __forceinline__ uint32_t load_bits(const uint32_t* bm, int64_t warp_row0, int lane) {
    uint32_t mask = 0;
    const uint32_t* w = bm + (warp_row0 >> 5) + (lane >> 3);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint32_t word = __ldg(w + u * (kUnitRows / 32));
        mask |= ((word >> ((lane & 7) * 4)) & 0xFu) << (u * 4);
    }
    return mask;
}

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------
// The kernels of one scan are launched back to back with programmaticStreamSerializationAllowed (mbc_scan.cu, MBC_PDL):
// a kernel lets its successor's CTAs be scheduled as soon as all of its own are running (pdl_trigger, first statement), and
// every successor blocks in pdl_wait (first statement) until its predecessor has completed and its writes are visible.  What
// is saved is the launch latency and the ramp between the kernels; no kernel reads anything before the wait.  Both are no-ops
// in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a) -----------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// wait for the phase of `bar` with the given parity; the hardware suspends the thread between polls
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
            : "memory");
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- the kernels --------------------------------------------------------------------------------
//
// The scan is two streaming passes with no dependency between CTAs (a single-pass version with a
// decoupled look-back was built first and measured: with ~300 persistent CTAs in flight on a B200 the
// look-back chain became a grid-wide barrier per wave of tiles and capped the low-selectivity case at
// ~1 TB/s; see DESIGN.md "What was tried").
//
//   filter_kernel   persistent CTAs, static tile assignment, ring of TMA-staged predicate columns:
//                   evaluates the CNF, emits the selection bitmap (BitSet order) and one count per tile.
//   tile_offsets_kernel  one block: exclusive scan of the tile counts -> global output offset of
//                   every tile (added to the running count, which is how chunked scans append).
//   write_kernel    one CTA per tile: ranks from the bitmap, rank->row list in shared memory, then
//                   one thread per SURVIVOR writes position / projected values and folds aggregates.

constexpr int kMaxStages = 4;
constexpr int kGatherBatch = 4;

template <typename T>
__device__ __forceinline__ T agg_combine(int kind, T a, T b) {
    return (kind == MBC_AGG_COUNT || kind == MBC_AGG_SUM) ? a + b : kind == MBC_AGG_MIN ? (a < b ? a : b) : (a > b ? a : b);
}

__device__ __forceinline__ unsigned long long agg_identity(const DevAgg& g) {
    const bool integral = g.kind == MBC_AGG_COUNT || g.type == MBC_ATTR_INTEGER;
    if (g.kind == MBC_AGG_MIN) return integral ? (unsigned long long)(long long)INT32_MAX : (unsigned long long)__double_as_longlong((double)INFINITY);
    if (g.kind == MBC_AGG_MAX) return integral ? (unsigned long long)(long long)INT32_MIN : (unsigned long long)__double_as_longlong((double)-INFINITY);
    return integral ? 0ull : (unsigned long long)__double_as_longlong(0.0);
}

__device__ __forceinline__ unsigned long long agg_merge(const DevAgg& g, unsigned long long a, unsigned long long b) {
    if (g.kind == MBC_AGG_COUNT || g.type == MBC_ATTR_INTEGER)
        return (unsigned long long)agg_combine<long long>(g.kind, (long long)a, (long long)b);
    return (unsigned long long)__double_as_longlong(agg_combine<double>(g.kind, __longlong_as_double((long long)a), __longlong_as_double((long long)b)));
}

// ---- pass 1: CNF -> selection bitmap + tile counts ------------------------------------------------------
#ifndef MBC_FILTER_CTAS
#define MBC_FILTER_CTAS 2
#endif
__global__ void __launch_bounds__(kScanThreads, MBC_FILTER_CTAS) filter_kernel(const __grid_constant__ ScanParams p) {
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t stage_mem[];          // [nstages][nstaged][kTileRows] uint32
    __shared__ __align__(8) uint64_t s_full[kMaxStages];
    __shared__ uint32_t s_wcnt[2][kWarpsPerCta];                   // by tile parity: see the note at the end of the loop

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int S = p.nstages;
    const uint32_t stage_bytes = (uint32_t)p.nstaged * kStageColBytes;

    auto issue_tile = [&](int slot, int tile) {                    // one elected thread
        mbar_arrive_expect_tx(&s_full[slot], stage_bytes);
        for (int c = 0; c < p.nstaged; ++c)
            tma_bulk_g2s(stage_mem + (size_t)slot * stage_bytes + (size_t)c * kStageColBytes,
                         reinterpret_cast<const uint8_t*>(p.staged_src[c]) + (size_t)tile * kStageColBytes, kStageColBytes,
                         &s_full[slot]);
    };

    if (tid == 0 && p.nstaged) {
        for (int s = 0; s < S; ++s) mbar_init(&s_full[s], 1);
        mbar_fence_init();
        for (int s = 0; s < S; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < p.ntiles) issue_tile(s, (int)t);
        }
    }
    __syncthreads();

    int slot = 0, flip = 0;
    uint32_t parity = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, flip ^= 1) {
        if (p.nstaged) mbar_wait(&s_full[slot], parity);
        const uint32_t* stage = reinterpret_cast<const uint32_t*>(stage_mem + (size_t)slot * stage_bytes);
        const int64_t warp_row0 = (int64_t)tile * kTileRows + warp * kWarpRows;
        const int64_t thread_row0 = warp_row0 + lane * kVec;
        const int tile_off = warp * kWarpRows + lane * kVec;       // this thread's first row within the tile

        uint32_t mask = 0xFFFFu;
        if (p.sel_bitmap) mask &= load_bits(p.sel_bitmap, warp_row0, lane);
        if (p.nterms > 0) {
            uint32_t acc = 0;
            for (int k = 0; k < p.nterms; ++k) {                   // warp-uniform term program
                const DevTerm& t = p.terms[k];
                acc |= (t.cmp_type == MBC_ATTR_STRING) ? eval_term_str(t, warp_row0, lane)
                                                       : eval_term32(t, thread_row0, stage, tile_off);
                if (t.end_conj) { mask &= acc; acc = 0; }          // OR inside, AND across (PredEval.java:164-176)
            }
        }
        if (p.deleted) mask &= ~load_bits(p.deleted, warp_row0, lane);   // TupleScan.java:85
        if (warp_row0 + kWarpRows > p.nrows) {                     // rows past the end of the table
#pragma unroll
            for (int u = 0; u < kUnits; ++u)
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    if (thread_row0 + u * kUnitRows + j >= p.nrows) mask &= ~(1u << (u * 4 + j));
        }
        // selection bitmap, java.util.BitSet order: bit p = word p/32, bit p%32
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
            uint32_t w = ((mask >> (u * 4)) & 0xFu) << ((lane & 7) * 4);
            w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
            w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
            w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
            if ((lane & 7) == 0) p.out_bitmap[(warp_row0 >> 5) + u * (kUnitRows / 32) + (lane >> 3)] = w;
        }
        const uint32_t wcnt = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(mask));
        if (lane == 0) s_wcnt[flip][warp] = wcnt;
        __syncthreads();                                           // every warp is done with the slot
        if (tid == 0) {
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < kWarpsPerCta; ++w) c += s_wcnt[flip][w];
            p.tile_counts[tile] = c;
            const long long next = tile + (long long)S * gridDim.x;
            if (p.nstaged && next < p.ntiles) issue_tile(slot, (int)next);
        }
        if (++slot == S) { slot = 0; parity ^= 1u; }
        // the other warps run ahead into the next tile while thread 0 is still summing: they write the other half of
        // s_wcnt, and this half is rewritten only two tiles on, with the next tile's barrier in between
    }
}

// ---- pass 1 of a scan that has no predicate terms, only a precomputed selection (bitmap CNF scans, join sides) -------
// One warp per tile, 128 bits per lane: selection AND NOT deleted, rows past the end cleared, tile count.  Independent
// warps instead of the filter pass's staged CTA loop: there is nothing to stage, a tile is 512 bytes of bitmap.
__global__ void __launch_bounds__(kScanThreads) select_bitmap_kernel(const uint32_t* __restrict__ sel, const uint32_t* __restrict__ deleted,
                                                                     int64_t nrows, int ntiles, uint32_t* __restrict__ out,
                                                                     uint32_t* __restrict__ tile_counts) {
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * kWarpsPerCta;
    for (int tile = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); tile < ntiles; tile += nwarps) {
        const size_t quad = (size_t)tile * (kTileRows / 128) + lane;
        uint4 q = __ldg(reinterpret_cast<const uint4*>(sel) + quad);
        if (deleted) {                                                 // TupleScan.java:85
            const uint4 d = __ldg(reinterpret_cast<const uint4*>(deleted) + quad);
            q.x &= ~d.x; q.y &= ~d.y; q.z &= ~d.z; q.w &= ~d.w;
        }
        const int64_t row0 = (int64_t)tile * kTileRows + lane * 128;
        if (row0 + 128 > nrows) {
            uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t valid = nrows - (row0 + 32 * j);
                if (valid <= 0) w[j] = 0u;
                else if (valid < 32) w[j] &= (1u << valid) - 1u;
            }
        }
        reinterpret_cast<uint4*>(out)[quad] = q;
        const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(__popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w)));
        if (lane == 0) tile_counts[tile] = c;
    }
}

// ---- pass 1.5: tile counts -> tile output offsets -----------------------------------------------------------
// Block b owns counts [b * 4096, (b + 1) * 4096), four per thread.  It sums every count before its range itself (the
// counts of even a 500 M-row table are a few hundred KB in L2), so the blocks are independent: no chain, no look-back.
// `start` is the running output offset left by the previous launch of a chunked scan (0 for the first launch); the
// last block publishes the new running offset in a DIFFERENT slot, since other blocks may still be reading the old one.
constexpr int kOffsetsPerBlock = 4096;
// How the write pass treats a group of 8 tiles, by its survivors: nothing to do / written whole by one CTA / gathered tile by
// tile / read whole through write_staged_kernel.  One byte per group, written next to the offsets: the CTAs of the write pass
// that have nothing to do (most of them, at either end of the density range) find out from a line that sits in their SM's L1
// instead of waiting ~1 us for the group's offsets from L2.
enum : uint8_t { kGroupEmpty = 0, kGroupSparse = 1, kGroupMid = 2, kGroupFull = 3 };
__global__ void __launch_bounds__(1024) tile_offsets_kernel(const uint32_t* counts, int ntiles, unsigned long long* tile_base,
                                                            const long long* running_in, long long* running_out,
                                                            unsigned int* work_counter, uint8_t* group_class, int sparse_max, int full_min) {
    pdl_trigger();
    pdl_wait();
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4* counts4 = reinterpret_cast<const uint4*>(counts);   // the buffer is padded: whole quads are readable
    const int first_quad = blockIdx.x * (kOffsetsPerBlock / 4);
    const unsigned long long start = running_in ? (unsigned long long)*running_in : 0ull;

    unsigned long long before = 0;                                   // everything ahead of this block
    for (int i = tid; i < first_quad; i += 1024) {
        const uint4 q = counts4[i];
        before += (unsigned long long)q.x + q.y + q.z + q.w;
    }
    const int i0 = (first_quad + tid) * 4;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (i0 < ntiles) q = counts4[first_quad + tid];
    if (i0 + 1 >= ntiles) q.y = 0u;
    if (i0 + 2 >= ntiles) q.z = 0u;
    if (i0 + 3 >= ntiles) q.w = 0u;
    const unsigned long long mine = (unsigned long long)q.x + q.y + q.z + q.w;
    {   // a group is the quads of two neighbouring threads (first_quad is even; kGroupTiles == 8 is asserted where it is defined)
        const unsigned long long gtot = mine + __shfl_xor_sync(0xFFFFFFFFu, mine, 1);
        if (!(tid & 1) && i0 < ntiles)
            group_class[i0 / 8] = gtot == 0 ? kGroupEmpty : gtot <= (unsigned long long)sparse_max ? kGroupSparse
                                              : gtot <= (unsigned long long)full_min ? kGroupMid : kGroupFull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = s_warp[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) s_prefix = v;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;                             // warp totals of the block's own counts
    __syncthreads();
    unsigned long long wsum = s_warp[lane], winc = wsum;             // every warp scans the 32 warp totals itself
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xFFFFFFFFu, winc, o);
        if (lane >= o) winc += n;
    }
    const unsigned long long warp_excl = __shfl_sync(0xFFFFFFFFu, winc - wsum, warp);
    const unsigned long long block_total = __shfl_sync(0xFFFFFFFFu, winc, 31);
    unsigned long long run = start + s_prefix + warp_excl + incl - mine;
    if (i0 < ntiles) {
        const uint32_t c[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < ntiles) { tile_base[i0 + j] = run; run += c[j]; }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const unsigned long long end = start + s_prefix + block_total;
        tile_base[ntiles] = end;                                     // end of the last tile: groups read [first, last + 1]
        *running_out = (long long)end;
        *work_counter = 0u;
    }
}

// ---- pass 2: ordered write of the survivors ----------------------------------------------------------------------
// write_kernel: one CTA per tile; tiles form GROUPS of kGroupTiles, classified by tile_offsets_kernel.  In a kGroupMid group
// (and a kGroupFull one when write_staged_kernel cannot take it) every CTA writes its own tile: ranks from the bitmap,
// rank -> row list in shared memory, one thread per SURVIVOR, gathers issued in batches before the dependent stores.  A
// kGroupSparse group (<= kSparseMax survivors) is written by its first CTA alone and the others exit: at low selectivity a
// tile holds a few dozen survivors and a CTA's load -> scan -> gather -> store chain is pure latency, so the group amortises
// it over 8x the rows.  There every thread owns kThreadWords consecutive words of the group's selection bitmap (128-bit
// loads, issued together with the group's offsets), a block scan of the popcounts ranks them, and the output columns go one
// per warp.
#ifndef MBC_WRITE_MIN_CTAS
#define MBC_WRITE_MIN_CTAS 4
#endif
#ifndef MBC_GROUP_TILES
#define MBC_GROUP_TILES 8
#endif
#ifndef MBC_SPARSE_MAX
#define MBC_SPARSE_MAX 4096
#endif
constexpr int kGroupTiles = MBC_GROUP_TILES;
constexpr int kGroupRows = kGroupTiles * kTileRows;                // <= 65536: list entries are uint16
constexpr int kGroupWords = kGroupRows / 32;
constexpr int kThreadWords = kGroupWords / kScanThreads;           // 4 (or 8) -> one (two) 128-bit loads per thread
constexpr int kListCap = MBC_SPARSE_MAX > kTileRows ? MBC_SPARSE_MAX : kTileRows;
constexpr int kSparseMax = MBC_SPARSE_MAX;                         // survivors per group handled item-per-warp
static_assert(kGroupTiles == 8, "tile_offsets_kernel classifies groups of two quads of tiles");
static_assert(kGroupRows <= 65536 && kThreadWords % 4 == 0 && kSparseMax <= kListCap && kTileRows <= kListCap && kWarpsPerCta <= 32, "group geometry");

template <typename V>
__device__ __forceinline__ void gather_store(const V* __restrict__ src, V* __restrict__ dst, const uint16_t* list, int first,
                                             int step, int n) {
    for (int k0 = first; k0 < n; k0 += step * kGatherBatch) {
        V v[kGatherBatch];
#pragma unroll
        for (int b = 0; b < kGatherBatch; ++b) {
            const int k = k0 + b * step;
            if (k < n) v[b] = __ldg(src + list[k]);
        }
#pragma unroll
        for (int b = 0; b < kGatherBatch; ++b) {
            const int k = k0 + b * step;
            if (k < n) dst[k] = v[b];
        }
    }
}

// gather of one 4-byte column that also feeds aggregates: every value is loaded ONCE, stored (dst may be null) and folded
// into the accumulators of the aggregates in `amask` (agg_fold4 is defined below)
__device__ __forceinline__ void agg_fold4(const DevAgg& g, unsigned long long& acc, const uint4 v, uint32_t bits);
static_assert(kMaxAgg <= 8 && kGatherBatch == 4, "aggregate masks are bytes; folds take four values");
__device__ __forceinline__ void gather_store_fold(const ScanParams& p, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                  const uint16_t* list, const int first, const int step, const int n, const uint32_t amask,
                                                  unsigned long long (&acc)[kMaxAgg]) {
    for (int k0 = first; k0 < n; k0 += step * kGatherBatch) {
        uint32_t v[kGatherBatch];
        uint32_t ok = 0;
#pragma unroll
        for (int b = 0; b < kGatherBatch; ++b) {
            const int k = k0 + b * step;
            v[b] = 0u;
            if (k < n) { v[b] = __ldg(src + list[k]); ok |= 1u << b; }
        }
        if (dst) {
#pragma unroll
            for (int b = 0; b < kGatherBatch; ++b)
                if ((ok >> b) & 1u) dst[k0 + b * step] = v[b];
        }
#pragma unroll
        for (int a = 0; a < kMaxAgg; ++a)
            if ((amask >> a) & 1u) agg_fold4(p.aggs[a], acc[a], make_uint4(v[0], v[1], v[2], v[3]), ok);
    }
}

__device__ __forceinline__ void gather_store_wide(const DevProj& pr, int64_t row0, long long out0, const uint16_t* list, int first,
                                                  int step, int n) {
    const int words = pr.stride >> 2;
    for (int k = first; k < n; k += step) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(pr.src) + (row0 + list[k]) * pr.stride);
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pr.dst) + (out0 + k) * pr.stride);
        for (int q = 0; q < words; ++q) dst[q] = __ldg(src + q);
    }
}

// One tile of a dense group: ranks from the bitmap in the filter pass's 16-rows-per-thread layout (the expansion is
// 16 predicated stores), rank -> row list, one thread per survivor.  Writes the tile's own partials.
// What a CTA needs before it can touch a column of its tile: the thread's selection bits, the tile's count and its output
// offset.  They are independent loads of data the earlier launches left in L2, but an L2 round trip under a saturated HBM
// is ~2 us: the callers issue them in ONE round (with the group's offsets, or while the previous tile is being written)
// instead of one after the other.
struct DenseTile {
    uint32_t mask;
    long long base;
    int T;
};
__device__ __forceinline__ DenseTile dense_tile_load(const ScanParams& p, const int tile) {
    DenseTile d;
    d.mask = load_bits(p.out_bitmap, (int64_t)tile * kTileRows + (threadIdx.x >> 5) * kWarpRows, threadIdx.x & 31);
    d.base = (long long)__ldg(p.tile_out + tile);
    d.T = (int)__ldg(p.tile_counts + tile);
    return d;
}

// Rank -> row list of one tile from the selection bits in the filter pass's 16-rows-per-thread layout (the expansion is 16
// predicated stores); s_wtot[w] is left holding the survivors of warp w's 512 rows.  Two barriers.
__device__ __forceinline__ void tile_list_build(const uint32_t mask, uint16_t* s_list, uint32_t* s_wtot) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    {
        const uint32_t packed = __popc(mask & 0xFu) | (__popc(mask & 0xF0u) << 8) | (__popc(mask & 0xF00u) << 16) |
                                (__popc(mask & 0xF000u) << 24);
        uint32_t incl = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);  // per-unit totals, each <= 128
        const uint32_t excl = incl - packed;
        uint32_t uoff[kUnits];
        uoff[0] = 0;
        uoff[1] = tot & 0xFFu;
        uoff[2] = uoff[1] + ((tot >> 8) & 0xFFu);
        uoff[3] = uoff[2] + ((tot >> 16) & 0xFFu);
        if (lane == 0) s_wtot[warp] = uoff[3] + (tot >> 24);
        __syncthreads();
        uint32_t wbase = (lane < warp) ? s_wtot[lane & (kWarpsPerCta - 1)] : 0u;   // warp < kWarpsPerCta <= 32 lanes hold the totals below
        wbase = __reduce_add_sync(0xFFFFFFFFu, wbase);
        if (mask) {
            uint16_t* list = s_list + wbase;
#pragma unroll
            for (int u = 0; u < kUnits; ++u) {
                uint32_t nib = (mask >> (u * 4)) & 0xFu;
                uint32_t r = uoff[u] + ((excl >> (8 * u)) & 0xFFu);
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    if ((nib >> j) & 1u) list[r++] = (uint16_t)(warp * kWarpRows + u * kUnitRows + lane * kVec + j);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void write_dense_tile(const ScanParams& p, const int tile, const DenseTile& d, uint16_t* s_list, uint32_t* s_wtot,
                                                 unsigned long long (*s_aggw)[kWarpsPerCta]) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int T = d.T;
    if (T == 0) {                                                  // block-uniform: nothing qualifies in this tile
        if (tid < p.nagg) p.partials[(size_t)tid * p.total_tiles + p.tile_base + tile] = agg_identity(p.aggs[tid]);
        return;
    }
    const long long base = d.base;
    const int64_t tile_row0 = (int64_t)tile * kTileRows;
    tile_list_build(d.mask, s_list, s_wtot);
    const uint16_t* list = s_list;
    if (p.out_pos)
        for (int k = tid; k < T; k += kScanThreads) p.out_pos[base + k] = p.pos_base + tile_row0 + list[k];
    // aggregates are folded from the values gathered for projection when their column is projected (one load per value),
    // else from one gather per distinct source column; per-thread accumulators, one butterfly per warp, warps combined in order
    unsigned long long acc[kMaxAgg];
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) acc[a] = a < p.nagg ? agg_identity(p.aggs[a]) : 0ull;
    for (int c = 0; c < p.nproj; ++c) {                            // iterator/Projection.java:103-144
        const DevProj& pr = p.proj[c];
        if (pr.stride == 4) {
            if (p.proj_aggs[c])
                gather_store_fold(p, reinterpret_cast<const uint32_t*>(pr.src) + tile_row0, reinterpret_cast<uint32_t*>(pr.dst) + base, list, tid,
                                  kScanThreads, T, p.proj_aggs[c], acc);
            else
                gather_store(reinterpret_cast<const uint32_t*>(pr.src) + tile_row0, reinterpret_cast<uint32_t*>(pr.dst) + base, list, tid, kScanThreads, T);
        } else if (pr.stride == 16) {
            gather_store(reinterpret_cast<const uint4*>(pr.src) + tile_row0, reinterpret_cast<uint4*>(pr.dst) + base, list, tid, kScanThreads, T);
        } else {
            gather_store_wide(pr, tile_row0, base, list, tid, kScanThreads, T);
        }
    }
    for (int a = 0; a < p.nagg; ++a)
        if (p.agg_group[a])
            gather_store_fold(p, reinterpret_cast<const uint32_t*>(p.aggs[a].src) + tile_row0, nullptr, list, tid, kScanThreads, T, p.agg_group[a], acc);
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) {
        if (a >= p.nagg) break;
        const DevAgg& g = p.aggs[a];
        if (g.kind == MBC_AGG_COUNT) continue;                     // the tile count is the count
        unsigned long long v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = agg_merge(g, v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
        if (lane == 0) s_aggw[a][warp] = v;
    }
    __syncthreads();
    if (tid < p.nagg) {
        const DevAgg& g = p.aggs[tid];
        unsigned long long v;
        if (g.kind == MBC_AGG_COUNT) {
            v = (unsigned long long)T;
        } else {
            v = s_aggw[tid][0];
            for (int x = 1; x < kWarpsPerCta; ++x) v = agg_merge(g, v, s_aggw[tid][x]);
        }
        p.partials[(size_t)tid * p.total_tiles + p.tile_base + tile] = v;
    }
}

// fold of four values of one aggregate source, bit j of `bits` says whether value j takes part (write_staged_kernel)
__device__ __forceinline__ void agg_fold4(const DevAgg& g, unsigned long long& acc, const uint4 v, uint32_t bits) {
    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
    if (g.type == MBC_ATTR_INTEGER) {
        long long a = (long long)acc;
        if (g.kind == MBC_AGG_SUM) {
#pragma unroll
            for (int j = 0; j < 4; ++j) a += ((bits >> j) & 1u) ? (long long)(int32_t)x[j] : 0ll;
        } else if (g.kind == MBC_AGG_MIN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((bits >> j) & 1u) a = min(a, (long long)(int32_t)x[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((bits >> j) & 1u) a = max(a, (long long)(int32_t)x[j]);
        }
        acc = (unsigned long long)a;
    } else {
        double a = __longlong_as_double((long long)acc);
        if (g.kind == MBC_AGG_SUM) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((bits >> j) & 1u) a += (double)__uint_as_float(x[j]);
        } else if (g.kind == MBC_AGG_MIN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((bits >> j) & 1u) a = fmin(a, (double)__uint_as_float(x[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((bits >> j) & 1u) a = fmax(a, (double)__uint_as_float(x[j]));
        }
        acc = (unsigned long long)__double_as_longlong(a);
    }
}

// one work item of write_sparse_group (one warp): a 4-byte column gathered once for its projection (dst may be null) and for
// the aggregates in amask; their partials of the group go to slot part0
__device__ __forceinline__ void group_fold_item(const ScanParams& p, const uint32_t* src, uint32_t* dst, const uint16_t* list, const int T,
                                                const uint32_t amask, const size_t part0) {
    const int lane = threadIdx.x & 31;
    unsigned long long acc[kMaxAgg];
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) acc[a] = ((amask >> a) & 1u) ? agg_identity(p.aggs[a]) : 0ull;
    gather_store_fold(p, src, dst, list, lane, 32, T, amask, acc);
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) {
        if (!((amask >> a) & 1u)) continue;
        const DevAgg& g = p.aggs[a];
        unsigned long long v = acc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = agg_merge(g, v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
        if (lane == 0) p.partials[(size_t)a * p.total_tiles + part0] = v;
    }
}

// A sparse group (0 < total <= kSparseMax survivors in ntl tiles from tile0), written by one CTA.
// The thread's words of a group's selection bitmap: loaded by the caller TOGETHER with the group's offsets (they do not depend
// on them), one memory round trip instead of two in front of a chain that is nothing but round trips.
struct GroupBits {
    uint4 q[kThreadWords / 4];
};
__device__ __forceinline__ GroupBits group_bits_load(const ScanParams& p, const int tile0, const int ntl) {
    GroupBits b;
    const int word0 = threadIdx.x * kThreadWords;
#pragma unroll
    for (int q = 0; q < kThreadWords; q += 4) {
        b.q[q / 4] = make_uint4(0u, 0u, 0u, 0u);
        if (word0 + q < ntl * (kTileRows / 32)) b.q[q / 4] = ldg128(p.out_bitmap + (size_t)tile0 * (kTileRows / 32) + word0 + q);
    }
    return b;
}

__device__ __forceinline__ void write_sparse_group(const ScanParams& p, const int tile0, const int ntl, const long long base, const int total,
                                                   const GroupBits& pre, uint16_t* s_list, uint32_t* s_wtot) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const size_t part0 = (size_t)p.tile_base + tile0;
    const int64_t row0 = (int64_t)tile0 * kTileRows;

    uint32_t w[kThreadWords];
    const int word0 = tid * kThreadWords;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < kThreadWords; q += 4) {
        const uint4 v = pre.q[q / 4];
        w[q] = v.x; w[q + 1] = v.y; w[q + 2] = v.z; w[q + 3] = v.w;
        cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_wtot[warp] = (uint32_t)incl;
    __syncthreads();
    uint32_t below = (lane < warp) ? s_wtot[lane & (kWarpsPerCta - 1)] : 0u;
#pragma unroll
    for (int o = 1; o < kWarpsPerCta; o <<= 1) below += __shfl_xor_sync(0xFFFFFFFFu, below, o);
    const int r0 = (int)__shfl_sync(0xFFFFFFFFu, below, 0) + incl - cnt;

    // expansion: every thread walks its words, rank -> row within the group
    if (cnt) {
        int r = r0;
#pragma unroll
        for (int i = 0; i < kThreadWords; ++i) {
            uint32_t bits = w[i];
            const int rowb = (word0 + i) * 32;
            while (bits) {
                s_list[r++] = (uint16_t)(rowb + __ffs(bits) - 1);
                bits &= bits - 1;
            }
        }
    }
    __syncthreads();
    // one work item (positions / one projected column / one aggregate) per warp, so the dependent load -> store
    // chains of the columns run side by side
    const uint16_t* list = s_list;
    const int T = total;
    const int nitems = 1 + p.nproj + p.nagg;
    for (int item = warp; item < nitems; item += kWarpsPerCta) {
        if (item == 0) {
            if (p.out_pos)
                for (int k = lane; k < T; k += 32) p.out_pos[base + k] = p.pos_base + row0 + list[k];
        } else if (item <= p.nproj) {                              // iterator/Projection.java:103-144
            const DevProj& pr = p.proj[item - 1];
            const uint32_t amask = p.proj_aggs[item - 1];
            if (pr.stride == 4 && amask)
                group_fold_item(p, reinterpret_cast<const uint32_t*>(pr.src) + row0, reinterpret_cast<uint32_t*>(pr.dst) + base, list, T, amask, part0);
            else if (pr.stride == 4)
                gather_store(reinterpret_cast<const uint32_t*>(pr.src) + row0, reinterpret_cast<uint32_t*>(pr.dst) + base, list, lane, 32, T);
            else if (pr.stride == 16)
                gather_store(reinterpret_cast<const uint4*>(pr.src) + row0, reinterpret_cast<uint4*>(pr.dst) + base, list, lane, 32, T);
            else
                gather_store_wide(pr, row0, base, list, lane, 32, T);
        } else {
            const int a = item - 1 - p.nproj;
            const DevAgg& g = p.aggs[a];
            if (g.kind == MBC_AGG_COUNT) {
                if (lane == 0) p.partials[(size_t)a * p.total_tiles + part0] = (unsigned long long)T;
            } else if (p.agg_group[a]) {                           // every aggregate of this source column, from one gather
                group_fold_item(p, reinterpret_cast<const uint32_t*>(g.src) + row0, nullptr, list, T, p.agg_group[a], part0);
            }
        }
    }
}


template <bool kPersistent>
__global__ void __launch_bounds__(kScanThreads, MBC_WRITE_MIN_CTAS) write_kernel(const __grid_constant__ ScanParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ uint16_t s_list[kListCap];                          // survivor rows within the group / tile, by rank
    __shared__ uint32_t s_wtot[kWarpsPerCta];
    __shared__ unsigned long long s_aggw[kMaxAgg][kWarpsPerCta];
    const int tid = threadIdx.x;
  if constexpr (kPersistent) {
    // Persistent form, for large tables: a few CTAs per SM take the groups by ticket (sparse ones are written whole),
    // then stride over the tiles (those of dense groups).  One CTA per tile costs ~0.6 ns per CTA of pure dispatch,
    // which a selective scan of 10^5 tiles, most of them skipped, would pay for nothing (C3 scan: 0.30 -> 0.21 ms).
    __shared__ uint32_t s_flags[kWarpsPerCta];
    __shared__ int s_group;
    const int ngroups = (p.ntiles + kGroupTiles - 1) / kGroupTiles;
    // groups are handed out by a counter (zeroed by tile_offsets_kernel): a few thousand groups over a few hundred
    // CTAs would otherwise quantise into whole rounds.  The next ticket is in flight while the current group is written.
    int ticket = 0;
    if (tid == 0) ticket = (int)atomicAdd(p.work_counter, 1u);
    for (;;) {
        if (tid == 0) s_group = ticket;
        __syncthreads();
        const int g = s_group;
        __syncthreads();                                           // everyone has the group before thread 0 moves on
        if (g >= ngroups) break;
        if (tid == 0) ticket = (int)atomicAdd(p.work_counter, 1u);
        const int tile0 = g * kGroupTiles;
        const int ntl = min(kGroupTiles, p.ntiles - tile0);
        const uint32_t cls = __ldg(p.group_class + g);
        if (cls >= kGroupMid) continue;                            // block-uniform: written tile by tile, below or by write_staged_kernel
        GroupBits bits;
        if (cls == kGroupSparse) bits = group_bits_load(p, tile0, ntl);   // one round of loads with the group's offsets
        const long long base = (long long)__ldg(p.tile_out + tile0);
        const int total = (int)((long long)__ldg(p.tile_out + tile0 + ntl) - base);
        for (int i = tid; i < p.nagg * ntl; i += kScanThreads) {   // one partial for the group (below), identity elsewhere
            const int a = i / ntl, t = i % ntl;
            if (t > 0 || total == 0) p.partials[(size_t)a * p.total_tiles + p.tile_base + tile0 + t] = agg_identity(p.aggs[a]);
        }
        if (total == 0) continue;
        write_sparse_group(p, tile0, ntl, base, total, bits, s_list, s_wtot);
    }
    const int niter = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int it0 = 0; it0 < niter; it0 += kScanThreads) {
        // which of this CTA's next 256 tiles belong to dense groups: one round of loads for all of them
        const int tile_t = blockIdx.x + (it0 + tid) * gridDim.x;
        bool dense = false;
        if (it0 + tid < niter) {
            const int tile0 = tile_t & ~(kGroupTiles - 1);
            const uint32_t cls = __ldg(p.group_class + tile0 / kGroupTiles);
            dense = cls == kGroupMid || (cls == kGroupFull && !p.dense_staged);   // the fullest go to write_staged_kernel
        }
        const uint32_t flags = __ballot_sync(0xFFFFFFFFu, dense);
        if (!__syncthreads_or(dense)) continue;                    // (the previous round is done with s_flags)
        if ((tid & 31) == 0) s_flags[tid >> 5] = flags;
        __syncthreads();
        const int n = min(kScanThreads, niter - it0);
        auto next_dense = [&](int j) {                             // block-uniform: first flagged tile at or after j (flags beyond n are 0)
            while (j < kScanThreads) {
                const uint32_t w = s_flags[j >> 5] >> (j & 31);
                if (w) return j + __ffs(w) - 1;
                j = (j | 31) + 1;
            }
            return n;
        };
        int j = next_dense(0);
        DenseTile cur = {0u, 0ll, 0};
        if (j < n) cur = dense_tile_load(p, blockIdx.x + (it0 + j) * gridDim.x);
        while (j < n) {
            const int jn = next_dense(j + 1);
            DenseTile nxt = {0u, 0ll, 0};
            if (jn < n) nxt = dense_tile_load(p, blockIdx.x + (it0 + jn) * gridDim.x);   // in flight while tile j is written
            __syncthreads();                                       // the shared arrays are reused from tile to tile
            write_dense_tile(p, blockIdx.x + (it0 + j) * gridDim.x, cur, s_list, s_wtot, s_aggw);
            cur = nxt;
            j = jn;
        }
    }
  } else {
    // One CTA per tile: the hardware scheduler balances them.  Most CTAs have nothing to write (the tiles of a sparse group
    // but its first, the tiles write_staged_kernel takes) and learn it from the group's class byte.
    const int tile = blockIdx.x;
    const int tile0 = tile & ~(kGroupTiles - 1);
    const uint32_t cls = __ldg(p.group_class + tile / kGroupTiles);
    if (cls == kGroupFull && p.dense_staged) return;               // block-uniform: write_staged_kernel's, partials included
    if (cls == kGroupEmpty || (cls == kGroupSparse && tile != tile0)) {   // one partial for a sparse group, the identity elsewhere
        if (tid < p.nagg) p.partials[(size_t)tid * p.total_tiles + p.tile_base + tile] = agg_identity(p.aggs[tid]);
        return;
    }
    if (cls != kGroupSparse) {                                     // every CTA of the group gathers its own tile
        write_dense_tile(p, tile, dense_tile_load(p, tile), s_list, s_wtot, s_aggw);
        return;
    }
    const int ntl = min(kGroupTiles, p.ntiles - tile0);            // tiles of this group
    const GroupBits bits = group_bits_load(p, tile0, ntl);         // one round of loads with the group's offsets
    const long long base = (long long)__ldg(p.tile_out + tile0);
    const int total = (int)((long long)__ldg(p.tile_out + tile0 + ntl) - base);
    write_sparse_group(p, tile0, ntl, base, total, bits, s_list, s_wtot);
  }
}

// ---- pass 2, dense groups: TMA-staged compaction ---------------------------------------------------------------------
// The tiles of the fullest groups (> p.stg_min survivors in kGroupTiles tiles) read every column of the write pass WHOLE:
// above ~1/3 density nearly every 32-byte sector of a column holds a survivor, so a gather fetches the column anyway -- through
// 4-byte loads whose addresses depend on the rank -> row list, which depends on the selection bits, which depend on the
// tile offsets: three L2/DRAM round trips (~2 us each under a saturated HBM) in front of every tile's first useful byte.
// Here persistent CTAs (two per SM) keep a ring of kStgRows-row stages filled by the TMA engine (cp.async.bulk, one bulk copy
// per column and stage, completion on an mbarrier); a producer warp runs ahead of the consumers by the depth of the ring and
// depends on none of those loads.  The eight consumer warps never wait for each other: warp w owns rows [128 w, 128 w + 128)
// of every stage, lane l rows 32 j + l of them (j = 0..3), so that
//   * the selection bits of an instruction's 32 rows are ONE word of the bitmap (the warp's four words come by shuffle from
//     the lane that loaded them; a lane holds four of the tile's 128 words, fetched one tile ahead together with the tile's
//     output offset),
//   * the rank of a row is the tile prefix of that word (one warp scan of popcounts per tile) plus popc(word & lanes below),
//   * shared-memory reads are conflict free (consecutive lanes, consecutive rows) and every store instruction writes the
//     survivors of its 32 rows back to back: compaction happens in the addresses, nothing is staged twice.
// A stage is handed back to the producer by one mbarrier arrival per warp.  Aggregates are folded from the same shared-memory
// values into per-lane accumulators that live for the whole kernel: one partial per CTA (in the slot of its first dense
// tile; its other dense tiles carry the identity, COUNT slots the tile's own count), so sums stay reproducible run to run.
// write_kernel (launched first, p.dense_staged = 1) writes the other groups.
constexpr int kStgRows = 1024;
constexpr int kStgSteps = kTileRows / kStgRows;
constexpr int kStgMaxStages = 6;
constexpr int kStgThreads = kScanThreads + 32;                     // eight consumer warps + the producer warp
constexpr int kStgWarpRows = kStgRows / kWarpsPerCta;              // 128: four 32-row instructions per warp and stage
static_assert(kStgWarpRows == 128 && kTileRows / 32 == 128 && kStgSteps * kWarpsPerCta == 32, "stage geometry: a lane holds 4 of the tile's 128 bitmap words");

__global__ void __launch_bounds__(kStgThreads, 2) write_staged_kernel(const __grid_constant__ ScanParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) uint8_t s_ring[];             // [stg_stages][stg_bytes]
    __shared__ __align__(8) uint64_t s_full[kStgMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kStgMaxStages];
    __shared__ unsigned long long s_aggw[kMaxAgg][kWarpsPerCta];
    __shared__ uint32_t s_flags[kWarpsPerCta];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int S = p.stg_stages;
    const uint32_t stage_bytes = (uint32_t)p.stg_bytes;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kWarpsPerCta); }
        mbar_fence_init();
    }
    int slot = 0;                                                  // this thread's cursor in the ring and its phase
    uint32_t phase = 0;
    int first_tile = -1;                                           // block-uniform: the CTA's first dense tile
    unsigned long long acc[kMaxAgg];
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) acc[a] = a < p.nagg ? agg_identity(p.aggs[a]) : 0ull;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const int niter = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int it0 = 0; it0 < niter; it0 += kScanThreads) {
        // which of this CTA's next 256 tiles belong to dense groups: one round of loads for all of them
        bool dense = false;
        if (tid < kScanThreads && it0 + tid < niter) {
            const int tile_t = blockIdx.x + (it0 + tid) * gridDim.x;
            dense = __ldg(p.group_class + tile_t / kGroupTiles) == kGroupFull;
        }
        const uint32_t flags = __ballot_sync(0xFFFFFFFFu, dense);
        if (!__syncthreads_or(dense)) continue;                    // (the previous round is over: the ring is empty, s_flags is free)
        if (warp < kWarpsPerCta && lane == 0) s_flags[warp] = flags;
        __syncthreads();
        const int n = min(kScanThreads, niter - it0);
        auto tile_of = [&](int j) { return (int)blockIdx.x + (it0 + j) * (int)gridDim.x; };
        auto next_dense = [&](int j) {                             // block-uniform: first flagged tile at or after j (flags beyond n are 0)
            while (j < kScanThreads) {
                const uint32_t w = s_flags[j >> 5] >> (j & 31);
                if (w) return j + __ffs(w) - 1;
                j = (j | 31) + 1;
            }
            return n;
        };
        if (first_tile < 0) {
            const int j0 = next_dense(0);
            if (j0 < n) first_tile = tile_of(j0);
        }
        if (warp == kWarpsPerCta) {
            // ---- producer: one lane feeds the ring, a stage as soon as all eight warps have handed it back
            if (lane == 0) {
                for (int j = next_dense(0); j < n; j = next_dense(j + 1)) {
                    const int64_t tile_row0 = (int64_t)tile_of(j) * kTileRows;
                    for (int step = 0; step < kStgSteps; ++step) {
                        mbar_wait(&s_empty[slot], phase ^ 1u);    // passes at once the first time round the ring
                        uint8_t* st = s_ring + (size_t)slot * stage_bytes;
                        const int64_t row0 = tile_row0 + step * kStgRows;
                        mbar_arrive_expect_tx(&s_full[slot], stage_bytes);
                        for (int c = 0; c < p.stg_n; ++c)
                            tma_bulk_g2s(st + p.stg_off[c], reinterpret_cast<const uint8_t*>(p.stg_src[c]) + row0 * p.stg_stride[c],
                                         (uint32_t)(kStgRows * p.stg_stride[c]), &s_full[slot]);
                        if (++slot == S) { slot = 0; phase ^= 1u; }
                    }
                }
            }
            continue;
        }
        // ---- consumers
        int j = next_dense(0);
        uint4 wcur = make_uint4(0u, 0u, 0u, 0u);                   // words 4 lane .. 4 lane + 3 of the tile's selection bitmap
        long long bcur = 0;
        if (j < n) {
            const int t = tile_of(j);
            wcur = ldg128(p.out_bitmap + (size_t)t * (kTileRows / 32) + lane * 4);
            bcur = (long long)__ldg(p.tile_out + t);
        }
        while (j < n) {
            const int jn = next_dense(j + 1);
            uint4 wnxt = make_uint4(0u, 0u, 0u, 0u);
            long long bnxt = 0;
            if (jn < n) {                                          // in flight while tile j is written
                const int t = tile_of(jn);
                wnxt = ldg128(p.out_bitmap + (size_t)t * (kTileRows / 32) + lane * 4);
                bnxt = (long long)__ldg(p.tile_out + t);
            }
            const int tile = tile_of(j);
            // survivors ahead of this lane's four words within the tile
            const int mine = __popc(wcur.x) + __popc(wcur.y) + __popc(wcur.z) + __popc(wcur.w);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - mine;
            const int T = __shfl_sync(0xFFFFFFFFu, incl, 31);      // the tile's count
            if (warp == 0 && lane < p.nagg) {                      // the CTA's one partial is written at the very end
                const DevAgg& g = p.aggs[lane];
                p.partials[(size_t)lane * p.total_tiles + p.tile_base + tile] = g.kind == MBC_AGG_COUNT ? (unsigned long long)T : agg_identity(g);
            }
#pragma unroll 1
            for (int step = 0; step < kStgSteps; ++step) {
                const int src = step * kWarpsPerCta + warp;        // the lane that holds this warp's 128 rows of the stage
                uint32_t q[4];
                q[0] = __shfl_sync(0xFFFFFFFFu, wcur.x, src);
                q[1] = __shfl_sync(0xFFFFFFFFu, wcur.y, src);
                q[2] = __shfl_sync(0xFFFFFFFFu, wcur.z, src);
                q[3] = __shfl_sync(0xFFFFFFFFu, wcur.w, src);
                long long out[4];                                  // output row of this lane's j-th row, if it survives
                out[0] = bcur + __shfl_sync(0xFFFFFFFFu, excl, src) + __popc(q[0] & lanes_below);
                out[1] = out[0] - __popc(q[0] & lanes_below) + __popc(q[0]) + __popc(q[1] & lanes_below);
                out[2] = out[1] - __popc(q[1] & lanes_below) + __popc(q[1]) + __popc(q[2] & lanes_below);
                out[3] = out[2] - __popc(q[2] & lanes_below) + __popc(q[2]) + __popc(q[3] & lanes_below);
                const uint32_t bits = ((q[0] >> lane) & 1u) | (((q[1] >> lane) & 1u) << 1) | (((q[2] >> lane) & 1u) << 2) | (((q[3] >> lane) & 1u) << 3);
                mbar_wait(&s_full[slot], phase);                   // the stage's columns have landed
                if ((q[0] | q[1] | q[2] | q[3]) != 0u) {           // warp-uniform
                    const uint8_t* st = s_ring + (size_t)slot * stage_bytes;
                    const int r0 = warp * kStgWarpRows + lane;     // this lane's rows of the stage: r0 + 32 i
                    if (p.out_pos) {
                        const int64_t pos0 = p.pos_base + (int64_t)tile * kTileRows + step * kStgRows + r0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if ((bits >> i) & 1u) p.out_pos[out[i]] = pos0 + 32 * i;
                    }
                    for (int c = 0; c < p.nproj; ++c) {            // iterator/Projection.java:103-144
                        const DevProj& pr = p.proj[c];
                        const uint8_t* col_s = st + p.stg_off[p.proj_stg[c]];
                        if (pr.stride == 4) {
                            const uint32_t* sv = reinterpret_cast<const uint32_t*>(col_s) + r0;
                            uint32_t* dst = reinterpret_cast<uint32_t*>(pr.dst);
                            uint32_t v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = sv[32 * i];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if ((bits >> i) & 1u) dst[out[i]] = v[i];
                        } else if (pr.stride == 16) {
                            const uint4* sv = reinterpret_cast<const uint4*>(col_s) + r0;
                            uint4* dst = reinterpret_cast<uint4*>(pr.dst);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if ((bits >> i) & 1u) dst[out[i]] = sv[32 * i];
                        } else {
                            const int words = pr.stride >> 2;
                            for (int i = 0; i < 4; ++i) {
                                if (!((bits >> i) & 1u)) continue;
                                const uint32_t* sv = reinterpret_cast<const uint32_t*>(col_s + (size_t)(r0 + 32 * i) * pr.stride);
                                uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pr.dst) + out[i] * pr.stride);
                                for (int w = 0; w < words; ++w) dst[w] = sv[w];
                            }
                        }
                    }
#pragma unroll
                    for (int a = 0; a < kMaxAgg; ++a) {
                        if (a >= p.nagg) break;
                        const DevAgg& g = p.aggs[a];
                        if (g.kind == MBC_AGG_COUNT) continue;     // the tile counts are the count
                        const uint32_t* sv = reinterpret_cast<const uint32_t*>(st + p.stg_off[p.agg_stg[a]]) + r0;
                        agg_fold4(g, acc[a], make_uint4(sv[0], sv[32], sv[64], sv[96]), bits);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[slot]);        // this warp is done with the stage
                if (++slot == S) { slot = 0; phase ^= 1u; }
            }
            wcur = wnxt;
            bcur = bnxt;
            j = jn;
        }
    }
    // the CTA's aggregate partial: lanes butterflied, warps combined in order
    if (p.nagg > 0 && first_tile >= 0) {                           // block-uniform
        if (warp < kWarpsPerCta) {
#pragma unroll
            for (int a = 0; a < kMaxAgg; ++a) {
                if (a >= p.nagg) break;
                const DevAgg& g = p.aggs[a];
                unsigned long long v = acc[a];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = agg_merge(g, v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
                if (lane == 0) s_aggw[a][warp] = v;
            }
        }
        __syncthreads();
        if (tid < p.nagg && p.aggs[tid].kind != MBC_AGG_COUNT) {
            const DevAgg& g = p.aggs[tid];
            unsigned long long v = s_aggw[tid][0];
            for (int x = 1; x < kWarpsPerCta; ++x) v = agg_merge(g, v, s_aggw[tid][x]);
            p.partials[(size_t)tid * p.total_tiles + p.tile_base + first_tile] = v;
        }
    }
}

// Reduce the per-tile partials of one aggregate (fixed association => reproducible sums): thread i folds
// entries i, i+1024, ... then a fixed tree combines the 1024 lanes.
struct AggList {
    DevAgg g[kMaxAgg];
};

__global__ void __launch_bounds__(1024) agg_finish_kernel(const unsigned long long* partials, int total_tiles,
                                                          int ntiles, const __grid_constant__ AggList list,
                                                          unsigned long long* out, const long long* count) {
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[kMaxAgg] = (unsigned long long)*count;   // aggregates + count leave in one copy
    const DevAgg g = list.g[blockIdx.x];
    const unsigned long long* src = partials + (size_t)blockIdx.x * total_tiles;
    __shared__ unsigned long long sh[1024];
    unsigned long long acc = agg_identity(g);
    int i = threadIdx.x;
    for (; i + 3 * 1024 < ntiles; i += 4 * 1024) {                 // four independent loads in flight
        unsigned long long x0 = src[i], x1 = src[i + 1024], x2 = src[i + 2048], x3 = src[i + 3072];
        acc = agg_merge(g, agg_merge(g, agg_merge(g, agg_merge(g, acc, x0), x1), x2), x3);
    }
    for (; i < ntiles; i += 1024) acc = agg_merge(g, acc, src[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = agg_merge(g, sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// COUNT is the only aggregate asked for: the running count is the answer, no partials to reduce
__global__ void agg_count_only_kernel(int nagg, unsigned long long* out, const long long* count) {
    if (threadIdx.x <= kMaxAgg) out[threadIdx.x] = (threadIdx.x < nagg || threadIdx.x == kMaxAgg) ? (unsigned long long)*count : 0ull;
}

struct TupleField {
    const void* src;
    int32_t type, width, stride, offset;
};
struct TupleParams {
    int64_t count;
    int32_t nfields, tuple_len, rows_per_cta, pad;
    uint8_t* out;
    TupleField f[kMaxProj];
};

__global__ void __launch_bounds__(128) tuple_encode_kernel(const __grid_constant__ TupleParams p) {
    extern __shared__ uint8_t sm[];
    const int64_t row0 = (int64_t)blockIdx.x * p.rows_per_cta;
    const int nrows = (int)min((long long)p.rows_per_cta, (long long)(p.count - row0));
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        uint8_t* t = sm + (size_t)r * p.tuple_len;
        const int64_t row = row0 + r;
        
        t[0] = (uint8_t)(p.nfields >> 8);
        t[1] = (uint8_t)p.nfields;
        for (int f = 0; f < p.nfields; ++f) {
            int off = p.f[f].offset;
            t[2 + 2 * f] = (uint8_t)(off >> 8);
            t[3 + 2 * f] = (uint8_t)off;
        }
        t[2 + 2 * p.nfields] = (uint8_t)(p.tuple_len >> 8);
        t[3 + 2 * p.nfields] = (uint8_t)p.tuple_len;
        
        for (int f = 0; f < p.nfields; ++f) {
            const TupleField& fd = p.f[f];
            uint8_t* d = t + fd.offset;
            if (fd.type != MBC_ATTR_STRING) {
                uint32_t v = reinterpret_cast<const uint32_t*>(fd.src)[row];
                d[0] = (uint8_t)(v >> 24); d[1] = (uint8_t)(v >> 16); d[2] = (uint8_t)(v >> 8); d[3] = (uint8_t)v;
            } else {
                const uint8_t* s = reinterpret_cast<const uint8_t*>(fd.src) + row * fd.stride;
                int len = 0;
                for (int k = 0; k < fd.width; ++k) {
                    uint8_t b = s[k];
                    d[2 + k] = b;
                    if (b != 0) len = k + 1;
                }
                d[0] = (uint8_t)(len >> 8);
                d[1] = (uint8_t)len;
            }
        }
    }
    __syncthreads();
    // contiguous, coalesced copy-out of the CTA's tuples
    const size_t bytes = (size_t)nrows * p.tuple_len;
    uint8_t* dst = p.out + (size_t)row0 * p.tuple_len;
    if ((((size_t)row0 * p.tuple_len) & 3) == 0) {
        const size_t words = bytes >> 2;
        for (size_t i = threadIdx.x; i < words; i += blockDim.x)
            reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(sm)[i];
        for (size_t i = (words << 2) + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = sm[i];
    } else {
        for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = sm[i];
    }
}
}  // namespace mbc
