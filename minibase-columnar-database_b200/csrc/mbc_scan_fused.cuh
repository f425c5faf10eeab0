// mbc_scan_fused.cuh -- K2+K5 in ONE residency: every column of the table is read from HBM once.
//
// Replaces the same reference loop as mbc_scan_kernels.cuh (iterator/ColumnarFileScan.java:156-172:
// while (scan.getNext) if (PredEval.Eval) Project), for scans whose projected / aggregated columns fit a
// shared-memory tile ring.  The two-pass engine (filter_kernel -> tile_offsets_kernel -> write_kernel) streams
// the predicate columns in pass 1 and then fetches the survivors' values with scattered 4..16-byte gathers, so a
// dense scan re-reads the predicate columns and pulls 64..128-byte DRAM lines around every gathered value
// (measured in round 1: 7.98 GB of DRAM traffic for 3.94 GB of useful bytes in the write pass).  Here:
//
//   * persistent CTAs (cooperative launch: the CTAs exchange tile counts, so they must be co-resident), tiles of kFR
//     rows assigned statically: CTA b owns tiles b, b + G, b + 2G, ...   ("wave" i = tiles [iG, iG + G));
//   * every CTA runs a three-stage software pipeline, one tile of each stage per iteration, all warps in every stage:
//       COUNT   tile i      the predicate columns arrive through a TMA ring (cp.async.bulk + mbarrier); the CNF (the same
//                           term program as filter_kernel) leaves the tile's selection words in a shared-memory mask ring
//                           and in the result bitmap, and the tile's count is PUBLISHED in global memory, tagged with the
//                           launch's epoch;
//       CONTROL tile i - a  (warp 0) the tile's output offset comes from the published counts WITHOUT a chain between
//                           CTAs: the warp sums every count of the wave itself (G words out of L2, requested one iteration
//                           early; they were published `a` tile-times ago), keeps the total of all earlier waves in a
//                           register, and decides how the survivors' values are fetched: a DENSE tile gets its projected
//                           columns bulk-copied whole into the payload ring (the predicate columns among them come out of
//                           L2: they were streamed a few microseconds earlier), a SPARSE tile is gathered later;
//       WRITE   tile i - a - l   ranks from the mask ring, rank -> row list, then one thread per SURVIVOR compacts the
//                           tile OUT OF SHARED MEMORY: coalesced stores in ascending position order; sparse tiles join a
//                           pending list that is gathered from global memory a few hundred survivors at a time, so the
//                           gather latency is paid once per batch and not once per tile;
//   * COUNT/SUM/MIN/MAX are folded into per-thread registers across all the tiles of the CTA and combined once, in a fixed
//     order, when the CTA finishes: one partial per CTA (reproducible run to run: the tile assignment is static).
//
// Output order is position order by construction (offsets, not atomics): bit-exact position lists.
#pragma once
#include "mbc_scan_kernels.cuh"

namespace mbc {

#ifndef MBC_FUSED_THREADS
#define MBC_FUSED_THREADS 256
#endif
constexpr int kFThreads = MBC_FUSED_THREADS;                      // threads of a CTA; a thread owns 4 consecutive rows of a tile
constexpr int kFWarps = kFThreads / 32;
constexpr int kFR = kFThreads * kVec;                             // rows per tile (1024)
constexpr int kFCtasPerSm = 512 / kFThreads;
constexpr int kFMaskWords = kFR / 32;
constexpr int kFMaskSlots = 8;                                    // mask / count ring: count runs at most 6 tiles ahead of write
constexpr int kFCtlSlots = 4;                                     // base / mode ring: control runs at most 3 tiles ahead of write
constexpr int kFMaxAhead = 3;
constexpr int kFMaxPredStages = 8;
constexpr int kFMaxPayStages = 4;
constexpr int kFMaxPay = 8;                                       // distinct projected / aggregated columns
constexpr int kFPendCap = kFR;                                    // survivors of sparse tiles waiting for one batched gather
constexpr int kFSegCap = 64;                                      // ... from at most this many tiles
constexpr int kFCountBits = 12;                                   // published word = epoch << 12 | count  (count <= kFR < 4096)
constexpr int kFPredColBytes = kFR * 4;
constexpr int kFWaveRegs = 10;                                    // warp 0 holds a wave's counts in registers: grid <= 320
static_assert(kFR <= 4095 && kPadRows % kFR == 0 && kFWarps <= 32 && kFPendCap <= 65536 && kFMaxAhead + kFMaxPayStages - 1 <= kFMaskSlots - 2 &&
                  kFMaxPayStages - 1 < kFCtlSlots,
              "fused tile geometry");

struct FusedParams {
    int32_t npay;                     // distinct payload columns
    int32_t pred_stages;              // depth of the predicate ring
    int32_t pay_stages;               // depth of the payload ring: control runs pay_stages - 1 tiles ahead of write
    int32_t dense_min;                // a tile with at least this many survivors is bulk-copied whole
    uint32_t epoch;                   // tag of this launch's published counts (1 .. 2^20 - 1)
    int32_t pay_stage_bytes;          // bytes of one payload stage
    int32_t ntiles;                   // tiles of kFR rows
    int32_t ahead;                    // count runs this many tiles ahead of control (0 .. kFMaxAhead)
    uint32_t* flags;                  // [ntiles] published counts
    long long* prof;                  // optional [gridDim.x][16] cycle counters (MBC_FUSED_PROF=1), NULL otherwise
    const void* pay_src[kFMaxPay];
    int32_t pay_stride[kFMaxPay];
    int32_t pay_off[kFMaxPay];        // byte offset of the column inside a payload stage
    int8_t proj_pay[kMaxProj];        // payload column of every projected field
    int8_t agg_pay[kMaxAgg];          // payload column of every aggregate (-1: COUNT)
};

// ---- small PTX helpers --------------------------------------------------------------------------------------------
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
            : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
            : "memory");
    }
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// cycle accounting (only when f.prof is set): acc += now - t0, t0 = now
__device__ __forceinline__ void prof_lap(bool on, long long& t0, long long& acc) {
    if (on) {
        const long long t = clock64();
        acc += t - t0;
        t0 = t;
    }
}

struct FusedSeg {
    long long row0;                   // first row of the tile
    long long out0;                   // output offset of the tile minus its first index in the pending list
};

// acc = combine(acc, value) for one survivor; `raw` is the 4-byte column value, acc a raw 64-bit accumulator (int64 or
// double bits).  The (kind, type) dispatch is warp-uniform; ints are sign-extended, reals widened to double (exact).
__device__ __forceinline__ void agg_step(const DevAgg& g, unsigned long long& acc, uint32_t raw) {
    if (g.type == MBC_ATTR_INTEGER) {
        const long long x = (long long)(int32_t)raw;
        const long long a = (long long)acc;
        acc = (unsigned long long)(g.kind == MBC_AGG_SUM ? a + x : g.kind == MBC_AGG_MIN ? (x < a ? x : a) : (x > a ? x : a));
    } else {
        const double x = (double)__uint_as_float(raw);
        const double a = __longlong_as_double((long long)acc);
        acc = (unsigned long long)__double_as_longlong(g.kind == MBC_AGG_SUM ? a + x : g.kind == MBC_AGG_MIN ? (x < a ? x : a) : (x > a ? x : a));
    }
}

__global__ void __launch_bounds__(kFThreads, kFCtasPerSm) fused_scan_kernel(const __grid_constant__ ScanParams p, const __grid_constant__ FusedParams f) {
    extern __shared__ __align__(128) uint8_t ring_mem[];           // [pred_stages][nstaged][kFR] u32, then [pay_stages][pay_stage_bytes]
    __shared__ __align__(8) uint64_t b_pred_full[kFMaxPredStages];
    __shared__ __align__(8) uint64_t b_pay_full[kFMaxPayStages];
    __shared__ uint32_t s_mask[kFMaskSlots][kFMaskWords];
    __shared__ uint32_t s_cnt[kFMaskSlots];
    __shared__ long long s_base[kFCtlSlots];
    __shared__ int s_mode[kFCtlSlots];                             // >= 0: payload stage of a dense tile; -1: sparse; -2: nothing to fetch
    __shared__ uint32_t s_wcnt[2][kFWarps];
    __shared__ uint32_t s_wtot[kFWarps];
    __shared__ uint16_t s_list[kFR];
    __shared__ uint32_t s_pend[kFPendCap];
    __shared__ FusedSeg s_seg[kFSegCap];
    __shared__ unsigned long long s_aggw[kMaxAgg][kFWarps];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int G = (int)gridDim.x;
    const int b = (int)blockIdx.x;
    const int P = f.pred_stages, S = f.pay_stages;
    const int A = f.ahead;                                         // count -> control distance
    const int L = S > 0 ? S - 1 : 0;                               // control -> write distance (payload prefetch)
    const uint32_t pred_stage_bytes = (uint32_t)p.nstaged * kFPredColBytes;
    uint8_t* const pay_mem = ring_mem + (size_t)P * pred_stage_bytes;
    const int my_tiles = b < f.ntiles ? (f.ntiles - b + G - 1) / G : 0;   // tiles this CTA owns (32-bit: no 64-bit divisions in the loop)
    const uint32_t kCountMask = (1u << kFCountBits) - 1u;

    auto issue_pred = [&](int slot, long long tile) {             // one elected thread
        mbar_arrive_expect_tx(&b_pred_full[slot], pred_stage_bytes);
        for (int c = 0; c < p.nstaged; ++c)
            tma_bulk_g2s(ring_mem + (size_t)slot * pred_stage_bytes + (size_t)c * kFPredColBytes,
                         reinterpret_cast<const uint8_t*>(p.staged_src[c]) + (size_t)tile * kFPredColBytes, kFPredColBytes, &b_pred_full[slot]);
    };

    if (tid == 0) {
        for (int s = 0; s < kFMaxPredStages; ++s) mbar_init(&b_pred_full[s], 1);
        for (int s = 0; s < kFMaxPayStages; ++s) mbar_init(&b_pay_full[s], 1);
        mbar_fence_init();
        if (p.nstaged)
            for (int s = 0; s < P && s < my_tiles; ++s) issue_pred(s, (long long)b + (long long)s * G);
    }
    __syncthreads();
    if (my_tiles == 0) return;

    // ---- state of the stages --------------------------------------------------------------------------------------
    // control (warp 0): the next wave's published counts are requested one iteration before they are used
    long long wave_base = (warp == 0 && p.count_in) ? *p.count_in : 0ll;
    int ctl_slot = 0;                                              // payload stage of the next dense tile (round robin)
    uint32_t held[kFWaveRegs];
    auto request = [&](uint32_t (&v)[kFWaveRegs], int i) {
        const int wave0 = i * G;
        const int nw = min(G, f.ntiles - wave0);
#pragma unroll
        for (int x = 0; x < kFWaveRegs; ++x) {
            const int k = x * 32 + lane;
            v[x] = k < nw ? ld_relaxed_gpu(f.flags + wave0 + k) : 0u;
        }
    };
#pragma unroll
    for (int x = 0; x < kFWaveRegs; ++x) held[x] = 0u;
    // write: payload stage counter, pending sparse batch, aggregate registers
    int wr_slot = 0;                                               // payload stage of the next dense tile to write, and the
    uint32_t wr_phase = 0;                                         // parity of its mbarrier phase
    int pred_slot = 0;                                             // predicate stage of the next tile to count, and the
    uint32_t pred_phase = 0;                                       // parity of its mbarrier phase
    int dense_written = 0;
    int npend = 0, nseg = 0;
    long long my_rows = 0;                                         // survivors this thread has written (COUNT)
    unsigned long long agg[kMaxAgg];
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) agg[a] = a < p.nagg ? agg_identity(p.aggs[a]) : 0ull;

    const bool prof = f.prof != nullptr && tid == 0;
    long long pt = prof ? clock64() : 0, pc_pred = 0, pc_count = 0, pc_flags = 0, pc_ctl = 0, pc_rank = 0, pc_payfull = 0, pc_dense = 0,
              pc_sparse = 0, pc_flush = 0;
    const long long pstart = pt;

    auto flush = [&]() {                                           // batched gather of the pending sparse survivors
        __syncthreads();                                           // list + segments are in place
        const int n = npend;
        for (int k0 = tid; k0 < n; k0 += kFThreads * kGatherBatch) {
            long long row[kGatherBatch], out[kGatherBatch];
#pragma unroll
            for (int x = 0; x < kGatherBatch; ++x) {
                const int k = k0 + x * kFThreads;
                row[x] = out[x] = -1;
                if (k < n) {
                    const uint32_t e = s_pend[k];
                    const FusedSeg sg = s_seg[e >> 16];
                    row[x] = sg.row0 + (e & 0xFFFFu);
                    out[x] = sg.out0 + k;
                    ++my_rows;
                    if (p.out_pos) p.out_pos[out[x]] = p.pos_base + row[x];
                }
            }
            for (int c = 0; c < p.nproj; ++c) {                    // iterator/Projection.java:103-144
                const DevProj& pr = p.proj[c];
                if (pr.stride == 4) {
                    uint32_t v[kGatherBatch];
#pragma unroll
                    for (int x = 0; x < kGatherBatch; ++x)
                        if (row[x] >= 0) v[x] = __ldg(reinterpret_cast<const uint32_t*>(pr.src) + row[x]);
#pragma unroll
                    for (int x = 0; x < kGatherBatch; ++x)
                        if (row[x] >= 0) reinterpret_cast<uint32_t*>(pr.dst)[out[x]] = v[x];
                } else if (pr.stride == 16) {
                    uint4 v[kGatherBatch];
#pragma unroll
                    for (int x = 0; x < kGatherBatch; ++x)
                        if (row[x] >= 0) v[x] = ldg128(reinterpret_cast<const uint4*>(pr.src) + row[x]);
#pragma unroll
                    for (int x = 0; x < kGatherBatch; ++x)
                        if (row[x] >= 0) reinterpret_cast<uint4*>(pr.dst)[out[x]] = v[x];
                } else {
                    const int words = pr.stride >> 2;
#pragma unroll
                    for (int x = 0; x < kGatherBatch; ++x) {
                        if (row[x] < 0) continue;
                        const uint32_t* src = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(pr.src) + row[x] * pr.stride);
                        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pr.dst) + out[x] * pr.stride);
                        for (int w = 0; w < words; ++w) dst[w] = __ldg(src + w);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < kMaxAgg; ++a) {
                if (a >= p.nagg) break;
                const DevAgg& g = p.aggs[a];
                if (g.kind == MBC_AGG_COUNT) continue;
                uint32_t v[kGatherBatch];
#pragma unroll
                for (int x = 0; x < kGatherBatch; ++x)
                    if (row[x] >= 0) v[x] = __ldg(reinterpret_cast<const uint32_t*>(g.src) + row[x]);
#pragma unroll
                for (int x = 0; x < kGatherBatch; ++x)
                    if (row[x] >= 0) agg_step(g, agg[a], v[x]);
            }
        }
        npend = 0;                                                 // the list and the segments are next written after the next
        nseg = 0;                                                  // iteration's barriers, which every thread reaches only after
    };                                                             // it has left this function

    const int niter = my_tiles + A + L;
    for (int it = 0; it < niter; ++it) {
        // ================= COUNT: tile `it` =========================================================================
        if (it < my_tiles) {
            const int tile = b + it * G;
            const int ps = pred_slot;
            const int ms = it % kFMaskSlots;
            prof_lap(prof, pt, pc_dense);
            if (p.nstaged) mbar_wait_sleep(&b_pred_full[ps], pred_phase);
            prof_lap(prof, pt, pc_pred);
            const uint32_t* stage = reinterpret_cast<const uint32_t*>(ring_mem + (size_t)ps * pred_stage_bytes);
            const int64_t warp_row0 = (int64_t)tile * kFR + warp * kUnitRows;
            const int64_t thread_row0 = warp_row0 + lane * kVec;
            const int tile_off = warp * kUnitRows + lane * kVec;

            uint32_t mask = 0xFu;
            if (p.sel_bitmap) mask &= load_bits<1>(p.sel_bitmap, warp_row0, lane);
            if (p.nterms > 0) {
                uint32_t acc = 0;
                for (int k = 0; k < p.nterms; ++k) {               // warp-uniform term program (PredEval.java:25-183)
                    const DevTerm& t = p.terms[k];
                    acc |= (t.cmp_type == MBC_ATTR_STRING) ? eval_term_str<1>(t, warp_row0, lane) : eval_term32<1>(t, thread_row0, stage, tile_off, kFR);
                    if (t.end_conj) { mask &= acc; acc = 0; }
                }
            }
            if (p.deleted) mask &= ~load_bits<1>(p.deleted, warp_row0, lane);   // TupleScan.java:85
            if (warp_row0 + kUnitRows > p.nrows) {
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    if (thread_row0 + j >= p.nrows) mask &= ~(1u << j);
            }
            {                                                      // java.util.BitSet order: bit p = word p/32, bit p%32
                uint32_t w = mask << ((lane & 7) * 4);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
                if ((lane & 7) == 0) {
                    p.out_bitmap[(warp_row0 >> 5) + (lane >> 3)] = w;
                    s_mask[ms][warp * (kUnitRows / 32) + (lane >> 3)] = w;
                }
            }
            const uint32_t wcnt = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(mask));
            if (lane == 0) s_wcnt[it & 1][warp] = wcnt;
            prof_lap(prof, pt, pc_count);
        }
        __syncthreads();   // (A) the predicate stage is free; masks and warp counts are in place; last iteration's write is complete
        if (it < my_tiles && tid == kFThreads - 1) {               // the last warp publishes (warp 0 is busy with control)
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < kFWarps; ++w) c += s_wcnt[it & 1][w];
            s_cnt[it % kFMaskSlots] = c;
            st_relaxed_gpu(f.flags + b + it * G, (f.epoch << kFCountBits) | c);
            const int next = it + P;
            if (p.nstaged && next < my_tiles) issue_pred(pred_slot, (long long)b + (long long)next * G);
        }
        if (it < my_tiles && ++pred_slot == P) { pred_slot = 0; pred_phase ^= 1u; }
        // ================= CONTROL: tile `it - A` (warp 0) ===========================================================
        const int ic = it - A;
        if (warp == 0 && ic >= 0 && ic < my_tiles) {
            const long long tile = (long long)b + (long long)ic * G;
            const int cs = ic % kFCtlSlots;
            if (ic == 0) request(held, 0);
            const int wave0 = ic * G;
            const int nw = min(G, f.ntiles - wave0);
            uint32_t before = 0, total = 0, c = 0;
#pragma unroll
            for (int x = 0; x < kFWaveRegs; ++x) {
                const int k = x * 32 + lane;
                if (k < nw) {
                    uint32_t v = held[x];
                    while ((v >> kFCountBits) != f.epoch) {        // not published yet: read it again
                        __nanosleep(20);
                        v = ld_relaxed_gpu(f.flags + wave0 + k);
                    }
                    v &= kCountMask;
                    total += v;
                    if (k < b) before += v;
                    if (k == b) c = v;
                }
            }
            before = __reduce_add_sync(0xFFFFFFFFu, before);
            total = __reduce_add_sync(0xFFFFFFFFu, total);
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            prof_lap(prof, pt, pc_flags);
            const bool dense = c > 0 && f.npay > 0 && (int)c >= f.dense_min;
            if (lane == 0) {
                int mode = -2;
                if (dense) {
                    // the stage was last used by the dense tile S issues ago, written at least one iteration (and one
                    // barrier) ago: control runs L = S - 1 tiles ahead of write
                    const int slot = ctl_slot;
                    mbar_arrive_expect_tx(&b_pay_full[slot], (uint32_t)f.pay_stage_bytes);
                    uint8_t* dst = pay_mem + (size_t)slot * f.pay_stage_bytes;
                    for (int k = 0; k < f.npay; ++k) {
                        const uint32_t bytes = (uint32_t)f.pay_stride[k] * kFR;
                        tma_bulk_g2s(dst + f.pay_off[k], reinterpret_cast<const uint8_t*>(f.pay_src[k]) + (size_t)tile * bytes, bytes, &b_pay_full[slot]);
                    }
                    mode = slot;
                } else if (c > 0 && f.npay > 0) {
                    mode = -1;
                }
                s_mode[cs] = mode;
                s_base[cs] = wave_base + before;
            }
            if (dense && ++ctl_slot == S) ctl_slot = 0;
            wave_base += total;
            if (ic + 1 < my_tiles) request(held, ic + 1);          // in flight during the write stage, used in the next iteration
            prof_lap(prof, pt, pc_ctl);
        }
        // ================= WRITE: tile `it - A - L` ===================================================================
        const int iw = it - A - L;
        if (iw < 0 || iw >= my_tiles) continue;
        if (L == 0) __syncthreads();                               // control of this very tile was written just above
        const long long tile = (long long)b + (long long)iw * G;
        const int ms = iw % kFMaskSlots;
        const int cs = iw % kFCtlSlots;
        const int T = (int)s_cnt[ms];
        if (T == 0) continue;                                      // CTA-uniform: nothing qualifies in this tile
        const long long base = s_base[cs];
        const int mode = s_mode[cs];
        const uint32_t bits = (s_mask[ms][tid >> 3] >> ((tid & 7) * 4)) & 0xFu;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_wtot[warp] = (uint32_t)incl;
        __syncthreads();   // (B)
        uint32_t below = (lane < warp) ? s_wtot[lane] : 0u;
        below = __reduce_add_sync(0xFFFFFFFFu, below);
        int r = (int)below + incl - cnt;
        const int64_t tile_row0 = (int64_t)tile * kFR;
        prof_lap(prof, pt, pc_rank);

        if (mode == -1) {
            // sparse tile: its survivors join the pending list, fetched later in one batch
            if (tid == 0) s_seg[nseg] = FusedSeg{(long long)tile_row0, base - (long long)npend};
            uint32_t bb = bits;
            while (bb) {
                const int j = __ffs(bb) - 1;
                bb &= bb - 1;
                s_pend[npend + r++] = ((uint32_t)nseg << 16) | (uint32_t)(tid * kVec + j);
            }
            npend += T;
            ++nseg;
            prof_lap(prof, pt, pc_sparse);
            if (npend + f.dense_min > kFPendCap || nseg == kFSegCap) {
                flush();
                prof_lap(prof, pt, pc_flush);
            }
            continue;
        }
        // dense tile (or nothing to fetch): rank -> row list, then one thread per survivor
        {
            uint32_t bb = bits;
            while (bb) {
                const int j = __ffs(bb) - 1;
                bb &= bb - 1;
                s_list[r++] = (uint16_t)(tid * kVec + j);
            }
        }
        __syncthreads();   // (C)
        const uint8_t* slot = nullptr;
        if (mode >= 0) {
            mbar_wait_sleep(&b_pay_full[mode], wr_phase);    // mode == wr_slot: control hands the stages out round robin too
            slot = pay_mem + (size_t)mode * f.pay_stage_bytes;
            ++dense_written;
            if (++wr_slot == S) { wr_slot = 0; wr_phase ^= 1u; }
        }
        prof_lap(prof, pt, pc_payfull);
        // this thread's survivors: ranks tid, tid + NT, ... (at most 4: T <= kFR = 4 NT)
        int row[kVec];
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const int k = tid + j * kFThreads;
            row[j] = k < T ? (int)s_list[k] : -1;
            my_rows += k < T ? 1 : 0;
        }
        if (p.out_pos) {
#pragma unroll
            for (int j = 0; j < kVec; ++j)
                if (row[j] >= 0) p.out_pos[base + tid + j * kFThreads] = p.pos_base + tile_row0 + row[j];
        }
        if (mode >= 0) {
            for (int c = 0; c < p.nproj; ++c) {                    // iterator/Projection.java:103-144
                const DevProj& pr = p.proj[c];
                const uint8_t* src = slot + f.pay_off[f.proj_pay[c]];
                if (pr.stride == 4) {
                    uint32_t v[kVec];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) v[j] = reinterpret_cast<const uint32_t*>(src)[row[j]];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) (reinterpret_cast<uint32_t*>(pr.dst) + base)[tid + j * kFThreads] = v[j];
                } else if (pr.stride == 16) {
                    uint4 v[kVec];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) v[j] = reinterpret_cast<const uint4*>(src)[row[j]];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) (reinterpret_cast<uint4*>(pr.dst) + base)[tid + j * kFThreads] = v[j];
                } else {
                    const int words = pr.stride >> 2;
#pragma unroll
                    for (int j = 0; j < kVec; ++j) {
                        if (row[j] < 0) continue;
                        const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)row[j] * pr.stride);
                        uint32_t* d = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pr.dst) + (base + tid + j * kFThreads) * pr.stride);
                        for (int w = 0; w < words; ++w) d[w] = s[w];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < kMaxAgg; ++a) {
                if (a >= p.nagg) break;
                const DevAgg& g = p.aggs[a];
                if (g.kind == MBC_AGG_COUNT) continue;
                const uint32_t* src = reinterpret_cast<const uint32_t*>(slot + f.pay_off[f.agg_pay[a]]);
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    if (row[j] >= 0) agg_step(g, agg[a], src[row[j]]);
            }
        }
        prof_lap(prof, pt, pc_dense);
        // the next iteration's barrier (A) separates these reads of the list and of the payload stage from their reuse
    }
    if (npend) flush();

    // ---- the CTA's aggregate partial: lanes butterflied, warps combined in order --------------------------------------
    if (p.nagg > 0) {
#pragma unroll
        for (int a = 0; a < kMaxAgg; ++a) {
            if (a >= p.nagg) break;
            const DevAgg& g = p.aggs[a];
            unsigned long long v = g.kind == MBC_AGG_COUNT ? (unsigned long long)my_rows : agg[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = agg_merge(g, v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
            if (lane == 0) s_aggw[a][warp] = v;
        }
        __syncthreads();
        if (tid < p.nagg) {
            const DevAgg& g = p.aggs[tid];
            unsigned long long v = s_aggw[tid][0];
            for (int x = 1; x < kFWarps; ++x) v = agg_merge(g, v, s_aggw[tid][x]);
            p.partials[(size_t)tid * p.total_tiles + p.tile_base + b] = v;
        }
    }
    if (b == 0 && tid == 0) *p.count_out = wave_base;              // CTA 0 owns a tile of every wave: its warp 0 has seen every count
    if (prof) {
        prof_lap(prof, pt, pc_dense);
        long long* o = f.prof + (size_t)b * 16;
        o[0] = pt - pstart; o[1] = pc_pred; o[2] = pc_count; o[3] = pc_flags; o[4] = pc_ctl; o[5] = pc_rank; o[6] = pc_payfull; o[7] = pc_dense;
        o[8] = pc_sparse; o[9] = pc_flush; o[10] = dense_written;
    }
}

}  // namespace mbc
