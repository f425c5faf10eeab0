// mbc_scan_fused.cuh -- K2+K5 in ONE residency: every column of the table is read from HBM once.
//
// Replaces the same reference loop as mbc_scan_kernels.cuh (iterator/ColumnarFileScan.java:156-172:
// while (scan.getNext) if (PredEval.Eval) Project), for scans whose projected / aggregated columns fit a
// shared-memory tile ring.  The two-pass engine (filter_kernel -> tile_offsets_kernel -> write_kernel) streams
// the predicate columns in pass 1 and then fetches the survivors' values with scattered 4..16-byte gathers, so a
// dense scan re-reads the predicate columns and pulls 64..128-byte DRAM lines around every gathered value
// (measured in round 1: 7.98 GB of DRAM traffic for 3.94 GB of useful bytes in the write pass).  Here:
//
//   * one persistent CTA per SM (cooperative launch: the CTAs exchange tile counts, so they must be co-resident),
//     tiles of kFR rows assigned statically: CTA b owns tiles b, b + G, b + 2G, ...   ("wave" i = tiles [iG, iG + G));
//   * three ROLES per CTA, decoupled by shared-memory rings and mbarriers, so that none of the long latencies of the scan
//     (HBM, and the ~2 us an L2 round trip takes while HBM is saturated) sits on another role's critical path:
//       COUNT warps   run AHEAD: they stream the predicate columns through a TMA ring (cp.async.bulk + mbarrier), evaluate
//                     the CNF (the term program of filter_kernel), leave the tile's selection words in a mask ring (and in
//                     the result bitmap), and PUBLISH the tile's count in global memory, tagged with the launch's epoch;
//       CONTROL warp  turns the published counts into the tile's output offset WITHOUT a chain between CTAs: the counts of
//                     a whole wave (G words) are bulk-copied into shared memory a few waves ahead of their use, the warp
//                     sums them itself and keeps the total of all earlier waves in a register; it then decides how the
//                     survivors' values are fetched: a DENSE tile gets its projected columns bulk-copied whole into the
//                     payload ring (the predicate columns among them come out of L2: they were streamed a few microseconds
//                     earlier), a SPARSE tile is gathered later;
//       WRITE warps   rank the survivors from the mask ring, build the rank -> row list, and compact the tile OUT OF SHARED
//                     MEMORY with one thread per SURVIVOR: coalesced stores in ascending position order; sparse tiles join a
//                     pending list that is gathered from global memory a few hundred survivors at a time, so the gather
//                     latency is paid once per batch and not once per tile;
//   * COUNT/SUM/MIN/MAX are folded into per-thread registers across all the tiles of the CTA and combined once, in a fixed
//     order, when the CTA finishes: one partial per CTA (reproducible run to run: the tile assignment is static).
//
// Output order is position order by construction (offsets, not atomics): bit-exact position lists.
#pragma once
#include "mbc_scan_kernels.cuh"

namespace mbc {

#ifndef MBC_FUSED_COUNT_GROUPS
#define MBC_FUSED_COUNT_GROUPS 1
#endif
constexpr int kFR = 2048;                                         // rows per tile
constexpr int kFGroupWarps = kFR / kWarpRows;                     // count warps per group: a warp owns 512 rows (16 per thread)
constexpr int kFGroups = MBC_FUSED_COUNT_GROUPS;                  // count groups take the CTA's tiles round robin
constexpr int kFGroupThreads = kFGroupWarps * 32;
constexpr int kFCountWarps = kFGroups * kFGroupWarps;
constexpr int kFWriteWarps = 8;
constexpr int kFWriteThreads = kFWriteWarps * 32;
constexpr int kFThreads = 32 * (1 + kFCountWarps + kFWriteWarps); // warp 0 = control
constexpr int kFRowsPerWriter = kFR / kFWriteThreads;             // consecutive rows ranked by one write thread (8)
constexpr int kFMaskSlots = 16;                                   // how far the count warps may run ahead of the write warps (tiles)
constexpr int kFMaskWords = kFR / 32;
constexpr int kFMaxPredStages = 8;
constexpr int kFMaxPayStages = 4;
constexpr int kFMaxPay = 8;                                       // distinct projected / aggregated columns
constexpr int kFPendCap = 2048;                                   // survivors of sparse tiles waiting for one batched gather
constexpr int kFSegCap = 128;                                     // ... from at most this many tiles
constexpr int kFCountBits = 12;                                   // published word = epoch << 12 | count  (count <= kFR < 4096)
constexpr int kFPredColBytes = kFR * 4;
constexpr int kFFlagStages = 4;                                   // waves of published counts in flight towards shared memory
constexpr int kFMaxGrid = 256;                                    // a wave's counts fit one flag stage
static_assert(kFR % kWarpRows == 0 && kFR <= 4095 && kFRowsPerWriter == 8 && kPadRows % kFR == 0 && kFFlagStages < kFMaskSlots, "fused tile geometry");

struct FusedParams {
    int32_t npay;                     // distinct payload columns
    int32_t pred_stages;              // depth of the predicate ring
    int32_t pay_stages;               // depth of the payload ring
    int32_t dense_min;                // a tile with at least this many survivors is bulk-copied whole
    uint32_t epoch;                   // tag of this launch's published counts (1 .. 2^20 - 1)
    int32_t pay_stage_bytes;          // bytes of one payload stage
    int32_t ntiles;                   // tiles of kFR rows
    int32_t pad;
    uint32_t* flags;                  // [ntiles, padded by one wave] published counts
    long long* prof;                  // optional [gridDim.x][24] cycle counters (MBC_FUSED_PROF=1), NULL otherwise
    long long* dbg;                   // debug builds: host-mapped record of the first failed check
    const void* pay_src[kFMaxPay];
    int32_t pay_stride[kFMaxPay];
    int32_t pay_off[kFMaxPay];        // byte offset of the column inside a payload stage
    int8_t proj_pay[kMaxProj];        // payload column of every projected field
    int8_t agg_pay[kMaxAgg];          // payload column of every aggregate (-1: COUNT)
};

// ---- small PTX helpers --------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
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

#ifdef MBC_FUSED_DEBUG
// debug builds: a failed check leaves (line, cta, thread, four values) in a host-mapped buffer and traps -- device printf
// output is lost when the context dies, host memory is not
#define MBC_FCHECK(cond, a0, a1, a2, a3)                                                                              \
    do {                                                                                                              \
        if (!(cond) && f.dbg) {                                                                                       \
            if (atomicCAS(reinterpret_cast<unsigned long long*>(f.dbg), 0ull, (unsigned long long)__LINE__) == 0ull) { \
                f.dbg[1] = blockIdx.x; f.dbg[2] = threadIdx.x; f.dbg[3] = (long long)(a0); f.dbg[4] = (long long)(a1);  \
                f.dbg[5] = (long long)(a2); f.dbg[6] = (long long)(a3);                                               \
                __threadfence_system();                                                                               \
            }                                                                                                         \
            __trap();                                                                                                 \
        }                                                                                                             \
    } while (0)
#else
#define MBC_FCHECK(cond, a0, a1, a2, a3) do { } while (0)
#endif

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

__global__ void __launch_bounds__(kFThreads, 1) fused_scan_kernel(const __grid_constant__ ScanParams p, const __grid_constant__ FusedParams f) {
    extern __shared__ __align__(128) uint8_t ring_mem[];           // [pred_stages][nstaged][kFR] u32, then [pay_stages][pay_stage_bytes]
    __shared__ __align__(8) uint64_t b_pred_full[kFMaxPredStages];
    __shared__ __align__(8) uint64_t b_pay_full[kFMaxPayStages], b_pay_free[kFMaxPayStages];
    __shared__ __align__(8) uint64_t b_mask_ready[kFMaskSlots], b_mask_free[kFMaskSlots], b_ctl_ready[kFMaskSlots];
    __shared__ __align__(8) uint64_t b_flag_full[kFFlagStages];
    __shared__ __align__(16) uint32_t s_flags[kFFlagStages][kFMaxGrid];
    __shared__ uint32_t s_mask[kFMaskSlots][kFMaskWords];
    __shared__ uint32_t s_cnt[kFMaskSlots];
    __shared__ long long s_base[kFMaskSlots];
    __shared__ int s_mode[kFMaskSlots];                            // >= 0: payload stage of a dense tile; -1: sparse; -2: nothing to fetch
    __shared__ uint32_t s_cwcnt[kFGroups][2][kFGroupWarps];
    __shared__ uint16_t s_list[kFR];
    __shared__ uint32_t s_pend[kFPendCap];
    __shared__ FusedSeg s_seg[kFSegCap];
    __shared__ uint32_t s_wtot[2][kFWriteWarps];
    __shared__ unsigned long long s_aggw[kMaxAgg][kFWriteWarps];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int G = (int)gridDim.x;
    const int b = (int)blockIdx.x;
    const int P = f.pred_stages, S = f.pay_stages;
    const uint32_t pred_stage_bytes = (uint32_t)p.nstaged * kFPredColBytes;
    uint8_t* const pay_mem = ring_mem + (size_t)P * pred_stage_bytes;
    const int my_tiles = b < f.ntiles ? (f.ntiles - b + G - 1) / G : 0;   // tiles this CTA owns (32-bit: no 64-bit divisions in the loops)
    // a wave's counts as a 16-byte multiple: G is a multiple of 4 (host) unless the whole table is one wave of < 4 tiles; the
    // flag buffer is padded by one wave, so the copy of the last (ragged) wave stays in bounds
    const uint32_t flag_bytes = (uint32_t)((G + 3) / 4 * 4) * 4u;

    auto issue_pred = [&](int slot, long long tile) {             // one elected thread
        mbar_arrive_expect_tx(&b_pred_full[slot], pred_stage_bytes);
        for (int c = 0; c < p.nstaged; ++c)
            tma_bulk_g2s(ring_mem + (size_t)slot * pred_stage_bytes + (size_t)c * kFPredColBytes,
                         reinterpret_cast<const uint8_t*>(p.staged_src[c]) + (size_t)tile * kFPredColBytes, kFPredColBytes, &b_pred_full[slot]);
    };
    auto issue_flags = [&](int slot, int wave) {                  // one elected thread: the wave's published counts -> shared memory
        mbar_arrive_expect_tx(&b_flag_full[slot], flag_bytes);
        tma_bulk_g2s(&s_flags[slot][0], f.flags + (size_t)wave * G, flag_bytes, &b_flag_full[slot]);
    };

    if (tid == 0) {
        for (int s = 0; s < kFMaxPredStages; ++s) mbar_init(&b_pred_full[s], 1);
        for (int s = 0; s < kFMaxPayStages; ++s) { mbar_init(&b_pay_full[s], 1); mbar_init(&b_pay_free[s], 1); }
        for (int s = 0; s < kFMaskSlots; ++s) { mbar_init(&b_mask_ready[s], 1); mbar_init(&b_mask_free[s], 1); mbar_init(&b_ctl_ready[s], 1); }
        for (int s = 0; s < kFFlagStages; ++s) mbar_init(&b_flag_full[s], 1);
        mbar_fence_init();
        if (p.nstaged)
            for (int s = 0; s < P && s < my_tiles; ++s) issue_pred(s, (long long)b + (long long)s * G);
    }
    __syncthreads();
    if (my_tiles == 0) return;

    if (warp == 0) {
        // ---- CONTROL: output offsets from the published counts, payload copies of dense tiles -------------------------
        // The counts of wave i are fetched into the flag ring when wave i - kFFlagStages is consumed, i.e. several tile-times
        // before they are needed; the count warps run further ahead than that (the mask ring is deeper), so the words are
        // normally published by then, and a word whose tag is still old is simply read again from global memory.
        long long wave_base = p.count_in ? *p.count_in : 0ll;
        const uint32_t kCountMask = (1u << kFCountBits) - 1u;
        const bool prof = f.prof != nullptr && lane == 0;
        long long pt = prof ? clock64() : 0, pc_flags = 0, pc_stale = 0, pc_payfree = 0, pc_rest = 0;
        const long long pstart = pt;
        int fslot = 0, pslot = 0, ms = 0;
        uint32_t fphase = 0, pfree_phase = 1;                      // a fresh "free" barrier passes a wait on parity 1
        int nstale = 0, ndense = 0;
        // the first waves' counts are requested when the count warps of THIS CTA have published theirs (a cheap proxy for
        // "the wave is probably published"); wave i + kFFlagStages is requested when wave i has been consumed
        if (lane == 0)
            for (int s = 0; s < kFFlagStages && s < my_tiles; ++s) {
                mbar_wait_sleep(&b_mask_ready[s], 0);
                issue_flags(s, s);
            }
        __syncwarp();
        for (int i = 0; i < my_tiles; ++i) {
            prof_lap(prof, pt, pc_rest);
            mbar_wait_sleep(&b_flag_full[fslot], fphase);
            prof_lap(prof, pt, pc_flags);
            const int wave0 = i * G;
            const int nw = min(G, f.ntiles - wave0);
            uint32_t before = 0, total = 0, c = 0;
            bool stale = false;
            for (int k = lane; k < nw; k += 32) {
                uint32_t v = s_flags[fslot][k];
                if ((v >> kFCountBits) != f.epoch) {
                    stale = true;
                    do {                                           // not published when the copy ran: read it from global memory
                        __nanosleep(100);
                        v = ld_relaxed_gpu(f.flags + wave0 + k);
                    } while ((v >> kFCountBits) != f.epoch);
                }
                v &= kCountMask;
                total += v;
                if (k < b) before += v;
                if (k == b) c = v;
            }
            before = __reduce_add_sync(0xFFFFFFFFu, before);
            total = __reduce_add_sync(0xFFFFFFFFu, total);
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            nstale += __any_sync(0xFFFFFFFFu, stale) ? 1 : 0;       // every lane has read the stage: it may be refilled
            prof_lap(prof, pt, pc_stale);
            MBC_FCHECK(c <= (uint32_t)kFR && total <= (uint32_t)kFR * (uint32_t)G && wave_base + before + c <= p.out_cap, i, c, total, wave_base + before);
            const bool dense = c > 0 && f.npay > 0 && (int)c >= f.dense_min;
            if (lane == 0) {
                if (i + kFFlagStages < my_tiles) issue_flags(fslot, i + kFFlagStages);
                int mode = -2;
                if (dense) {
                    mbar_wait_sleep(&b_pay_free[pslot], pfree_phase);
                    prof_lap(prof, pt, pc_payfree);
                    mbar_arrive_expect_tx(&b_pay_full[pslot], (uint32_t)f.pay_stage_bytes);
                    uint8_t* dst = pay_mem + (size_t)pslot * f.pay_stage_bytes;
                    const long long tile = (long long)b + (long long)i * G;
                    for (int k = 0; k < f.npay; ++k) {
                        const uint32_t bytes = (uint32_t)f.pay_stride[k] * kFR;
                        tma_bulk_g2s(dst + f.pay_off[k], reinterpret_cast<const uint8_t*>(f.pay_src[k]) + (size_t)tile * bytes, bytes, &b_pay_full[pslot]);
                    }
                    mode = pslot;
                } else if (c > 0 && f.npay > 0) {
                    mode = -1;
                }
                // the mask slot's s_mode / s_base are free: this CTA's own count of tile i is published only after its count
                // warps waited for the write warps to release the slot
                s_mode[ms] = mode;
                s_base[ms] = wave_base + before;
                mbar_arrive(&b_ctl_ready[ms]);
            }
            if (dense) {
                ++ndense;
                if (++pslot == S) { pslot = 0; pfree_phase ^= 1u; }
            }
            wave_base += total;
            if (++fslot == kFFlagStages) { fslot = 0; fphase ^= 1u; }
            if (++ms == kFMaskSlots) ms = 0;
            __syncwarp();
        }
        if (b == 0 && lane == 0) *p.count_out = wave_base;         // CTA 0 owns a tile of every wave: it has seen every count
        if (prof) {
            prof_lap(prof, pt, pc_rest);
            long long* o = f.prof + (size_t)b * 24;
            o[0] = pt - pstart; o[1] = pc_flags; o[2] = pc_stale; o[3] = pc_payfree; o[4] = ndense; o[5] = nstale;
        }
        return;
    }

    if (warp <= kFCountWarps) {
        // ---- COUNT: CNF over the predicate ring -> mask ring + result bitmap + published count -----------------------
        const int g = (warp - 1) / kFGroupWarps;                   // group
        const int cw = (warp - 1) % kFGroupWarps;                  // warp within the group
        const int ct = cw * 32 + lane;
        const bool prof = f.prof != nullptr && ct == 0 && g == 0;
        long long pt = prof ? clock64() : 0, pc_maskfree = 0, pc_pred = 0, pc_eval = 0;
        const long long pstart = pt;
        // ring positions of this group's tiles g, g + kFGroups, ...: advanced by kFGroups per iteration
        int ps = g % P, ms = g % kFMaskSlots;
        uint32_t pphase = (uint32_t)(g / P) & 1u, mphase = (uint32_t)(g / kFMaskSlots) & 1u;
        int flip = 0;
        for (int i = g; i < my_tiles; i += kFGroups) {
            const long long tile = (long long)b + (long long)i * G;
            prof_lap(prof, pt, pc_eval);
            mbar_wait_sleep(&b_mask_free[ms], mphase ^ 1u);        // the write warps are done with the tile that used this slot
            prof_lap(prof, pt, pc_maskfree);
            if (p.nstaged) mbar_wait_sleep(&b_pred_full[ps], pphase);
            prof_lap(prof, pt, pc_pred);
            const uint32_t* stage = reinterpret_cast<const uint32_t*>(ring_mem + (size_t)ps * pred_stage_bytes);
            const int64_t warp_row0 = (int64_t)tile * kFR + cw * kWarpRows;
            const int64_t thread_row0 = warp_row0 + lane * kVec;
            const int tile_off = cw * kWarpRows + lane * kVec;

            uint32_t mask = 0xFFFFu;
            if (p.sel_bitmap) mask &= load_bits(p.sel_bitmap, warp_row0, lane);
            if (p.nterms > 0) {
                uint32_t acc = 0;
                for (int k = 0; k < p.nterms; ++k) {               // warp-uniform term program (PredEval.java:25-183)
                    const DevTerm& t = p.terms[k];
                    acc |= (t.cmp_type == MBC_ATTR_STRING) ? eval_term_str(t, warp_row0, lane) : eval_term32(t, thread_row0, stage, tile_off, kFR);
                    if (t.end_conj) { mask &= acc; acc = 0; }
                }
            }
            if (p.deleted) mask &= ~load_bits(p.deleted, warp_row0, lane);   // TupleScan.java:85
            if (warp_row0 + kWarpRows > p.nrows) {
#pragma unroll
                for (int u = 0; u < kUnits; ++u)
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (thread_row0 + u * kUnitRows + j >= p.nrows) mask &= ~(1u << (u * 4 + j));
            }
#pragma unroll
            for (int u = 0; u < kUnits; ++u) {                     // java.util.BitSet order: bit p = word p/32, bit p%32
                uint32_t w = ((mask >> (u * 4)) & 0xFu) << ((lane & 7) * 4);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
                if ((lane & 7) == 0) {
                    const int wi = u * (kUnitRows / 32) + (lane >> 3);
                    p.out_bitmap[(warp_row0 >> 5) + wi] = w;
                    s_mask[ms][cw * (kWarpRows / 32) + wi] = w;
                }
            }
            const uint32_t wcnt = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(mask));
            if (lane == 0) s_cwcnt[g][flip][cw] = wcnt;
            named_bar(1 + g, kFGroupThreads);                      // the group is done with the predicate stage; masks are in place
            if (ct == 0) {
                uint32_t c = 0;
#pragma unroll
                for (int w = 0; w < kFGroupWarps; ++w) c += s_cwcnt[g][flip][w];
                s_cnt[ms] = c;
                st_relaxed_gpu(f.flags + tile, (f.epoch << kFCountBits) | c);
                mbar_arrive(&b_mask_ready[ms]);
                const int next = i + P;
                if (p.nstaged && next < my_tiles) issue_pred(ps, (long long)b + (long long)next * G);
            }
            // the other warps run ahead; s_cwcnt[g][flip] is rewritten two tiles of this group on, with a barrier in between
            flip ^= 1;
            for (int x = 0; x < kFGroups; ++x) {                   // advance the ring positions by kFGroups tiles
                if (++ps == P) { ps = 0; pphase ^= 1u; }
                if (++ms == kFMaskSlots) { ms = 0; mphase ^= 1u; }
            }
        }
        if (prof) {
            prof_lap(prof, pt, pc_eval);
            long long* o = f.prof + (size_t)b * 24;
            o[8] = pt - pstart; o[9] = pc_maskfree; o[10] = pc_pred; o[11] = pc_eval;
        }
        return;
    }

    // ---- WRITE: rank -> compact out of shared memory -> ordered, coalesced stores; aggregates -------------------------
    const int wt = tid - 32 * (1 + kFCountWarps);
    const int ww = wt >> 5;
    constexpr int kWriteBar = 1 + kFGroups;
    int npend = 0, nseg = 0;
    long long my_rows = 0;                                         // survivors this thread has written (COUNT)
    unsigned long long agg[kMaxAgg];
#pragma unroll
    for (int a = 0; a < kMaxAgg; ++a) agg[a] = a < p.nagg ? agg_identity(p.aggs[a]) : 0ull;
    const bool prof = f.prof != nullptr && wt == 0;
    long long pt = prof ? clock64() : 0, pc_maskready = 0, pc_ctl = 0, pc_rank = 0, pc_payfull = 0, pc_dense = 0, pc_sparse = 0, pc_flush = 0;
    const long long pstart = pt;

    auto flush = [&]() {                                           // batched gather of the pending sparse survivors
        named_bar(kWriteBar, kFWriteThreads);                      // list + segments are in place
        const int n = npend;
        for (int k0 = wt; k0 < n; k0 += kFWriteThreads * kGatherBatch) {
            long long row[kGatherBatch], out[kGatherBatch];
#pragma unroll
            for (int x = 0; x < kGatherBatch; ++x) {
                const int k = k0 + x * kFWriteThreads;
                row[x] = out[x] = -1;
                if (k < n) {
                    const uint32_t e = s_pend[k];
                    MBC_FCHECK((int)(e >> 16) < nseg, k, n, e >> 16, nseg);
                    const FusedSeg sg = s_seg[e >> 16];
                    row[x] = sg.row0 + (e & 0xFFFFu);
                    out[x] = sg.out0 + k;
                    MBC_FCHECK(row[x] >= 0 && row[x] < p.nrows && out[x] >= 0 && out[x] < p.out_cap, row[x], out[x], k, n);
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
        nseg = 0;                                                  // tile's rank barrier, which every write thread reaches only
    };                                                             // after it has left this function

    int ms = 0, wslot = 0;
    uint32_t mphase = 0, wphase = 0;
    int ndense = 0;
    for (int i = 0; i < my_tiles; ++i) {
        const long long tile = (long long)b + (long long)i * G;
        prof_lap(prof, pt, pc_dense);
        mbar_wait_sleep(&b_mask_ready[ms], mphase);
        prof_lap(prof, pt, pc_maskready);
        mbar_wait_sleep(&b_ctl_ready[ms], mphase);
        prof_lap(prof, pt, pc_ctl);
        const int T = (int)s_cnt[ms];
        const long long base = s_base[ms];
        const int mode = s_mode[ms];
        MBC_FCHECK(T >= 0 && T <= kFR && base >= 0 && base + T <= p.out_cap && mode >= -2 && mode < max(S, 1), tile, T, base, mode);
        const uint32_t bits = (s_mask[ms][wt >> 2] >> ((wt & 3) * 8)) & 0xFFu;   // this thread's 8 consecutive rows
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_wtot[i & 1][ww] = (uint32_t)incl;
        named_bar(kWriteBar, kFWriteThreads);                      // every write thread holds its bits, T, base and mode now
        if (wt == 0) mbar_arrive(&b_mask_free[ms]);
        if (++ms == kFMaskSlots) { ms = 0; mphase ^= 1u; }
        if (T == 0) continue;                                      // nothing qualifies in this tile
        uint32_t below = (lane < ww) ? s_wtot[i & 1][lane] : 0u;
        below = __reduce_add_sync(0xFFFFFFFFu, below);
        int r = (int)below + incl - cnt;
        const int64_t tile_row0 = (int64_t)tile * kFR;
        MBC_FCHECK(r >= 0 && r + cnt <= T && (wt != kFWriteThreads - 1 || r + cnt == T), tile, r, cnt, T);
        prof_lap(prof, pt, pc_rank);

        if (mode == -1) {
            // sparse tile: its survivors join the pending list, fetched later in one batch
            MBC_FCHECK(nseg < kFSegCap && npend + T <= kFPendCap, nseg, npend, T, 0);
            if (wt == 0) s_seg[nseg] = FusedSeg{(long long)tile_row0, base - (long long)npend};
            uint32_t bb = bits;
            while (bb) {
                const int j = __ffs(bb) - 1;
                bb &= bb - 1;
                s_pend[npend + r++] = ((uint32_t)nseg << 16) | (uint32_t)(wt * kFRowsPerWriter + j);
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
                s_list[r++] = (uint16_t)(wt * kFRowsPerWriter + j);
            }
        }
        named_bar(kWriteBar, kFWriteThreads);
        const uint8_t* slot = nullptr;
        prof_lap(prof, pt, pc_dense);
        if (mode >= 0) {
            mbar_wait_sleep(&b_pay_full[mode], wphase);            // mode == wslot: control hands the stages out round robin too
            slot = pay_mem + (size_t)mode * f.pay_stage_bytes;
        }
        prof_lap(prof, pt, pc_payfull);
        // this thread's survivors: ranks wt, wt + 256, ... in batches of 4 (at most 8: T <= kFR)
        for (int k0 = wt; k0 < T; k0 += kFWriteThreads * kVec) {
            int row[kVec];
#pragma unroll
            for (int j = 0; j < kVec; ++j) {
                const int k = k0 + j * kFWriteThreads;
                row[j] = k < T ? (int)s_list[k] : -1;
                my_rows += k < T ? 1 : 0;
            }
            if (p.out_pos) {
#pragma unroll
                for (int j = 0; j < kVec; ++j)
                    if (row[j] >= 0) p.out_pos[base + k0 + j * kFWriteThreads] = p.pos_base + tile_row0 + row[j];
            }
            if (mode < 0) continue;
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
                        if (row[j] >= 0) (reinterpret_cast<uint32_t*>(pr.dst) + base)[k0 + j * kFWriteThreads] = v[j];
                } else if (pr.stride == 16) {
                    uint4 v[kVec];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) v[j] = reinterpret_cast<const uint4*>(src)[row[j]];
#pragma unroll
                    for (int j = 0; j < kVec; ++j)
                        if (row[j] >= 0) (reinterpret_cast<uint4*>(pr.dst) + base)[k0 + j * kFWriteThreads] = v[j];
                } else {
                    const int words = pr.stride >> 2;
#pragma unroll
                    for (int j = 0; j < kVec; ++j) {
                        if (row[j] < 0) continue;
                        const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)row[j] * pr.stride);
                        uint32_t* d = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(pr.dst) + (base + k0 + j * kFWriteThreads) * pr.stride);
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
        named_bar(kWriteBar, kFWriteThreads);                      // the payload stage and the list have been read by everyone
        if (mode >= 0) {
            if (wt == 0) mbar_arrive(&b_pay_free[mode]);
            ++ndense;
            if (++wslot == S) { wslot = 0; wphase ^= 1u; }
        }
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
            if (lane == 0) s_aggw[a][ww] = v;
        }
        named_bar(kWriteBar, kFWriteThreads);
        if (wt < p.nagg) {
            const DevAgg& g = p.aggs[wt];
            unsigned long long v = s_aggw[wt][0];
            for (int x = 1; x < kFWriteWarps; ++x) v = agg_merge(g, v, s_aggw[wt][x]);
            p.partials[(size_t)wt * p.total_tiles + p.tile_base + b] = v;
        }
    }
    if (prof) {
        prof_lap(prof, pt, pc_dense);
        long long* o = f.prof + (size_t)b * 24;
        o[12] = pt - pstart; o[13] = pc_maskready; o[14] = pc_ctl; o[15] = pc_rank; o[16] = pc_payfull; o[17] = pc_dense; o[18] = pc_sparse;
        o[19] = pc_flush; o[20] = ndense;
    }
}

}  // namespace mbc
