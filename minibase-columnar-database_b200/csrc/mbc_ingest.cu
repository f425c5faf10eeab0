// mbc_ingest.cu -- K1: decode a reference-format MiniBase DB file image into device-resident columns.
//
// Replaces, for a whole Columnarfile at once, the per-row page walk of
//     heap/Scan.java:84-114 (getNext / nextDataPage), heap/HFPage.java:543-573 (getRecord),
//     global/Convert.java:18-126 (big-endian int / float / modified-UTF-8 string decode),
//     columnar/TupleScan.java:55-89 (zip of the per-column scans) and
//     heap/Heapfile.java:262-289,349-417 (position <-> RID arithmetic).
//
// On-disk format (all scalars big-endian, pages of 1024 bytes):
//   page 0 / directory chain   diskmgr/DB.java:866-871,985-1000: [next:int][nEntries:int] then entries of
//                              56 bytes {firstPage:int, name: UTF (2-byte length + bytes)}
//   heapfile                   linked list of DIRECTORY HFPages whose records are 8-byte DataPageInfo
//                              {availspace:short, recct:short, pageId:int} (heap/DataPageInfo.java:19-29)
//   HFPage                     heap/HFPage.java:31-40: {slotCnt:short@0, usedPtr@2, freeSpace@4, type@6,
//                              prev:int@8, next:int@12, cur:int@16}, slot directory from byte 20
//                              {length:short, offset:short} (length -1 = empty), records packed from 1024 down
//   <cf>.hdr                   columnar/Columnarfile.java:257-323: rec0 numCols, rec1 types, rec2 sizes, rec3 names ...
//   <cf>.<i>                   one heapfile per column, record = 4 bytes (int/real) or size+2 bytes (string)
//   <cf>.md                    deleted bitmap: BitSet.toByteArray() cut into 1000-byte records, one per chained page
//                              (bitmap/BM.java:64-129,179-215)
//   position                   recsPerDataPage * (dirPageIndex * 83 + dirSlot) + slotNo  (Heapfile.java:262-273,349-417)
//
// The host only follows the (small) directory structure; every data page is decoded on the GPU, one warp
// per page out of a shared-memory copy of the page, from the uploaded file image.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>

#include "mbc_internal.cuh"

namespace mbc {

constexpr int kPage = 1024;
constexpr int kDpFixed = 20;
constexpr int kDirRecs = (kPage - kDpFixed) / (4 + 8);     // DataPageInfo records per directory page = 83

struct PageRef {
    int32_t page_id;      // data page in the file image
    int32_t page_index;   // dirIdx * 83 + dirSlot: the page's rank in position space
};

__device__ __forceinline__ uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

// One warp per data page, kDecodeWarps pages per CTA and iteration.  The page is first copied whole into shared memory with
// two 128-bit coalesced loads per lane (1 KB per warp instruction pair: the slot directory, the record offsets and the
// records are then read out of shared memory instead of as scattered byte loads from HBM), then lanes take slots lane,
// lane + 32, ...: consecutive slots are consecutive positions, so the column stores are coalesced.
// `present` (bit p = word p/32, bit p%32) receives one bit per OCCUPIED slot: the Java purge deletes the heap records of
// marked rows and then clears their markedDeleted bits (columnar/Columnarfile.java:874,912-914), so a purged position is an
// empty slot (or an empty directory slot) that no Scan / TupleScan ever returns -- it must not come back as a live row.
constexpr int kDecodeWarps = 8;

__device__ __forceinline__ uint32_t sm_be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

__global__ void __launch_bounds__(kDecodeWarps * 32) decode_pages_kernel(const uint8_t* __restrict__ db, const PageRef* __restrict__ pages, int64_t npages,
                                                                        int per_page, int type, int width, int stride, uint8_t* __restrict__ dst,
                                                                        int64_t nrows, uint32_t* __restrict__ present) {
    __shared__ __align__(16) uint8_t s_page[kDecodeWarps][kPage];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* pg = s_page[warp];
    for (int64_t w = (int64_t)blockIdx.x * kDecodeWarps + warp; w < npages; w += (int64_t)gridDim.x * kDecodeWarps) {
        const PageRef ref = pages[w];
        const uint4* src = reinterpret_cast<const uint4*>(db + (int64_t)ref.page_id * kPage);   // the image is 256-byte aligned, pages 1 KB
        const uint4 a = __ldg(src + lane), b = __ldg(src + 32 + lane);
        __syncwarp();                                               // the previous page of this warp has been decoded
        reinterpret_cast<uint4*>(pg)[lane] = a;
        reinterpret_cast<uint4*>(pg)[32 + lane] = b;
        __syncwarp();
        const int slot_cnt = min((int)(int16_t)sm_be16(pg), per_page);
        for (int s0 = 0; s0 < slot_cnt; s0 += 32) {                 // warp-uniform trip count
            const int s = s0 + lane;
            const int64_t pos = (int64_t)ref.page_index * per_page + s;
            bool occupied = false;
            if (s < slot_cnt && pos < nrows) {
                const uint8_t* sl = pg + kDpFixed + 4 * s;
                const int len = (int)(int16_t)sm_be16(sl);
                const int off = (int)sm_be16(sl + 2);
                if (len >= 0 && off + len <= kPage) {               // len < 0: EMPTY_SLOT (HFPage.java:300)
                    occupied = true;
                    const uint8_t* rec = pg + off;
                    if (type != MBC_ATTR_STRING) {
                        // Convert.getIntValue / getFloValue: 4 bytes big-endian
                        const uint32_t v = (off & 3) == 0 ? __byte_perm(*reinterpret_cast<const uint32_t*>(rec), 0, 0x0123)
                                                          : ((uint32_t)rec[0] << 24) | ((uint32_t)rec[1] << 16) | ((uint32_t)rec[2] << 8) | rec[3];
                        reinterpret_cast<uint32_t*>(dst)[pos] = v;
                    } else {
                        // Convert.getStrValue: [length:2][modified UTF-8 bytes]; the column keeps the bytes zero padded
                        int n = (int)sm_be16(rec);
                        n = min(n, min(width, len - 2));
                        uint32_t* d = reinterpret_cast<uint32_t*>(dst + pos * stride);       // stride is a multiple of 4
                        for (int k = 0; k < n; k += 4) {
                            uint32_t word = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (k + j < n) word |= (uint32_t)rec[2 + k + j] << (8 * j);
                            d[k >> 2] = word;                       // the rest of the row keeps the table's zero fill
                        }
                    }
                }
            }
            // the warp's 32 consecutive positions straddle at most two words of the presence bitmap
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, occupied);
            const int64_t pos0 = (int64_t)ref.page_index * per_page + s0;
            const int sh = (int)(pos0 & 31);
            if (lane == 0 && (m << sh)) atomicOr(present + (pos0 >> 5), m << sh);
            if (lane == 1 && sh && (m >> (32 - sh))) atomicOr(present + (pos0 >> 5) + 1, m >> (32 - sh));
        }
    }
}

// deleted |= ~present (rows of the table only); *any is raised when a deleted bit exists
__global__ void absent_rows_kernel(const uint32_t* present, uint32_t* deleted, int64_t nrows, int* any) {
    const int64_t nwords = (nrows + 31) >> 5;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t a = ~present[i];
        if (i == nwords - 1 && (nrows & 31)) a &= (1u << (nrows & 31)) - 1u;
        const uint32_t d = deleted[i] | a;
        deleted[i] = d;
        if (d) *any = 1;
    }
}

// deleted |= extra (the <cf>.md bits)
__global__ void or_words_kernel(uint32_t* deleted, const uint32_t* extra, int64_t nwords, int* any) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = deleted[i] | extra[i];
        deleted[i] = d;
        if (d) *any = 1;
    }
}

// ---- host-side walk of the directory structure -------------------------------------------------------------

struct Image {
    const uint8_t* b;
    int64_t len;
    bool page_ok(int64_t pid) const { return pid >= 0 && (pid + 1) * kPage <= len; }
    const uint8_t* page(int64_t pid) const { return b + pid * kPage; }
};

static int32_t rd32(const uint8_t* p) { return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]); }
static int rd16(const uint8_t* p) { return (int16_t)((p[0] << 8) | p[1]); }

static int32_t file_entries(const Image& im, std::map<std::string, int32_t>* out) {
    int64_t pid = 0;
    int guard = 0;
    while (pid != -1) {
        if (!im.page_ok(pid) || ++guard > 1 << 20) MBC_FAIL(MBC_ERR_FORMAT, "DB directory chain leaves the file image at page %lld", (long long)pid);
        const uint8_t* pg = im.page(pid);
        const int32_t next = rd32(pg);
        const int32_t n = rd32(pg + 4);
        if (n < 0 || 8 + (int64_t)n * 56 > kPage) MBC_FAIL(MBC_ERR_FORMAT, "DB directory page %lld claims %d entries", (long long)pid, n);
        for (int i = 0; i < n; ++i) {
            const uint8_t* e = pg + 8 + i * 56;
            const int32_t first = rd32(e);
            if (first == -1) continue;
            const int ln = ((e[4] << 8) | e[5]);
            if (ln > 50) MBC_FAIL(MBC_ERR_FORMAT, "DB directory entry with a %d-byte name", ln);
            (*out)[std::string((const char*)e + 6, ln)] = first;
        }
        pid = next;
    }
    return MBC_OK;
}

// every record of a (small) heapfile, in scan order -- used for <cf>.hdr only
static int32_t heap_records(const Image& im, int32_t first_dir, std::vector<std::vector<uint8_t>>* out) {
    int64_t dpid = first_dir;
    int guard = 0;
    while (dpid != -1) {
        if (!im.page_ok(dpid) || ++guard > 1 << 24) MBC_FAIL(MBC_ERR_FORMAT, "heapfile directory chain leaves the file image");
        const uint8_t* d = im.page(dpid);
        const int cnt = rd16(d);
        for (int s = 0; s < cnt; ++s) {
            const int len = rd16(d + kDpFixed + 4 * s), off = (uint16_t)rd16(d + kDpFixed + 4 * s + 2);
            if (len < 0) continue;
            if (len != 8 || off + 8 > kPage) MBC_FAIL(MBC_ERR_FORMAT, "directory record of %d bytes", len);
            const int32_t data_pid = rd32(d + off + 4);
            if (!im.page_ok(data_pid)) MBC_FAIL(MBC_ERR_FORMAT, "data page %d outside the file image", data_pid);
            const uint8_t* pg = im.page(data_pid);
            const int pc = rd16(pg);
            for (int ps = 0; ps < pc; ++ps) {
                const int rl = rd16(pg + kDpFixed + 4 * ps), ro = (uint16_t)rd16(pg + kDpFixed + 4 * ps + 2);
                if (rl < 0) continue;
                if (ro + rl > kPage) MBC_FAIL(MBC_ERR_FORMAT, "record leaves its page");
                out->emplace_back(pg + ro, pg + ro + rl);
            }
        }
        dpid = rd32(d + 12);
    }
    return MBC_OK;
}

// data pages of a column heapfile with their position-space rank (Heapfile.loadPositionBuffer :349-417)
static int32_t column_pages(const Image& im, int32_t first_dir, int per_page, std::vector<PageRef>* pages, int64_t* nrows) {
    int64_t dpid = first_dir;
    int dir_idx = 0, guard = 0;
    *nrows = 0;
    while (dpid != -1) {
        if (!im.page_ok(dpid) || ++guard > 1 << 26) MBC_FAIL(MBC_ERR_FORMAT, "column directory chain leaves the file image");
        const uint8_t* d = im.page(dpid);
        const int cnt = rd16(d);
        if (cnt < 0 || kDpFixed + 4 * cnt > kPage) MBC_FAIL(MBC_ERR_FORMAT, "directory page with %d slots", cnt);
        for (int s = 0; s < cnt; ++s) {
            const int len = rd16(d + kDpFixed + 4 * s), off = (uint16_t)rd16(d + kDpFixed + 4 * s + 2);
            if (len < 0) continue;                                  // deleted directory slot: leaves a gap in position space
            if (len != 8 || off + 8 > kPage) MBC_FAIL(MBC_ERR_FORMAT, "directory record of %d bytes", len);
            const int32_t data_pid = rd32(d + off + 4);
            if (!im.page_ok(data_pid)) MBC_FAIL(MBC_ERR_FORMAT, "data page %d outside the file image", data_pid);
            const int32_t page_index = dir_idx * kDirRecs + s;
            pages->push_back({data_pid, page_index});
            // highest occupied slot of the page bounds the row count
            const uint8_t* pg = im.page(data_pid);
            const int pc = rd16(pg);
            if (pc < 0 || kDpFixed + 4 * pc > kPage) MBC_FAIL(MBC_ERR_FORMAT, "data page %d with %d slots", data_pid, pc);
            for (int ps = pc - 1; ps >= 0; --ps) {
                if (rd16(pg + kDpFixed + 4 * ps) >= 0) {
                    *nrows = std::max<int64_t>(*nrows, (int64_t)page_index * per_page + ps + 1);
                    break;
                }
            }
        }
        dpid = rd32(d + 12);
        ++dir_idx;
    }
    return MBC_OK;
}

}  // namespace mbc

using namespace mbc;

extern "C" int32_t mbc_table_ingest_dbfile(mbc_ctx* ctx, const uint8_t* db_bytes, int64_t db_len, const char* cf_name,
                                           mbc_table** out) {
    if (!ctx || !db_bytes || !cf_name || !out || db_len < kPage) MBC_FAIL(MBC_ERR_ARG, "mbc_table_ingest_dbfile: bad argument");
    *out = nullptr;
    MBC_CUDA(cudaSetDevice(ctx->device));
    Image im{db_bytes, db_len};
    std::map<std::string, int32_t> files;
    MBC_TRY(file_entries(im, &files));
    const std::string name(cf_name);
    auto hdr = files.find(name + ".hdr");
    if (hdr == files.end()) MBC_FAIL(MBC_ERR_FORMAT, "Columnar File does not exist: %s", cf_name);   // Columnarfile.java:253

    // ---- schema (Columnarfile.java:257-300) ----
    std::vector<std::vector<uint8_t>> recs;
    MBC_TRY(heap_records(im, hdr->second, &recs));
    if (recs.size() < 4 || recs[0].size() < 4) MBC_FAIL(MBC_ERR_FORMAT, "%s.hdr has %zu records", cf_name, recs.size());
    const int ncols = rd32(recs[0].data());
    if (ncols <= 0 || ncols > 256 || (int)recs[1].size() < 4 * ncols || (int)recs[2].size() < 4 * ncols)
        MBC_FAIL(MBC_ERR_FORMAT, "%s.hdr describes %d columns", cf_name, ncols);
    std::vector<mbc_coldesc> descs(ncols);
    for (int c = 0; c < ncols; ++c) {
        descs[c].type = rd32(recs[1].data() + 4 * c);
        descs[c].width = rd32(recs[2].data() + 4 * c);
    }

    // ---- page lists ----
    std::vector<std::vector<PageRef>> pages(ncols);
    int64_t nrows = -1;
    for (int c = 0; c < ncols; ++c) {
        auto f = files.find(name + "." + std::to_string(c));
        if (f == files.end()) MBC_FAIL(MBC_ERR_FORMAT, "column heapfile %s.%d is missing", cf_name, c);
        const int rec = descs[c].type == MBC_ATTR_STRING ? descs[c].width + 2 : descs[c].width;
        const int per_page = (kPage - kDpFixed) / (4 + rec);         // Heapfile.java:528
        int64_t n = 0;
        MBC_TRY(column_pages(im, f->second, per_page, &pages[c], &n));
        if (nrows >= 0 && n != nrows) MBC_FAIL(MBC_ERR_FORMAT, "columns disagree on the row count (%lld vs %lld)", (long long)n, (long long)nrows);
        nrows = n;
    }

    mbc_table* t = nullptr;
    MBC_TRY(mbc_table_create(ctx, ncols, descs.data(), nrows, 0, &t));

    // ---- upload the image once, decode every data page on the device ----
    // A position is live only if EVERY column holds a record for it: each column's decode marks its occupied slots in a
    // presence bitmap, and the complement (within [0, nrows)) is ORed into the table's deleted mask.
    uint8_t* d_db = nullptr;
    PageRef* d_pages = nullptr;
    uint32_t* d_present = nullptr;
    int* d_any = nullptr;
    size_t total_pages = 0;
    std::vector<size_t> page_off(ncols, 0);
    for (int c = 0; c < ncols; ++c) { page_off[c] = total_pages; total_pages += pages[c].size(); }
    std::vector<PageRef> all_pages;
    all_pages.reserve(std::max<size_t>(total_pages, 1));
    for (auto& p : pages) all_pages.insert(all_pages.end(), p.begin(), p.end());
    const size_t mask_bytes = (size_t)t->words_pad * 4;
    const unsigned mask_grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((t->words_pad + 255) / 256, ctx->sm_count * 8));
    int32_t s = dev_alloc(ctx, (void**)&d_db, (size_t)db_len, false);
    if (s == MBC_OK) s = dev_alloc(ctx, (void**)&d_pages, std::max<size_t>(total_pages, 1) * sizeof(PageRef), false);
    if (s == MBC_OK) s = dev_alloc(ctx, (void**)&d_present, mask_bytes, false);
    if (s == MBC_OK) s = dev_alloc(ctx, (void**)&d_any, sizeof(int), true);
    if (s == MBC_OK && !t->d_deleted) s = dev_alloc(ctx, (void**)&t->d_deleted, mask_bytes, true);
    // one upload of the image and of every column's page list, then the decodes queue back to back (nothing waits in between)
    if (s == MBC_OK && cudaMemcpyAsync(d_db, db_bytes, (size_t)db_len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) s = MBC_ERR_CUDA;
    if (s == MBC_OK && total_pages &&
        cudaMemcpyAsync(d_pages, all_pages.data(), total_pages * sizeof(PageRef), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) s = MBC_ERR_CUDA;
    begin_timing(ctx);
    for (int c = 0; c < ncols && s == MBC_OK; ++c) {
        const Column& col = t->cols[c];
        const int rec = col.type == MBC_ATTR_STRING ? col.width + 2 : col.width;
        const int per_page = (kPage - kDpFixed) / (4 + rec);
        if (cudaMemsetAsync(d_present, 0, mask_bytes, ctx->stream) != cudaSuccess) { s = MBC_ERR_CUDA; break; }
        if (!pages[c].empty()) {
            const int64_t np = (int64_t)pages[c].size();
            const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((np + kDecodeWarps - 1) / kDecodeWarps, (int64_t)ctx->sm_count * 8));
            decode_pages_kernel<<<grid, kDecodeWarps * 32, 0, ctx->stream>>>(d_db, d_pages + page_off[c], np, per_page, col.type, col.width, col.stride,
                                                                            (uint8_t*)col.d, nrows, d_present);
            ctx->launches++;
        }
        if (nrows > 0) {
            absent_rows_kernel<<<mask_grid, 256, 0, ctx->stream>>>(d_present, t->d_deleted, nrows, d_any);
            ctx->launches++;
        }
    }
    end_timing(ctx);
    if (s == MBC_OK && (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)) s = MBC_ERR_CUDA;   // `all_pages` is pageable

    // ---- markedDeleted (BM.readBitSet :179-215): first record of every page of the <cf>.md chain, ORed in ----
    auto md = files.find(name + ".md");
    if (s == MBC_OK && md != files.end()) {
        std::vector<uint8_t> bytes;
        int64_t pid = md->second;
        int guard = 0;
        while (pid != -1 && im.page_ok(pid) && ++guard < 1 << 24) {
            const uint8_t* pg = im.page(pid);
            if (rd16(pg) > 0) {
                const int len = rd16(pg + kDpFixed), off = (uint16_t)rd16(pg + kDpFixed + 2);
                if (len > 0 && off + len <= kPage) bytes.insert(bytes.end(), pg + off, pg + off + len);
            }
            pid = rd32(pg + 12);
        }
        // BitSet.valueOf(byte[]) is little-endian: byte k = bits 8k..8k+7, the same numbering as the uint32 words; bits at
        // or beyond nrows are irrelevant (every scan masks rows >= nrows)
        bytes.resize(std::min(bytes.size(), mask_bytes));
        bool any = false;
        for (uint8_t b : bytes) any |= b != 0;
        if (any) {
            bytes.resize(mask_bytes, 0);
            if (cudaMemcpyAsync(d_present, bytes.data(), mask_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) s = MBC_ERR_CUDA;
            if (s == MBC_OK) {
                or_words_kernel<<<mask_grid, 256, 0, ctx->stream>>>(t->d_deleted, d_present, t->words_pad, d_any);
                ctx->launches++;
                if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) s = MBC_ERR_CUDA;   // `bytes` is pageable host memory
            }
        }
    }
    int any_deleted = 0;
    if (s == MBC_OK && (cudaMemcpyAsync(&any_deleted, d_any, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                        cudaStreamSynchronize(ctx->stream) != cudaSuccess)) s = MBC_ERR_CUDA;
    dev_free(ctx, d_db);
    dev_free(ctx, d_pages);
    dev_free(ctx, d_present);
    dev_free(ctx, d_any);
    if (s != MBC_OK) {
        set_error("mbc_table_ingest_dbfile: device decode failed");
        mbc_table_free(t);
        return s;
    }
    t->has_deleted = any_deleted != 0;
    *out = t;
    return MBC_OK;
}
