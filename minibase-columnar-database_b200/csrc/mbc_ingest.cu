// mbc_ingest.cu -- K1 heap-page decode (placeholder until the kernels land in this round).
#include "mbc_internal.cuh"
extern "C" int32_t mbc_table_ingest_dbfile(mbc_ctx*, const uint8_t*, int64_t, const char*, mbc_table**) {
    MBC_FAIL(MBC_ERR_UNSUPPORTED, "mbc_table_ingest_dbfile: not built yet");
}
