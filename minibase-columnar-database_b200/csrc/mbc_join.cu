// mbc_join.cu -- K6 bitmap equi-join (placeholder until the kernels land in this round).
#include "mbc_internal.cuh"
extern "C" int32_t mbc_bitmap_join(mbc_table*, mbc_table*, const mbc_result*, const mbc_result*, const mbc_term*, int32_t,
                                   const mbc_projspec*, int32_t, uint32_t, const mbc_aggspec*, int32_t, mbc_result**) {
    MBC_FAIL(MBC_ERR_UNSUPPORTED, "mbc_bitmap_join: not built yet");
}
