// mbc_synth.cu -- stateless counter-RNG synthetic columns, generated in HBM.
//
// Every value is a pure function of (seed, column, global position), so any shard of any table can
// be regenerated on the CPU (oracle/oracle.py: synth_*) and on any GPU without materialising the
// table on the host (BASELINE config 5: 176 GB across 8 GPUs).
#include <algorithm>

#include "mbc_internal.cuh"

namespace mbc {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t synth_draw(uint64_t seed, int col, int64_t pos) {
    return splitmix64((seed ^ ((uint64_t)(col + 1) * 0x9E3779B97F4A7C15ull)) + (uint64_t)pos);
}

// kind 0: int uniform [0,domain)   kind 1: real uniform [0,1000) in steps of 1000/2^24
// kind 3: int (pos * 2654435761 + 12345) mod domain  (a permutation of [0,domain) when nrows == domain)
__global__ void synth32_kernel(uint32_t* out, int64_t nrows, int64_t pos_base, int kind, uint64_t seed, int col,
                               uint64_t domain) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += step) {
        const int64_t pos = pos_base + i;
        uint32_t v;
        if (kind == 0) {
            v = (uint32_t)(synth_draw(seed, col, pos) % domain);
        } else if (kind == 1) {
            double d = (double)(synth_draw(seed, col, pos) >> 40) * (1000.0 / 16777216.0);
            v = __float_as_uint((float)d);
        } else {
            v = (uint32_t)(((uint64_t)pos * 2654435761ull + 12345ull) % domain);
        }
        out[i] = v;
    }
}

// kind 2: char(width): byte k = 0x21 + ((splitmix64(x ^ ((k/8+1) * C)) >> 8*(k%8)) & 0xFF) % 94
__global__ void synth_str_kernel(uint8_t* out, int64_t nrows, int64_t pos_base, uint64_t seed, int col, int width,
                                 int stride) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrows; i += step) {
        const uint64_t x = synth_draw(seed, col, pos_base + i);
        uint8_t* row = out + i * stride;
        uint64_t y = 0;
        for (int k = 0; k < stride; ++k) {
            if ((k & 7) == 0) y = splitmix64(x ^ ((uint64_t)(k / 8 + 1) * 0xD1B54A32D192ED03ull));
            row[k] = k < width ? (uint8_t)(0x21 + ((y >> (8 * (k & 7))) & 0xFF) % 94) : 0;
        }
    }
}

}  // namespace mbc

using namespace mbc;

extern "C" int32_t mbc_table_generate(mbc_table* t, int32_t col, int32_t kind, uint64_t seed, int64_t domain) {
    if (!t || col < 0 || col >= (int)t->cols.size()) MBC_FAIL(MBC_ERR_ARG, "mbc_table_generate: bad column");
    mbc_ctx* ctx = t->ctx;
    MBC_CUDA(cudaSetDevice(ctx->device));
    Column& c = t->cols[col];
    const int threads = 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((t->nrows + threads - 1) / threads, (int64_t)ctx->sm_count * 16));
    if (kind == 2) {
        if (c.type != MBC_ATTR_STRING) MBC_FAIL(MBC_ERR_ARG, "mbc_table_generate: kind 2 needs a string column");
        synth_str_kernel<<<grid, threads, 0, ctx->stream>>>((uint8_t*)c.d, t->nrows, t->pos_base, seed, col, c.width, c.stride);
    } else if (kind == 0 || kind == 1 || kind == 3) {
        if (c.type == MBC_ATTR_STRING || (kind == 1) != (c.type == MBC_ATTR_REAL))
            MBC_FAIL(MBC_ERR_ARG, "mbc_table_generate: kind %d does not match column type %d", kind, c.type);
        if (kind != 1 && (domain <= 0 || domain > (int64_t)INT32_MAX + 1))
            MBC_FAIL(MBC_ERR_ARG, "mbc_table_generate: domain %lld", (long long)domain);
        synth32_kernel<<<grid, threads, 0, ctx->stream>>>((uint32_t*)c.d, t->nrows, t->pos_base, kind, seed, col,
                                                          (uint64_t)std::max<int64_t>(domain, 1));
    } else {
        MBC_FAIL(MBC_ERR_ARG, "mbc_table_generate: unknown kind %d", kind);
    }
    ctx->launches++;
    MBC_CUDA(cudaGetLastError());
    MBC_CUDA(cudaStreamSynchronize(ctx->stream));
    return MBC_OK;
}
