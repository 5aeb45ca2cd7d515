// Exclusive prefix sums (uint8/uint32 -> uint64) used by the compaction steps of the filter and of
// read linking. Three launches (block reduce, scan of block sums, block scan + offset): pure
// streaming, 1 read + 1 read + 1 write of the data.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) {
            v += t;
        }
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total) {
    __shared__ uint64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = warp_incl_scan(v, lane);
    if (lane == 31) {
        warp_sums[warp] = incl;
    }
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        uint64_t wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) {
            warp_sums[lane] = wi - w; // exclusive
        }
        if (lane == SCAN_THREADS / 32 - 1) {
            *total = wi;
        }
    }
    __syncthreads();
    uint64_t r = warp_sums[warp] + incl - v;
    __syncthreads();
    return r;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T *__restrict__ in, uint64_t n,
                                                                  uint64_t *__restrict__ block_sums) {
    __shared__ uint64_t total;
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint64_t i = base + static_cast<uint64_t>(k) * SCAN_THREADS + threadIdx.x;
        if (i < n) {
            s += in[i];
        }
    }
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) {
        block_sums[blockIdx.x] = total;
    }
}

// single block: exclusive scan of block_sums in place, grand total to *grand
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint64_t *__restrict__ block_sums, uint64_t nb,
                                                                uint64_t *__restrict__ grand) {
    __shared__ uint64_t total;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < nb; base += SCAN_THREADS) {
        uint64_t i = base + threadIdx.x;
        uint64_t v = i < nb ? block_sums[i] : 0;
        uint64_t ex = block_excl_scan(v, &total);
        if (i < nb) {
            block_sums[i] = carry + ex;
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *grand = carry;
    }
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const T *__restrict__ in, uint64_t n,
                                                                 const uint64_t *__restrict__ block_sums,
                                                                 uint64_t *__restrict__ out) {
    __shared__ uint64_t total;
    // thread-contiguous items so that the per-thread serial scan is over consecutive elements
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_TILE + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t s = 0;
    if (sizeof(T) == 1 && SCAN_ITEMS == 8 && base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(in) & 7) == 0) {
        // 8 consecutive bytes per thread in one load
        const uint64_t w = *reinterpret_cast<const uint64_t *>(in + base);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (w >> (8 * k)) & 0xFF;
            s += v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            uint64_t i = base + k;
            v[k] = i < n ? static_cast<uint64_t>(in[i]) : 0;
            s += v[k];
        }
    }
    uint64_t ex = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint64_t i = base + k;
        if (i < n) {
            out[i] = ex;
        }
        ex += v[k];
    }
}

template <typename T>
int scan_impl(sgpu_ctx *ctx, const T *in, uint64_t *out, uint64_t n) {
    cudaStream_t st = ctx->stream;
    if (n == 0) {
        SGPU_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(uint64_t), st));
        return SGPU_OK;
    }
    const uint64_t nb = ceil_div_u64(n, SCAN_TILE);
    DevBuf<uint64_t> sums;
    SGPU_CUDA(ctx, sums.alloc(nb, ctx));
    SGPU_LAUNCH(ctx, (scan_reduce_kernel<T><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, st>>>(in, n, sums.p)));
    SGPU_LAUNCH(ctx, (scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(sums.p, nb, out + n)));
    SGPU_LAUNCH(ctx, (scan_apply_kernel<T><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, st>>>(in, n, sums.p, out)));
    SGPU_CUDA(ctx, cudaGetLastError());
    return SGPU_OK;
}

} // namespace

int sgpu_scan_u8_u64(sgpu_ctx *ctx, const uint8_t *in, uint64_t *out, uint64_t n) {
    return scan_impl<uint8_t>(ctx, in, out, n);
}

int sgpu_scan_u32_u64(sgpu_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n) {
    return scan_impl<uint32_t>(ctx, in, out, n);
}
