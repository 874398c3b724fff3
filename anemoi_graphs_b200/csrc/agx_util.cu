// Error buffer, launch accounting, exclusive scan and max-positive reduction.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "agx_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void agx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void agx_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Scratch buffers come from the stream-ordered allocator.  Its default release threshold is 0: every
// synchronisation hands freed blocks back to the driver and the next cudaMallocAsync pays for a fresh
// mapping (milliseconds for the 100 MB binning buffers).  Keep the pool warm instead.
void agx_pool_keep_warm(void) {
    static std::atomic<unsigned> done_mask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return;
    unsigned bit = 1u << dev;
    if (done_mask.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_mask.fetch_or(bit, std::memory_order_relaxed);
}

extern "C" const char* agx_last_error(void) { return g_err; }
extern "C" int agx_abi_version(void) { return AGX_ABI_VERSION; }
extern "C" int64_t agx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// exclusive scan int32 -> int64 (reduce / scan-of-sums / scan), 3 launches
// ------------------------------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v, int64_t* total_out) {
    // exclusive scan of one value per thread over a SCAN_THREADS block
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1 && total_out) *total_out = wi;
    }
    __syncthreads();
    int64_t res = warp_sums[warp] + incl - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                                  int64_t* __restrict__ tile_sums) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    __shared__ int64_t total;
    block_exclusive_scan_i64(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_of_sums(int64_t* __restrict__ tile_sums, int64_t n_tiles,
                                                                int64_t* __restrict__ grand_total) {
    __shared__ int64_t carry_s, total;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < n_tiles ? tile_sums[i] : 0;
        int64_t ex = block_exclusive_scan_i64(v, &total);
        int64_t carry = carry_s;
        if (i < n_tiles) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const int32_t* __restrict__ in, int64_t n,
                                                              const int64_t* __restrict__ tile_sums,
                                                              int64_t* __restrict__ out) {
    // thread t owns SCAN_ITEMS consecutive items so the per-thread prefix is a register loop
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int64_t ex = block_exclusive_scan_i64(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = ex;  // owner of the last item writes the total
}

extern "C" int agx_exclusive_scan(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_exclusive_scan: n < 0");
    if (n == 0) {
        AGX_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(int64_t), stream));
        if (total) {
            AGX_CUDA_OK(cudaStreamSynchronize(stream));
            *total = 0;
        }
        return AGX_OK;
    }
    int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* tmp = nullptr;
    agx_pool_keep_warm();
    AGX_CUDA_OK(cudaMallocAsync(&tmp, (n_tiles + 1) * sizeof(int64_t), stream));
    k_scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS, 0, stream>>>(counts, n, tmp);
    k_scan_of_sums<<<1, SCAN_THREADS, 0, stream>>>(tmp, n_tiles, tmp + n_tiles);
    k_scan_final<<<(unsigned)n_tiles, SCAN_THREADS, 0, stream>>>(counts, n, tmp, offsets);
    AGX_LAUNCH_OK();
    agx_note_launch(3);
    if (total) {
        AGX_CUDA_OK(cudaMemcpyAsync(total, tmp + n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
        AGX_CUDA_OK(cudaFreeAsync(tmp, stream));
        AGX_CUDA_OK(cudaStreamSynchronize(stream));
    } else {
        AGX_CUDA_OK(cudaFreeAsync(tmp, stream));
    }
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------------
// max over strictly positive float64 values (+ lowest flat index attaining it)
// ------------------------------------------------------------------------------------------------
struct MaxPos {
    double v;
    long long i;
};

__device__ __forceinline__ MaxPos maxpos_merge(MaxPos a, MaxPos b) {
    if (b.v > a.v || (b.v == a.v && b.i >= 0 && (a.i < 0 || b.i < a.i))) return b;
    return a;
}

__global__ void __launch_bounds__(256) k_max_positive(const double* __restrict__ v, int64_t n, MaxPos* __restrict__ part,
                                                       int final_pass, const MaxPos* __restrict__ part_in) {
    MaxPos best = {0.0, -1};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        MaxPos c;
        if (final_pass) {
            c = part_in[i];
        } else {
            double x = v[i];
            c.v = x > 0.0 ? x : 0.0;
            c.i = x > 0.0 ? i : -1;
        }
        best = maxpos_merge(best, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxPos t;
        t.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        t.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        best = maxpos_merge(best, t);
    }
    __shared__ MaxPos sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = maxpos_merge(best, sm[w]);
        part[blockIdx.x] = best;
    }
}

extern "C" int agx_max_positive(const double* values, int64_t n, double* out_value, int64_t* out_index, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0 && out_value && out_index, AGX_ERR_ARG, "agx_max_positive: bad arguments");
    int blocks = agx_grid(n, 256, 4);
    MaxPos* part = nullptr;
    agx_pool_keep_warm();
    AGX_CUDA_OK(cudaMallocAsync(&part, (blocks + 1) * sizeof(MaxPos), stream));
    k_max_positive<<<blocks, 256, 0, stream>>>(values, n, part, 0, nullptr);
    k_max_positive<<<1, 256, 0, stream>>>(nullptr, blocks, part + blocks, 1, part);
    AGX_LAUNCH_OK();
    agx_note_launch(2);
    MaxPos host;
    AGX_CUDA_OK(cudaMemcpyAsync(&host, part + blocks, sizeof(MaxPos), cudaMemcpyDeviceToHost, stream));
    AGX_CUDA_OK(cudaFreeAsync(part, stream));
    AGX_CUDA_OK(cudaStreamSynchronize(stream));
    *out_value = host.v;
    *out_index = host.i;
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------------
// node pruning (RemoveUnconnectedNodes, /root/reference/src/anemoi/graphs/processors/post_process.py:45-60,
// 133-149): mark the endpoints that occur in an edge row; relabel a row through new_index = exclusive scan of
// the keep flags.  Replaces a python dict + Tensor.apply_ over every edge.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mark_nodes(const int32_t* __restrict__ row, int64_t n, int64_t n_nodes,
                                                     int32_t* __restrict__ flags, int* __restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int v = row[i];
        if (v < 0 || v >= n_nodes)
            *bad = 1;
        else
            flags[v] = 1;  // benign race: every writer stores the same value
    }
}

__global__ void __launch_bounds__(256) k_relabel_nodes(int32_t* __restrict__ row, int64_t n, const int64_t* __restrict__ new_index) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        row[i] = (int32_t)new_index[row[i]];
}

extern "C" int agx_mark_nodes(const int32_t* row, int64_t n, int64_t n_nodes, int32_t* flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0 && n_nodes >= 0, AGX_ERR_ARG, "agx_mark_nodes: negative size");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(row && flags, AGX_ERR_ARG, "agx_mark_nodes: NULL buffer");
    int* bad = nullptr;
    agx_pool_keep_warm();
    AGX_CUDA_OK(cudaMallocAsync(&bad, sizeof(int), stream));
    AGX_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    k_mark_nodes<<<agx_grid(n, 256, 8), 256, 0, stream>>>(row, n, n_nodes, flags, bad);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    int host = 0;
    AGX_CUDA_OK(cudaMemcpyAsync(&host, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    AGX_CUDA_OK(cudaFreeAsync(bad, stream));
    AGX_CUDA_OK(cudaStreamSynchronize(stream));
    AGX_REQUIRE(host == 0, AGX_ERR_ARG, "agx_mark_nodes: an edge endpoint is outside [0, %lld)", (long long)n_nodes);
    return AGX_OK;
}

extern "C" int agx_relabel_nodes(int32_t* row, int64_t n, const int64_t* new_index, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_relabel_nodes: n < 0");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(row && new_index, AGX_ERR_ARG, "agx_relabel_nodes: NULL buffer");
    k_relabel_nodes<<<agx_grid(n, 256, 8), 256, 0, stream>>>(row, n, new_index);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}
