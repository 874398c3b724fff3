// Error buffer, launch accounting, exclusive scan and max-positive reduction.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "agx_common.cuh"

#include <math.h>

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

// Small device -> host read-backs (a scan total, a reduction result) must not queue behind the bulk device -> host
// copies of the previous edge set on the copy engine (3.7 ms for the 208 MB of the O1280 cut-off edges): a
// one-thread kernel stores the words straight into mapped pinned memory over PCIe, and the host waits for the
// stream only.
__global__ void k_store_to_host(const unsigned long long* __restrict__ src, volatile unsigned long long* dst, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

int agx_readback(void* host_dst, const void* dev_src, int words, cudaStream_t stream) {
    static thread_local unsigned long long* slot = nullptr;  // 32 words of mapped pinned memory per host thread
    AGX_REQUIRE(words > 0 && words <= 32, AGX_ERR_ARG, "agx_readback: 1..32 words");
    if (!slot) AGX_CUDA_OK(cudaHostAlloc((void**)&slot, 32 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
    unsigned long long* dev_view = nullptr;
    AGX_CUDA_OK(cudaHostGetDevicePointer((void**)&dev_view, slot, 0));
    k_store_to_host<<<1, 32, 0, stream>>>((const unsigned long long*)dev_src, dev_view, words);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    AGX_CUDA_OK(cudaStreamSynchronize(stream));
    memcpy(host_dst, slot, words * sizeof(unsigned long long));
    return AGX_OK;
}

// Stream gates (agx_b200.h): work queued behind agx_gate_wait starts when the gate word - 4 bytes of page-locked host
// memory - becomes 1, either through agx_gate_open queued on ANOTHER stream (ordered behind that stream's work) or
// through a plain store by the CPU.  Driver stream memory operations (cuStreamWaitValue32 / cuStreamWriteValue32),
// looked up through the runtime so that the library does not link libcuda: no SM is held while waiting.
#include <cuda.h>
typedef CUresult (*agx_memop32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static int agx_memop(const char* symbol, agx_memop32_fn* out) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
    AGX_CUDA_OK(cudaGetDriverEntryPoint(symbol, &fn, cudaEnableDefault, &found));
    AGX_REQUIRE(found == cudaDriverEntryPointSuccess && fn, AGX_ERR_CUDA, "%s is not available in this driver", symbol);
    *out = (agx_memop32_fn)fn;
    return AGX_OK;
}

static int agx_gate_device_pointer(const void* gate, CUdeviceptr* out) {
    AGX_REQUIRE(gate && ((uintptr_t)gate & 3) == 0, AGX_ERR_ARG, "stream gate: NULL or unaligned word");
    void* dev_view = nullptr;
    AGX_CUDA_OK(cudaHostGetDevicePointer(&dev_view, const_cast<void*>(gate), 0));  // fails for pageable memory
    *out = (CUdeviceptr)(uintptr_t)dev_view;
    return AGX_OK;
}

extern "C" int agx_gate_wait(const uint32_t* gate, void* stream);
extern "C" int agx_gate_open(uint32_t* gate, void* stream);

// One functional self-test per device, remembered: a write through a stream opens a gate, a wait behind it passes.
// (CUDA 12 drivers enable the stream memory operations everywhere they run; a virtualised or restricted device that
// refuses them makes the callers keep their un-gated path instead of failing a build.)
extern "C" int agx_gate_supported(void) {
    static std::atomic<int> verdict[32];  // 0 unknown, 1 yes, 2 no
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return 0;
    int known = verdict[dev].load(std::memory_order_relaxed);
    if (known) return known == 1;
    int ok = 0;
    uint32_t* word = nullptr;
    cudaStream_t probe = nullptr;
    if (cudaHostAlloc((void**)&word, sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
        cudaStreamCreateWithFlags(&probe, cudaStreamNonBlocking) == cudaSuccess) {
        *word = 0;
        ok = agx_gate_open(word, probe) == AGX_OK && agx_gate_wait(word, probe) == AGX_OK &&
             cudaStreamSynchronize(probe) == cudaSuccess && *(volatile uint32_t*)word == 1u;
    }
    if (probe) cudaStreamDestroy(probe);
    if (word) cudaFreeHost(word);
    cudaGetLastError();  // a refused operation must not linger as this thread's last error
    verdict[dev].store(ok ? 1 : 2, std::memory_order_relaxed);
    return ok;
}

extern "C" int agx_gate_wait(const uint32_t* gate, void* stream) {
    static agx_memop32_fn wait_value = nullptr;
    if (!wait_value) {
        int rc = agx_memop("cuStreamWaitValue32", &wait_value);
        if (rc) return rc;
    }
    CUdeviceptr addr;
    int rc = agx_gate_device_pointer(gate, &addr);
    if (rc) return rc;
    CUresult res = wait_value((CUstream)stream, addr, 1u, CU_STREAM_WAIT_VALUE_EQ);
    AGX_REQUIRE(res == CUDA_SUCCESS, AGX_ERR_CUDA, "cuStreamWaitValue32 failed: CUresult %d", (int)res);
    return AGX_OK;
}

extern "C" int agx_gate_open(uint32_t* gate, void* stream) {
    static agx_memop32_fn write_value = nullptr;
    if (!write_value) {
        int rc = agx_memop("cuStreamWriteValue32", &write_value);
        if (rc) return rc;
    }
    CUdeviceptr addr;
    int rc = agx_gate_device_pointer(gate, &addr);
    if (rc) return rc;
    CUresult res = write_value((CUstream)stream, addr, 1u, CU_STREAM_WRITE_VALUE_DEFAULT);
    AGX_REQUIRE(res == CUDA_SUCCESS, AGX_ERR_CUDA, "cuStreamWriteValue32 failed: CUresult %d", (int)res);
    return AGX_OK;
}

// Query-order decision of the tile kernels (agx_tile.cuh agx_query_order): a caller that searches ONE query set in
// several chunks lets the first call decide (sampling costs a stream synchronisation) and pins that decision for the
// remaining chunks, so that the host can run ahead of the device.
static thread_local int g_order_mode = -1;  // -1 auto, 0 as given, 1 binned
static thread_local int g_order_last = 0;
int agx_order_mode(void) { return g_order_mode; }
void agx_note_order(int binned) { g_order_last = binned ? 1 : 0; }
extern "C" void agx_set_query_order_mode(int mode) { g_order_mode = mode < 0 ? -1 : (mode ? 1 : 0); }
extern "C" int agx_last_query_order(void) { return g_order_last; }

// Output label maps of the searches (agx_b200.h agx_set_output_maps): thread-local, read at launch time.
static thread_local const int64_t* g_src_map = nullptr;
static thread_local const int64_t* g_dst_map = nullptr;
extern "C" void agx_set_output_maps(const int64_t* src_map, const int64_t* dst_map) {
    g_src_map = src_map;
    g_dst_map = dst_map;
}
void agx_output_maps(const int64_t** src_map, const int64_t** dst_map) {
    *src_map = g_src_map;
    *dst_map = g_dst_map;
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
        int rc = agx_readback(total, tmp + n_tiles, 1, stream);
        cudaFreeAsync(tmp, stream);
        if (rc) return rc;
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
    int rc = agx_readback(&host, part + blocks, 2, stream);
    cudaFreeAsync(part, stream);
    if (rc) return rc;
    *out_value = host.v;
    *out_index = host.i;
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------------
// The host half of the grid reference distance (utils.py:62-63): sklearn's compiled HaversineDistance64 calls the
// C library's sin / cos (sklearn/metrics/_dist_metrics.pyx.tp:2639-2648), so the handful of candidate nodes the
// GPU search nominates are re-evaluated HERE with the same libm calls and the reference's exact-compare
// semantics: per node the smallest distance to its neighbours (column 1 of the k = 2 self query), over nodes the
// largest strictly positive one.  Plain host code (no CUDA): the value carries glibc's bits, not the GPU's.
// ------------------------------------------------------------------------------------------------
static double host_rdist64(float lat1f, float lon1f, float lat2f, float lon2f) {
    volatile double lat1 = lat1f, lon1 = lon1f, lat2 = lat2f, lon2 = lon2f;
    volatile double sin_0 = sin(0.5 * (lat1 - lat2));
    volatile double sin_1 = sin(0.5 * (lon1 - lon2));
    volatile double a = sin_0 * sin_0;  // volatile: no contraction into FMAs whatever the host flags are
    volatile double cc = cos(lat1) * cos(lat2);
    volatile double b = cc * sin_1;
    volatile double c = b * sin_1;
    return a + c;
}

extern "C" int agx_host_reference_rdist(const float* q_latlon, const float* nb_latlon, int64_t n_cand, int n_nb,
                                        double* out_rdist) {
    AGX_REQUIRE(n_cand >= 0 && n_nb > 0 && out_rdist, AGX_ERR_ARG, "agx_host_reference_rdist: bad arguments");
    AGX_REQUIRE(n_cand == 0 || (q_latlon && nb_latlon), AGX_ERR_ARG, "agx_host_reference_rdist: NULL buffer");
    double best = 0.0;
    for (int64_t c = 0; c < n_cand; ++c) {
        double nearest = 1.0e300;
        for (int j = 0; j < n_nb; ++j) {
            const float* p = nb_latlon + 2 * (c * n_nb + j);
            double r = host_rdist64(q_latlon[2 * c], q_latlon[2 * c + 1], p[0], p[1]);
            if (r < nearest) nearest = r;
        }
        if (nearest > best && nearest < 1.0e300) best = nearest;
    }
    *out_rdist = best;
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

// The device half of get_coordinates_ordering (generate/utils.py:30-33) once the host has produced its two argsorts:
// order = arange(n)[index_latitude][index_longitude[::-1]], its inverse, and the re-ordered coordinates, in one pass.
__global__ void __launch_bounds__(256) k_order_resolve(const int64_t* __restrict__ index_latitude,
                                                       const int64_t* __restrict__ index_longitude, int64_t n,
                                                       const float2* __restrict__ x_in, float2* __restrict__ x_out,
                                                       int64_t* __restrict__ order, int64_t* __restrict__ rank) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = index_latitude[index_longitude[n - 1 - i]];
        order[i] = o;
        rank[o] = i;
        x_out[i] = x_in[o];
    }
}

extern "C" int agx_order_resolve(const int64_t* index_latitude, const int64_t* index_longitude, int64_t n,
                                 const float* x_in, float* x_out, int64_t* order, int64_t* rank, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_order_resolve: n < 0");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(index_latitude && index_longitude && x_in && x_out && order && rank, AGX_ERR_ARG, "agx_order_resolve: NULL buffer");
    k_order_resolve<<<agx_grid(n, 256, 8), 256, 0, stream>>>(index_latitude, index_longitude, n, (const float2*)x_in,
                                                             (float2*)x_out, order, rank);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

extern "C" int agx_mark_nodes(const int32_t* row, int64_t n, int64_t n_nodes, int32_t* flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0 && n_nodes >= 0, AGX_ERR_ARG, "agx_mark_nodes: negative size");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(row && flags, AGX_ERR_ARG, "agx_mark_nodes: NULL buffer");
    int* bad = nullptr;
    agx_pool_keep_warm();
    AGX_CUDA_OK(cudaMallocAsync(&bad, 2 * sizeof(int), stream));
    AGX_CUDA_OK(cudaMemsetAsync(bad, 0, 2 * sizeof(int), stream));
    k_mark_nodes<<<agx_grid(n, 256, 8), 256, 0, stream>>>(row, n, n_nodes, flags, bad);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    int host2[2] = {0, 0};
    int rc = agx_readback(host2, bad, 1, stream);
    cudaFreeAsync(bad, stream);
    if (rc) return rc;
    const int host = host2[0];
    AGX_REQUIRE(host == 0, AGX_ERR_ARG, "agx_mark_nodes: an edge endpoint is outside [0, %lld)", (long long)n_nodes);
    return AGX_OK;
}


// ------------------------------------------------------------------------------------------------
// ascending list of the set bytes of a flag array (deterministic: count per tile, scan, fill)
// ------------------------------------------------------------------------------------------------
#define FLAG_TILE 1024  // flags per block: 256 threads x one 32-bit word
__device__ __forceinline__ unsigned flag_word(const uint8_t* __restrict__ flags, int64_t n, int64_t i) {
    if (i + 4 <= n) return *reinterpret_cast<const unsigned*>(flags + i);
    unsigned w = 0;
    for (int b = 0; b < 4; ++b)
        if (i + b < n) w |= (unsigned)flags[i + b] << (8 * b);
    return w;
}
__device__ __forceinline__ int flag_word_count(unsigned w) {
    return ((w & 0xffu) != 0) + ((w & 0xff00u) != 0) + ((w & 0xff0000u) != 0) + ((w & 0xff000000u) != 0);
}

__global__ void __launch_bounds__(256) k_flag_counts(const uint8_t* __restrict__ flags, int64_t n, int32_t* __restrict__ counts) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    int c = i < n ? flag_word_count(flag_word(flags, n, i)) : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int warp_c[8];
    if ((threadIdx.x & 31) == 0) warp_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += warp_c[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_flag_fill(const uint8_t* __restrict__ flags, int64_t n, const int64_t* __restrict__ offsets,
                                                   int32_t* __restrict__ list) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    unsigned w = i < n ? flag_word(flags, n, i) : 0u;
    int c = flag_word_count(w);
    // exclusive scan of c over the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ int warp_tot[8];
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int k = 0; k < warp; ++k) before += warp_tot[k];
    int64_t pos = offsets[blockIdx.x] + before + incl - c;
    for (int b = 0; b < 4; ++b)
        if ((w >> (8 * b)) & 0xffu) list[pos++] = (int32_t)(i + b);
}

extern "C" int agx_compact_flags(const uint8_t* flags, int64_t n, int32_t* list, int64_t* count, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0 && n < (int64_t)2147483647, AGX_ERR_ARG, "agx_compact_flags: n out of range");
    AGX_REQUIRE(count != nullptr, AGX_ERR_ARG, "agx_compact_flags: count is NULL");
    if (n == 0) {
        AGX_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int64_t), stream));
        return AGX_OK;
    }
    AGX_REQUIRE(flags && list, AGX_ERR_ARG, "agx_compact_flags: NULL buffer");
    AGX_REQUIRE(((uintptr_t)flags & 3) == 0, AGX_ERR_ARG, "agx_compact_flags: flags must be 4-byte aligned");
    const int64_t n_tiles = (n + FLAG_TILE - 1) / FLAG_TILE;
    int32_t* counts = nullptr;
    int64_t* offsets = nullptr;
    agx_pool_keep_warm();
    AGX_CUDA_OK(cudaMallocAsync(&counts, n_tiles * sizeof(int32_t), stream));
    AGX_CUDA_OK(cudaMallocAsync(&offsets, (n_tiles + 1) * sizeof(int64_t), stream));
    k_flag_counts<<<(unsigned)n_tiles, 256, 0, stream>>>(flags, n, counts);
    agx_note_launch(1);
    int rc = agx_exclusive_scan(counts, n_tiles, offsets, nullptr, stream);
    if (rc == AGX_OK) {
        k_flag_fill<<<(unsigned)n_tiles, 256, 0, stream>>>(flags, n, offsets, list);
        agx_note_launch(1);
        AGX_CUDA_OK(cudaMemcpyAsync(count, offsets + n_tiles, sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
    }
    cudaFreeAsync(counts, stream);
    cudaFreeAsync(offsets, stream);
    if (rc) return rc;
    AGX_LAUNCH_OK();
    return AGX_OK;
}

#define RELABEL_MAX_ROWS 8
struct RelabelRows {
    int32_t* row[RELABEL_MAX_ROWS];
    int64_t len[RELABEL_MAX_ROWS];
    int64_t end[RELABEL_MAX_ROWS];  // exclusive prefix ends of the concatenated work space (units of four entries)
    int n_rows;
};

// rows are walked in units of four entries (one 128-bit load / store when the row is 16-byte aligned; row starts of a
// (2, E) list are only 4-byte aligned, so unaligned rows fall back to scalar accesses of the same four entries)
__global__ void __launch_bounds__(256) k_relabel_rows(RelabelRows r, const int64_t* __restrict__ new_index) {
    const int64_t total = r.end[r.n_rows - 1];  // in units of four entries (per row: ceil(len / 4))
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        int which = 0;
#pragma unroll
        for (int w = 0; w < RELABEL_MAX_ROWS - 1; ++w) which += (w < r.n_rows - 1 && u >= r.end[w]) ? 1 : 0;
        int32_t* row = r.row[0];
        int64_t len = r.len[0], first = 0;
#pragma unroll
        for (int w = 1; w < RELABEL_MAX_ROWS; ++w)
            if (w == which) {
                row = r.row[w];
                len = r.len[w];
                first = r.end[w - 1];
            }
        const int64_t j = (u - first) * 4;
        if (j + 4 <= len && ((uintptr_t)(row + j) & 15) == 0) {
            int4 v = *reinterpret_cast<int4*>(row + j);
            v.x = (int32_t)new_index[v.x];
            v.y = (int32_t)new_index[v.y];
            v.z = (int32_t)new_index[v.z];
            v.w = (int32_t)new_index[v.w];
            *reinterpret_cast<int4*>(row + j) = v;
        } else {
            for (int64_t i = j; i < len && i < j + 4; ++i) row[i] = (int32_t)new_index[row[i]];
        }
    }
}

extern "C" int agx_relabel_rows(int32_t* const* rows, const int64_t* lens, int n_rows, const int64_t* new_index,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n_rows >= 0 && n_rows <= RELABEL_MAX_ROWS, AGX_ERR_ARG, "agx_relabel_rows: n_rows out of [0, 8]");
    RelabelRows r;
    r.n_rows = 0;
    int64_t total = 0;
    for (int i = 0; i < n_rows; ++i) {
        AGX_REQUIRE(lens[i] >= 0, AGX_ERR_ARG, "agx_relabel_rows: negative length");
        if (lens[i] == 0) continue;
        AGX_REQUIRE(rows[i] != nullptr, AGX_ERR_ARG, "agx_relabel_rows: NULL row");
        total += (lens[i] + 3) / 4;
        r.row[r.n_rows] = rows[i];
        r.len[r.n_rows] = lens[i];
        r.end[r.n_rows] = total;
        r.n_rows++;
    }
    if (r.n_rows == 0) return AGX_OK;
    AGX_REQUIRE(new_index != nullptr, AGX_ERR_ARG, "agx_relabel_rows: new_index is NULL");
    for (int i = r.n_rows; i < RELABEL_MAX_ROWS; ++i) {
        r.row[i] = nullptr;
        r.len[i] = 0;
        r.end[i] = total;
    }
    k_relabel_rows<<<agx_grid(total, 256, 8), 256, 0, stream>>>(r, new_index);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
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
