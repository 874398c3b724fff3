// Tile machinery shared by the KNN and the low-degree cut-off kernels: one WARP per tile of 32 queries, the union of
// the lanes' cap windows staged through shared memory with 1-D bulk async copies (TMA), completion on an mbarrier.
#pragma once
#include <stdlib.h>

#include "agx_search.cuh"

#define AGX_TILE_WARPS 4
#define AGX_TILE_STAGE 320  // candidate records per warp stage (5 KB)
#define AGX_TILE_MAX_ROWS 32

// The stage plan of one tile (a warp of 32 queries): the union of the lanes' cap windows, one window row per lane.
// On return lane l owns the contiguous record run [s, s + cnt) of the cell-sorted array (cnt = 0: no row), `incl`
// is the inclusive prefix of cnt over the lanes and m_total the number of candidates.  false = the tile cannot be
// staged (more than AGX_TILE_MAX_ROWS rows or AGX_TILE_STAGE candidates: queries far apart, or a cap that covers the sphere).
__device__ __forceinline__ bool agx_tile_plan(const int* __restrict__ cell_start, int cells, float3 qv, const AgxCap& cap,
                                              int lane, int& s, int& cnt, int& incl, int& m_total) {
    int n_rows = 0;  // rows of the union window over all faces (uniform across the warp)
    int my_row = -1, my_j0 = 0, my_j1 = 0;
    if (cap.everything) return false;
    bool fits = true;
    // fast path: the whole tile lives on ONE cube face and every lane's cap ends at least a cell inside it
    bool fast = false;
    {
        const int face = agx_major_face(qv.x, qv.y, qv.z);
        if (__all_sync(0xffffffffu, face == __shfl_sync(0xffffffffu, face, 0))) {
            float fa, fb, fc;
            agx_face_frame(face, qv.x, qv.y, qv.z, fa, fb, fc);
            int i0, i1, j0, j1;
            bool ok = agx_axis_window_major(fa, fc, cap, cells, i0, i1);
            ok = agx_axis_window_major(fb, fc, cap, cells, j0, j1) && ok;
            if (__all_sync(0xffffffffu, ok)) {
                fast = true;
                i0 = __reduce_min_sync(0xffffffffu, i0);
                j0 = __reduce_min_sync(0xffffffffu, j0);
                i1 = __reduce_max_sync(0xffffffffu, i1);
                j1 = __reduce_max_sync(0xffffffffu, j1);
                n_rows = i1 - i0 + 1;
                if (lane < n_rows) {
                    my_row = (face * cells + i0 + lane) * cells;
                    my_j0 = j0;
                    my_j1 = j1;
                }
                fits = n_rows <= AGX_TILE_MAX_ROWS;
            }
        }
    }
    if (!fast) {
        for (int face = 0; face < 6; ++face) {
            int i0, i1, j0, j1;
            bool ok = agx_face_window(face, qv, cap, cells, i0, i1, j0, j1);
            if (!__any_sync(0xffffffffu, ok)) continue;
            i0 = __reduce_min_sync(0xffffffffu, ok ? i0 : 0x7fffffff);
            j0 = __reduce_min_sync(0xffffffffu, ok ? j0 : 0x7fffffff);
            i1 = __reduce_max_sync(0xffffffffu, ok ? i1 : -1);
            j1 = __reduce_max_sync(0xffffffffu, ok ? j1 : -1);
            int rows = i1 - i0 + 1;
            int slot = lane - n_rows;
            if (slot >= 0 && slot < rows) {
                my_row = (face * cells + i0 + slot) * cells;
                my_j0 = j0;
                my_j1 = j1;
            }
            n_rows += rows;
        }
        fits = n_rows <= AGX_TILE_MAX_ROWS;
    }
    if (!fits) return false;
    s = 0;
    cnt = 0;
    if (my_row >= 0) {
        s = __ldg(cell_start + my_row + my_j0);
        cnt = __ldg(cell_start + my_row + my_j1 + 1) - s;
    }
    // exclusive prefix of the row lengths = each row's offset in the stage
    incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    m_total = __shfl_sync(0xffffffffu, incl, 31);
    return m_total <= AGX_TILE_STAGE;
}

// ---- mbarrier / bulk-copy primitives (PTX) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(phase)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; 16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- incoherent query order --------------------------------------------------------------------------------
// The tile search relies on consecutive queries being neighbours (grids are).  When they are not (a shuffled point
// cloud), every tile falls back to the per-thread search - 5x slower.  agx_knn samples the tile plans first; if
// most sampled tiles cannot be staged it bins the queries on a coarse cube-sphere grid (counting sort, the order
// inside a bin fixed by query index so that runs are reproducible) and walks them in bin order.
#define AGX_SAMPLE_TILES 1024

static __global__ void __launch_bounds__(128) k_tile_sample(const int* __restrict__ cell_start, int cells, const float2* __restrict__ q_latlon,
                                                            int64_t nq, float chord2_cap, int64_t tile_stride, int n_samples,
                                                            int* __restrict__ n_fit) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_samples) return;
    const int64_t slot = (int64_t)w * tile_stride * 32 + lane;
    const float3 qv = agx_search_xyz(q_latlon[slot < nq ? slot : nq - 1]);
    const AgxCap cap = agx_make_cap(chord2_cap);
    int s, cnt, incl, m_total;
    bool fits = agx_tile_plan(cell_start, cells, qv, cap, lane, s, cnt, incl, m_total);
    if (lane == 0 && fits) atomicAdd(n_fit, 1);
}

static __global__ void __launch_bounds__(256) k_query_bins(const float2* __restrict__ q_latlon, int64_t nq, int bins,
                                                     int* __restrict__ bin_of, int* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        float3 p = agx_search_xyz(q_latlon[i]);
        int b = agx_cell_of(p.x, p.y, p.z, bins);
        bin_of[i] = b;
        atomicAdd(&hist[b], 1);
    }
}

static __global__ void __launch_bounds__(256) k_query_scatter(const int* __restrict__ bin_of, int64_t nq,
                                                        const int64_t* __restrict__ start, int* __restrict__ fill,
                                                        int32_t* __restrict__ perm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        int b = bin_of[i];
        perm[start[b] + atomicAdd(&fill[b], 1)] = (int32_t)i;
    }
}

// one warp per bin: ascending query index inside the bin (rank by counting).  Bins beyond 1024 entries (heavily
// clustered queries) keep the scatter order - still correct, only the processing order varies between runs.
static __global__ void __launch_bounds__(256) k_query_sort_bins(const int32_t* __restrict__ in, const int64_t* __restrict__ start,
                                                          int n_bins, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_bins; b += warps) {
        const int64_t s0 = start[b];
        const int m = (int)(start[b + 1] - s0);
        if (m > 1024) {
            for (int e = lane; e < m; e += 32) out[s0 + e] = in[s0 + e];
            continue;
        }
        for (int e = lane; e < m; e += 32) {
            int key = in[s0 + e], rank = 0;
            for (int j = 0; j < m; ++j) rank += (in[s0 + j] < key);
            out[s0 + rank] = key;
        }
    }
}

// perm = the queries in bin order.  Scratch comes from the stream-ordered pool; the caller frees perm.
static int agx_bin_queries(const float2* q_latlon, int64_t nq, int32_t** perm_out, cudaStream_t stream) {
    double c = sqrt((double)nq / 144.0);  // ~24 queries per bin on a uniform sphere
    int bins = (int)(c < 1.0 ? 1.0 : (c > 1024.0 ? 1024.0 : c));
    int n_bins = 6 * bins * bins;
    int *bin_of = nullptr, *hist = nullptr;
    int64_t* start = nullptr;
    int32_t *tmp = nullptr, *perm = nullptr;
    AGX_CUDA_OK(cudaMallocAsync(&bin_of, nq * sizeof(int), stream));
    AGX_CUDA_OK(cudaMallocAsync(&hist, 2 * (size_t)n_bins * sizeof(int), stream));
    AGX_CUDA_OK(cudaMallocAsync(&start, ((size_t)n_bins + 1) * sizeof(int64_t), stream));
    AGX_CUDA_OK(cudaMallocAsync(&tmp, nq * sizeof(int32_t), stream));
    AGX_CUDA_OK(cudaMallocAsync(&perm, nq * sizeof(int32_t), stream));
    AGX_CUDA_OK(cudaMemsetAsync(hist, 0, 2 * (size_t)n_bins * sizeof(int), stream));
    int grid = agx_grid(nq, 256, 8);
    k_query_bins<<<grid, 256, 0, stream>>>(q_latlon, nq, bins, bin_of, hist);
    int rc = agx_exclusive_scan(hist, n_bins, start, nullptr, stream);
    if (rc != AGX_OK) return rc;
    k_query_scatter<<<grid, 256, 0, stream>>>(bin_of, nq, start, hist + n_bins, tmp);
    k_query_sort_bins<<<agx_grid((int64_t)n_bins * 32, 256, 8), 256, 0, stream>>>(tmp, start, n_bins, perm);
    AGX_LAUNCH_OK();
    agx_note_launch(3);
    AGX_CUDA_OK(cudaFreeAsync(bin_of, stream));
    AGX_CUDA_OK(cudaFreeAsync(hist, stream));
    AGX_CUDA_OK(cudaFreeAsync(start, stream));
    AGX_CUDA_OK(cudaFreeAsync(tmp, stream));
    *perm_out = perm;
    return AGX_OK;
}

// fraction of sampled tiles whose candidates can be staged (1.0 for small inputs, which are not worth sampling)
static int agx_sample_coherence(const agx_index_t* ix, const float2* q_latlon, int64_t nq, float chord2_cap, double* frac,
                                cudaStream_t stream) {
    int64_t n_tiles = (nq + 31) / 32;
    int n_samples = (int)(n_tiles < AGX_SAMPLE_TILES ? n_tiles : AGX_SAMPLE_TILES);
    int64_t stride = n_tiles / n_samples;
    int* n_fit = nullptr;
    AGX_CUDA_OK(cudaMallocAsync(&n_fit, 2 * sizeof(int), stream));
    AGX_CUDA_OK(cudaMemsetAsync(n_fit, 0, 2 * sizeof(int), stream));
    k_tile_sample<<<(n_samples * 32 + 127) / 128, 128, 0, stream>>>(ix->cell_start, ix->cells, q_latlon, nq, chord2_cap, stride,
                                                                       n_samples, n_fit);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    int host[2] = {0, 0};
    int rc = agx_readback(host, n_fit, 1, stream);  // not through the copy engine: bulk D2H copies may be queued there
    cudaFreeAsync(n_fit, stream);
    if (rc) return rc;
    *frac = (double)host[0] / (double)n_samples;
    return AGX_OK;
}

// Processing order of a query set for the tile kernels: NULL (as given) or a binned permutation the caller frees
// with cudaFreeAsync.  "auto" samples the tile plans of large inputs and bins when fewer than 3/4 of the sampled
// tiles can be staged; env_name (AGX_KNN_BIN / AGX_RADIUS_BIN) = 0 / 1 forces the choice.
static int agx_query_order(const agx_index_t* ix, const float2* q_latlon, int64_t nq, float chord2_cap, const char* env_name,
                           int32_t** perm, cudaStream_t stream) {
    *perm = nullptr;
    int mode = -1;
    if (const char* env = getenv(env_name)) mode = atoi(env);
    if (mode < 0) mode = agx_order_mode();  // pinned by a caller that searches one query set in chunks
    bool bin = mode == 1;
    if (mode < 0 && nq >= 262144) {
        double frac = 1.0;
        int rc = agx_sample_coherence(ix, q_latlon, nq, chord2_cap, &frac, stream);
        if (rc != AGX_OK) return rc;
        bin = frac < 0.75;
    }
    agx_note_order(bin);
    if (!bin) return AGX_OK;
    agx_pool_keep_warm();
    return agx_bin_queries(q_latlon, nq, perm, stream);
}
