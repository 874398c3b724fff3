// K0/K1: cube-sphere cell binning of a reference point set.
//
// Replaces the BallTree build of `NearestNeighbors(metric="haversine").fit(...)`
// (/root/reference/src/anemoi/graphs/edges/builder.py:259-260,364-365).  Layout in HBM: one float4
// (x, y, z, original index) per point, sorted by (cell, original index); `cell_start` gives the run
// of every cell, and cells of one face row are adjacent, so a query's (row, j0..j1) window is ONE
// contiguous run of 16-byte records - every neighbour-search load is a full-sector LDG.128.
#include "agx_common.cuh"

__global__ void __launch_bounds__(256) k_index_cells(const float2* __restrict__ latlon, int64_t n, int cells,
                                                      float4* __restrict__ rec, int* __restrict__ cell_of,
                                                      int* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 p = agx_search_xyz(latlon[i]);
        int cell = agx_cell_of(p.x, p.y, p.z, cells);
        rec[i] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
        cell_of[i] = cell;
        atomicAdd(&hist[cell], 1);
    }
}

// the search vectors on their own (tests pin their error bound, which the FP32 filter margin is derived from)
__global__ void __launch_bounds__(256) k_search_vectors(const float2* __restrict__ latlon, int64_t n, float* __restrict__ xyz) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 p = agx_search_xyz(latlon[i]);
        xyz[3 * i] = p.x;
        xyz[3 * i + 1] = p.y;
        xyz[3 * i + 2] = p.z;
    }
}

extern "C" int agx_search_vectors(const float* latlon, int64_t n, float* xyz, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_search_vectors: n < 0");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(latlon && xyz, AGX_ERR_ARG, "agx_search_vectors: NULL buffer");
    k_search_vectors<<<agx_grid(n, 256, 8), 256, 0, stream>>>((const float2*)latlon, n, xyz);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

__global__ void __launch_bounds__(256) k_index_scatter(const float4* __restrict__ rec, const int* __restrict__ cell_of,
                                                        int64_t n, const int64_t* __restrict__ cell_start64,
                                                        int* __restrict__ fill, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int cell = cell_of[i];
        int slot = atomicAdd(&fill[cell], 1);
        out[cell_start64[cell] + slot] = rec[i];
    }
}

// One warp per cell: order the cell's records by original index (rank by counting), so the binned
// array - and with it every downstream edge order - is deterministic despite the atomic scatter.
__global__ void __launch_bounds__(256) k_index_sort_cells(const float4* __restrict__ in, const int64_t* __restrict__ cell_start64,
                                                           int n_cells, float4* __restrict__ out,
                                                           int* __restrict__ cell_start32) {
    int lane = threadIdx.x & 31;
    int warps = (gridDim.x * blockDim.x) >> 5;
    for (int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cell <= n_cells; cell += warps) {
        int64_t s = cell_start64[cell];
        if (lane == 0) cell_start32[cell] = (int)s;
        if (cell == n_cells) break;
        int m = (int)(cell_start64[cell + 1] - s);
        if (m == 0) continue;
        if (m <= 32) {
            float4 mine = lane < m ? in[s + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            int key = lane < m ? __float_as_int(mine.w) : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += (__shfl_sync(0xffffffffu, key, j) < key);
            if (lane < m) out[s + rank] = mine;
        } else {
            for (int e = lane; e < m; e += 32) {
                float4 mine = in[s + e];
                int key = __float_as_int(mine.w);
                int rank = 0;
                for (int j = 0; j < m; ++j) rank += (__float_as_int(in[s + j].w) < key);
                out[s + rank] = mine;
            }
        }
    }
}

static int choose_cells(int64_t n, int hint_k, double hint_radius) {
    const double half_pi = 1.5707963267948966;
    double width;
    if (hint_radius > 0.0) {
        width = hint_radius;  // radius search: a 3x3 window of radius-wide cells covers the cap
    } else if (hint_k > 0) {
        width = 2.0 * sqrt((double)hint_k / (double)(n > 0 ? n : 1));  // ~ 0.6x the k-NN search radius
    } else {
        width = sqrt(12.566370614359172 / (double)(n > 0 ? n : 1)) * 1.5;
    }
    double c = floor(half_pi / width);
    // never more cells than ~4 per point: empty cells cost cell_start traffic and sort-warp time
    double cap = ceil(sqrt((double)(n > 0 ? n : 1) * 4.0 / 6.0));
    if (c > cap) c = cap;
    if (c > 2048.0) c = 2048.0;
    if (c < 1.0) c = 1.0;
    return (int)c;
}

extern "C" int agx_index_build(const float* latlon, int64_t n, int cells_per_face, int hint_k, double hint_radius,
                               void* stream_, agx_index_t** out) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(out != nullptr, AGX_ERR_ARG, "agx_index_build: out is NULL");
    *out = nullptr;
    AGX_REQUIRE(n > 0 && latlon != nullptr, AGX_ERR_ARG, "agx_index_build: need n > 0 reference points (got %lld)", (long long)n);
    AGX_REQUIRE(n < (int64_t)2147483647, AGX_ERR_ARG, "agx_index_build: n must fit int32 (edge_index is int32)");
    AGX_REQUIRE(cells_per_face >= 0 && cells_per_face <= 2048, AGX_ERR_ARG, "agx_index_build: cells_per_face out of [0, 2048]");
    agx_pool_keep_warm();
    int dev = 0;
    AGX_CUDA_OK(cudaGetDevice(&dev));
    int cells = cells_per_face > 0 ? cells_per_face : choose_cells(n, hint_k, hint_radius);
    int n_cells = 6 * cells * cells;

    agx_index* ix = new agx_index();
    ix->n = n;
    ix->cells = cells;
    ix->device = dev;
    ix->latlon = (const float2*)latlon;
    ix->pts = nullptr;
    ix->cell_start = nullptr;
    ix->chord2_typ = (float)(12.566370614359172 / (double)n);
    ix->order_q = nullptr;
    ix->order_nq = 0;
    ix->order_perm = nullptr;

    float4 *rec = nullptr, *scat = nullptr;
    int *cell_of = nullptr, *hist = nullptr;
    int64_t* start64 = nullptr;
    cudaError_t e;
#define IDX_TRY(expr)                                                                      \
    if ((e = (expr)) != cudaSuccess) {                                                     \
        agx_set_error("%s failed: %s", #expr, cudaGetErrorString(e));                      \
        cudaFreeAsync(rec, stream); cudaFreeAsync(scat, stream); cudaFreeAsync(cell_of, stream); \
        cudaFreeAsync(hist, stream); cudaFreeAsync(start64, stream);                       \
        cudaFreeAsync(ix->pts, stream); cudaFreeAsync(ix->cell_start, stream);             \
        delete ix;                                                                         \
        return AGX_ERR_CUDA;                                                               \
    }
    IDX_TRY(cudaMallocAsync(&ix->pts, n * sizeof(float4), stream));
    IDX_TRY(cudaMallocAsync(&ix->cell_start, ((size_t)n_cells + 1) * sizeof(int), stream));
    IDX_TRY(cudaMallocAsync(&rec, n * sizeof(float4), stream));
    IDX_TRY(cudaMallocAsync(&scat, n * sizeof(float4), stream));
    IDX_TRY(cudaMallocAsync(&cell_of, n * sizeof(int), stream));
    IDX_TRY(cudaMallocAsync(&hist, 2 * (size_t)n_cells * sizeof(int), stream));
    IDX_TRY(cudaMallocAsync(&start64, ((size_t)n_cells + 1) * sizeof(int64_t), stream));
    IDX_TRY(cudaMemsetAsync(hist, 0, 2 * (size_t)n_cells * sizeof(int), stream));

    int grid = agx_grid(n, 256, 8);
    k_index_cells<<<grid, 256, 0, stream>>>((const float2*)latlon, n, cells, rec, cell_of, hist);
    agx_note_launch(1);
    int rc = agx_exclusive_scan(hist, n_cells, start64, nullptr, stream);
    if (rc != AGX_OK) {
        e = cudaErrorUnknown;
        IDX_TRY(e);
    }
    k_index_scatter<<<grid, 256, 0, stream>>>(rec, cell_of, n, start64, hist + n_cells, scat);
    int sort_grid = agx_grid((int64_t)(n_cells + 1) * 32, 256, 8);
    k_index_sort_cells<<<sort_grid, 256, 0, stream>>>(scat, start64, n_cells, ix->pts, ix->cell_start);
    agx_note_launch(2);
    IDX_TRY(cudaGetLastError());
    IDX_TRY(cudaFreeAsync(rec, stream)); rec = nullptr;
    IDX_TRY(cudaFreeAsync(scat, stream)); scat = nullptr;
    IDX_TRY(cudaFreeAsync(cell_of, stream)); cell_of = nullptr;
    IDX_TRY(cudaFreeAsync(hist, stream)); hist = nullptr;
    IDX_TRY(cudaFreeAsync(start64, stream)); start64 = nullptr;
#undef IDX_TRY
    *out = ix;
    return AGX_OK;
}

extern "C" int agx_index_free(agx_index_t* ix, void* stream_) {
    if (!ix) return AGX_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaFreeAsync(ix->pts, stream);
    cudaFreeAsync(ix->cell_start, stream);
    if (ix->order_perm) cudaFreeAsync(ix->order_perm, stream);
    delete ix;
    return AGX_OK;
}

extern "C" int agx_index_info(const agx_index_t* ix, int64_t* n, int* cells_per_face) {
    AGX_REQUIRE(ix != nullptr, AGX_ERR_ARG, "agx_index_info: NULL index");
    if (n) *n = ix->n;
    if (cells_per_face) *cells_per_face = ix->cells;
    return AGX_OK;
}
