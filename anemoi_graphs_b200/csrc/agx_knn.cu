// K2: k-nearest-neighbour search over the cell-binned index.
//
// Replaces sklearn's BallTree.query as driven by KNNEdges.get_adjacency_matrix
// (/root/reference/src/anemoi/graphs/edges/builder.py:259-265) and utils.get_grid_reference_distance
// (utils.py:62).  One thread per query:
//   1. FP32 filter: scan the cells the search cap touches, squared chord on the FMA pipe, keep the
//      k+1 best (d, index) in registers.  The cap grows until it provably holds the k-th neighbour.
//   2. If the (k+1)-th candidate is within the FP32 error margin of the k-th, the set is decided in
//      float64 with the reference's own haversine formula (tie rule: lower index within 2^-40).
// The edge SET is what parity is judged on, so the filter only has to isolate the k members;
// their float64 distances are evaluated only on request (out_rdist).
#include "agx_search.cuh"

template <int CAP>
struct TopF {  // ascending (d, idx), CAP entries in registers
    float d[CAP];
    int id[CAP];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < CAP; ++s) {
            d[s] = __int_as_float(0x7f800000);
            id[s] = 0x7fffffff;
        }
    }
    // register-resident reads at a runtime position (a dynamic subscript would demote the arrays to local memory)
    __device__ __forceinline__ float d_at(int i) const {
        float r = d[0];
#pragma unroll
        for (int s = 1; s < CAP; ++s) r = (i == s) ? d[s] : r;
        return r;
    }
    __device__ __forceinline__ int id_at(int i) const {
        int r = id[0];
#pragma unroll
        for (int s = 1; s < CAP; ++s) r = (i == s) ? id[s] : r;
        return r;
    }
    __device__ __forceinline__ void insert(float cd, int ci) {
#pragma unroll
        for (int s = 0; s < CAP; ++s) {
            bool lt = (cd < d[s]) || (cd == d[s] && ci < id[s]);
            float td = d[s];
            int ti = id[s];
            d[s] = lt ? cd : td;
            id[s] = lt ? ci : ti;
            cd = lt ? td : cd;
            ci = lt ? ti : ci;
        }
    }
};

template <int CAP>
struct TopD {  // ascending under agx_tie_less
    double r[CAP];
    int id[CAP];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < CAP; ++s) {
            r[s] = 1e300;
            id[s] = 0x7fffffff;
        }
    }
    __device__ __forceinline__ double r_at(int i) const {
        double v = r[0];
#pragma unroll
        for (int s = 1; s < CAP; ++s) v = (i == s) ? r[s] : v;
        return v;
    }
    __device__ __forceinline__ int id_at(int i) const {
        int v = id[0];
#pragma unroll
        for (int s = 1; s < CAP; ++s) v = (i == s) ? id[s] : v;
        return v;
    }
    __device__ __forceinline__ void insert(double cr, int ci) {
#pragma unroll
        for (int s = 0; s < CAP; ++s) {
            bool lt = agx_tie_less(cr, ci, r[s], id[s]);
            double tr = r[s];
            int ti = id[s];
            r[s] = lt ? cr : tr;
            id[s] = lt ? ci : ti;
            cr = lt ? tr : cr;
            ci = lt ? ti : ci;
        }
    }
};

// CAP >= k + 1.  Entries beyond k + 1 are carried but never read.
template <int CAP>
__global__ void __launch_bounds__(128) k_knn(const float4* __restrict__ pts, const int* __restrict__ cell_start,
                                             const float2* __restrict__ ref_latlon, int cells, float chord2_init,
                                             const float2* __restrict__ q_latlon, int64_t nq, int k,
                                             int32_t* __restrict__ out_src, int32_t* __restrict__ out_dst,
                                             int64_t dst_base, double* __restrict__ out_rdist,
                                             unsigned long long* __restrict__ stats) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const float2 ql = q_latlon[q];
        const float3 qv = agx_search_xyz(ql);
        TopF<CAP> top;
        float t2 = chord2_init;
        bool widened = false;
        AgxCap cap;
        while (true) {
            top.reset();
            cap = agx_make_cap(t2);
            for (int face = 0; face < 6; ++face) {
                int i0, i1, j0, j1;
                if (!agx_face_window(face, qv, cap, cells, i0, i1, j0, j1)) continue;
                for (int i = i0; i <= i1; ++i) {
                    int row = (face * cells + i) * cells;
                    int s = __ldg(cell_start + row + j0), e = __ldg(cell_start + row + j1 + 1);
                    for (int p = s; p < e; ++p) {
                        float4 c = __ldg(pts + p);
                        float dx = qv.x - c.x, dy = qv.y - c.y, dz = qv.z - c.z;
                        float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        int ci = __float_as_int(c.w);
                        if (d < top.d[CAP - 1] || (d == top.d[CAP - 1] && ci < top.id[CAP - 1])) top.insert(d, ci);
                    }
                }
            }
            if (cap.everything) break;
            float dk = top.d_at(k - 1);
            bool have_k = top.id_at(k - 1) != 0x7fffffff;
            float need = dk + 3.75f * agx_chord2_margin(dk);  // d_k + 3 margins (margin taken 1.25x)
            if (have_k && need <= t2) break;
            widened = true;
            t2 = have_k ? fmaxf(need * 1.0001f, t2 * 1.5f) : t2 * 4.0f;
        }
        // ---- decide the set ------------------------------------------------------------------
        float dk = top.d_at(k - 1);
        float amb = dk + 2.5f * agx_chord2_margin(dk);  // d_k + 2 margins
        bool ambiguous = (CAP > 1) && (top.id_at(k) != 0x7fffffff) && (top.d_at(k) <= amb);
        int32_t* os = out_src + q * k;
        if (!ambiguous) {
            if (out_rdist == nullptr) {
#pragma unroll
                for (int s = 0; s < CAP - 1; ++s)
                    if (s < k) os[s] = top.id[s];
            } else {
                TopD<CAP> fin;
                fin.reset();
#pragma unroll
                for (int s = 0; s < CAP - 1; ++s)
                    if (s < k) fin.insert(agx_rdist64(ql, ref_latlon[top.id[s]]), top.id[s]);
#pragma unroll
                for (int s = 0; s < CAP - 1; ++s)
                    if (s < k) {
                        os[s] = fin.id[s];
                        out_rdist[q * k + s] = fin.r[s];
                    }
            }
        } else {
            // float64 re-decision over every candidate the filter could not separate
            TopD<CAP> fin;
            fin.reset();
            for (int face = 0; face < 6; ++face) {
                int i0, i1, j0, j1;
                if (!agx_face_window(face, qv, cap, cells, i0, i1, j0, j1)) continue;
                for (int i = i0; i <= i1; ++i) {
                    int row = (face * cells + i) * cells;
                    int s = __ldg(cell_start + row + j0), e = __ldg(cell_start + row + j1 + 1);
                    for (int p = s; p < e; ++p) {
                        float4 c = __ldg(pts + p);
                        float dx = qv.x - c.x, dy = qv.y - c.y, dz = qv.z - c.z;
                        float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        if (d <= amb) {
                            int ci = __float_as_int(c.w);
                            double r = agx_rdist64(ql, ref_latlon[ci]);
                            if (agx_tie_less(r, ci, fin.r[CAP - 1], fin.id[CAP - 1])) fin.insert(r, ci);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < CAP - 1; ++s)
                if (s < k) {
                    os[s] = fin.id[s];
                    if (out_rdist) out_rdist[q * k + s] = fin.r[s];
                }
            if (stats) {
                atomicAdd(stats + 0, 1ull);
                double rk = fin.r_at(k - 1), rn = fin.r_at(k);
                if (fin.id_at(k) != 0x7fffffff && fabs(rn - rk) <= AGX_TIE_TAU * fmax(rn, rk)) atomicAdd(stats + 1, 1ull);
            }
        }
        if (stats && widened) atomicAdd(stats + 2, 1ull);
        if (out_dst) {
            int32_t* od = out_dst + q * k;
            int32_t t = (int32_t)(dst_base + q);
            for (int s = 0; s < k; ++s) od[s] = t;
        }
    }
}

template <int CAP>
static void launch_knn(const agx_index* ix, float chord2_init, const float* q, int64_t nq, int k, int32_t* out_src,
                       int32_t* out_dst, int64_t dst_base, double* out_rdist, int64_t* stats, cudaStream_t stream) {
    int grid = agx_grid(nq, 128, 16);
    k_knn<CAP><<<grid, 128, 0, stream>>>(ix->pts, ix->cell_start, ix->latlon, ix->cells, chord2_init, (const float2*)q,
                                         nq, k, out_src, out_dst, dst_base, out_rdist, (unsigned long long*)stats);
}

extern "C" int agx_knn(const agx_index_t* ix, const float* q_latlon, int64_t nq, int k, int32_t* out_src,
                       int32_t* out_dst, int64_t dst_base, double* out_rdist, int64_t* stats, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(ix != nullptr, AGX_ERR_ARG, "agx_knn: NULL index");
    AGX_REQUIRE(nq >= 0, AGX_ERR_ARG, "agx_knn: nq < 0");
    AGX_REQUIRE(k > 0, AGX_ERR_ARG, "agx_knn: k must be positive (got %d)", k);
    // sklearn raises "Expected n_neighbors <= n_samples_fit" (neighbors/_base.py kneighbors)
    AGX_REQUIRE((int64_t)k <= ix->n, AGX_ERR_ARG, "Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld",
                k, (long long)ix->n);
    AGX_REQUIRE(k <= 64, AGX_ERR_UNSUPPORTED, "agx_knn: k = %d > 64 is not built yet", k);
    AGX_REQUIRE(dst_base + nq < (int64_t)2147483647, AGX_ERR_ARG, "agx_knn: target index exceeds int32");
    if (nq == 0) return AGX_OK;
    AGX_REQUIRE(q_latlon && out_src, AGX_ERR_ARG, "agx_knn: NULL buffer");
    // first cap: ~1.5x the k-NN radius at mean density, chord^2 ~ rho^2 = 9 * (k+1) / n
    double t2 = 9.0 * (double)(k + 1) / (double)ix->n;
    if (t2 > 4.0) t2 = 4.0;
    float chord2_init = (float)t2;
    if (k <= 3)
        launch_knn<4>(ix, chord2_init, q_latlon, nq, k, out_src, out_dst, dst_base, out_rdist, stats, stream);
    else if (k <= 7)
        launch_knn<8>(ix, chord2_init, q_latlon, nq, k, out_src, out_dst, dst_base, out_rdist, stats, stream);
    else if (k <= 16)
        launch_knn<17>(ix, chord2_init, q_latlon, nq, k, out_src, out_dst, dst_base, out_rdist, stats, stream);
    else if (k <= 32)
        launch_knn<33>(ix, chord2_init, q_latlon, nq, k, out_src, out_dst, dst_base, out_rdist, stats, stream);
    else
        launch_knn<65>(ix, chord2_init, q_latlon, nq, k, out_src, out_dst, dst_base, out_rdist, stats, stream);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}
