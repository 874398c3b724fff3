// K3: cut-off (radius) search, two-pass count -> scan -> fill.
//
// Replaces `radius_neighbors_graph(target, radius)` of CutOffEdges.get_adjacency_matrix
// (/root/reference/src/anemoi/graphs/edges/builder.py:364-366; sklearn _binary_tree.pxi.tp:1902-1979:
// a point is a neighbour iff rdist <= sin^2(radius/2), inclusive).
// One warp per query; the lanes stride over the contiguous record run of each window row (coalesced
// LDG.128), classify on the FP32 FMA pipe against chord^2 = 4 sin^2(r/2) -/+ the FP32 error margin, and
// only the pairs inside the margin evaluate the reference's float64 haversine.  The fill pass repeats
// the scan and writes (src, dst) with a ballot prefix, so the output order is the scan order:
// deterministic, grouped by query.
#include "agx_search.cuh"

template <bool FILL>
__global__ void __launch_bounds__(256) k_radius(const float4* __restrict__ pts, const int* __restrict__ cell_start,
                                                const float2* __restrict__ ref_latlon, int cells,
                                                const float2* __restrict__ q_latlon, int64_t nq, float chord2_thr,
                                                float chord2_margin, double rdist_thr, int32_t* __restrict__ counts,
                                                const int64_t* __restrict__ offsets, int32_t* __restrict__ out_src,
                                                int32_t* __restrict__ out_dst, int64_t dst_base,
                                                unsigned long long* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float t_in = chord2_thr - chord2_margin, t_out = chord2_thr + chord2_margin;
    const AgxCap cap = agx_make_cap(t_out);
    unsigned long long n_f64 = 0, n_boundary = 0;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += warps) {
        const float2 ql = q_latlon[q];
        const float3 qv = agx_search_xyz(ql);
        int64_t out_pos = FILL ? offsets[q] : 0;
        const int32_t dst = (int32_t)(dst_base + q);
        int found = 0;
        // single-face fast path (the cap ends at least a cell inside the query's major face), else all six faces
        const int mface = agx_major_face(qv.x, qv.y, qv.z);
        int fi0, fi1, fj0, fj1;
        bool fast;
        {
            float fa, fb, fc;
            agx_face_frame(mface, qv.x, qv.y, qv.z, fa, fb, fc);
            fast = !cap.everything && agx_axis_window_major(fa, fc, cap, cells, fi0, fi1);
            fast = fast && agx_axis_window_major(fb, fc, cap, cells, fj0, fj1);
        }
        const int face_end = fast ? mface + 1 : 6;
        for (int face = fast ? mface : 0; face < face_end; ++face) {
            int i0 = fi0, i1 = fi1, j0 = fj0, j1 = fj1;
            if (!fast && !agx_face_window(face, qv, cap, cells, i0, i1, j0, j1)) continue;
            for (int i = i0; i <= i1; ++i) {
                int row = (face * cells + i) * cells;
                int s = __ldg(cell_start + row + j0), e = __ldg(cell_start + row + j1 + 1);
                for (int base = s; base < e; base += 32) {
                    int p = base + lane;
                    bool hit = false;
                    int ci = 0;
                    if (p < e) {
                        float4 c = __ldg(pts + p);
                        ci = __float_as_int(c.w);
                        float dx = qv.x - c.x, dy = qv.y - c.y, dz = qv.z - c.z;
                        float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        if (d <= t_in) {
                            hit = true;
                        } else if (d <= t_out) {
                            double r = agx_rdist64(ql, ref_latlon[ci]);
                            hit = r <= rdist_thr;
                            if (FILL) {
                                ++n_f64;
                                if (fabs(r - rdist_thr) <= AGX_TIE_TAU * rdist_thr) ++n_boundary;
                            }
                        }
                    }
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (FILL) {
                        if (hit) {
                            int64_t w = out_pos + __popc(m & ((1u << lane) - 1u));
                            out_src[w] = ci;
                            out_dst[w] = dst;
                        }
                        out_pos += __popc(m);
                    } else {
                        found += __popc(m);
                    }
                }
            }
        }
        if (!FILL && lane == 0) counts[q] = found;
    }
    if (FILL && stats) {
        if (n_f64) atomicAdd(stats + 0, n_f64);
        if (n_boundary) atomicAdd(stats + 1, n_boundary);
    }
}

static int radius_params(double radius, float* chord2_thr, float* margin, double* rdist_thr) {
    AGX_REQUIRE(radius >= 0.0, AGX_ERR_ARG, "radius must be non-negative (got %g)", radius);
    // sklearn: reduced radius = sin(0.5 r)^2 (HaversineDistance64._dist_to_rdist)
    double s = sin(0.5 * (radius < 3.141592653589793 ? radius : 3.141592653589793));
    *rdist_thr = s * s;
    double c2 = 4.0 * (*rdist_thr);
    *chord2_thr = (float)c2;
    // same bound as agx_chord2_margin plus the rounding of the threshold itself
    *margin = (float)(4.2e-7 * sqrt(c2) + 1.2e-6 * c2 + 1.0e-13);
    return AGX_OK;
}

extern "C" int agx_radius_count(const agx_index_t* ix, const float* q_latlon, int64_t nq, double radius,
                                int32_t* counts, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(ix != nullptr, AGX_ERR_ARG, "agx_radius_count: NULL index");
    AGX_REQUIRE(nq >= 0, AGX_ERR_ARG, "agx_radius_count: nq < 0");
    float c2, m;
    double thr;
    int rc = radius_params(radius, &c2, &m, &thr);
    if (rc) return rc;
    if (nq == 0) return AGX_OK;
    AGX_REQUIRE(q_latlon && counts, AGX_ERR_ARG, "agx_radius_count: NULL buffer");
    int grid = agx_grid(nq * 32, 256, 8);
    k_radius<false><<<grid, 256, 0, stream>>>(ix->pts, ix->cell_start, ix->latlon, ix->cells, (const float2*)q_latlon,
                                             nq, c2, m, thr, counts, nullptr, nullptr, nullptr, 0, nullptr);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

extern "C" int agx_radius_fill(const agx_index_t* ix, const float* q_latlon, int64_t nq, double radius,
                               const int64_t* offsets, int32_t* out_src, int32_t* out_dst, int64_t dst_base,
                               int64_t* stats, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(ix != nullptr, AGX_ERR_ARG, "agx_radius_fill: NULL index");
    AGX_REQUIRE(nq >= 0, AGX_ERR_ARG, "agx_radius_fill: nq < 0");
    AGX_REQUIRE(dst_base + nq < (int64_t)2147483647, AGX_ERR_ARG, "agx_radius_fill: target index exceeds int32");
    float c2, m;
    double thr;
    int rc = radius_params(radius, &c2, &m, &thr);
    if (rc) return rc;
    if (nq == 0) return AGX_OK;
    AGX_REQUIRE(q_latlon && offsets && out_src && out_dst, AGX_ERR_ARG, "agx_radius_fill: NULL buffer");
    int grid = agx_grid(nq * 32, 256, 8);
    k_radius<true><<<grid, 256, 0, stream>>>(ix->pts, ix->cell_start, ix->latlon, ix->cells, (const float2*)q_latlon, nq,
                                            c2, m, thr, nullptr, offsets, out_src, out_dst, dst_base,
                                            (unsigned long long*)stats);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}
