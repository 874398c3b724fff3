// K5: fused edge attributes (EdgeLength + EdgeDirection) with global normalisation.
//
// Replaces EdgeLength.compute / EdgeDirection.compute
// (/root/reference/src/anemoi/graphs/edges/attributes.py:42-157; utils.py:84-103;
// edges/directional.py:19-94; generate/transforms.py:91-140; normalise.py:20-55).
//
// Arithmetic follows the reference stage by stage:
//  * float32 stage (numpy): latlon -> xyz and the haversine `a` term use numpy's float32 sin/cos
//    reproduced bit for bit (agx_np_sincosf) with un-fused float32 multiplies/adds.
//  * float64 stage (numpy/scipy): cross products, arccos, scipy Rotation.from_rotvec (small-angle
//    series for angle <= 1e-3), as_matrix, apply, the two epsilon nudges of direction_vec.
// Everything that depends on ONE node only (xyz, cos lat, the rotation quaternion of a target) is
// tabulated per node; the per-edge kernel is gather + ~60 flops + 12-byte store, HBM-bound.
#include "agx_common.cuh"

#define ATTR_THREADS 256
#define ATTR_MAX_BLOCKS 4096
#define ATTR_STAT_FIELDS 8  // len: sum, sumsq, min, max; dir: sum, sumsq, min, max

extern "C" int64_t agx_edge_attrs_workspace(void) { return (int64_t)ATTR_STAT_FIELDS * (ATTR_MAX_BLOCKS + 2); }

// ------------------------------------------------------------------------------------------------
// per-node tables
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_node_tables(const float2* __restrict__ latlon, int64_t n,
                                                      float4* __restrict__ xyzc, double* __restrict__ quat) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float2 ll = latlon[i];
        float sl, cl, so, co;
        agx_np_sincosf(ll.x, sl, cl);
        agx_np_sincosf(ll.y, so, co);
        // latlon_rad_to_cartesian (generate/transforms.py:106-110), radius = 1.0: float32 products
        float x = __fmul_rn(cl, co), y = __fmul_rn(cl, so), z = sl;
        xyzc[i] = make_float4(x, y, z, cl);
        if (quat != nullptr) {
            // get_rotation_from_unit_vecs(points=this node as TARGET, reference=(0,0,1))
            // direction_vec: v = cross(p, z^) = (p_y, -p_x, 0) in float64 from the float32 components
            double v0 = (double)y, v1 = -(double)x;
            double vn = __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
            float pz = z;
            if (vn < 10e-11) {  // generate/transforms.py:135-139: float32 in-place nudge of all components
                const float eps32 = (float)10e-11;
                float xn = __fadd_rn(x, eps32), yn = __fadd_rn(y, eps32);
                pz = __fadd_rn(z, eps32);
                v0 = (double)yn;
                v1 = -(double)xn;
                vn = __dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1));
            }
            double inv = sqrt(vn);
            double u0 = v0 / inv, u1 = v1 / inv;
            double theta = acos((double)pz);  // arccos(dot(points, reference)) = arccos(p_z)
            double r0 = __dmul_rn(u0, theta), r1 = __dmul_rn(u1, theta);
            // scipy Rotation.from_rotvec
            double angle = sqrt(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)));
            double scale;
            if (angle <= 1e-3) {
                double a2 = angle * angle;
                scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
            } else {
                scale = sin(angle / 2.0) / angle;
            }
            double4 qd = make_double4(scale * r0, scale * r1, cos(angle / 2.0), 0.0);
            reinterpret_cast<double4*>(quat)[i] = qd;
        }
    }
}

extern "C" int agx_node_tables(const float* latlon, int64_t n, float* xyzc, double* quat, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_node_tables: n < 0");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(latlon && xyzc, AGX_ERR_ARG, "agx_node_tables: NULL buffer");
    int grid = agx_grid(n, 256, 8);
    k_node_tables<<<grid, 256, 0, stream>>>((const float2*)latlon, n, (float4*)xyzc, quat);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------------
// per-edge raw values
// ------------------------------------------------------------------------------------------------
// utils.haversine_distance in float32 (numpy), last step in float64 then rounded.
__device__ __forceinline__ float edge_length_raw(float2 s, float cs, float2 t, float ct) {
    float dlat = __fsub_rn(t.x, s.x), dlon = __fsub_rn(t.y, s.y);
    float sh_lat, sh_lon, unused;
    agx_np_sincosf(__fmul_rn(dlat, 0.5f), sh_lat, unused);
    agx_np_sincosf(__fmul_rn(dlon, 0.5f), sh_lon, unused);
    float a = __fadd_rn(__fmul_rn(sh_lat, sh_lat), __fmul_rn(__fmul_rn(cs, ct), __fmul_rn(sh_lon, sh_lon)));
    float ra = __fsqrt_rn(a), rb = __fsqrt_rn(__fsub_rn(1.0f, a));
    return __fmul_rn(2.0f, (float)atan2((double)ra, (double)rb));
}

// compute_directions (edges/directional.py:40-65) for one edge; q = R(target) * source_xyz
__device__ __forceinline__ void edge_direction_rotated(float4 sxyz, double4 tq, double& o0, double& o1) {
    double x = tq.x, y = tq.y, w = tq.z;  // z component of the quaternion is exactly 0
    double x2 = x * x, y2 = y * y, w2 = w * w, xy = x * y, yw = y * w, xw = x * w;
    double sx = (double)sxyz.x, sy = (double)sxyz.y, sz = (double)sxyz.z;
    // scipy as_matrix rows 0 and 1 with z = 0
    double m00 = x2 - y2 + w2, m01 = 2.0 * xy, m02 = 2.0 * yw;
    double m10 = 2.0 * xy, m11 = -x2 + y2 + w2, m12 = -2.0 * xw;
    double qx = m00 * sx + m01 * sy + m02 * sz;
    double qy = m10 * sx + m11 * sy + m12 * sz;
    // direction_vec(q, z^): v = (q_y, -q_x, 0); nudge in float64 when |v|^2 < 1e-10
    double vn = qy * qy + qx * qx;
    if (vn < 10e-11) {
        qx += 10e-11;
        qy += 10e-11;
        vn = qy * qy + qx * qx;
    }
    double inv = rsqrt(vn);
    double d0 = qy * inv, d1 = -qx * inv;
    double inv2 = rsqrt(d0 * d0 + d1 * d1);  // the final renormalisation (edges/directional.py:65)
    o0 = d0 * inv2;
    o1 = d1 * inv2;
}

struct Stat4 {
    double sum, sumsq, mn, mx;
    __device__ __forceinline__ void init() {
        sum = 0.0;
        sumsq = 0.0;
        mn = 1e300;
        mx = -1e300;
    }
    __device__ __forceinline__ void add(double v) {
        sum += v;
        sumsq += v * v;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    __device__ __forceinline__ void merge(const Stat4& o) {
        sum += o.sum;
        sumsq += o.sumsq;
        mn = fmin(mn, o.mn);
        mx = fmax(mx, o.mx);
    }
};

__device__ __forceinline__ Stat4 warp_reduce(Stat4 s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Stat4 t;
        t.sum = __shfl_down_sync(0xffffffffu, s.sum, o);
        t.sumsq = __shfl_down_sync(0xffffffffu, s.sumsq, o);
        t.mn = __shfl_down_sync(0xffffffffu, s.mn, o);
        t.mx = __shfl_down_sync(0xffffffffu, s.mx, o);
        s.merge(t);
    }
    return s;
}

struct NormParams {  // out = (v - shift) * mul   (float64 path)   /   (v - shift) / div  (float32 path)
    double shift, mul;
    float shift32, div32;
};

template <bool STATS>
__global__ void __launch_bounds__(ATTR_THREADS) k_edge_attrs(
    const int32_t* __restrict__ esrc, const int32_t* __restrict__ edst, int64_t n_edges,
    const float2* __restrict__ s_ll, const float4* __restrict__ s_xyzc, const float2* __restrict__ t_ll,
    const float4* __restrict__ t_xyzc, const double2* __restrict__ t_quat, int want_len, int len_invert,
    float* __restrict__ out_len, int want_dir, int dir_rotated, float* __restrict__ out_dir,
    double* __restrict__ ws) {
    Stat4 st_len, st_dir;
    st_len.init();
    st_dir.init();
    NormParams np_len, np_dir;
    if (!STATS) {
        np_len.shift32 = (float)ws[0];
        np_len.div32 = (float)ws[1];
        np_dir.shift = ws[2];
        np_dir.mul = ws[3];
        np_dir.shift32 = (float)ws[4];
        np_dir.div32 = (float)ws[5];
    }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        int s = __ldg(esrc + e), t = __ldg(edst + e);
        float2 sl = __ldg(s_ll + s), tl = __ldg(t_ll + t);
        float4 sx = __ldg(s_xyzc + s);
        if (want_len) {
            float ct = __ldg(&t_xyzc[t].w);
            float v = edge_length_raw(sl, sx.w, tl, ct);
            if (STATS) {
                st_len.add((double)v);
            } else {
                v = __fdiv_rn(__fsub_rn(v, np_len.shift32), np_len.div32);
                if (len_invert) v = __fsub_rn(1.0f, v);
                out_len[e] = v;
            }
        }
        if (want_dir) {
            if (dir_rotated) {
                double2 qa = __ldg(t_quat + 2 * (int64_t)t), qb = __ldg(t_quat + 2 * (int64_t)t + 1);
                double4 tq = make_double4(qa.x, qa.y, qb.x, 0.0);
                double d0, d1;
                edge_direction_rotated(sx, tq, d0, d1);
                if (STATS) {
                    st_dir.add(d0);
                    st_dir.add(d1);
                } else {
                    float2 o = make_float2((float)((d0 - np_dir.shift) * np_dir.mul), (float)((d1 - np_dir.shift) * np_dir.mul));
                    reinterpret_cast<float2*>(out_dir)[e] = o;
                }
            } else {
                // directional_edge_features(..., relative_to_rotated_target=False): loc2 - loc1 in float32
                float d0 = __fsub_rn(tl.x, sl.x), d1 = __fsub_rn(tl.y, sl.y);
                if (STATS) {
                    st_dir.add((double)d0);
                    st_dir.add((double)d1);
                } else {
                    float2 o = make_float2(__fdiv_rn(__fsub_rn(d0, np_dir.shift32), np_dir.div32),
                                           __fdiv_rn(__fsub_rn(d1, np_dir.shift32), np_dir.div32));
                    reinterpret_cast<float2*>(out_dir)[e] = o;
                }
            }
        }
    }
    if (STATS) {
        __shared__ Stat4 sm[2][ATTR_THREADS / 32];
        st_len = warp_reduce(st_len);
        st_dir = warp_reduce(st_dir);
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
            sm[0][warp] = st_len;
            sm[1][warp] = st_dir;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < ATTR_THREADS / 32; ++w) {
                st_len.merge(sm[0][w]);
                st_dir.merge(sm[1][w]);
            }
            double* p = ws + (int64_t)ATTR_STAT_FIELDS * (2 + blockIdx.x);
            p[0] = st_len.sum; p[1] = st_len.sumsq; p[2] = st_len.mn; p[3] = st_len.mx;
            p[4] = st_dir.sum; p[5] = st_dir.sumsq; p[6] = st_dir.mn; p[7] = st_dir.mx;
        }
    }
}

// Fold the per-block partials into stats[8] = {len sum, sumsq, min, max, dir sum, sumsq, min, max}.
// One 256-thread block: thread t folds partials t, t+256, ... in order, then a fixed shuffle/shared tree -
// the association order depends only on n_blocks, so the result is reproducible run to run.
__global__ void __launch_bounds__(256) k_attr_fold(const double* __restrict__ ws, int n_blocks, double* __restrict__ stats) {
    Stat4 L, D;
    L.init();
    D.init();
    for (int b = threadIdx.x; b < n_blocks; b += 256) {
        const double* p = ws + (int64_t)ATTR_STAT_FIELDS * (2 + b);
        Stat4 l = {p[0], p[1], p[2], p[3]}, d = {p[4], p[5], p[6], p[7]};
        L.merge(l);
        D.merge(d);
    }
    __shared__ Stat4 sm[2][8];
    L = warp_reduce(L);
    D = warp_reduce(D);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sm[0][warp] = L;
        sm[1][warp] = D;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            L.merge(sm[0][w]);
            D.merge(sm[1][w]);
        }
        stats[0] = L.sum; stats[1] = L.sumsq; stats[2] = L.mn; stats[3] = L.mx;
        stats[4] = D.sum; stats[5] = D.sumsq; stats[6] = D.mn; stats[7] = D.mx;
    }
}

// Global statistics -> normalisation parameters (normalise.py:33-52).
// ws[0..1] = length (shift, div); ws[2..5] = direction (shift, 1/div, shift, div).
__global__ void k_attr_params(double* __restrict__ ws, const double* __restrict__ stats, int64_t n_edges, int len_norm,
                              int dir_norm) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int which = 0; which < 2; ++which) {
        int norm = which ? dir_norm : len_norm;
        double shift = 0.0, div = 1.0;
        if (norm > 0) {
            const double* S = stats + 4 * which;  // sum, sumsq, min, max
            double count = (double)n_edges * (which ? 2.0 : 1.0);
            if (norm == AGX_NORM_L1) div = S[0];
            else if (norm == AGX_NORM_L2) div = sqrt(S[1]);
            else if (norm == AGX_NORM_UNIT_MAX) div = S[3];
            else if (norm == AGX_NORM_UNIT_RANGE) { shift = S[2]; div = S[3] - S[2]; }
            else if (norm == AGX_NORM_UNIT_STD) {
                double mean = S[0] / count;
                double var = S[1] / count - mean * mean;
                double sd = var > 0.0 ? sqrt(var) : 0.0;
                div = sd == 0.0 ? 1.0 : sd;  // normalise.py:46-50: skipped when std == 0
            }
        }
        if (which == 0) {
            ws[0] = shift;
            ws[1] = div;
        } else {
            ws[2] = shift;
            ws[3] = 1.0 / div;
            ws[4] = shift;
            ws[5] = div;
        }
    }
}

static int attrs_check(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges, const float* src_latlon,
                       const float* src_xyzc, const float* dst_latlon, const float* dst_xyzc, const double* dst_quat,
                       int want_dir, int dir_rotated, const double* workspace) {
    AGX_REQUIRE(n_edges >= 0, AGX_ERR_ARG, "agx_edge_attrs: n_edges < 0");
    if (n_edges == 0) return AGX_OK;
    AGX_REQUIRE(edge_src && edge_dst && src_latlon && src_xyzc && dst_latlon && dst_xyzc && workspace, AGX_ERR_ARG,
                "agx_edge_attrs: NULL buffer");
    AGX_REQUIRE(!(want_dir && dir_rotated) || dst_quat, AGX_ERR_ARG, "agx_edge_attrs: rotated directions need dst_quat");
    return AGX_OK;
}

static inline int attrs_grid(int64_t n_edges) {
    int grid = agx_grid(n_edges, ATTR_THREADS, 8);
    return grid > ATTR_MAX_BLOCKS ? ATTR_MAX_BLOCKS : grid;
}

extern "C" int agx_edge_attrs_stats(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                                    const float* src_latlon, const float* src_xyzc, const float* dst_latlon,
                                    const float* dst_xyzc, const double* dst_quat, int want_len, int want_dir,
                                    int dir_rotated, double* stats, double* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(stats != nullptr, AGX_ERR_ARG, "agx_edge_attrs_stats: stats is NULL");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_latlon, src_xyzc, dst_latlon, dst_xyzc, dst_quat, want_dir,
                         dir_rotated, workspace);
    if (rc) return rc;
    int grid = 0;
    if (n_edges > 0 && (want_len || want_dir)) {
        grid = attrs_grid(n_edges);
        k_edge_attrs<true><<<grid, ATTR_THREADS, 0, stream>>>(
            edge_src, edge_dst, n_edges, (const float2*)src_latlon, (const float4*)src_xyzc, (const float2*)dst_latlon,
            (const float4*)dst_xyzc, (const double2*)dst_quat, want_len, 0, nullptr, want_dir, dir_rotated, nullptr,
            workspace);
        agx_note_launch(1);
    }
    k_attr_fold<<<1, 256, 0, stream>>>(workspace, grid, stats);  // grid == 0: the empty statistics (a rank with no edges)
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

extern "C" int agx_edge_attrs_apply(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                                    const float* src_latlon, const float* src_xyzc, const float* dst_latlon,
                                    const float* dst_xyzc, const double* dst_quat, int len_norm, int len_invert,
                                    float* out_len, int dir_norm, int dir_rotated, float* out_dir, const double* stats,
                                    int64_t n_edges_global, double* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int want_len = len_norm >= 0, want_dir = dir_norm >= 0;
    AGX_REQUIRE(len_norm <= AGX_NORM_UNIT_STD && dir_norm <= AGX_NORM_UNIT_STD, AGX_ERR_ARG, "agx_edge_attrs: unknown norm code");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_latlon, src_xyzc, dst_latlon, dst_xyzc, dst_quat, want_dir,
                         dir_rotated, workspace);
    if (rc) return rc;
    if (n_edges == 0 || (!want_len && !want_dir)) return AGX_OK;
    AGX_REQUIRE(!want_len || out_len, AGX_ERR_ARG, "agx_edge_attrs: out_len is NULL");
    AGX_REQUIRE(!want_dir || out_dir, AGX_ERR_ARG, "agx_edge_attrs: out_dir is NULL");
    bool need_stats = (want_len && len_norm > 0) || (want_dir && dir_norm > 0);
    AGX_REQUIRE(!need_stats || stats, AGX_ERR_ARG, "agx_edge_attrs_apply: this normalisation needs the global statistics");
    AGX_REQUIRE(n_edges_global >= n_edges, AGX_ERR_ARG, "agx_edge_attrs_apply: n_edges_global < n_edges");
    k_attr_params<<<1, 32, 0, stream>>>(workspace, stats, n_edges_global, want_len ? len_norm : 0, want_dir ? dir_norm : 0);
    k_edge_attrs<false><<<attrs_grid(n_edges), ATTR_THREADS, 0, stream>>>(
        edge_src, edge_dst, n_edges, (const float2*)src_latlon, (const float4*)src_xyzc, (const float2*)dst_latlon,
        (const float4*)dst_xyzc, (const double2*)dst_quat, want_len, len_invert, out_len, want_dir, dir_rotated, out_dir,
        workspace);
    AGX_LAUNCH_OK();
    agx_note_launch(2);
    return AGX_OK;
}

extern "C" int agx_edge_attrs(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                              const float* src_latlon, const float* src_xyzc, const float* dst_latlon,
                              const float* dst_xyzc, const double* dst_quat, int len_norm, int len_invert,
                              float* out_len, int dir_norm, int dir_rotated, float* out_dir, double* workspace,
                              void* stream_) {
    int want_len = len_norm >= 0, want_dir = dir_norm >= 0;
    bool need_stats = (want_len && len_norm > 0) || (want_dir && dir_norm > 0);
    double* stats = workspace ? workspace + 6 : nullptr;  // ws[6..13]: between the parameters and the block partials
    if (need_stats && n_edges > 0) {
        int rc = agx_edge_attrs_stats(edge_src, edge_dst, n_edges, src_latlon, src_xyzc, dst_latlon, dst_xyzc, dst_quat,
                                      want_len && len_norm > 0, want_dir && dir_norm > 0, dir_rotated, stats, workspace,
                                      stream_);
        if (rc) return rc;
    }
    return agx_edge_attrs_apply(edge_src, edge_dst, n_edges, src_latlon, src_xyzc, dst_latlon, dst_xyzc, dst_quat,
                                len_norm, len_invert, out_len, dir_norm, dir_rotated, out_dir, stats, n_edges, workspace,
                                stream_);
}
