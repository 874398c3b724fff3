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
// tabulated per node in ONE 32-byte record per role, so an edge costs exactly two gathered sectors:
//   source record  float[8]  = (x, y, z, cos lat, lat, lon, 0, 0)
//   target record  double[4] = (quat x, quat y, quat w, bits(lat, lon))   (quat z is exactly 0)
// The per-edge kernel works on warp tiles of 32*J consecutive edges: J coalesced index loads per lane, then
// all 2*J record gathers (LDG.128 pairs) in flight before the arithmetic - the kernel is latency-bound on
// the dependent index -> record chain, so memory-level parallelism per warp is what buys bandwidth.
#include <stdlib.h>

#include "agx_common.cuh"

#define ATTR_THREADS 256
#define ATTR_MAX_BLOCKS 4096
#define ATTR_STAT_FIELDS 8  // len: sum, sumsq, min, max; dir: sum, sumsq, min, max
#define AGX_ATTR_FLAGS_SKIP 1
#define AGX_ATTR_FLAGS_ONLY 2

extern "C" int64_t agx_edge_attrs_workspace(void) { return (int64_t)ATTR_STAT_FIELDS * (ATTR_MAX_BLOCKS + 2); }

// ------------------------------------------------------------------------------------------------
// per-node tables
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_node_tables(const float2* __restrict__ latlon, int64_t n,
                                                      float4* __restrict__ src_rec, double4* __restrict__ dst_rec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float2 ll = latlon[i];
        float sl, cl, so, co;
        agx_np_sincosf(ll.x, sl, cl);
        agx_np_sincosf(ll.y, so, co);
        // latlon_rad_to_cartesian (generate/transforms.py:106-110), radius = 1.0: float32 products
        float x = __fmul_rn(cl, co), y = __fmul_rn(cl, so), z = sl;
        if (src_rec != nullptr) {
            src_rec[2 * i] = make_float4(x, y, z, cl);
            src_rec[2 * i + 1] = make_float4(ll.x, ll.y, 0.0f, 0.0f);
        }
        if (dst_rec != nullptr) {
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
            long long packed = ((long long)__float_as_int(ll.y) << 32) | (unsigned int)__float_as_int(ll.x);
            dst_rec[i] = make_double4(scale * r0, scale * r1, cos(angle / 2.0), __longlong_as_double(packed));
        }
    }
}

extern "C" int agx_node_tables(const float* latlon, int64_t n, float* src_rec, double* dst_rec, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n >= 0, AGX_ERR_ARG, "agx_node_tables: n < 0");
    if (n == 0) return AGX_OK;
    AGX_REQUIRE(latlon && (src_rec || dst_rec), AGX_ERR_ARG, "agx_node_tables: NULL buffer");
    AGX_REQUIRE(((uintptr_t)src_rec & 15) == 0 && ((uintptr_t)dst_rec & 31) == 0, AGX_ERR_ARG,
                "agx_node_tables: records must be 16-byte (source) / 32-byte (target) aligned");
    int grid = agx_grid(n, 256, 8);
    k_node_tables<<<grid, 256, 0, stream>>>((const float2*)latlon, n, (float4*)src_rec, (double4*)dst_rec);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

// ------------------------------------------------------------------------------------------------
// per-edge raw values
// ------------------------------------------------------------------------------------------------
// atan2(ra, rb) for ra, rb >= 0 in float64.  Edges are short, so ra/rb = tan(d/2) is tiny: below 1/16 the odd
// Taylor series to t^11 is exact to < 3e-16 relative (next term t^12/13) and costs a handful of DFMAs instead of
// the library's ~150-instruction path.  The quotient uses a float reciprocal seed + two Newton steps (1e-14
// relative); the result is rounded to float32 by the caller, so this cannot be told from the exact quotient.
__device__ __forceinline__ double atan2_short(float ra32, float rb32) {
    double ra = (double)ra32, rb = (double)rb32;
    if (ra <= 0.0625 * rb) {
        float seed;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"(rb32));  // 2^-23 relative; two Newton steps -> 2^-90
        double r = (double)seed;
        r = r * (2.0 - rb * r);
        r = r * (2.0 - rb * r);
        double t = ra * r, t2 = t * t;
        double p = fma(t2, -1.0 / 11.0, 1.0 / 9.0);
        p = fma(t2, -p, 1.0 / 7.0);
        p = fma(t2, -p, 1.0 / 5.0);
        p = fma(t2, -p, 1.0 / 3.0);
        p = fma(t2, -p, 1.0);
        return t * p;
    }
    return atan2(ra, rb);  // long edges, and NaN (a > 1 in float32: antipodal pairs, as in the reference)
}

// utils.haversine_distance in float32 (numpy), last step in float64 then rounded.
__device__ __forceinline__ float edge_length_raw(float2 s, float cs, float2 t, float ct) {
    float dlat = __fsub_rn(t.x, s.x), dlon = __fsub_rn(t.y, s.y);
    float sh_lat = agx_np_sinf_short(__fmul_rn(dlat, 0.5f));
    float sh_lon = agx_np_sinf_short(__fmul_rn(dlon, 0.5f));
    float a = __fadd_rn(__fmul_rn(sh_lat, sh_lat), __fmul_rn(__fmul_rn(cs, ct), __fmul_rn(sh_lon, sh_lon)));
    float ra = __fsqrt_rn(a), rb = __fsqrt_rn(__fsub_rn(1.0f, a));
    return __fmul_rn(2.0f, (float)atan2_short(ra, rb));
}

// compute_directions (edges/directional.py:40-65) for one edge; q = R(target) * source_xyz.
// For the unit quaternion (x, y, 0, w) scipy's as_matrix rows 0 and 1 are (1 - 2y^2, 2xy, 2yw) and
// (2xy, 1 - 2x^2, -2xw) (it writes them as x^2 - y^2 + w^2 etc., equal to 1e-16), so with
// T = x s_y - y s_x + w s_z:  q_x = s_x + 2 y T,  q_y = s_y - 2 x T  - seven float64 operations.  For nearby
// endpoints q_x, q_y are O(separation) differences of O(1) terms; each rounding is <= 1.1e-16 absolute, i.e.
// <= 1e-12 relative at the shortest edges of an O1280 graph (separation 1e-3).
__device__ __forceinline__ void edge_direction_rotated(float4 sxyz, double x, double y, double w, double& o0, double& o1) {
    double sx = (double)sxyz.x, sy = (double)sxyz.y, sz = (double)sxyz.z;
    double T2 = 2.0 * fma(x, sy, fma(-y, sx, w * sz));
    double qx = fma(y, T2, sx);
    double qy = fma(-x, T2, sy);
    // direction_vec(q, z^): v = (q_y, -q_x, 0); nudge in float64 when |v|^2 < 1e-10
    double vn = fma(qy, qy, qx * qx);
    if (vn < 10e-11) {
        qx += 10e-11;
        qy += 10e-11;
        vn = fma(qy, qy, qx * qx);
    }
    // v / |v|; the reference normalises once more (edges/directional.py:65), which moves a unit vector by at most
    // an ulp of float64 - invisible after the float32 cast.  NaN inputs propagate as they do there.
    double inv = rsqrt(vn);
    o0 = qy * inv;
    o1 = -qx * inv;
}

// Running statistics of one attribute.  Sums are float64 (the reference reduces a float64 array for directions and
// numpy's pairwise float32 sum for lengths - both within 1e-7 of this).  Minimum / maximum are tracked on the
// float32 value that is stored, one FMNMX each: for lengths that IS the reference's value; for rotated directions
// the reference takes them on the float64 array, whose float32 rounding differs by <= 6e-8 relative.
struct Stat4 {
    double sum, sumsq;
    float mn, mx;
    __device__ __forceinline__ void init() {
        sum = 0.0;
        sumsq = 0.0;
        mn = __int_as_float(0x7f800000);
        mx = __int_as_float(0xff800000);
    }
    __device__ __forceinline__ void add(double v, float v32) {
        sum += v;
        sumsq = fma(v, v, sumsq);
        mn = fminf(mn, v32);
        mx = fmaxf(mx, v32);
    }
    __device__ __forceinline__ void merge(const Stat4& o) {
        sum += o.sum;
        sumsq += o.sumsq;
        mn = fminf(mn, o.mn);
        mx = fmaxf(mx, o.mx);
    }
    // the workspace / C-ABI form: {sum, sumsq, min, max} as doubles, +-1e300 for "no value"
    __device__ __forceinline__ void store(double* p) const {
        p[0] = sum;
        p[1] = sumsq;
        p[2] = mn == __int_as_float(0x7f800000) ? 1e300 : (double)mn;
        p[3] = mx == __int_as_float(0xff800000) ? -1e300 : (double)mx;
    }
};

// the same four fields in float64, for folding per-block partials
struct Stat4D {
    double sum, sumsq, mn, mx;
    __device__ __forceinline__ void init() {
        sum = 0.0;
        sumsq = 0.0;
        mn = 1e300;
        mx = -1e300;
    }
    __device__ __forceinline__ void merge(const Stat4D& o) {
        sum += o.sum;
        sumsq += o.sumsq;
        mn = fmin(mn, o.mn);
        mx = fmax(mx, o.mx);
    }
};

template <class S>
__device__ __forceinline__ S warp_reduce(S s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        S t;
        t.sum = __shfl_down_sync(0xffffffffu, s.sum, o);
        t.sumsq = __shfl_down_sync(0xffffffffu, s.sumsq, o);
        t.mn = __shfl_down_sync(0xffffffffu, s.mn, o);
        t.mx = __shfl_down_sync(0xffffffffu, s.mx, o);
        s.merge(t);
    }
    return s;
}

// STATS: reduce the raw values (a shard's statistics, or the single-GPU first pass).
// WRITE: store the raw float32 values - the normalisation follows as an in-place scaling pass (k_attr_scale), so
//        the trigonometry and the gathers run ONCE per edge.
// J edges per lane, interleaved by 32 so that every index load / store of the warp is one contiguous run.
template <bool STATS, bool WRITE, int J>
__global__ void __launch_bounds__(ATTR_THREADS, J == 4 ? 2 : (J == 2 ? 3 : 4)) k_edge_attrs(
    const int32_t* __restrict__ esrc, const int32_t* __restrict__ edst, int64_t n_edges,
    const float4* __restrict__ s_rec, const double2* __restrict__ t_rec, int want_len, int len_invert_now,
    float* __restrict__ out_len, int want_dir, int dir_rotated, float* __restrict__ out_dir, double* __restrict__ ws,
    const uint8_t* __restrict__ dst_flags, int flag_mode, const int32_t* __restrict__ only_list,
    const int64_t* __restrict__ only_count, int regular_k) {
    // flag_mode (with dst_flags, one byte per TARGET node): AGX_ATTR_FLAGS_SKIP = edges into a flagged target are
    // written but left out of the statistics; AGX_ATTR_FLAGS_ONLY = only those edges are evaluated at all (the rest is
    // neither read beyond its target index, nor written, nor counted).  KNN edges whose source set is re-decided
    // once the node order is known use the pair: everything now, the re-decided queries again later.
    Stat4 st_len, st_dir;
    st_len.init();
    st_dir.init();
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // only_list (with regular_k: edges of target t are [t k, (t + 1) k), a KNN result): the work items are the edges of
    // the listed targets only - item w is edge only_list[w / k] * k + w % k - instead of all n_edges
    const int64_t n_work = only_list ? *only_count * regular_k : n_edges;
    const int64_t n_tiles = (n_work + 32 * J - 1) / (32 * J);
    for (int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < n_tiles; tile += n_warps) {
        const int64_t base = tile * (32 * J) + lane;
        int s[J], t[J];
        int64_t eid[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            int64_t w = base + 32 * j;
            w = w < n_work ? w : n_work - 1;  // tail lanes repeat the last item (never stored, never counted)
            int64_t e = only_list ? (int64_t)only_list[w / regular_k] * regular_k + (w % regular_k) : w;
            eid[j] = e;
            s[j] = __ldg(esrc + e);
            t[j] = __ldg(edst + e);
        }
        bool flagged[J];
#pragma unroll
        for (int j = 0; j < J; ++j) flagged[j] = flag_mode != 0 && dst_flags[t[j]] != 0;
        if (flag_mode == AGX_ATTR_FLAGS_ONLY) {
            bool any = false;
#pragma unroll
            for (int j = 0; j < J; ++j) any |= flagged[j] && (base + 32 * j < n_work);
            if (!__any_sync(0xffffffffu, any)) continue;
        }
        float4 sx[J];
        float2 sl[J];
        double2 qa[J], qb[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            sx[j] = __ldg(s_rec + 2 * (int64_t)s[j]);
            sl[j] = __ldg(reinterpret_cast<const float2*>(s_rec + 2 * (int64_t)s[j] + 1));
            qa[j] = __ldg(t_rec + 2 * (int64_t)t[j]);
            qb[j] = __ldg(t_rec + 2 * (int64_t)t[j] + 1);
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t e = eid[j];
            const bool live = (base + 32 * j < n_work) && (flag_mode != AGX_ATTR_FLAGS_ONLY || flagged[j]);
            const bool counted = live && (flag_mode != AGX_ATTR_FLAGS_SKIP || !flagged[j]);
            long long packed = __double_as_longlong(qb[j].y);
            float2 tl = make_float2(__int_as_float((int)(packed & 0xffffffffll)), __int_as_float((int)(packed >> 32)));
            if (want_len) {
                float st_unused, ct;
                agx_np_sincosf(tl.x, st_unused, ct);  // numpy's float32 cos(lat) of the target
                float v = edge_length_raw(sl[j], sx[j].w, tl, ct);
                if (STATS && counted) st_len.add((double)v, v);
                if (WRITE && live) out_len[e] = len_invert_now ? __fsub_rn(1.0f, v) : v;
            }
            if (want_dir) {
                if (dir_rotated) {
                    double d0, d1;
                    edge_direction_rotated(sx[j], qa[j].x, qa[j].y, qb[j].x, d0, d1);
                    float f0 = (float)d0, f1 = (float)d1;
                    if (STATS && counted) {
                        st_dir.add(d0, f0);
                        st_dir.add(d1, f1);
                    }
                    if (WRITE && live) reinterpret_cast<float2*>(out_dir)[e] = make_float2(f0, f1);
                } else {
                    // directional_edge_features(..., relative_to_rotated_target=False): loc2 - loc1 in float32
                    float d0 = __fsub_rn(tl.x, sl[j].x), d1 = __fsub_rn(tl.y, sl[j].y);
                    if (STATS && counted) {
                        st_dir.add((double)d0, d0);
                        st_dir.add((double)d1, d1);
                    }
                    if (WRITE && live) reinterpret_cast<float2*>(out_dir)[e] = make_float2(d0, d1);
                }
            }
        }
    }
    if (STATS) {
        __shared__ Stat4 sm[2][ATTR_THREADS / 32];
        st_len = warp_reduce(st_len);
        st_dir = warp_reduce(st_dir);
        int warp = threadIdx.x >> 5;
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
            st_len.store(p);
            st_dir.store(p + 4);
        }
    }
}

// In-place normalisation of the raw float32 attributes: out = (v - shift) / div (+ optional 1 - v).
// EdgeLength and non-rotated directions are float32 arrays in the reference, so the division is float32
// (normalise.py on a float32 array); rotated directions are float64 there and cast last, so the arithmetic is
// float64 on the stored float32 value (one extra rounding, <= 6e-8 relative).  A flat array of n floats is
// processed as 128-bit vectors between an aligned head and tail.
template <bool F64>
__device__ __forceinline__ float scale_one(float v, float shift32, float div32, double shift, double mul, int invert) {
    float r = F64 ? (float)(((double)v - shift) * mul) : __fdiv_rn(__fsub_rn(v, shift32), div32);
    return invert ? __fsub_rn(1.0f, r) : r;
}

template <bool F64>
__device__ __forceinline__ void scale_array(float* __restrict__ a, int64_t n, float shift32, float div32, double shift,
                                            double mul, int invert) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    int64_t head = ((16 - ((uintptr_t)a & 15)) & 15) >> 2;  // floats before the first 16-byte boundary
    if (head > n) head = n;
    int64_t n_vec = (n - head) >> 2;
    float4* v4 = reinterpret_cast<float4*>(a + head);
    for (int64_t i = tid; i < n_vec; i += stride) {
        float4 v = v4[i];
        v.x = scale_one<F64>(v.x, shift32, div32, shift, mul, invert);
        v.y = scale_one<F64>(v.y, shift32, div32, shift, mul, invert);
        v.z = scale_one<F64>(v.z, shift32, div32, shift, mul, invert);
        v.w = scale_one<F64>(v.w, shift32, div32, shift, mul, invert);
        v4[i] = v;
    }
    int64_t tail0 = head + (n_vec << 2);
    for (int64_t i = tid; i < head + (n - tail0); i += stride) {
        int64_t j = i < head ? i : tail0 + (i - head);
        a[j] = scale_one<F64>(a[j], shift32, div32, shift, mul, invert);
    }
}

__global__ void __launch_bounds__(256) k_attr_scale(float* __restrict__ out_len, int64_t n_len, int len_scale,
                                                     int len_invert, float* __restrict__ out_dir, int64_t n_dir,
                                                     int dir_scale, int dir_f64, const double* __restrict__ ws) {
    if (out_len != nullptr && (len_scale || len_invert))
        scale_array<false>(out_len, n_len, len_scale ? (float)ws[0] : 0.0f, len_scale ? (float)ws[1] : 1.0f, 0.0, 1.0,
                           len_invert);
    if (out_dir != nullptr && dir_scale) {
        if (dir_f64)
            scale_array<true>(out_dir, n_dir, 0.0f, 1.0f, ws[2], ws[3], 0);
        else
            scale_array<false>(out_dir, n_dir, (float)ws[4], (float)ws[5], 0.0, 1.0, 0);
    }
}

// Fold the per-block partials into stats[8] = {len sum, sumsq, min, max, dir sum, sumsq, min, max}.
// One 256-thread block: thread t folds partials t, t+256, ... in order, then a fixed shuffle/shared tree -
// the association order depends only on n_blocks, so the result is reproducible run to run.
__global__ void __launch_bounds__(256) k_attr_fold(const double* __restrict__ ws, int n_blocks, double* __restrict__ stats) {
    Stat4D L, D;
    L.init();
    D.init();
    for (int b = threadIdx.x; b < n_blocks; b += 256) {
        const double* p = ws + (int64_t)ATTR_STAT_FIELDS * (2 + b);
        Stat4D l = {p[0], p[1], p[2], p[3]}, d = {p[4], p[5], p[6], p[7]};
        L.merge(l);
        D.merge(d);
    }
    __shared__ Stat4D sm[2][8];
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
__global__ void k_attr_params(double* __restrict__ ws, const double* __restrict__ stats_sets, int n_sets, int64_t n_edges,
                              int len_norm, int dir_norm) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // fold the per-shard statistics in order (one set on a single GPU): the same association on every rank
    double stats[8] = {0.0, 0.0, 1e300, -1e300, 0.0, 0.0, 1e300, -1e300};
    for (int r = 0; r < n_sets; ++r)
        for (int b = 0; b < 8; b += 4) {
            const double* p = stats_sets + 8 * r + b;
            stats[b] += p[0];
            stats[b + 1] += p[1];
            stats[b + 2] = fmin(stats[b + 2], p[2]);
            stats[b + 3] = fmax(stats[b + 3], p[3]);
        }
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

static int attrs_check(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges, const float* src_rec,
                       const double* dst_rec, const double* workspace) {
    AGX_REQUIRE(n_edges >= 0, AGX_ERR_ARG, "agx_edge_attrs: n_edges < 0");
    if (n_edges == 0) return AGX_OK;
    AGX_REQUIRE(edge_src && edge_dst && src_rec && dst_rec && workspace, AGX_ERR_ARG, "agx_edge_attrs: NULL buffer");
    AGX_REQUIRE(((uintptr_t)src_rec & 15) == 0 && ((uintptr_t)dst_rec & 15) == 0, AGX_ERR_ARG,
                "agx_edge_attrs: node records must be 16-byte aligned");
    return AGX_OK;
}

// edges per lane and tile (AGX_ATTR_J=1|2|4 overrides, for tuning)
static inline int attrs_j() {
    static int j = 0;
    if (j == 0) {
        j = 2;
        if (const char* env = getenv("AGX_ATTR_J")) {
            int v = atoi(env);
            if (v == 1 || v == 2 || v == 4) j = v;
        }
    }
    return j;
}

static inline int attrs_grid(int64_t n_edges, int j) {
    int grid = agx_grid((n_edges + j - 1) / j, ATTR_THREADS, j == 4 ? 2 : (j == 2 ? 3 : 4));
    return grid > ATTR_MAX_BLOCKS ? ATTR_MAX_BLOCKS : grid;
}

#define ATTR_KERNEL_ARGS edge_src, edge_dst, n_edges, (const float4*)src_rec, (const double2*)dst_rec

template <int J>
static void attrs_launch(bool stats, bool write, int grid, cudaStream_t stream, const int32_t* edge_src,
                         const int32_t* edge_dst, int64_t n_edges, const float* src_rec, const double* dst_rec,
                         int want_len, int len_invert_now, float* out_len, int want_dir, int dir_rotated, float* out_dir,
                         double* workspace, const uint8_t* dst_flags, int flag_mode, const int32_t* only_list = nullptr,
                         const int64_t* only_count = nullptr, int regular_k = 0) {
    if (stats && write)
        k_edge_attrs<true, true, J><<<grid, ATTR_THREADS, 0, stream>>>(ATTR_KERNEL_ARGS, want_len, len_invert_now, out_len,
                                                                      want_dir, dir_rotated, out_dir, workspace, dst_flags,
                                                                      flag_mode, only_list, only_count, regular_k);
    else if (stats)
        k_edge_attrs<true, false, J><<<grid, ATTR_THREADS, 0, stream>>>(ATTR_KERNEL_ARGS, want_len, 0, nullptr, want_dir,
                                                                       dir_rotated, nullptr, workspace, dst_flags,
                                                                       flag_mode, only_list, only_count, regular_k);
    else
        k_edge_attrs<false, true, J><<<grid, ATTR_THREADS, 0, stream>>>(ATTR_KERNEL_ARGS, want_len, len_invert_now, out_len,
                                                                       want_dir, dir_rotated, out_dir, workspace, dst_flags,
                                                                       flag_mode, only_list, only_count, regular_k);
}

// pass A: raw values of the local edges -> out_* (float32) and/or stats[8]
static int attrs_raw(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges, const float* src_rec,
                     const double* dst_rec, int want_len, int len_invert_now, float* out_len, int want_dir,
                     int dir_rotated, float* out_dir, bool write, double* stats, double* workspace, cudaStream_t stream,
                     const uint8_t* dst_flags = nullptr, int flag_mode = 0) {
    int grid = 0;
    if (n_edges > 0 && (want_len || want_dir)) {
        int j = attrs_j();
        grid = attrs_grid(n_edges, j);
#define ATTR_LAUNCH(J)                                                                                                \
    attrs_launch<J>(stats != nullptr, write, grid, stream, edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len,      \
                    len_invert_now, out_len, want_dir, dir_rotated, out_dir, workspace, dst_flags, flag_mode)
        if (j == 1)
            ATTR_LAUNCH(1);
        else if (j == 2)
            ATTR_LAUNCH(2);
        else
            ATTR_LAUNCH(4);
#undef ATTR_LAUNCH
        agx_note_launch(1);
    }
    if (stats) {
        k_attr_fold<<<1, 256, 0, stream>>>(workspace, grid, stats);  // grid == 0: the empty statistics
        agx_note_launch(1);
    }
    AGX_LAUNCH_OK();
    return AGX_OK;
}

// pass B: statistics -> parameters, in-place scaling of the raw values
static int attrs_scale(int64_t n_edges, int len_norm, int len_invert, float* out_len, int dir_norm, int dir_rotated,
                       float* out_dir, const double* stats, int n_stat_sets, int64_t n_edges_global, double* workspace,
                       cudaStream_t stream) {
    int want_len = len_norm >= 0, want_dir = dir_norm >= 0;
    bool len_scale = want_len && len_norm > 0, dir_scale = want_dir && dir_norm > 0;
    if (n_edges == 0 || !(len_scale || dir_scale || (want_len && len_invert))) return AGX_OK;
    k_attr_params<<<1, 32, 0, stream>>>(workspace, stats, n_stat_sets, n_edges_global, want_len ? len_norm : 0,
                                        want_dir ? dir_norm : 0);
    int64_t work = (n_edges * (want_dir ? 2 : 1) + 3) / 4;
    k_attr_scale<<<agx_grid(work, 256, 8), 256, 0, stream>>>(want_len ? out_len : nullptr, n_edges, len_scale, len_invert,
                                                           want_dir ? out_dir : nullptr, 2 * n_edges, dir_scale,
                                                           dir_rotated, workspace);
    AGX_LAUNCH_OK();
    agx_note_launch(2);
    return AGX_OK;
}

extern "C" int agx_edge_attrs_stats(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                                    const float* src_rec, const double* dst_rec, int want_len, int want_dir,
                                    int dir_rotated, float* out_len, float* out_dir, double* stats, double* workspace,
                                    void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(stats != nullptr, AGX_ERR_ARG, "agx_edge_attrs_stats: stats is NULL");
    AGX_REQUIRE(workspace != nullptr, AGX_ERR_ARG, "agx_edge_attrs_stats: workspace is NULL");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_rec, dst_rec, workspace);
    if (rc) return rc;
    bool write = (want_len && out_len) || (want_dir && out_dir);
    AGX_REQUIRE(!write || ((!want_len || out_len) && (!want_dir || out_dir)), AGX_ERR_ARG,
                "agx_edge_attrs_stats: give every requested output buffer or none");
    return attrs_raw(edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len, 0, out_len, want_dir, dir_rotated, out_dir,
                     write, stats, workspace, stream);
}

extern "C" int agx_edge_attrs_stats_flagged(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                                            const float* src_rec, const double* dst_rec, int want_len, int want_dir,
                                            int dir_rotated, float* out_len, float* out_dir, double* stats,
                                            double* workspace, const uint8_t* dst_flags, int flag_mode, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(stats != nullptr, AGX_ERR_ARG, "agx_edge_attrs_stats_flagged: stats is NULL");
    AGX_REQUIRE(workspace != nullptr, AGX_ERR_ARG, "agx_edge_attrs_stats_flagged: workspace is NULL");
    AGX_REQUIRE(flag_mode == 0 || ((flag_mode == AGX_ATTR_FLAGS_SKIP || flag_mode == AGX_ATTR_FLAGS_ONLY) && dst_flags),
                AGX_ERR_ARG, "agx_edge_attrs_stats_flagged: flag_mode must be 0, 1 (skip) or 2 (only) with dst_flags");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_rec, dst_rec, workspace);
    if (rc) return rc;
    AGX_REQUIRE((!want_len || out_len) && (!want_dir || out_dir), AGX_ERR_ARG,
                "agx_edge_attrs_stats_flagged: give every requested output buffer");
    return attrs_raw(edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len, 0, out_len, want_dir, dir_rotated, out_dir,
                     true, stats, workspace, stream, dst_flags, flag_mode);
}

// Raw values + statistics of the edges of a LIST of targets of a regular-k edge list (a KNN result: the edges of target
// t are [t k, (t + 1) k)); the list length lives in device memory.  One small launch, sized for a few thousand targets.
extern "C" int agx_edge_attrs_stats_list(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges, int regular_k,
                                         const int32_t* list, const int64_t* count, const float* src_rec,
                                         const double* dst_rec, int want_len, int want_dir, int dir_rotated,
                                         float* out_len, float* out_dir, double* stats, double* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(stats && workspace && list && count, AGX_ERR_ARG, "agx_edge_attrs_stats_list: NULL buffer");
    AGX_REQUIRE(regular_k > 0 && n_edges % regular_k == 0, AGX_ERR_ARG, "agx_edge_attrs_stats_list: n_edges is not a multiple of k");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_rec, dst_rec, workspace);
    if (rc) return rc;
    AGX_REQUIRE((!want_len || out_len) && (!want_dir || out_dir), AGX_ERR_ARG,
                "agx_edge_attrs_stats_list: give every requested output buffer");
    int grid = 0;
    if (n_edges > 0 && (want_len || want_dir)) {
        grid = 32;
        attrs_launch<1>(true, true, grid, stream, edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len, 0, out_len,
                        want_dir, dir_rotated, out_dir, workspace, nullptr, 0, list, count, regular_k);
        agx_note_launch(1);
    }
    k_attr_fold<<<1, 256, 0, stream>>>(workspace, grid, stats);
    agx_note_launch(1);
    AGX_LAUNCH_OK();
    return AGX_OK;
}

extern "C" int agx_edge_attrs_apply(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges,
                                    const float* src_rec, const double* dst_rec, int len_norm, int len_invert,
                                    float* out_len, int dir_norm, int dir_rotated, float* out_dir, const double* stats,
                                    int n_stat_sets, int64_t n_edges_global, int raw_present, double* workspace,
                                    void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int want_len = len_norm >= 0, want_dir = dir_norm >= 0;
    AGX_REQUIRE(len_norm <= AGX_NORM_UNIT_STD && dir_norm <= AGX_NORM_UNIT_STD, AGX_ERR_ARG, "agx_edge_attrs: unknown norm code");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_rec, dst_rec, workspace);
    if (rc) return rc;
    if (n_edges == 0 || (!want_len && !want_dir)) return AGX_OK;
    AGX_REQUIRE(!want_len || out_len, AGX_ERR_ARG, "agx_edge_attrs: out_len is NULL");
    AGX_REQUIRE(!want_dir || out_dir, AGX_ERR_ARG, "agx_edge_attrs: out_dir is NULL");
    bool need_stats = (want_len && len_norm > 0) || (want_dir && dir_norm > 0);
    AGX_REQUIRE(!need_stats || (stats && n_stat_sets >= 1), AGX_ERR_ARG,
                "agx_edge_attrs_apply: this normalisation needs the statistics of every shard");
    AGX_REQUIRE(n_edges_global >= n_edges, AGX_ERR_ARG, "agx_edge_attrs_apply: n_edges_global < n_edges");
    if (!raw_present) {
        rc = attrs_raw(edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len, 0, out_len, want_dir, dir_rotated, out_dir,
                       true, nullptr, workspace, stream);
        if (rc) return rc;
    }
    return attrs_scale(n_edges, len_norm, len_invert, out_len, dir_norm, dir_rotated, out_dir, stats, n_stat_sets,
                       n_edges_global, workspace, stream);
}

extern "C" int agx_edge_attrs(const int32_t* edge_src, const int32_t* edge_dst, int64_t n_edges, const float* src_rec,
                              const double* dst_rec, int len_norm, int len_invert, float* out_len, int dir_norm,
                              int dir_rotated, float* out_dir, double* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int want_len = len_norm >= 0, want_dir = dir_norm >= 0;
    AGX_REQUIRE(len_norm <= AGX_NORM_UNIT_STD && dir_norm <= AGX_NORM_UNIT_STD, AGX_ERR_ARG, "agx_edge_attrs: unknown norm code");
    int rc = attrs_check(edge_src, edge_dst, n_edges, src_rec, dst_rec, workspace);
    if (rc) return rc;
    if (n_edges == 0 || (!want_len && !want_dir)) return AGX_OK;
    AGX_REQUIRE(!want_len || out_len, AGX_ERR_ARG, "agx_edge_attrs: out_len is NULL");
    AGX_REQUIRE(!want_dir || out_dir, AGX_ERR_ARG, "agx_edge_attrs: out_dir is NULL");
    bool need_stats = (want_len && len_norm > 0) || (want_dir && dir_norm > 0);
    double* stats = need_stats ? workspace + 6 : nullptr;  // ws[6..13]: between the parameters and the block partials
    // no normalisation: the raw pass writes the final values (an inversion needs no statistics)
    rc = attrs_raw(edge_src, edge_dst, n_edges, src_rec, dst_rec, want_len, need_stats ? 0 : len_invert, out_len, want_dir,
                   dir_rotated, out_dir, true, stats, workspace, stream);
    if (rc || !need_stats) return rc;
    return attrs_scale(n_edges, len_norm, len_invert, out_len, dir_norm, dir_rotated, out_dir, stats, 1, n_edges, workspace,
                       stream);
}
