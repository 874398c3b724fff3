// K6/K7: icosphere subdivision and multi-scale (x_hops) edges for global TriNodes.
//
// K6 replaces trimesh.creation.icosphere as used by the reference
// (/root/reference/src/anemoi/graphs/generate/tri_icosahedron.py:121,173) - trimesh is a third-party
// dependency; its published algorithm (icosahedron table, Trimesh.subdivide with midpoints numbered
// by np.unique over (min | max << 32), refine_spherical after every level) is restated in
// oracle/trimesh_icosphere.py and re-designed here without any sort: a midpoint's number is
//     V + (number of edges whose larger endpoint is < u) + (rank of v among u's smaller neighbours)
// for the edge (v < u), i.e. a per-vertex count, one exclusive scan and a <= 6-entry local sort.
// All vertex arithmetic uses explicit round-to-nearest float64 intrinsics in numpy's operation order,
// so vertices are bit-identical to the CPU restatement.
//
// K7 replaces tri_icosahedron.add_edges_to_nx_graph + nx.to_scipy_sparse_array
// (generate/tri_icosahedron.py:138-224, edges/builder.py:412-455): per level a fixed-width adjacency
// table (degree <= 6), then one thread per graph node runs the <= x_hops frontier expansion on every
// requested level the vertex exists in, unions the results, relabels through rank_of_vertex and
// sorts - the output is already in canonical (dst, src) order.  count -> scan -> fill.
#include <math.h>

#include "agx_common.cuh"

static inline int64_t ico_nv(int level) { return 10 * ((int64_t)1 << (2 * level)) + 2; }
static inline int64_t ico_nf(int level) { return 20 * ((int64_t)1 << (2 * level)); }
static inline int64_t ico_face_offset(int level) { return 20 * ((((int64_t)1 << (2 * level)) - 1) / 3); }

// ---------------------------------------------------------------------------------------------
// subdivision
// ---------------------------------------------------------------------------------------------
// Every phase is a __device__ body over a (first, stride) index range so that it serves both the grid-wide
// kernels of the large levels and the single-CTA kernel that runs all small levels back to back.
__device__ __forceinline__ void ico_lower_neighbours(const int32_t* __restrict__ faces, int64_t nf, int* __restrict__ cnt,
                                                     int32_t* __restrict__ lower, int64_t first, int64_t stride) {
    for (int64_t f = first; f < nf; f += stride) {
        int v[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            int a = v[e], b = v[(e + 1) % 3];
            if (a > b) {  // each undirected edge has exactly one half-edge with a > b
                int slot = atomicAdd(&cnt[a], 1);
                lower[6 * (int64_t)a + slot] = b;
            }
        }
    }
}

__device__ __forceinline__ void ico_sort_lower(const int* __restrict__ cnt, int32_t* __restrict__ lower, int64_t nv,
                                               int64_t first, int64_t stride) {
    for (int64_t u = first; u < nv; u += stride) {
        int m = cnt[u];
        int32_t* l = lower + 6 * u;
        for (int i = 1; i < m; ++i) {
            int x = l[i], j = i - 1;
            while (j >= 0 && l[j] > x) {
                l[j + 1] = l[j];
                --j;
            }
            l[j + 1] = x;
        }
    }
}

__device__ __forceinline__ void ico_midpoints(const int* __restrict__ cnt, const int32_t* __restrict__ lower,
                                              const int64_t* __restrict__ base, int64_t nv, double* __restrict__ vert,
                                              int64_t first, int64_t stride) {
    for (int64_t u = first; u < nv; u += stride) {
        int m = cnt[u];
        for (int s = 0; s < m; ++s) {
            int64_t v = lower[6 * u + s];
            int64_t id = nv + base[u] + s;
#pragma unroll
            for (int c = 0; c < 3; ++c)  // vertices[edges[unique]].mean(axis=1): (p_min + p_max) / 2
                vert[3 * id + c] = __ddiv_rn(__dadd_rn(vert[3 * v + c], vert[3 * u + c]), 2.0);
        }
    }
}

__device__ __forceinline__ int ico_mid(int a, int b, const int* cnt, const int32_t* lower, const int64_t* base, int64_t nv) {
    int u = max(a, b), v = min(a, b);
    int m = cnt[u];
    int pos = 0;
    for (int s = 0; s < m; ++s)
        if (lower[6 * (int64_t)u + s] == v) pos = s;
    return (int)(nv + base[u] + pos);
}

__device__ __forceinline__ void ico_faces(const int32_t* __restrict__ faces, int64_t nf, const int* __restrict__ cnt,
                                          const int32_t* __restrict__ lower, const int64_t* __restrict__ base, int64_t nv,
                                          int32_t* __restrict__ out, int64_t first, int64_t stride) {
    for (int64_t f = first; f < nf; f += stride) {
        int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
        int m0 = ico_mid(a, b, cnt, lower, base, nv), m1 = ico_mid(b, c, cnt, lower, base, nv),
            m2 = ico_mid(c, a, cnt, lower, base, nv);
        int32_t* o = out + 12 * f;
        o[0] = a;  o[1] = m0; o[2] = m2;
        o[3] = m0; o[4] = b;  o[5] = m1;
        o[6] = m2; o[7] = m1; o[8] = c;
        o[9] = m0; o[10] = m1; o[11] = m2;
    }
}

// icosphere's refine_spherical: scalar = sqrt(dot(v**2, [1,1,1])); v += (v / scalar) * (radius - scalar)
__device__ __forceinline__ void ico_refine(double* __restrict__ vert, int64_t nv, int64_t first, int64_t stride) {
    for (int64_t i = first; i < nv; i += stride) {
        double x = vert[3 * i], y = vert[3 * i + 1], z = vert[3 * i + 2];
        double scalar = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
        double off = __dsub_rn(1.0, scalar);
        vert[3 * i] = __dadd_rn(x, __dmul_rn(__ddiv_rn(x, scalar), off));
        vert[3 * i + 1] = __dadd_rn(y, __dmul_rn(__ddiv_rn(y, scalar), off));
        vert[3 * i + 2] = __dadd_rn(z, __dmul_rn(__ddiv_rn(z, scalar), off));
    }
}

#define GRID_RANGE (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x

__global__ void k_ico_lower_neighbours(const int32_t* __restrict__ faces, int64_t nf, int* __restrict__ cnt,
                                       int32_t* __restrict__ lower) {
    ico_lower_neighbours(faces, nf, cnt, lower, GRID_RANGE);
}
__global__ void k_ico_sort_lower(const int* __restrict__ cnt, int32_t* __restrict__ lower, int64_t nv) {
    ico_sort_lower(cnt, lower, nv, GRID_RANGE);
}
__global__ void k_ico_midpoints(const int* __restrict__ cnt, const int32_t* __restrict__ lower,
                                const int64_t* __restrict__ base, int64_t nv, double* __restrict__ vert) {
    ico_midpoints(cnt, lower, base, nv, vert, GRID_RANGE);
}
__global__ void k_ico_faces(const int32_t* __restrict__ faces, int64_t nf, const int* __restrict__ cnt,
                            const int32_t* __restrict__ lower, const int64_t* __restrict__ base, int64_t nv,
                            int32_t* __restrict__ out) {
    ico_faces(faces, nf, cnt, lower, base, nv, out, GRID_RANGE);
}
__global__ void k_ico_refine(double* __restrict__ vert, int64_t nv) { ico_refine(vert, nv, GRID_RANGE); }

// Levels 0 .. n_steps of the subdivision in ONE CTA (the small levels are pure launch latency otherwise: nine
// launches of a few microseconds per level).  Phases are separated by __syncthreads(), which also orders the
// CTA's global-memory writes.  Step l reads level l and writes level l + 1; n_steps <= ICO_FUSED_STEPS.
#define ICO_FUSED_THREADS 1024
#define ICO_FUSED_STEPS 5  // up to level 5: 2 562 -> 10 242 vertices in the last fused step
#define ICO_FUSED_ITEMS 3  // ceil(2562 / 1024) items of the per-vertex scan per thread

struct IcoBase {
    double v[36];
    int32_t f[60];
};

__global__ void __launch_bounds__(ICO_FUSED_THREADS) k_ico_small_levels(IcoBase b, int n_steps, double* __restrict__ vert,
                                                                        int32_t* __restrict__ faces_all, int* __restrict__ cnt,
                                                                        int32_t* __restrict__ lower, int64_t* __restrict__ base) {
    __shared__ int warp_sums[ICO_FUSED_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t < 36) vert[t] = b.v[t];
    if (t < 60) faces_all[t] = b.f[t];
    __syncthreads();
    int64_t nv = 12, nf = 20, face_off = 0;
    for (int level = 0; level < n_steps; ++level) {
        const int32_t* faces = faces_all + 3 * face_off;
        int32_t* faces_next = faces_all + 3 * (face_off + nf);
        for (int64_t i = t; i < nv; i += ICO_FUSED_THREADS) cnt[i] = 0;
        __syncthreads();
        ico_lower_neighbours(faces, nf, cnt, lower, t, ICO_FUSED_THREADS);
        __syncthreads();
        ico_sort_lower(cnt, lower, nv, t, ICO_FUSED_THREADS);
        // exclusive scan of cnt[0 .. nv) -> base[0 .. nv]: thread t owns items [3t, 3t + 3)
        int c[ICO_FUSED_ITEMS], mine = 0;
#pragma unroll
        for (int k = 0; k < ICO_FUSED_ITEMS; ++k) {
            int64_t i = (int64_t)t * ICO_FUSED_ITEMS + k;
            c[k] = i < nv ? cnt[i] : 0;
            mine += c[k];
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += v;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        int64_t ex = warp_sums[warp] + incl - mine;
#pragma unroll
        for (int k = 0; k < ICO_FUSED_ITEMS; ++k) {
            int64_t i = (int64_t)t * ICO_FUSED_ITEMS + k;
            if (i <= nv) base[i] = ex;  // base[nv] = total (c[] is 0 beyond nv)
            ex += c[k];
        }
        __syncthreads();
        ico_midpoints(cnt, lower, base, nv, vert, t, ICO_FUSED_THREADS);
        ico_faces(faces, nf, cnt, lower, base, nv, faces_next, t, ICO_FUSED_THREADS);
        __syncthreads();
        face_off += nf;
        nv = 4 * nv - 6;  // 10 * 4^(l+1) + 2
        nf *= 4;
        ico_refine(vert, nv, t, ICO_FUSED_THREADS);
        __syncthreads();
    }
}

// cartesian_to_latlon_rad (generate/transforms.py:50-52): lat = arcsin(z / sum(xyz^2)), lon = arctan2(y, x)
__global__ void k_ico_latlon(const double* __restrict__ vert, int64_t nv, float2* __restrict__ latlon) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        double x = vert[3 * i], y = vert[3 * i + 1], z = vert[3 * i + 2];
        double n2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
        latlon[i] = make_float2((float)asin(__ddiv_rn(z, n2)), (float)atan2(y, x));
    }
}

extern "C" int agx_icosphere(int max_level, double* vertices, int32_t* faces_all, float* latlon, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(max_level >= 0 && max_level <= 12, AGX_ERR_ARG, "agx_icosphere: level %d out of [0, 12]", max_level);
    AGX_REQUIRE(vertices && faces_all, AGX_ERR_ARG, "agx_icosphere: NULL buffer");
    // trimesh.creation.icosahedron: the tables travel as a kernel argument (no host staging, no sync)
    IcoBase base0;
    {
        const double t = (1.0 + sqrt(5.0)) / 2.0;
        const double s = sqrt(2.0 + t);
        const double raw[36] = {-1, t, 0, 1, t, 0, -1, -t, 0, 1, -t, 0, 0, -1, t, 0, 1, t,
                                0, -1, -t, 0, 1, -t, t, 0, -1, t, 0, 1, -t, 0, -1, -t, 0, 1};
        static const int32_t f0[60] = {0, 11, 5, 0, 5, 1, 0, 1, 7, 0, 7, 10, 0, 10, 11, 1, 5, 9, 5, 11, 4, 11, 10, 2, 10, 7, 6, 7, 1, 8,
                                       3, 9, 4, 3, 4, 2, 3, 2, 6, 3, 6, 8, 3, 8, 9, 4, 9, 5, 2, 4, 11, 6, 2, 10, 8, 6, 7, 9, 8, 1};
        for (int i = 0; i < 36; ++i) base0.v[i] = raw[i] / s;
        for (int i = 0; i < 60; ++i) base0.f[i] = f0[i];
    }
    int64_t nv_max = ico_nv(max_level);
    agx_pool_keep_warm();
    int* cnt = nullptr;
    int32_t* lower = nullptr;
    int64_t* base = nullptr;
    {   // scratch sized for the coarser of every pair of levels (and at least level 0 for the fused kernel's init)
        int64_t nv_prev = ico_nv(max_level > 0 ? max_level - 1 : 0);
        AGX_CUDA_OK(cudaMallocAsync(&cnt, nv_prev * sizeof(int), stream));
        AGX_CUDA_OK(cudaMallocAsync(&lower, 6 * nv_prev * sizeof(int32_t), stream));
        AGX_CUDA_OK(cudaMallocAsync(&base, (nv_prev + 1) * sizeof(int64_t), stream));
    }
    const int fused_steps = max_level < ICO_FUSED_STEPS ? max_level : ICO_FUSED_STEPS;
    k_ico_small_levels<<<1, ICO_FUSED_THREADS, 0, stream>>>(base0, fused_steps, vertices, faces_all, cnt, lower, base);
    agx_note_launch(1);
    for (int level = fused_steps; level < max_level; ++level) {
        int64_t nv = ico_nv(level), nf = ico_nf(level);
        const int32_t* faces = faces_all + 3 * ico_face_offset(level);
        int32_t* faces_next = faces_all + 3 * ico_face_offset(level + 1);
        AGX_CUDA_OK(cudaMemsetAsync(cnt, 0, nv * sizeof(int), stream));
        int gf = agx_grid(nf, 256, 8), gv = agx_grid(nv, 256, 8);
        k_ico_lower_neighbours<<<gf, 256, 0, stream>>>(faces, nf, cnt, lower);
        k_ico_sort_lower<<<gv, 256, 0, stream>>>(cnt, lower, nv);
        int rc = agx_exclusive_scan(cnt, nv, base, nullptr, stream);
        if (rc) return rc;
        k_ico_midpoints<<<gv, 256, 0, stream>>>(cnt, lower, base, nv, vertices);
        k_ico_faces<<<gf, 256, 0, stream>>>(faces, nf, cnt, lower, base, nv, faces_next);
        k_ico_refine<<<agx_grid(ico_nv(level + 1), 256, 8), 256, 0, stream>>>(vertices, ico_nv(level + 1));
        agx_note_launch(5);
    }
    if (latlon) {
        k_ico_latlon<<<agx_grid(nv_max, 256, 8), 256, 0, stream>>>(vertices, nv_max, (float2*)latlon);
        agx_note_launch(1);
    }
    AGX_LAUNCH_OK();
    AGX_CUDA_OK(cudaFreeAsync(cnt, stream));
    AGX_CUDA_OK(cudaFreeAsync(lower, stream));
    AGX_CUDA_OK(cudaFreeAsync(base, stream));
    return AGX_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-scale edges
// ---------------------------------------------------------------------------------------------
#define MS_MAX_LEVELS 16

struct MsLevels {
    const int32_t* nb[MS_MAX_LEVELS];    // adjacency table of each requested level: nb[6*v + s]
    const int* deg[MS_MAX_LEVELS];
    const int32_t* vmap[MS_MAX_LEVELS];  // level vertex -> graph node position, or -1 (vertex not valid)
    const int32_t* inv[MS_MAX_LEVELS];   // graph node position -> level vertex, or -1
    int n;
};

__global__ void k_ms_adjacency(const int32_t* __restrict__ faces, int64_t nf, int* __restrict__ deg,
                               int32_t* __restrict__ nb) {
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        int v[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
#pragma unroll
        for (int e = 0; e < 3; ++e) {  // the half-edges of all faces are every directed edge exactly once
            int a = v[e], b = v[(e + 1) % 3];
            int slot = atomicAdd(&deg[a], 1);
            nb[6 * (int64_t)a + slot] = b;
        }
    }
}

// scratch element i of node t lives at scratch[i * n_nodes + t] (coalesced across threads)
// walk_all = 0: the frontier only crosses valid vertices (tri meshes, tri_icosahedron.py:214-215);
// walk_all = 1: the whole disk is walked and invalid cells are dropped afterwards (h3.k_ring(idx, k) & nodes,
// hex_icosahedron.py:147).
__global__ void __launch_bounds__(128) k_ms_expand(MsLevels lv, int x_hops, int walk_all, int64_t n_nodes, int hop_cap,
                                                    int32_t* __restrict__ counts, int32_t* __restrict__ scratch) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_nodes; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t* uni = scratch + t;                                      // union list (graph positions), stride n_nodes
        int32_t* que = scratch + (int64_t)lv.n * hop_cap * n_nodes + t;  // BFS queue (vertex ids), stride n_nodes
        int n_uni = 0;
        for (int l = 0; l < lv.n; ++l) {
            const int u = lv.inv[l][t];
            if (u < 0) continue;  // this graph node is not a (valid) vertex of the level
            const int32_t* nb = lv.nb[l];
            const int* deg = lv.deg[l];
            const int32_t* vmap = lv.vmap[l];
            int n_que = 1, frontier_begin = 0;
            que[0] = u;
            for (int hop = 0; hop < x_hops; ++hop) {
                int frontier_end = n_que;
                for (int i = frontier_begin; i < frontier_end; ++i) {
                    int w = que[(int64_t)i * n_nodes];
                    int d = deg[w];
                    for (int s = 0; s < d; ++s) {
                        int x = nb[6 * (int64_t)w + s];
                        if (!walk_all && vmap[x] < 0) continue;  // mesh edges need both endpoints valid
                        bool seen = false;
                        for (int j = 0; j < n_que; ++j) seen |= (que[(int64_t)j * n_nodes] == x);
                        if (!seen) que[(int64_t)(n_que++) * n_nodes] = x;
                    }
                }
                frontier_begin = frontier_end;
            }
            for (int i = 1; i < n_que; ++i) {  // i = 0 is the centre (center=False)
                int r = vmap[que[(int64_t)i * n_nodes]];
                if (r < 0) continue;
                bool seen = (r == (int)t);  // two vertices of one level never share a node; guard self loops anyway
                for (int j = 0; j < n_uni; ++j) seen |= (uni[(int64_t)j * n_nodes] == r);
                if (!seen) uni[(int64_t)(n_uni++) * n_nodes] = r;
            }
        }
        for (int i = 1; i < n_uni; ++i) {  // ascending source position
            int x = uni[(int64_t)i * n_nodes], j = i - 1;
            while (j >= 0 && uni[(int64_t)j * n_nodes] > x) {
                uni[(int64_t)(j + 1) * n_nodes] = uni[(int64_t)j * n_nodes];
                --j;
            }
            uni[(int64_t)(j + 1) * n_nodes] = x;
        }
        counts[t] = n_uni;
    }
}

__global__ void k_ms_fill(int64_t n_nodes, const int32_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                          const int32_t* __restrict__ scratch, int32_t* __restrict__ out_src,
                          int32_t* __restrict__ out_dst) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_nodes; t += (int64_t)gridDim.x * blockDim.x) {
        int m = counts[t];
        int64_t o = offsets[t];
        for (int i = 0; i < m; ++i) {
            out_src[o + i] = scratch[(int64_t)i * n_nodes + t];
            out_dst[o + i] = (int32_t)t;
        }
    }
}

// vmap / inv of one level for GLOBAL TriNodes: vmap[v] = rank_of_vertex[v], inv[t] = node_ordering[t] if it is a
// vertex of the level (lower levels are prefixes of the finest one).
__global__ void k_ms_global_maps(const int32_t* __restrict__ node_ordering, int64_t n_nodes, int64_t nv_level,
                                 int32_t* __restrict__ inv) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_nodes; t += (int64_t)gridDim.x * blockDim.x) {
        int u = node_ordering[t];
        inv[t] = u < nv_level ? u : -1;
    }
}

static inline int ms_hop_cap(int x_hops) { return 3 * x_hops * (x_hops + 1) + 1; }

extern "C" int64_t agx_multiscale_scratch_per_node(int n_levels, int x_hops) {
    return (int64_t)(n_levels + 1) * ms_hop_cap(x_hops);
}

static int ms_run(int max_level, const int32_t* faces_all, const int32_t* levels, int n_levels, int x_hops,
                  int64_t n_nodes, const int32_t* const* vmap, const int32_t* const* inv, int32_t* counts,
                  int32_t* scratch, cudaStream_t stream) {
    MsLevels lv;
    lv.n = n_levels;
    agx_pool_keep_warm();
    int* deg_all = nullptr;
    int32_t* nb_all = nullptr;
    int64_t tot_v = 0;
    for (int l = 0; l < n_levels; ++l) tot_v += ico_nv(levels[l]);
    AGX_CUDA_OK(cudaMallocAsync(&deg_all, tot_v * sizeof(int), stream));
    AGX_CUDA_OK(cudaMallocAsync(&nb_all, 6 * tot_v * sizeof(int32_t), stream));
    AGX_CUDA_OK(cudaMemsetAsync(deg_all, 0, tot_v * sizeof(int), stream));
    int64_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        int64_t nv = ico_nv(levels[l]), nf = ico_nf(levels[l]);
        lv.nb[l] = nb_all + 6 * off;
        lv.deg[l] = deg_all + off;
        lv.vmap[l] = vmap[l];
        lv.inv[l] = inv[l];
        k_ms_adjacency<<<agx_grid(nf, 256, 8), 256, 0, stream>>>(faces_all + 3 * ico_face_offset(levels[l]), nf,
                                                                 deg_all + off, nb_all + 6 * off);
        off += nv;
    }
    (void)max_level;
    k_ms_expand<<<agx_grid(n_nodes, 128, 16), 128, 0, stream>>>(lv, x_hops, 0, n_nodes, ms_hop_cap(x_hops), counts, scratch);
    AGX_LAUNCH_OK();
    agx_note_launch(n_levels + 1);
    AGX_CUDA_OK(cudaFreeAsync(deg_all, stream));
    AGX_CUDA_OK(cudaFreeAsync(nb_all, stream));
    return AGX_OK;
}

static int ms_check(const char* who, int max_level, const int32_t* levels, int n_levels, int x_hops) {
    AGX_REQUIRE(max_level >= 0 && max_level <= 12, AGX_ERR_ARG, "%s: bad max_level %d", who, max_level);
    AGX_REQUIRE(n_levels > 0 && n_levels <= MS_MAX_LEVELS, AGX_ERR_ARG, "%s: 1..%d levels supported", who, MS_MAX_LEVELS);
    AGX_REQUIRE(x_hops > 0, AGX_ERR_ARG, "x_hops == 0, graph would have no edges ...");
    AGX_REQUIRE(x_hops <= 8, AGX_ERR_UNSUPPORTED, "%s: x_hops = %d > 8 is not built yet", who, x_hops);
    AGX_REQUIRE(levels != nullptr, AGX_ERR_ARG, "%s: levels is NULL", who);
    for (int l = 0; l < n_levels; ++l)
        AGX_REQUIRE(levels[l] >= 0 && levels[l] <= max_level, AGX_ERR_ARG, "%s: level %d out of [0, %d]", who, levels[l], max_level);
    return AGX_OK;
}

extern "C" int agx_multiscale_tri_count(int max_level, const int32_t* faces_all, const int32_t* levels, int n_levels,
                                        int x_hops, const int32_t* node_ordering, const int32_t* rank_of_vertex,
                                        int32_t* counts, int32_t* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = ms_check("agx_multiscale_tri_count", max_level, levels, n_levels, x_hops);
    if (rc) return rc;
    AGX_REQUIRE(faces_all && node_ordering && rank_of_vertex && counts && scratch, AGX_ERR_ARG,
                "agx_multiscale_tri_count: NULL buffer");
    int64_t n_nodes = ico_nv(max_level);
    int32_t* inv_all = nullptr;
    AGX_CUDA_OK(cudaMallocAsync(&inv_all, (int64_t)n_levels * n_nodes * sizeof(int32_t), stream));
    const int32_t *vmap[MS_MAX_LEVELS], *inv[MS_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        vmap[l] = rank_of_vertex;  // every vertex of a global mesh is a graph node
        inv[l] = inv_all + (int64_t)l * n_nodes;
        k_ms_global_maps<<<agx_grid(n_nodes, 256, 8), 256, 0, stream>>>(node_ordering, n_nodes, ico_nv(levels[l]),
                                                                        inv_all + (int64_t)l * n_nodes);
    }
    agx_note_launch(n_levels);
    rc = ms_run(max_level, faces_all, levels, n_levels, x_hops, n_nodes, vmap, inv, counts, scratch, stream);
    cudaFreeAsync(inv_all, stream);
    return rc;
}

extern "C" int agx_multiscale_tri_count_mapped(int max_level, const int32_t* faces_all, const int32_t* levels,
                                               int n_levels, int x_hops, int64_t n_nodes, const int32_t* vertex_map,
                                               const int32_t* node_vertex, int32_t* counts, int32_t* scratch,
                                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = ms_check("agx_multiscale_tri_count_mapped", max_level, levels, n_levels, x_hops);
    if (rc) return rc;
    AGX_REQUIRE(n_nodes >= 0, AGX_ERR_ARG, "agx_multiscale_tri_count_mapped: n_nodes < 0");
    if (n_nodes == 0) return AGX_OK;
    AGX_REQUIRE(faces_all && vertex_map && node_vertex && counts && scratch, AGX_ERR_ARG,
                "agx_multiscale_tri_count_mapped: NULL buffer");
    const int32_t *vmap[MS_MAX_LEVELS], *inv[MS_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
        vmap[l] = vertex_map;  // lower levels are prefixes of the finest one: one map serves every level
        inv[l] = node_vertex + (int64_t)l * n_nodes;
    }
    return ms_run(max_level, faces_all, levels, n_levels, x_hops, n_nodes, vmap, inv, counts, scratch, stream);
}

extern "C" int agx_multiscale_adj_count(int n_levels, const int32_t* const* nb, const int32_t* const* deg,
                                        const int32_t* const* cell_node, const int32_t* const* node_cell, int x_hops,
                                        int walk_all, int64_t n_nodes, int32_t* counts, int32_t* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AGX_REQUIRE(n_levels > 0 && n_levels <= MS_MAX_LEVELS, AGX_ERR_ARG, "agx_multiscale_adj_count: 1..%d levels supported",
                MS_MAX_LEVELS);
    AGX_REQUIRE(x_hops > 0, AGX_ERR_ARG, "x_hops == 0, graph would have no edges ...");
    AGX_REQUIRE(x_hops <= 8, AGX_ERR_UNSUPPORTED, "agx_multiscale_adj_count: x_hops = %d > 8 is not built yet", x_hops);
    AGX_REQUIRE(n_nodes >= 0, AGX_ERR_ARG, "agx_multiscale_adj_count: n_nodes < 0");
    if (n_nodes == 0) return AGX_OK;
    AGX_REQUIRE(nb && deg && cell_node && node_cell && counts && scratch, AGX_ERR_ARG, "agx_multiscale_adj_count: NULL buffer");
    MsLevels lv;
    lv.n = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        AGX_REQUIRE(nb[l] && deg[l] && cell_node[l] && node_cell[l], AGX_ERR_ARG, "agx_multiscale_adj_count: NULL table at level %d", l);
        lv.nb[l] = nb[l];
        lv.deg[l] = deg[l];
        lv.vmap[l] = cell_node[l];
        lv.inv[l] = node_cell[l];
    }
    k_ms_expand<<<agx_grid(n_nodes, 128, 16), 128, 0, stream>>>(lv, x_hops, walk_all ? 1 : 0, n_nodes, ms_hop_cap(x_hops),
                                                                counts, scratch);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}

extern "C" int agx_multiscale_tri_fill(int64_t n_nodes, const int32_t* counts, const int64_t* offsets,
                                       const int32_t* scratch, int64_t scratch_per_node, int32_t* out_src,
                                       int32_t* out_dst, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)scratch_per_node;
    AGX_REQUIRE(n_nodes >= 0, AGX_ERR_ARG, "agx_multiscale_tri_fill: n_nodes < 0");
    if (n_nodes == 0) return AGX_OK;
    AGX_REQUIRE(counts && offsets && scratch && out_src && out_dst, AGX_ERR_ARG, "agx_multiscale_tri_fill: NULL buffer");
    k_ms_fill<<<agx_grid(n_nodes, 256, 8), 256, 0, stream>>>(n_nodes, counts, offsets, scratch, out_src, out_dst);
    AGX_LAUNCH_OK();
    agx_note_launch(1);
    return AGX_OK;
}
