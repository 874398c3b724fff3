// Cap -> cell windows: which cells of each cube face can hold a point within angular radius rho of q.
//
// For a face with in-face axes (a, b) and normal component c, cell columns are bands of
// alpha = atan2(a, c) ("longitude" about the b axis) and cell rows of the other face angle are bands of
// beta = atan2(b, c).  All points within angle rho of q have alpha in [alpha_q - D, alpha_q + D] with
// sin D = sin(rho) / sqrt(a^2 + c^2) (longitude extent of a small circle), likewise for beta.  The
// window is that box clipped to the face; every bound is inflated so float32 rounding can only add cells.
#pragma once
#include "agx_common.cuh"

struct AgxCap {
    float sin_rho;   // inflated
    float cos_rho;
    bool everything; // rho >= pi/2: scan all cells
};

__device__ __forceinline__ AgxCap agx_make_cap(float chord2) {
    AgxCap cap;
    cap.everything = chord2 >= 1.9f;
    float t = fminf(chord2, 2.0f);
    cap.sin_rho = sqrtf(t * (1.0f - 0.25f * t)) * 1.00002f + 3e-7f;  // sin(rho), chord = 2 sin(rho/2)
    cap.cos_rho = 1.0f - 0.5f * t;
    return cap;
}

// One axis of the window: returns false if the band misses the face.
__device__ __forceinline__ bool agx_axis_window(float a, float c, const AgxCap& cap, int cells, int& lo, int& hi) {
    float len = sqrtf(a * a + c * c);
    if (cap.sin_rho >= len * 0.99999f) {
        lo = 0;
        hi = cells - 1;
        return true;
    }
    float x = cap.sin_rho / len;
    // asin(x) <= x (1 + 0.18 x^2) for x <= 1/4 (series: 1/6 + 3 x^2 / 40 + ... < 0.172)
    float d = (x <= 0.25f ? x * fmaf(0.18f * x, x, 1.0f) : asinf(x)) * 1.00002f + 2e-6f;
    float ang = atan2f(a, c);
    float l = ang - d, h = ang + d;
    if (h < -AGX_QUARTER_PI_F || l > AGX_QUARTER_PI_F) return false;
    lo = agx_angle_to_cell(l, cells);
    hi = agx_angle_to_cell(h, cells);
    return true;
}

// atan(t) for |t| <= 1 with absolute error < 4e-6: least-squares fit of atan(t)/t as a degree-5 polynomial in t^2,
// float32 Horner (3.4e-6 measured on 2 M points).  Cells are >= 7.7e-4 rad wide, the window bounds are inflated by 1.2e-5.
__device__ __forceinline__ float agx_atan_unit(float t) {
    float u = t * t;
    float p = fmaf(-0.013422180f, u, 0.057330552f);
    p = fmaf(p, u, -0.12110946f);
    p = fmaf(p, u, 0.19558905f);
    p = fmaf(p, u, -0.33298862f);
    p = fmaf(p, u, 0.99999553f);
    return p * t;
}

// One axis of the window on the query's MAJOR face (c >= |a|, c >= 0.577).  true only if the band is narrow and
// ends at least one cell away from the face border - then no point of the cap can lie on another face (it would
// need |alpha| > pi/4), and the caller can skip the other five faces.  No inverse trigonometry: asin by its
// series bound, atan by a polynomial on [-1, 1].
__device__ __forceinline__ bool agx_axis_window_major(float a, float c, const AgxCap& cap, int cells, int& lo, int& hi) {
    float x = cap.sin_rho * rsqrtf(a * a + c * c) * 1.000002f;
    if (x > 0.25f) return false;
    float d = x * fmaf(0.18f * x, x, 1.0f) * 1.00002f + 1.2e-5f;
    float ang = agx_atan_unit(__fdividef(a, c));
    lo = agx_angle_to_cell(ang - d, cells);
    hi = agx_angle_to_cell(ang + d, cells);
    return lo > 0 && hi < cells - 1;
}

// Window of `face` for the cap around q.  false = face not touched.
__device__ __forceinline__ bool agx_face_window(int face, float3 q, const AgxCap& cap, int cells, int& i0, int& i1,
                                                int& j0, int& j1) {
    float a, b, c;
    agx_face_frame(face, q.x, q.y, q.z, a, b, c);
    if (cap.everything) {
        i0 = j0 = 0;
        i1 = j1 = cells - 1;
        return true;
    }
    // every point of the face lies within 54.7357 deg of its normal: reject if q is further than that + rho
    if (c < 0.57735026f * cap.cos_rho - 0.81649658f * cap.sin_rho - 1e-5f) return false;
    if (!agx_axis_window(a, c, cap, cells, i0, i1)) return false;
    if (!agx_axis_window(b, c, cap, cells, j0, j1)) return false;
    return true;
}
