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
