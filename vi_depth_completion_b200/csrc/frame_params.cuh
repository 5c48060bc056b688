// frame_params.cuh -- per-frame rotation / homography / canvas-scale parameters.
//
// Restates, with the exact fp32 roundings of the reference executed on CPU (see exact_math.cuh and
// DESIGN.md "accumulate schemes"), networks/warping_2dof_alignment.py:
//   :35-58    _build_homography           (q = g x a, q4 = cos(atan2(|q|, a.g)/2), R, H, Hinv)
//   :125-140  corner projection, bbox, 4:3-fit scale  (repeated at :168-194 and :226-240)
// One thread computes one frame; ~150 fp32 ops, one fp64 cosine.  No host synchronisation
// (the reference performs dozens of device->host reads per frame in this block).
#pragma once
#include "exact_math.cuh"
#include "../../include/vidc_b200.h"

namespace vidc {

// torch.max / torch.min over a 1-D tensor propagate NaN
VIDC_HD float tmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
VIDC_HD float tmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// ref :125-140 given H.  corners: (0,0), (W-1,0), (0,H-1), (W-1,H-1), homogeneous (ref :18)
VIDC_HD void frame_scale(const vidc_camera& cam, const float* Hm, vidc_frame_params& p) {
    const float Wm = (float)(cam.W - 1), Hmm = (float)(cam.H - 1);
    const float cxs[4] = {0.0f, Wm, 0.0f, Wm};
    const float cys[4] = {0.0f, 0.0f, Hmm, Hmm};
    float px[4], py[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float c0 = dot3_021(Hm[0], Hm[1], Hm[2], cxs[j], cys[j], 1.0f);
        const float c1 = dot3_021(Hm[3], Hm[4], Hm[5], cxs[j], cys[j], 1.0f);
        const float c2 = dot3_021(Hm[6], Hm[7], Hm[8], cxs[j], cys[j], 1.0f);
        px[j] = c0 / c2;                                                    // :126
        py[j] = c1 / c2;
    }
    const float px_max = tmax(tmax(tmax(px[0], px[1]), px[2]), px[3]);      // :127-130
    const float px_min = tmin(tmin(tmin(px[0], px[1]), px[2]), px[3]);
    const float py_max = tmax(tmax(tmax(py[0], py[1]), py[2]), py[3]);
    const float py_min = tmin(tmin(tmin(py[0], py[1]), py[2]), py[3]);
    const float h_max = py_max - py_min;                                    // :132
    const float w_max = px_max - px_min;                                    // :133
    const float Wf = (float)cam.W, Hf = (float)cam.H;
    float kw, kh;
    // python `scalar / tensor` is tensor.reciprocal() * scalar: two roundings
    if (w_max > (4.0f * h_max) / 3.0f) {                                    // :135
        kw = (1.0f / w_max) * Wf;                                           // :136
        kh = (1.0f / ((3.0f * w_max) / 4.0f)) * Hf;                         // :137
    } else {
        kh = (1.0f / h_max) * Hf;                                           // :139
        kw = (1.0f / ((4.0f * h_max) / 3.0f)) * Wf;                         // :140
    }
    p.px_min = px_min; p.py_min = py_min;
    p.kw = kw; p.kh = kh;
    p.ikw = (1.0f / kw) * 1.0f;                                             // "1./kw" :142
    p.ikh = (1.0f / kh) * 1.0f;                                             // "1./kh" :143
    p.w_max = w_max; p.h_max = h_max;
}

// ref :35-58 followed by :125-140
VIDC_HD void frame_params_from_gravity(const vidc_camera& cam, const float* g, const float* a,
                                       vidc_frame_params& p) {
    // :41-42  q = (-[a]x) g  (bmm, mul+add)
    float q0 = dot3_muladd(-0.0f, a[2], -a[1], g[0], g[1], g[2]);
    float q1 = dot3_muladd(-a[2], -0.0f, a[0], g[0], g[1], g[2]);
    float q2 = dot3_muladd(a[1], -a[0], -0.0f, g[0], g[1], g[2]);
    const float d = dot3_muladd(a[0], a[1], a[2], g[0], g[1], g[2]);        // :43
    float ss = q0 * q0;                                                     // :44 norm(dim=1)
    ss = fmaf(q1, q1, ss);
    ss = fmaf(q2, q2, ss);
    const float n = sqrtf(ss);
    const float q4 = mkl_cosf_ha(0.5f * glibc_atan2f(n, d));                // :48
    const float two_q4 = 2.0f * q4;
    q0 = q0 / two_q4; q1 = q1 / two_q4; q2 = q2 / two_q4;                   // :51
    const float S[9] = {0.0f, -q2, q1, q2, 0.0f, -q0, -q1, q0, 0.0f};       // :52
    float* R = p.R;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float I3 = (r == c) ? 1.0f : 0.0f;
            const float t1 = two_q4 * S[3 * r + c];
            const float t2 = dot3_fma(2.0f * S[3 * r], 2.0f * S[3 * r + 1], 2.0f * S[3 * r + 2],
                                      S[c], S[3 + c], S[6 + c]);            // (2.*S) @ S
            R[3 * r + c] = (I3 + t1) + t2;                                  // :53-54
        }
    float KR[9], KRt[9];
    const float* K = cam.K;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            KR[3 * r + c]  = dot3_muladd(K[3 * r], K[3 * r + 1], K[3 * r + 2], R[c], R[3 + c], R[6 + c]);
            KRt[3 * r + c] = dot3_muladd(K[3 * r], K[3 * r + 1], K[3 * r + 2], R[3 * c], R[3 * c + 1], R[3 * c + 2]);
        }
    const float* Ki = cam.Kinv;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p.H[3 * r + c]    = dot3_fma(KR[3 * r], KR[3 * r + 1], KR[3 * r + 2], Ki[c], Ki[3 + c], Ki[6 + c]);   // :55
            p.Hinv[3 * r + c] = dot3_fma(KRt[3 * r], KRt[3 * r + 1], KRt[3 * r + 2], Ki[c], Ki[3 + c], Ki[6 + c]); // :56-57
        }
    frame_scale(cam, p.H, p);
}

// Conservative 'this 32x32 canvas tile lies entirely outside the source footprint' (kernels_params.cuh explains why it is
// safe; forward warp with input and canvas of the same size).  Not part of any result: it only decides which CTAs may
// write zeros without computing coordinates.
VIDC_HD bool tile_certainly_exterior(const vidc_frame_params& p, const vidc_camera& cam, int tx, int ty) {
    const float Wf = (float)cam.W, Hf = (float)cam.H;
    const float X[2] = {(float)(tx * 32), fminf((float)(tx * 32 + 31), Wf - 1.0f)};
    const float Y[2] = {(float)(ty * 32), fminf((float)(ty * 32 + 31), Hf - 1.0f)};
    float ix_lo = 3.0e38f, ix_hi = -3.0e38f, iy_lo = 3.0e38f, iy_hi = -3.0e38f, s_lo = 3.0e38f, s_hi = -3.0e38f;
    bool ok = true;
    for (int c = 0; c < 4; ++c) {
        const float px = p.ikw * X[c & 1] + p.px_min, py = p.ikh * Y[c >> 1] + p.py_min;
        const float t0 = p.Hinv[6] * px, t1 = p.Hinv[7] * py;
        const float s = t0 + t1 + p.Hinv[8];
        const float u = p.Hinv[0] * px + p.Hinv[1] * py + p.Hinv[2];
        const float v = p.Hinv[3] * px + p.Hinv[4] * py + p.Hinv[5];
        ok = ok && fabsf(s) > 1e-3f * (fabsf(t0) + fabsf(t1) + fabsf(p.Hinv[8]));
        const float ix = ((u / s - cam.cx) * cam.inv_half_w + 1.0f) * Wf * 0.5f - 0.5f;
        const float iy = ((v / s - cam.cy) * cam.inv_half_h + 1.0f) * Hf * 0.5f - 0.5f;
        ok = ok && fabsf(ix) < 1e30f && fabsf(iy) < 1e30f;          // also rejects NaN
        ix_lo = fminf(ix_lo, ix); ix_hi = fmaxf(ix_hi, ix); iy_lo = fminf(iy_lo, iy); iy_hi = fmaxf(iy_hi, iy);
        s_lo = fminf(s_lo, s); s_hi = fmaxf(s_hi, s);
    }
    ok = ok && (s_lo > 0.0f || s_hi < 0.0f);
    const float mx = 4.0f + 2e-3f * fmaxf(fabsf(ix_lo), fabsf(ix_hi)), my = 4.0f + 2e-3f * fmaxf(fabsf(iy_lo), fabsf(iy_hi));
    return ok && (ix_hi < -1.0f - mx || ix_lo > Wf + mx || iy_hi < -1.0f - my || iy_lo > Hf + my);
}


}  // namespace vidc
