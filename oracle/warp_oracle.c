/*
 * warp_oracle.c -- CPU restatement of the reference's gravity warp / unwarp path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (vi_depth_completion_b200/)
 * links, imports or calls this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it, and only as the
 * checker.
 *
 * What it restates (reference = MARSLab-UMN/vi_depth_completion):
 *   networks/warping_2dof_alignment.py:6-24    camera constants
 *   networks/warping_2dof_alignment.py:35-58   _build_homography
 *   networks/warping_2dof_alignment.py:108-156 warp_with_gravity_center_aligned
 *   networks/warping_2dof_alignment.py:158-214 image_sampler_forward_inverse
 *   networks/warping_2dof_alignment.py:216-255 inverse_warp_normal_image_with_gravity_center_aligned
 *   networks/surface_normal.py:150-156,170     validity mask, pyramid masks, renormalise
 *   normal_utils.py:20-34                      masked angular statistics
 * plus the third-party pieces the reference calls on this path and which are not in
 * its tree: ATen grid_sampler_2d (torch 2.11, align_corners=False, zeros padding),
 * MKL sgemm accumulate orders for the 3x3 products, glibc atan2f and MKL VML vmsCos(HA)
 * (the functions torch's CPU backend dispatches to for the 0-dim tensors at ref :48).
 *
 * Parity pin: the reference ships NO tests or golden vectors for this path
 * (SURVEY.md section 4).  This restatement is pinned instead against the reference
 * source itself, executed on CPU in the build container (oracle/ref_loader.py):
 * tests/test_oracle_vs_reference.py (container only) and the committed fixtures in
 * tests/golden/ made by oracle/make_golden.py.  Every float operation below is a
 * separately rounded IEEE binary32 op unless written as fmaf(); compile with
 * -ffp-contract=off (see oracle/Makefile).
 *
 * One documented divergence from the CPU-executed reference: a NON-FINITE sampling
 * coordinate (non-finite gravity input only) reads as out of bounds here, as in the CUDA
 * build of ATen (GridSampler.cuh:140-147) -- the CPU build returns NaN in bilinear mode.
 * tests/test_oracle_vs_reference.py::test_degenerate_gravity_against_live_reference.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

typedef struct {
    int32_t W, H;          /* ref :13-14  ceil(2cx), ceil(2cy)                   */
    float K[9], Kinv[9];   /* ref :15-20  fp64 K and inv(K) rounded to fp32      */
    float cx, cy;          /* python doubles rounded once to fp32 (ref :149-150) */
    float inv_half_w;      /* float(1. / (W / 2))  ref :149                      */
    float inv_half_h;      /* float(1. / (H / 2))  ref :150                      */
    float corners[12];     /* 3x4 homogeneous corners, row-major values ref :18  */
} oracle_camera;

/* torch.cos on CPU == MKL VML vmsCos HA; restated near the end of this file */
float oracle_cosf_mkl_ha(float d);
float oracle_sinf_mkl_ha(float d);

/* ref :6-24 */
API void vidc_oracle_camera_init(double fx, double fy, double cx, double cy, oracle_camera *cam)
{
    cam->W = (int32_t)ceil(2.0 * cx);
    cam->H = (int32_t)ceil(2.0 * cy);
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    /* np.linalg.inv == LAPACK getrf/getri; for this upper-triangular K no pivoting
       happens and trtri gives inv = [[1/fx,0,-(cx*(1/fx))],[0,1/fy,-(cy*(1/fy))],[0,0,1]]. */
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double Ki[9] = {ifx, 0, -(cx * ifx), 0, ify, -(cy * ify), 0, 0, 1};
    for (int i = 0; i < 9; ++i) { cam->K[i] = (float)K[i]; cam->Kinv[i] = (float)Ki[i]; }
    cam->cx = (float)cx;
    cam->cy = (float)cy;
    cam->inv_half_w = (float)(1.0 / ((double)cam->W / 2.0));
    cam->inv_half_h = (float)(1.0 / ((double)cam->H / 2.0));
    const float Wm = (float)(cam->W - 1), Hm = (float)(cam->H - 1);
    const float c[12] = {0, Wm, 0, Wm, 0, 0, Hm, Hm, 1, 1, 1, 1};
    memcpy(cam->corners, c, sizeof c);
}

/* dot products of length 3 in the accumulate orders MKL uses for each call site
   (discovered by brute force against the executed reference, DESIGN.md "schemes") */
static inline float dot3_muladd(const float *a, int sa, const float *b, int sb)
{   /* (a0 b0 + a1 b1) + a2 b2, every op rounded: bmm / broadcast-matmul class */
    float p0 = a[0] * b[0], p1 = a[sa] * b[sb], p2 = a[2 * sa] * b[2 * sb];
    float s = p0 + p1;
    return s + p2;
}
static inline float dot3_fma(const float *a, int sa, const float *b, int sb)
{   /* k-ascending FMA chain: mm class */
    float s = a[0] * b[0];
    s = fmaf(a[sa], b[sb], s);
    return fmaf(a[2 * sa], b[2 * sb], s);
}
static inline float dot3_021(const float *a, int sa, const float *b, int sb)
{   /* (a0 b0 + a2 b2) + a1 b1, no FMA: mm with a column-major right operand (ref :18,:125) */
    float p0 = a[0] * b[0], p1 = a[sa] * b[sb], p2 = a[2 * sa] * b[2 * sb];
    float s = p0 + p2;
    return s + p1;
}

static void skew(const float *x, float *S)
{   /* ref :26-32 */
    S[0] = 0.0f;  S[1] = -x[2]; S[2] = x[1];
    S[3] = x[2];  S[4] = 0.0f;  S[5] = -x[0];
    S[6] = -x[1]; S[7] = x[0];  S[8] = 0.0f;
}

/* ref :35-58.  Ig, Ia: (B,3).  Outputs (B,3,3) row-major each. */
API void vidc_oracle_build_homography(const oracle_camera *cam, const float *Ig, const float *Ia,
                                      int B, float *Hm, float *Rm, float *Hinv)
{
    for (int i = 0; i < B; ++i) {
        const float *g = Ig + 3 * i, *a = Ia + 3 * i;
        float nS[9], S[9], q[3];
        skew(a, S);
        for (int k = 0; k < 9; ++k) nS[k] = -S[k];                       /* :41 */
        for (int r = 0; r < 3; ++r) q[r] = dot3_muladd(nS + 3 * r, 1, g, 1); /* :42 bmm */
        const float d = dot3_muladd(a, 1, g, 1);                          /* :43 bmm */
        /* :44 Tensor.norm(dim=1) over 3 elements: sqrt of a k-ascending FMA chain */
        float ss = q[0] * q[0];
        ss = fmaf(q[1], q[1], ss);
        ss = fmaf(q[2], q[2], ss);
        const float n = sqrtf(ss);
        /* :48 atan2 on two 0-dim tensors -> scalar loop -> glibc atan2f;
               0.5*t ; cos on a 0-dim tensor -> MKL VML vmsCos (HA) */
        const float q4 = oracle_cosf_mkl_ha(0.5f * atan2f(n, d));
        /* :49-50 identity branch is overwritten at :53 -> no effect */
        const float two_q4 = 2.0f * q4;
        for (int r = 0; r < 3; ++r) q[r] = q[r] / two_q4;                 /* :51 */
        skew(q, S);                                                        /* :52 */
        float *R = Rm + 9 * i;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const float I3 = (r == c) ? 1.0f : 0.0f;
                const float t1 = two_q4 * S[3 * r + c];                    /* 2.*q4*S      */
                float S2r[3] = {2.0f * S[3 * r], 2.0f * S[3 * r + 1], 2.0f * S[3 * r + 2]};
                const float t2 = dot3_fma(S2r, 1, S + c, 3);               /* (2.*S) @ S mm */
                R[3 * r + c] = (I3 + t1) + t2;                             /* :53-54       */
            }
        /* :55 H = (K @ R) @ Kinv ; :56-57 Hinv = (K @ R^T) @ Kinv */
        float KR[9], KRt[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                KR[3 * r + c]  = dot3_muladd(cam->K + 3 * r, 1, R + c, 3);
                KRt[3 * r + c] = dot3_muladd(cam->K + 3 * r, 1, R + 3 * c, 1);
            }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                Hm[9 * i + 3 * r + c]   = dot3_fma(KR + 3 * r, 1, cam->Kinv + c, 3);
                Hinv[9 * i + 3 * r + c] = dot3_fma(KRt + 3 * r, 1, cam->Kinv + c, 3);
            }
    }
}

/* ref :125-140 (identically :168-194, :226-240).
   out[8] = px_min, py_min, kw, kh, ikw, ikh, w_max, h_max */
API void vidc_oracle_frame_scale(const oracle_camera *cam, const float *Hm, float *out)
{
    float px[4], py[4];
    for (int j = 0; j < 4; ++j) {
        const float c0 = dot3_021(Hm + 0, 1, cam->corners + j, 4);
        const float c1 = dot3_021(Hm + 3, 1, cam->corners + j, 4);
        const float c2 = dot3_021(Hm + 6, 1, cam->corners + j, 4);
        px[j] = c0 / c2;
        py[j] = c1 / c2;
    }
    /* torch.max/min over 4 values; NaN propagates in torch */
    float px_max = px[0], px_min = px[0], py_max = py[0], py_min = py[0];
    for (int j = 1; j < 4; ++j) {
        if (px[j] > px_max || px[j] != px[j]) px_max = (px_max != px_max) ? px_max : px[j];
        if (px[j] < px_min || px[j] != px[j]) px_min = (px_min != px_min) ? px_min : px[j];
        if (py[j] > py_max || py[j] != py[j]) py_max = (py_max != py_max) ? py_max : py[j];
        if (py[j] < py_min || py[j] != py[j]) py_min = (py_min != py_min) ? py_min : py[j];
    }
    const float h_max = py_max - py_min, w_max = px_max - px_min;
    const float Wf = (float)cam->W, Hf = (float)cam->H;
    float kw, kh;
    /* scalar / tensor == tensor.reciprocal() * scalar (torch/_tensor.py __rdiv__) */
    if (w_max > (4.0f * h_max) / 3.0f) {
        kw = (1.0f / w_max) * Wf;
        kh = (1.0f / ((3.0f * w_max) / 4.0f)) * Hf;
    } else {
        kh = (1.0f / h_max) * Hf;
        kw = (1.0f / ((4.0f * h_max) / 3.0f)) * Wf;
    }
    out[0] = px_min; out[1] = py_min; out[2] = kw; out[3] = kh;
    out[4] = (1.0f / kw) * 1.0f;   /* "1./kw" ref :142 */
    out[5] = (1.0f / kh) * 1.0f;   /* "1./kh" ref :143 */
    out[6] = w_max;  out[7] = h_max;
}

/* ref :142-150: canvas pixel (X,Y) -> normalised source coordinates.  grid (H,W,2). */
static void forward_grid_frame(const oracle_camera *cam, const float *Hinv, const float *sc, float *grid)
{
    const float px_min = sc[0], py_min = sc[1], ikw = sc[4], ikh = sc[5];
    for (int Y = 0; Y < cam->H; ++Y)
        for (int X = 0; X < cam->W; ++X) {
            const float px = ikw * (float)X + px_min;
            const float py = ikh * (float)Y + py_min;
            /* (3,3)@(3,WH) mm: k-ascending FMA chain; the third row of P is ones */
            const float u = fmaf(Hinv[2], 1.0f, fmaf(Hinv[1], py, Hinv[0] * px));
            const float v = fmaf(Hinv[5], 1.0f, fmaf(Hinv[4], py, Hinv[3] * px));
            const float s = fmaf(Hinv[8], 1.0f, fmaf(Hinv[7], py, Hinv[6] * px));
            const float sx = u / s, sy = v / s;
            float *o = grid + 2 * ((size_t)Y * cam->W + X);
            o[0] = cam->inv_half_w * (sx - cam->cx);
            o[1] = cam->inv_half_h * (sy - cam->cy);
        }
}

/* ref :242-249: camera pixel (X,Y) -> normalised canvas coordinates */
static void inverse_grid_frame(const oracle_camera *cam, const float *Hm, const float *sc, float *grid)
{
    const float px_min = sc[0], py_min = sc[1], kw = sc[2], kh = sc[3];
    for (int Y = 0; Y < cam->H; ++Y)
        for (int X = 0; X < cam->W; ++X) {
            const float fx = (float)X, fy = (float)Y;
            const float u = fmaf(Hm[2], 1.0f, fmaf(Hm[1], fy, Hm[0] * fx));
            const float v = fmaf(Hm[5], 1.0f, fmaf(Hm[4], fy, Hm[3] * fx));
            const float s = fmaf(Hm[8], 1.0f, fmaf(Hm[7], fy, Hm[6] * fx));
            const float tx = u / s, ty = v / s;
            const float cxp = kw * (tx - px_min);
            const float cyp = kh * (ty - py_min);
            float *o = grid + 2 * ((size_t)Y * cam->W + X);
            o[0] = cam->inv_half_w * (cxp - cam->cx);
            o[1] = cam->inv_half_h * (cyp - cam->cy);
        }
}

static void identity_grid_frame(const oracle_camera *cam, float *grid)
{   /* ref :181-186 */
    for (int Y = 0; Y < cam->H; ++Y)
        for (int X = 0; X < cam->W; ++X) {
            float *o = grid + 2 * ((size_t)Y * cam->W + X);
            o[0] = cam->inv_half_w * ((float)X - cam->cx);
            o[1] = cam->inv_half_h * ((float)Y - cam->cy);
        }
}

API void vidc_oracle_forward_grid(const oracle_camera *cam, const float *Hm, const float *Hinv, int B, float *grid)
{
    const size_t fs = (size_t)cam->H * cam->W * 2;
    for (int i = 0; i < B; ++i) {
        float sc[8];
        vidc_oracle_frame_scale(cam, Hm + 9 * i, sc);
        forward_grid_frame(cam, Hinv + 9 * i, sc, grid + fs * i);
    }
}

API void vidc_oracle_inverse_grid(const oracle_camera *cam, const float *Hm, int B, float *grid)
{
    const size_t fs = (size_t)cam->H * cam->W * 2;
    for (int i = 0; i < B; ++i) {
        float sc[8];
        vidc_oracle_frame_scale(cam, Hm + 9 * i, sc);
        inverse_grid_frame(cam, Hm + 9 * i, sc, grid + fs * i);
    }
}

/* ref :158-214: returns R^T (or I3 under the aspect guard), forward and inverse grids */
API void vidc_oracle_sampler_forward_inverse(const oracle_camera *cam, const float *Ig, const float *Ia, int B,
                                             float *Rt_ret, float *grid, float *inv_grid)
{
    float *Hm = malloc(sizeof(float) * 27 * (size_t)B), *Rm = Hm + 9 * (size_t)B, *Hi = Rm + 9 * (size_t)B;
    vidc_oracle_build_homography(cam, Ig, Ia, B, Hm, Rm, Hi);
    const size_t fs = (size_t)cam->H * cam->W * 2;
    for (int i = 0; i < B; ++i) {
        float sc[8];
        vidc_oracle_frame_scale(cam, Hm + 9 * i, sc);
        const float sigma = sc[6] / sc[7];                               /* :178 */
        if (sigma < 0.8f || sigma > 2.2f) {                              /* :179, python doubles -> fp32 compare */
            for (int k = 0; k < 9; ++k) Rt_ret[9 * i + k] = (k % 4 == 0) ? 1.0f : 0.0f;
            identity_grid_frame(cam, grid + fs * i);
            identity_grid_frame(cam, inv_grid + fs * i);
            continue;
        }
        forward_grid_frame(cam, Hi + 9 * i, sc, grid + fs * i);
        inverse_grid_frame(cam, Hm + 9 * i, sc, inv_grid + fs * i);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rt_ret[9 * i + 3 * r + c] = Rm[9 * i + 3 * c + r];
    }
    free(Hm);
}

/* ---------------------------------------------------------------------------------------
 * ATen grid_sampler_2d, align_corners=False, padding_mode='zeros' (torch 2.11).
 * Helper source visible at torch/include/ATen/native/cuda/GridSampler.cuh:23-31 (unnormalise),
 * :140-147 (safe_downgrade_to_int_range), :220-222 (within_bounds_2d); the loop body is
 * restated from upstream GridSampler.cu / GridSamplerKernel.cpp.
 * mode: 0 = bilinear, 1 = nearest.
 * Non-finite or |coord| > INT_MAX-1 is treated as out of bounds -> contributes 0
 * (the CUDA kernel's behaviour; the CPU kernel would produce NaN there).
 * ------------------------------------------------------------------------------------- */
static inline float unnormalize(float g, int size)
{   /* ((g + 1) * size - 1) / 2 as compiled: both the CPU build (GCC -ffp-contract=fast on
       "(in + 1) * (size/2) - 0.5") and the CUDA build (nvcc -fmad=true) contract the
       multiply-subtract into ONE fma; fma(g+1, size, -1)/2 == fma(g+1, size/2, -0.5) exactly. */
    return fmaf(g + 1.0f, (float)size, -1.0f) / 2.0f;
}
static inline float safe_downgrade(float x)
{
    if (x > 2147483646.0f || x < -2147483648.0f || !isfinite(x)) return -100.0f;
    return x;
}

/* cubic convolution coefficients for the taps at -1, 0, +1, +2 (ATen get_cubic_upsample_coefficients, A = -0.75) */
static void bicubic_coefficients(float t, float c[4])
{
    const float A = -0.75f;
    float x = t + 1.0f;
    c[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    x = t;
    c[1] = fmaf(fmaf(A + 2.0f, x, -(A + 3.0f)) * x, x, 1.0f);
    x = 1.0f - t;
    c[2] = fmaf(fmaf(A + 2.0f, x, -(A + 3.0f)) * x, x, 1.0f);
    x = 2.0f - t;
    c[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

API void vidc_oracle_grid_sample(const float *x, int B, int C, int Hin, int Win,
                                 const float *grid, int Hout, int Wout, int mode, float *out)
{
    for (int b = 0; b < B; ++b)
        for (int Y = 0; Y < Hout; ++Y)
            for (int X = 0; X < Wout; ++X) {
                const float *g = grid + 2 * (((size_t)b * Hout + Y) * Wout + X);
                const float ix = safe_downgrade(unnormalize(g[0], Win));
                const float iy = safe_downgrade(unnormalize(g[1], Hin));
                const size_t opix = (size_t)Y * Wout + X;
                if (mode == 0) {
                    const float x0f = floorf(ix), y0f = floorf(iy);
                    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
                    const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
                    const float w_nw = (x1f - ix) * (y1f - iy);
                    const float w_ne = (ix - x0f) * (y1f - iy);
                    const float w_sw = (x1f - ix) * (iy - y0f);
                    const float w_se = (ix - x0f) * (iy - y0f);
                    const int in_x0 = x0 >= 0 && x0 < Win, in_x1 = x1 >= 0 && x1 < Win;
                    const int in_y0 = y0 >= 0 && y0 < Hin, in_y1 = y1 >= 0 && y1 < Hin;
                    for (int c = 0; c < C; ++c) {
                        const float *p = x + ((size_t)b * C + c) * Hin * Win;
                        /* ATen gathers 0 for an out-of-bounds tap and still runs it through the chain
                           nw*w + ne*w + sw*w + se*w (first term a product, the rest contracted to FMAs): this
                           decides the SIGN of a zero result (all-(-0) taps give -0, any padded tap gives +0) */
                        const float v_nw = (in_y0 && in_x0) ? p[(size_t)y0 * Win + x0] : 0.0f;
                        const float v_ne = (in_y0 && in_x1) ? p[(size_t)y0 * Win + x1] : 0.0f;
                        const float v_sw = (in_y1 && in_x0) ? p[(size_t)y1 * Win + x0] : 0.0f;
                        const float v_se = (in_y1 && in_x1) ? p[(size_t)y1 * Win + x1] : 0.0f;
                        float acc = v_nw * w_nw;
                        acc = fmaf(v_ne, w_ne, acc);
                        acc = fmaf(v_sw, w_sw, acc);
                        acc = fmaf(v_se, w_se, acc);
                        out[((size_t)b * C + c) * Hout * Wout + opix] = acc;
                    }
                } else if (mode == 2) {
                    /* bicubic (A = -0.75), ATen GridSamplerKernel.cpp / UpSample.h as this torch build rounds it (found by
                       search against F.grid_sample on CPU, 0 mismatches): coordinates are NOT clipped before the weights
                       (a non-finite coordinate gives NaN weights, hence NaN, on the CPU and the CUDA build alike);
                       conv2(x) = ((A x - 5A) x + 8A) x - 4A with every operation rounded,
                       conv1(x) = fma(fma(A + 2, x, -(A + 3)) * x, x, 1);
                       row = (fma(c0, v0, c1 v1) + c2 v2) + c3 v3;  result = fma chain over the four rows from cy0 * row0. */
                    const float rx = unnormalize(g[0], Win), ry = unnormalize(g[1], Hin);
                    const float x0f = floorf(rx), y0f = floorf(ry);
                    const float tx = rx - x0f, ty = ry - y0f;
                    float cx[4], cy[4];
                    bicubic_coefficients(tx, cx);
                    bicubic_coefficients(ty, cy);
                    for (int c = 0; c < C; ++c) {
                        const float *p = x + ((size_t)b * C + c) * Hin * Win;
                        float rows[4];
                        for (int i = 0; i < 4; ++i) {
                            float v[4];
                            for (int j = 0; j < 4; ++j) {
                                const float xf = safe_downgrade(x0f + (float)(j - 1)), yf = safe_downgrade(y0f + (float)(i - 1));
                                const int xi = (int)xf, yi = (int)yf;
                                v[j] = (xi >= 0 && xi < Win && yi >= 0 && yi < Hin) ? p[(size_t)yi * Win + xi] : 0.0f;
                            }
                            float acc = fmaf(cx[0], v[0], cx[1] * v[1]);
                            acc = acc + cx[2] * v[2];
                            acc = acc + cx[3] * v[3];
                            rows[i] = acc;
                        }
                        float acc = cy[0] * rows[0];
                        acc = fmaf(cy[1], rows[1], acc);
                        acc = fmaf(cy[2], rows[2], acc);
                        acc = fmaf(cy[3], rows[3], acc);
                        out[((size_t)b * C + c) * Hout * Wout + opix] = acc;
                    }
                } else {
                    const float xr = nearbyintf(ix), yr = nearbyintf(iy);
                    const int xn = (int)xr, yn = (int)yr;
                    const int in = xn >= 0 && xn < Win && yn >= 0 && yn < Hin;
                    for (int c = 0; c < C; ++c) {
                        const float *p = x + ((size_t)b * C + c) * Hin * Win;
                        out[((size_t)b * C + c) * Hout * Wout + opix] = in ? p[(size_t)yn * Win + xn] : 0.0f;
                    }
                }
            }
}

/* ref :108-156.  x (B,C,Hin,Win) -> y (B,C,H,W); also returns H (B,3,3). */
API void vidc_oracle_warp_forward(const oracle_camera *cam, const float *x, int B, int C, int Hin, int Win,
                                  const float *Ig, const float *Ia, int mode, float *Hm_out, float *y)
{
    float *Hm = malloc(sizeof(float) * 27 * (size_t)B), *Rm = Hm + 9 * (size_t)B, *Hi = Rm + 9 * (size_t)B;
    const size_t fs = (size_t)cam->H * cam->W * 2;
    float *grid = malloc(sizeof(float) * fs * (size_t)B);
    vidc_oracle_build_homography(cam, Ig, Ia, B, Hm, Rm, Hi);
    vidc_oracle_forward_grid(cam, Hm, Hi, B, grid);
    vidc_oracle_grid_sample(x, B, C, Hin, Win, grid, cam->H, cam->W, mode, y);
    if (Hm_out) memcpy(Hm_out, Hm, sizeof(float) * 9 * (size_t)B);
    free(grid); free(Hm);
}

/* ref :216-255.  x (B,3,H,W) normals in the aligned frame -> z = R^T * sample(x) (B,3,H,W).
   z is NOT normalised here (that is surface_normal.py:170, vidc_oracle_normalize). */
API void vidc_oracle_inverse_warp_normals(const oracle_camera *cam, const float *x, int B,
                                          const float *Ig, const float *Ia, float *Hm_out, float *z)
{
    const int Hh = cam->H, Ww = cam->W;
    const size_t hw = (size_t)Hh * Ww;
    float *Hm = malloc(sizeof(float) * 27 * (size_t)B), *Rm = Hm + 9 * (size_t)B, *Hi = Rm + 9 * (size_t)B;
    float *grid = malloc(sizeof(float) * hw * 2 * (size_t)B);
    float *y = malloc(sizeof(float) * hw * 3 * (size_t)B);
    vidc_oracle_build_homography(cam, Ig, Ia, B, Hm, Rm, Hi);
    vidc_oracle_inverse_grid(cam, Hm, B, grid);
    vidc_oracle_grid_sample(x, B, 3, Hh, Ww, grid, Hh, Ww, 0, y);
    for (int b = 0; b < B; ++b) {
        const float *R = Rm + 9 * b;              /* C_R_Cg = R^T: z_c = sum_k R[k][c] y_k  (:253 bmm) */
        const float *yb = y + 3 * hw * b;
        float *zb = z + 3 * hw * b;
        for (size_t p = 0; p < hw; ++p)
            for (int c = 0; c < 3; ++c) {
                float s = fmaf(R[c], yb[p], 0.0f);   /* the GEMM accumulator starts at +0: I * (-0) = +0 */
                s = fmaf(R[3 + c], yb[hw + p], s);
                s = fmaf(R[6 + c], yb[2 * hw + p], s);
                zb[c * hw + p] = s;
            }
    }
    if (Hm_out) memcpy(Hm_out, Hm, sizeof(float) * 9 * (size_t)B);
    free(y); free(grid); free(Hm);
}

/* networks/surface_normal.py:170  F.normalize(z, dim=1): z / max(||z||_2, 1e-12) */
API void vidc_oracle_normalize(const float *z, int B, int C, size_t hw, float *out)
{
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < hw; ++p) {
            const float *zb = z + (size_t)b * C * hw + p;
            /* CPU norm kernel: squares summed left to right, no FMA (found by brute force) */
            float ss = 0.0f;
            for (int c = 0; c < C; ++c) { const float sq = zb[c * hw] * zb[c * hw]; ss = (c == 0) ? sq : ss + sq; }
            float n = sqrtf(ss);
            if (n < 1e-12f) n = 1e-12f;
            for (int c = 0; c < C; ++c) out[(size_t)b * C * hw + c * hw + p] = zb[c * hw] / n;
        }
}

/* networks/surface_normal.py:151  mask = (R + G) + B > float(1e-2), on the rounded fp32 values */
API void vidc_oracle_mask(const float *x1, int B, size_t hw, uint8_t *mask)
{
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < hw; ++p) {
            const float *xb = x1 + (size_t)b * 3 * hw + p;
            const float s = (xb[0] + xb[hw]) + xb[2 * hw];
            mask[(size_t)b * hw + p] = s > 0.01f;
        }
}

/* networks/surface_normal.py:153-156  F.interpolate(mask, size, 'nearest'):
   src = min(floor(dst * (in/out as float)), in-1) */
API void vidc_oracle_mask_nearest(const uint8_t *mask, int B, int Hin, int Win, int Hout, int Wout, uint8_t *out)
{
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < Hout; ++y) {
            int sy = (int)floorf((float)y * sh); if (sy > Hin - 1) sy = Hin - 1;
            for (int x = 0; x < Wout; ++x) {
                int sx = (int)floorf((float)x * sw); if (sx > Win - 1) sx = Win - 1;
                out[((size_t)b * Hout + y) * Wout + x] = mask[((size_t)b * Hin + sy) * Win + sx];
            }
        }
}

/* normal_utils.py:20-34 with Normalize := F.normalize(., dim=1) (undefined in the reference).
   Returns double-accumulated statistics: out[0] = sum over pixels of angle_deg * mask,
   out[1] = sum(mask), out[2] = L1 sum |norms*mask - gt*mask| (divide by out[1] for the loss).
   The per-pixel angle is fp32 as in the reference; only the final reductions are order-free here. */
API void vidc_oracle_normal_stats(const float *gt, const float *pred, const float *mask,
                                  int B, size_t hw, int normalize_prediction, double *out)
{
    double angle_sum = 0.0, msum = 0.0, l1 = 0.0;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < hw; ++p) {
            const float *pb = pred + (size_t)b * 3 * hw + p, *gb = gt + (size_t)b * 3 * hw + p;
            const float m = mask[(size_t)b * hw + p];
            float n0 = pb[0], n1 = pb[hw], n2 = pb[2 * hw];
            if (normalize_prediction) {
                float ss = (n0 * n0 + n1 * n1) + n2 * n2;
                float nn = sqrtf(ss); if (nn < 1e-12f) nn = 1e-12f;
                n0 /= nn; n1 /= nn; n2 /= nn;
            }
            float dp = (n0 * gb[0] + n1 * gb[hw]) + n2 * gb[2 * hw];
            if (dp < -1.0f) dp = -1.0f;
            if (dp > 1.0f) dp = 1.0f;
            const float ang = (float)((double)acosf(dp) / 3.14159265358979323846 * 180.0);
            angle_sum += (double)(ang * m);
            msum += m;
            l1 += fabs((double)(n0 * m) - (double)(gb[0] * m)) + fabs((double)(n1 * m) - (double)(gb[hw] * m))
                + fabs((double)(n2 * m) - (double)(gb[2 * hw] * m));
        }
    out[0] = angle_sum; out[1] = msum; out[2] = l1;
}

/* ---------------------------------------------------------------------------------------
 * torch.cos on a CPU float tensor (any numel, including the 0-dim tensor at ref :48) is
 * dispatched by ATen to MKL VML vmsCos(n, in, out, VML_HA | ...)  (ATen/cpu/vml.h).
 * MKL is closed source; the routine below restates the main path (|x| <= 10000) of
 * MKL 2024.2's AVX-512 "HA" kernel as observed in the build container:
 *   N  = rint((|x| + pi/2) / pi)            (fp32, via the 1.5*2^23 shifter)
 *   r  = |x| - (N - 0.5) * pi               (fp64, two-term pi, FMA)
 *   sin(r) ~ r + r * (r^2 * P(r^2))         (fp64, degree-3 P, FMA Horner)
 *   cos(x) = (-1)^N * (float) sin(r)
 * Verified bit-for-bit against torch.cos for EVERY float in [0, 1.6] (1.07e9 values) by
 * oracle/check_math_vs_torch.py in the build container.
 * ------------------------------------------------------------------------------------- */
float oracle_cosf_mkl_ha(float x)
{
    const float HALFPI = 0x1.921fb6p+0f, INVPI = 0x1.45f306p-2f, SHIFTER = 0x1.8p+23f;
    const double PI_HI = 0x1.921fb5444p+1, PI_LO = 0x1.68c234c4c6629p-38;
    const double C3 = -0x1.55554bc836587p-3, C5 = 0x1.110ed3804ca96p-7,
                 C7 = -0x1.9f6ffeea73463p-13, C9 = 0x1.5dbdf0e4c7deep-19;
    const float ax = fabsf(x);
    const float t = ax + HALFPI;
    const float y = fmaf(t, INVPI, SHIFTER);
    uint32_t ybits; memcpy(&ybits, &y, 4);
    float n = y - SHIFTER;
    n = n - 0.5f;
    const double dn = (double)n;
    double r = (double)ax;
    r = fma(-PI_HI, dn, r);
    r = fma(-dn, PI_LO, r);
    const double r2 = r * r;
    double p = fma(C9, r2, C7);
    p = fma(r2, p, C5);
    p = fma(r2, p, C3);
    const double q = p * r2;
    const float f = (float)fma(r, q, r);
    uint32_t fb; memcpy(&fb, &f, 4);
    fb ^= ybits << 31;
    float out; memcpy(&out, &fb, 4);
    return out;
}


/* ---------------------------------------------------------------------------------------
 * Batch-parallel wrappers (frames are independent): the reference's CPU path runs with all
 * host threads torch/MKL give it, so the cpu_baseline / --impl reference legs of bench.py use
 * these to spread frames over `nthreads` POSIX threads.  Results are identical to the serial
 * functions (same per-frame code).
 * ------------------------------------------------------------------------------------- */
typedef struct {
    const oracle_camera *cam; int B0, B1;
    const float *rgb, *depth, *normals, *Ig, *Ia;
    float *rgb_w, *depth_w, *normals_cam; uint8_t *mask;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    const oracle_camera *cam = j->cam;
    const size_t hw = (size_t)cam->H * cam->W;
    const int n = j->B1 - j->B0;
    if (n <= 0) return NULL;
    const size_t o = (size_t)j->B0;
    vidc_oracle_warp_forward(cam, j->rgb + 3 * hw * o, n, 3, cam->H, cam->W, j->Ig + 3 * o, j->Ia + 3 * o, 0, NULL, j->rgb_w + 3 * hw * o);
    if (j->depth) vidc_oracle_warp_forward(cam, j->depth + hw * o, n, 1, cam->H, cam->W, j->Ig + 3 * o, j->Ia + 3 * o, 0, NULL, j->depth_w + hw * o);
    vidc_oracle_mask(j->rgb_w + 3 * hw * o, n, hw, j->mask + hw * o);
    float *tmp = malloc(sizeof(float) * 3 * hw * (size_t)n);
    vidc_oracle_inverse_warp_normals(cam, j->normals + 3 * hw * o, n, j->Ig + 3 * o, j->Ia + 3 * o, NULL, tmp);
    vidc_oracle_normalize(tmp, n, 3, hw, j->normals_cam + 3 * hw * o);
    free(tmp);
    return NULL;
}

/* One "frame step" of the hot path for B frames: warp RGB, warp depth (bilinear), mask,
   inverse-warp normals, renormalise -- exactly what bench.py times on the GPU. */
API void vidc_oracle_warp_unwarp_mt(const oracle_camera *cam, int B, int nthreads,
                                    const float *rgb, const float *depth, const float *normals,
                                    const float *Ig, const float *Ia,
                                    float *rgb_w, float *depth_w, uint8_t *mask, float *normals_cam)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    mt_job *jobs = malloc(sizeof(mt_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        mt_job j = {cam, (int)((long long)B * t / nthreads), (int)((long long)B * (t + 1) / nthreads),
                    rgb, depth, normals, Ig, Ia, rgb_w, depth_w, normals_cam, mask};
        jobs[t] = j;
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
}


/* torch.sin on a CPU float tensor == MKL VML vmsSin(VML_HA); same structure and constants as the cosine:
   N = rint(|x| / pi), r = |x| - N pi (fp64), sin x = sign(x) (-1)^N (float)(r + r (r^2 P(r^2))). */
float oracle_sinf_mkl_ha(float x)
{
    const float INVPI = 0x1.45f306p-2f, SHIFTER = 0x1.8p+23f;
    const double PI_HI = 0x1.921fb5444p+1, PI_LO = 0x1.68c234c4c6629p-38;
    const double C3 = -0x1.55554bc836587p-3, C5 = 0x1.110ed3804ca96p-7,
                 C7 = -0x1.9f6ffeea73463p-13, C9 = 0x1.5dbdf0e4c7deep-19;
    uint32_t xb; memcpy(&xb, &x, 4);
    const float ax = fabsf(x);
    const float y = fmaf(ax, INVPI, SHIFTER);
    uint32_t ybits; memcpy(&ybits, &y, 4);
    const float n = y - SHIFTER;
    const double dn = (double)n;
    double r = (double)ax;
    r = fma(-PI_HI, dn, r);
    r = fma(-dn, PI_LO, r);
    const double r2 = r * r;
    double p = fma(C9, r2, C7);
    p = fma(r2, p, C5);
    p = fma(r2, p, C3);
    const double q = p * r2;
    const float f = (float)fma(r, q, r);
    uint32_t fb; memcpy(&fb, &f, 4);
    fb ^= (ybits << 31) ^ (xb & 0x80000000u);
    float out; memcpy(&out, &fb, 4);
    return out;
}
API void vidc_oracle_sinf_array(const float *in, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = oracle_sinf_mkl_ha(in[i]);
}

/* dataset.py gravity conditioning.
   rule 0: Azure / Demo loaders (:334-345, :472-483): flip y,z; psi < 1e-4 -> a = [0,1,0];
           cos(pitch) > 0.707 -> [0,1,0] else [0, cos(pitch), sin(pitch)].
   rule 1: ScanNet compute_alignment_tensor (:45-55), no flip: psi < 1e-6 -> a = g; cos(pitch) > 0.3 -> [0,1,0] else a = g. */
API void vidc_oracle_condition_gravity(const float *raw, int B, int rule, float *Ig, float *Ia)
{
    for (int i = 0; i < B; ++i) {
        float g0 = raw[3 * i], g1 = raw[3 * i + 1], g2 = raw[3 * i + 2];
        if (rule == 0) { g1 = -g1; g2 = -g2; }
        const float a1 = g1 * g1, a2 = g2 * g2;
        const float psi = a1 + a2;
        float a[3] = {0.0f, 1.0f, 0.0f};
        if (rule == 0) {
            if (!(psi < 1e-4f)) {
                const float pitch = atan2f(g2, g1);
                const float c = oracle_cosf_mkl_ha(pitch);
                if (!(c > 0.707f)) { a[0] = 0.0f; a[1] = c; a[2] = oracle_sinf_mkl_ha(pitch); }
            }
        } else {
            if (psi < 1e-6f) { a[0] = g0; a[1] = g1; a[2] = g2; }
            else {
                const float pitch = atan2f(g2, g1);
                if (!(oracle_cosf_mkl_ha(pitch) > 0.3f)) { a[0] = g0; a[1] = g1; a[2] = g2; }
            }
        }
        Ig[3 * i] = g0; Ig[3 * i + 1] = g1; Ig[3 * i + 2] = g2;
        Ia[3 * i] = a[0]; Ia[3 * i + 1] = a[1]; Ia[3 * i + 2] = a[2];
    }
}

/* array wrappers so the tests can sweep the scalar math functions */
API void vidc_oracle_cosf_array(const float *in, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = oracle_cosf_mkl_ha(in[i]);
}
API void vidc_oracle_atan2f_array(const float *y, const float *x, size_t n, float *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = atan2f(y[i], x[i]);
}
