// kernels_params.cuh -- per-frame kernels: parameters, gravity conditioning, sparse-depth rasterisation
// Part of libvidc_b200.so (one translation unit, vidc_kernels.cu); compiled with -fmad=false.
#pragma once
#include "device_common.cuh"

namespace vidc_k {

// ------------------------------------------------------------------------------------------
__global__ void frame_params_kernel(vidc_camera cam, const float* __restrict__ Ig, const float* __restrict__ Ia,
                                    int B, vidc_frame_params* __restrict__ out, float* __restrict__ H_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float g[3] = {Ig[3 * i], Ig[3 * i + 1], Ig[3 * i + 2]};
    const float a[3] = {Ia[3 * i], Ia[3 * i + 1], Ia[3 * i + 2]};
    vidc_frame_params p;
    vidc::frame_params_from_gravity(cam, g, a, p);
    // Orientation of the gather: when the source x coordinate changes much faster along a canvas COLUMN than along a canvas
    // row (roll beyond ~76 deg; the row-major kernels fall off a cliff near 90 deg, profiles/r1_history.md) the kernels
    // switch to their column-major tile path.
    p.fwd_col_major = fabsf(p.Hinv[1] * p.ikh) > 4.0f * fabsf(p.Hinv[0] * p.ikw) ? 1.0f : 0.0f;
    p.inv_col_major = fabsf(p.H[1]) > 4.0f * fabsf(p.H[0]) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 10; ++k) p.reserved[k] = 0.0f;
    // reserved[10]: 1.0 when the inverse warp's shared-reciprocal division is provably exact for every pixel of this frame
    // (vidc::inv_division_proven) -- the inverse kernels then skip their per-pixel window test
    p.reserved[10] = vidc::inv_division_proven(p, cam) ? 1.0f : 0.0f;
    out[i] = p;
    if (H_out) {                       // the Cg_H_C every reference method returns (:153-156, :255)
#pragma unroll
        for (int k = 0; k < 9; ++k) H_out[9 * i + k] = p.H[k];
    }
}

// ---- exterior-tile bitmap ------------------------------------------------------------------------------------------------
// About 40 % of the forward canvas lies outside the source footprint, most of it in whole 32x32 tiles.  One CTA per frame:
// thread 0 builds the parameters exactly as frame_params_kernel does, then thread t classifies canvas tile t and the
// verdicts go into vidc_frame_params::reserved as a bitmap (<= 320 tiles: 640x480 has 300).  A set bit means CERTAINLY
// exterior, by a conservative test; the forward kernels then write the tile's zeros without computing a coordinate.
//
// Why the test is safe.  Over a tile the projective denominator s is affine in (X, Y): if its four corner values share a
// sign it keeps that sign over the tile, the map is continuous there and takes the tile into the convex quadrilateral of
// the four mapped corners -- every pixel's exact source coordinate lies inside the corners' bounding box.  The fp32 values
// the kernels compute differ from the exact ones by at most a few ulps of the terms of u, v, s; with the conditioning
// demanded below (|s| > 1e-3 of the sum of its terms' magnitudes) that is < 1e-4 of the coordinate's magnitude plus a
// fraction of a pixel, at the corners here and at every pixel there.  The margin 4 px + 2e-3 |coordinate| is 20x that.
// Anything that fails a condition (sign change, ill-conditioned or non-finite corner) is simply not marked.
// The test itself is vidc::tile_certainly_exterior (frame_params.cuh, host / device: tests/ run it on the CPU).
__global__ void __launch_bounds__(320) frame_params_tiles_kernel(vidc_camera cam, const float* __restrict__ Ig,
                                                                 const float* __restrict__ Ia, int B,
                                                                 vidc_frame_params* __restrict__ out, float* __restrict__ H_out,
                                                                 uint4* __restrict__ src_boxes) {
    __shared__ vidc_frame_params sp;
    const int i = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        const float g[3] = {Ig[3 * i], Ig[3 * i + 1], Ig[3 * i + 2]};
        const float a[3] = {Ia[3 * i], Ia[3 * i + 1], Ia[3 * i + 2]};
        vidc_frame_params p;
        vidc::frame_params_from_gravity(cam, g, a, p);
        p.fwd_col_major = fabsf(p.Hinv[1] * p.ikh) > 4.0f * fabsf(p.Hinv[0] * p.ikw) ? 1.0f : 0.0f;
        p.inv_col_major = fabsf(p.H[1]) > 4.0f * fabsf(p.H[0]) ? 1.0f : 0.0f;
#pragma unroll
        for (int k = 0; k < 10; ++k) p.reserved[k] = 0.0f;
        p.reserved[10] = vidc::inv_division_proven(p, cam) ? 1.0f : 0.0f;      // as frame_params_kernel: prepared workspaces serve both directions
        sp = p;
    }
    __syncthreads();
    const int tiles_x = (cam.W + 31) / 32, tiles_y = (cam.H + 31) / 32, nt = tiles_x * tiles_y;
    const bool ext = nt <= 320 && t < nt && vidc::tile_certainly_exterior(sp, cam, t % tiles_x, t / tiles_x);
    const unsigned int bal = __ballot_sync(0xffffffffu, ext);
    float* __restrict__ o = reinterpret_cast<float*>(out + i);
    const float* spf = reinterpret_cast<const float*>(&sp);
    if (t < 37) o[t] = spf[t];                                       // everything before reserved[]
    if ((t & 31) == 0) o[37 + (t >> 5)] = __uint_as_float(bal);      // reserved[0..9]: 320 tile bits
    if (t == 1) o[47] = sp.reserved[10];
    if (H_out && t < 9) H_out[9 * i + t] = sp.H[t];
    if (src_boxes) {                                                 // prefetch hints of the sheared forward kernels
        for (int k = t; k < nt; k += blockDim.x) {
            uint32_t e[4];
            vidc::fwd_tile_src_box(sp, cam, k % tiles_x, k / tiles_x, e);
            src_boxes[(size_t)i * nt + k] = make_uint4(e[0], e[1], e[2], e[3]);
        }
    }
}

// dataset.py gravity conditioning on device: raw IMU gravity -> (I_g, I_a)
__global__ void condition_gravity_kernel(const float* __restrict__ raw, int B, int rule, float* __restrict__ Ig, float* __restrict__ Ia) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float r[3] = {raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]};
    float g[3], a[3];
    vidc::condition_gravity(r, rule, g, a);
#pragma unroll
    for (int k = 0; k < 3; ++k) { Ig[3 * i + k] = g[k]; Ia[3 * i + k] = a[k]; }
}

// Sparse-depth rasterisation on device (dataset.py:496-510 Demo, :316-329 Azure).
// tracks: (B, N, cols >= 4) fp64 rows [id, x, y, z, ...] as np.loadtxt yields them; the reference walks them in order, so the LAST
// point that lands on a pixel wins: pass 1 records the largest point index per pixel, pass 2 writes that point's depth.
__global__ void rasterize_index_kernel(const double* __restrict__ tracks, const int* __restrict__ counts, int B, int N, int cols,
                                       double fc0, double fc1, double cc0, double cc1, int H, int W, int* __restrict__ winner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= N || (counts && i >= counts[b])) return;
    const double* t = tracks + ((long long)b * N + i) * cols;
    const double u = t[1] / t[3], v = t[2] / t[3];           // :503-504
    const double px = fc0 * u + cc0, py = fc1 * v + cc1;     // :505-506 (numpy: separate multiply and add)
    if (!(fabs(px) < 2.0e9) || !(fabs(py) < 2.0e9)) return;  // int() of nan / inf raises in Python; such rows are skipped here
    const int col = (int)px, row = (int)py;                  // int(): truncation toward zero  :507-508
    if (row >= 0 && row < H && col >= 0 && col < W) atomicMax(winner + ((long long)b * H + row) * W + col, i);
}
__global__ void rasterize_write_kernel(const double* __restrict__ tracks, int N, int cols, long long hw, const int* __restrict__ winner,
                                       float* __restrict__ depth, long long total) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int i = winner[p];
    depth[p] = i < 0 ? 0.0f : (float)tracks[((p / hw) * N + i) * cols + 3];    // klt_depth_tensor[0,row,col] = klt_tracks[i,3]  :510
}

// explicit homographies (ref :292-310): NON-uniform kw, kh; inverse in fp64
__global__ void frame_params_from_h_kernel(vidc_camera cam, const float* __restrict__ Hm, int B,
                                           vidc_frame_params* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    vidc_frame_params p;
    double h[9];
    for (int k = 0; k < 9; ++k) { p.H[k] = Hm[9 * i + k]; h[k] = (double)p.H[k]; p.R[k] = (k % 4 == 0) ? 1.0f : 0.0f; }
    // fp64 corners / bbox, :293-300
    const double Wm = cam.W - 1, Hmm = cam.H - 1;
    const double cxs[4] = {0, Wm, 0, Wm}, cys[4] = {0, 0, Hmm, Hmm};
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int j = 0; j < 4; ++j) {
        const double c2 = h[6] * cxs[j] + h[7] * cys[j] + h[8];
        const double x = (h[0] * cxs[j] + h[1] * cys[j] + h[2]) / c2, y = (h[3] * cxs[j] + h[4] * cys[j] + h[5]) / c2;
        xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
    }
    const double kw = cam.W / (xmax - xmin), kh = cam.H / (ymax - ymin);
    // adjugate inverse in fp64 (:301 np.linalg.inv)
    const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]);
    const double id = 1.0 / det;
    const double inv[9] = {(h[4] * h[8] - h[5] * h[7]) * id, (h[2] * h[7] - h[1] * h[8]) * id, (h[1] * h[5] - h[2] * h[4]) * id,
                           (h[5] * h[6] - h[3] * h[8]) * id, (h[0] * h[8] - h[2] * h[6]) * id, (h[2] * h[3] - h[0] * h[5]) * id,
                           (h[3] * h[7] - h[4] * h[6]) * id, (h[1] * h[6] - h[0] * h[7]) * id, (h[0] * h[4] - h[1] * h[3]) * id};
    for (int k = 0; k < 9; ++k) p.Hinv[k] = (float)inv[k];
    p.px_min = (float)xmin; p.py_min = (float)ymin;
    p.kw = (float)kw; p.kh = (float)kh; p.ikw = (float)(1.0 / kw); p.ikh = (float)(1.0 / kh);
    p.w_max = (float)(xmax - xmin); p.h_max = (float)(ymax - ymin);
    p.fwd_col_major = fabsf(p.Hinv[1] * p.ikh) > 4.0f * fabsf(p.Hinv[0] * p.ikw) ? 1.0f : 0.0f;
    p.inv_col_major = 0.0f;
    for (int k = 0; k < 11; ++k) p.reserved[k] = 0.0f;
    out[i] = p;
}

__global__ void scatter_homography_kernel(const vidc_frame_params* __restrict__ prm, int B,
                                          float* __restrict__ Hm, float* __restrict__ Rm, float* __restrict__ Hi,
                                          float* __restrict__ Rt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 9) return;
    const int b = i / 9, k = i % 9;
    if (Hm) Hm[i] = prm[b].H[k];
    if (Rm) Rm[i] = prm[b].R[k];
    if (Hi) Hi[i] = prm[b].Hinv[k];
    if (Rt) Rt[i] = prm[b].R[3 * (k % 3) + k / 3];
}

}  // namespace vidc_k
