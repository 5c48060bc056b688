// kernels_packed.cuh -- packed RGBD (channels-last, C = 4) forward kernel
// Part of libvidc_b200.so (one translation unit, vidc_kernels.cu); compiled with -fmad=false.
#pragma once
#include "kernels_fast.cuh"

namespace vidc_k {

// ------------------------------------------------------------------------------------------
// Packed RGBD forward warp: pixels are interleaved (B, H, W, 4) = channels-last with C = 4, so every bilinear tap is
// ONE 128-bit load carrying all four channels and every output pixel ONE 128-bit store (the gather costs 4 LSU
// requests per pixel instead of 16).  Opt-in layout for callers that can hand RGB + depth over packed; same
// arithmetic per channel as the planar kernels.
struct PackedArgs {
    const vidc_frame_params* prm; CamConst cam;
    const float4* in; long long in_sn; int Hin, Win;      // strides in pixels (float4)
    float4* out; long long out_sn;
    int mode_d; unsigned char* mask; unsigned int* coverage;
};

__device__ __forceinline__ float4 bilerp_px(const float4 nw, const float4 ne, const float4 sw, const float4 se, const Pos& t) {
    float4 o;
    o.x = bilerp(nw.x, ne.x, sw.x, se.x, t);
    o.y = bilerp(nw.y, ne.y, sw.y, se.y, t);
    o.z = bilerp(nw.z, ne.z, sw.z, se.z, t);
    o.w = bilerp(nw.w, ne.w, sw.w, se.w, t);
    return o;
}

__global__ void __launch_bounds__(256, VIDC_MIN_BLOCKS)
warp_rgbd_nhwc4_kernel(const __grid_constant__ PackedArgs a) {
    const int W = a.cam.W, H = a.cam.H, Win = a.Win, Hin = a.Hin;
    const int b = blockIdx.z;
    const int lane = threadIdx.x;
    const PixelMap pm = pixel_map();
    const int X = pm.X, Y0 = pm.Y0;
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    const float* Hi = pr + 2;
    const float px_min = pr[11], py_min = pr[12], ikw = pr[15], ikh = pr[16];
    const float px = ikw * (float)X + px_min;
    const float u0 = Hi[0] * px, v0 = Hi[3] * px, s0 = Hi[6] * px;
    const float Winf = (float)Win, Hinf = (float)Hin;
    const float4* __restrict__ in = a.in + (long long)b * a.in_sn;
    float4* __restrict__ o = a.out + ((long long)b * a.out_sn + (long long)Y0 * W + X);
    unsigned char* __restrict__ o_mask = a.mask ? a.mask + (((long long)b * H + Y0) * W + X) : nullptr;
    const bool xlive = X < W;
    unsigned int cov = 0;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll kUnroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int Y = Y0 + j * PATCH_H;
        const bool live = xlive && Y < H;
        const float py = ikh * (float)Y + py_min;
        const float u = fmaf(Hi[1], py, u0) + Hi[2];
        const float v = fmaf(Hi[4], py, v0) + Hi[5];
        const float s = fmaf(Hi[7], py, s0) + Hi[8];
        float sx, sy;
        div2_rn(u, v, s, sx, sy);
        const float gx = a.cam.inv_half_w * (sx - a.cam.cx);
        const float gy = a.cam.inv_half_h * (sy - a.cam.cy);
        const float ix = unnormalize(gx, Winf);
        const float iy = unnormalize(gy, Hinf);
        Pos t = make_pos(ix, iy, Hin, Win);
        t.touch = t.touch && live;
        float4 r = zero4;
        if (__any_sync(0xffffffffu, t.touch)) {
            if (__all_sync(0xffffffffu, t.interior)) {
                const float4* __restrict__ p = in + (t.y0 * Win + t.x0);
                r = bilerp_px(__ldg(p), __ldg(p + 1), __ldg(p + Win), __ldg(p + Win + 1), t);
            } else {
                const bool in_x0 = (unsigned)t.x0 < (unsigned)Win, in_x1 = (unsigned)(t.x0 + 1) < (unsigned)Win;
                const bool in_y0 = (unsigned)t.y0 < (unsigned)Hin, in_y1 = (unsigned)(t.y0 + 1) < (unsigned)Hin;
                const float4* __restrict__ p = in + (t.y0 * Win + t.x0);
                const float4 nw = (t.touch && in_x0 && in_y0) ? __ldg(p) : zero4;
                const float4 ne = (t.touch && in_x1 && in_y0) ? __ldg(p + 1) : zero4;
                const float4 sw = (t.touch && in_x0 && in_y1) ? __ldg(p + Win) : zero4;
                const float4 se = (t.touch && in_x1 && in_y1) ? __ldg(p + Win + 1) : zero4;
                r = t.touch ? bilerp_px(nw, ne, sw, se, t) : zero4;      // non-finite coordinate: NaN weights, reads as 0
            }
            if (a.mode_d != VIDC_BILINEAR) {   // depth channel by nearest neighbour
                const int xn = (int)rintf(ix), yn = (int)rintf(iy);
                const bool inn = t.touch && (unsigned)xn < (unsigned)Win && (unsigned)yn < (unsigned)Hin;
                r.w = inn ? __ldg(reinterpret_cast<const float*>(in + (yn * Win + xn)) + 3) : 0.0f;
            }
        }
        const bool m = (r.x + r.y) + r.z > 0.01f;
        if (live) {
            *o = r;
            if (a.mask) *o_mask = m ? 1 : 0;
        }
        o += PATCH_H * W;
        if (a.mask) o_mask += PATCH_H * W;
        if (a.coverage) cov += __popc(__ballot_sync(0xffffffffu, m && live));
    }
    if (a.coverage) {
        __shared__ unsigned int cta_count;
        const int tid = threadIdx.y * 32 + lane;
        if (tid == 0) cta_count = 0;
        __syncthreads();
        if (lane == 0 && cov) atomicAdd(&cta_count, cov);
        __syncthreads();
        if (tid == 0 && cta_count) atomicAdd(a.coverage + b, cta_count);
    }
}

}  // namespace vidc_k
