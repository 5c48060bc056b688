// kernels_tma.cuh -- opt-in TMA-staged variants (cp.async.bulk.tensor + mbarrier ring)
// Part of libvidc_b200.so (one translation unit, vidc_kernels.cu); compiled with -fmad=false.
#pragma once
#include "kernels_fast.cuh"
#include "tma_stage.cuh"

namespace vidc_k {

// ------------------------------------------------------------------------------------------
// TMA-staged forward warp.  CTA = 32 x (8 * TMA_ROWS) canvas pixels.  Warp 0 derives the bounding box of
// the tile's source footprint from its four corner pixels (a homography maps the tile to a convex
// quadrilateral, so the corners bound it; +-1 px of slack covers rounding and the +1 bilinear tap), one
// thread issues the bulk tensor copies of that box for all planes, and every pixel then takes its taps
// from shared memory.  Anything that does not fit (box larger than 64 x 48, non-finite corners, a pixel
// whose taps leave the staged box) falls back to the global-memory row path, so the result never depends
// on the box estimate.  Same arithmetic as every other kernel.
#ifndef VIDC_TMA_ROWS
#define VIDC_TMA_ROWS 2
#endif
constexpr int TMA_ROWS = VIDC_TMA_ROWS, TMA_TILE_H = 8 * TMA_ROWS;
enum { TILE_FALLBACK = 0, TILE_EXTERIOR = 1, TILE_STAGED = 2 };

struct RowPos { int x0, y0; float w_nw, w_ne, w_sw, w_se, ix, iy; };

template <bool HAS_D>
__global__ void __launch_bounds__(256, 4)
warp_rgbd_tma_kernel(const __grid_constant__ FwdArgs a, const __grid_constant__ TmaMaps maps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int s_info[4];                                    // mode, x_lo, y_lo, box height
    float* __restrict__ stage = reinterpret_cast<float*>(smem_raw);

    const int W = a.cam.W, H = a.cam.H, Win = a.Win, Hin = a.Hin;
    const int in_sh = a.in_sh, rgb_sc = a.rgb_sc;
    const int b = blockIdx.z;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int X = blockIdx.x * 32 + lane;
    const int Yt = blockIdx.y * TMA_TILE_H;
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    const float* Hi = pr + 2;
    const float px_min = pr[11], py_min = pr[12], ikw = pr[15], ikh = pr[16];
    const float Winf = (float)Win, Hinf = (float)Hin;

    // ---- warp 0: footprint box of the tile, bulk tensor copy issued as early as possible ----------
    if (warp == 0) {
        const int cxp = min(blockIdx.x * 32 + ((lane & 1) ? 31 : 0), W - 1);
        const int cyp = min(Yt + ((lane & 2) ? TMA_TILE_H - 1 : 0), H - 1);
        float ix, iy;
        forward_coords(Hi, px_min, py_min, ikw, ikh, a.cam, (float)cxp, (float)cyp, Winf, Hinf, ix, iy);
        bool fin = fabsf(ix) < 1.0e8f && fabsf(iy) < 1.0e8f;
        float xmn = ix, xmx = ix, ymn = iy, ymx = iy;
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            xmn = fminf(xmn, __shfl_xor_sync(0xffffffffu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xffffffffu, xmx, o));
            ymn = fminf(ymn, __shfl_xor_sync(0xffffffffu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, o));
        }
        fin = __all_sync(0xffffffffu, fin);
        if (lane == 0) {
            int mode = TILE_FALLBACK, x_lo = 0, y_lo = 0, bh = 0;
            if (fin) {
                x_lo = ((int)floorf(xmn) - 1) & ~3;   // TMA: innermost coordinate * 4 B must be 16-byte aligned
                y_lo = (int)floorf(ymn) - 1;
                const int x_hi = (int)floorf(xmx) + 2, y_hi = (int)floorf(ymx) + 2;
                const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;
                if (x_hi < 0 || x_lo >= Win || y_hi < 0 || y_lo >= Hin) {
                    mode = TILE_EXTERIOR;
                } else if (need_w <= TMA_BW && need_h <= TMA_BH_MAX) {
                    const int cls = need_h <= 24 ? 0 : need_h <= 32 ? 1 : need_h <= 40 ? 2 : 3;
                    bh = tma_box_h(cls);
                    mode = TILE_STAGED;
                    mbar_init(&bar, 1);
                    const uint32_t plane_bytes = (uint32_t)(TMA_BW * bh * 4);
                    mbar_expect_tx(&bar, plane_bytes * (HAS_D ? 4u : 3u));
                    tma_load_4d(stage, &maps.a[cls], &bar, x_lo, y_lo, 0, b);
                    if (HAS_D) tma_load_4d(stage + 3 * TMA_BW * bh, &maps.d[cls], &bar, x_lo, y_lo, 0, b);
                }
            }
            s_info[0] = mode; s_info[1] = x_lo; s_info[2] = y_lo; s_info[3] = bh;
        }
    }

    // ---- phase A (overlaps the copy): sampling positions of this thread's rows ---------------------
    const float px = ikw * (float)X + px_min;
    const float u0 = Hi[0] * px, v0 = Hi[3] * px, s0 = Hi[6] * px;
    const int Y0 = Yt + warp * TMA_ROWS;
    const bool xlive = X < W;
    RowPos rp[TMA_ROWS];
#pragma unroll
    for (int j = 0; j < TMA_ROWS; ++j) {
        const float py = ikh * (float)(Y0 + j) + py_min;
        const float u = fmaf(Hi[1], py, u0) + Hi[2];
        const float v = fmaf(Hi[4], py, v0) + Hi[5];
        const float s = fmaf(Hi[7], py, s0) + Hi[8];
        float sx, sy;
        div2_rn(u, v, s, sx, sy);
        const float gx = a.cam.inv_half_w * (sx - a.cam.cx);
        const float gy = a.cam.inv_half_h * (sy - a.cam.cy);
        const float ix = unnormalize(gx, Winf), iy = unnormalize(gy, Hinf);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
        const bool fin = fabsf(ix) <= 2147483648.0f && fabsf(iy) <= 2147483648.0f;
        rp[j].x0 = fin ? __float2int_rd(ix) : -0x40000000;       // non-finite: far outside every box and every image
        rp[j].y0 = fin ? __float2int_rd(iy) : -0x40000000;
        rp[j].w_nw = wx0 * wy0; rp[j].w_ne = wx1 * wy0; rp[j].w_sw = wx0 * wy1; rp[j].w_se = wx1 * wy1;
        rp[j].ix = ix; rp[j].iy = iy;
    }
    __syncthreads();
    const int mode = s_info[0], x_lo = s_info[1], y_lo = s_info[2], bh = s_info[3];
    const int plane = TMA_BW * bh;
    const float* __restrict__ in_rgb = a.rgb + (long long)b * a.rgb_sn;
    const float* __restrict__ in_dep = HAS_D ? a.dep + (long long)b * a.dep_sn : nullptr;
    float* __restrict__ o_rgb = a.rgb_o + ((long long)b * a.rgbo_sn + Y0 * a.rgbo_sh + X);
    float* __restrict__ o_dep = HAS_D ? a.dep_o + ((long long)b * a.depo_sn + Y0 * a.depo_sh + X) : nullptr;
    unsigned char* __restrict__ o_mask = a.mask ? a.mask + (((long long)b * H + Y0) * W + X) : nullptr;
    unsigned int cov = 0;
    if (mode == TILE_STAGED) mbar_wait(&bar, 0);

    // ---- phase B: taps from shared memory ------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < TMA_ROWS; ++j) {
        const bool live = xlive && (Y0 + j) < H;
        Pos t;
        t.x0 = rp[j].x0; t.y0 = rp[j].y0;
        t.w_nw = rp[j].w_nw; t.w_ne = rp[j].w_ne; t.w_sw = rp[j].w_sw; t.w_se = rp[j].w_se;
        const int rx = t.x0 - x_lo, ry = t.y0 - y_lo;
        const bool inbox = !live || ((unsigned)rx <= (unsigned)(TMA_BW - 2) && (unsigned)ry <= (unsigned)(bh - 2));
        Px4 o = {0.0f, 0.0f, 0.0f, 0.0f};
        if (mode == TILE_STAGED && __all_sync(0xffffffffu, inbox)) {
            const float* __restrict__ p = stage + (live ? ry * TMA_BW + rx : 0);
            o.r = bilerp(p[0], p[1], p[TMA_BW], p[TMA_BW + 1], t);
            o.g = bilerp(p[plane], p[plane + 1], p[plane + TMA_BW], p[plane + TMA_BW + 1], t);
            o.b = bilerp(p[2 * plane], p[2 * plane + 1], p[2 * plane + TMA_BW], p[2 * plane + TMA_BW + 1], t);
            if (HAS_D) {
                if (a.mode_d == VIDC_BILINEAR) {
                    o.d = bilerp(p[3 * plane], p[3 * plane + 1], p[3 * plane + TMA_BW], p[3 * plane + TMA_BW + 1], t);
                } else {
                    const int xn = (int)rintf(rp[j].ix) - x_lo, yn = (int)rintf(rp[j].iy) - y_lo;
                    o.d = live ? stage[3 * plane + yn * TMA_BW + xn] : 0.0f;
                }
            }
        } else {
            // general path: classification against the image, taps from global memory
            t.interior = (unsigned)t.x0 < (unsigned)(Win - 1) && (unsigned)t.y0 < (unsigned)(Hin - 1);
            t.touch = live && (unsigned)(t.x0 + 1) <= (unsigned)Win && (unsigned)(t.y0 + 1) <= (unsigned)Hin;
            if (!(mode == TILE_EXTERIOR && __all_sync(0xffffffffu, !t.touch)))
                o = fwd_sample_row<HAS_D>(in_rgb, in_dep, in_sh, rgb_sc, Hin, Win, a.mode_d, rp[j].ix, rp[j].iy, t);
        }
        const bool m = (o.r + o.g) + o.b > 0.01f;
        if (live) {
            o_rgb[0] = o.r; o_rgb[a.rgbo_sc] = o.g; o_rgb[2 * a.rgbo_sc] = o.b;
            if (HAS_D) *o_dep = o.d;
            if (a.mask) *o_mask = m ? 1 : 0;
        }
        o_rgb += a.rgbo_sh;
        if (HAS_D) o_dep += a.depo_sh;
        if (a.mask) o_mask += W;
        if (a.coverage) cov += __popc(__ballot_sync(0xffffffffu, m && live));
    }
    if (a.coverage) {
        __shared__ unsigned int cta_count;
        const int tid = warp * 32 + lane;
        if (tid == 0) cta_count = 0;
        __syncthreads();
        if (lane == 0 && cov) atomicAdd(&cta_count, cov);
        __syncthreads();
        if (tid == 0 && cta_count) atomicAdd(a.coverage + b, cta_count);
    }
}

// ------------------------------------------------------------------------------------------
// Persistent, warp-specialised, TMA-pipelined inverse warp.  One producer warp per CTA walks the CTA's tiles one
// stage ahead: it bounds the tile's canvas footprint from its four corner pixels and issues ONE bulk tensor copy
// (3 planes) into the next ring slot; eight consumer warps take their taps from shared memory (immediate offsets,
// no bounds tests, zero fill = zeros padding), rotate, renormalise and store.  full[]/empty[] mbarriers form the
// ring.  Tiles whose box does not fit, and rows whose taps leave the box, use the global-memory row path.
constexpr int INV_BW = 48, INV_STAGES = 2, INV_NH = 3;
__host__ __device__ constexpr int inv_box_h(int cls) { return cls == 0 ? 32 : cls == 1 ? 40 : 44; }
constexpr int INV_BH_MAX = 44;
constexpr int INV_STAGE_FLOATS = INV_BW * INV_BH_MAX * 3;
struct InvTmaMaps { CUtensorMap m[INV_NH]; };

template <bool NORMALIZE>
__global__ void __launch_bounds__(288, 4)
unwarp_normals_tma_kernel(const __grid_constant__ InvArgs a, const __grid_constant__ InvTmaMaps maps, int tiles_x, int tiles_y, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[INV_STAGES], empty_bar[INV_STAGES];
    __shared__ int s_info[INV_STAGES][4];                        // mode, x_lo, y_lo, box height
    float* __restrict__ ring = reinterpret_cast<float*>(smem_raw);
    const int W = a.cam.W, H = a.cam.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta, t_end = min(n_tiles, t_begin + per_cta);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < INV_STAGES; ++s) { mbar_init_only(&full_bar[s], 1); mbar_init_only(&empty_bar[s], 8); }
        fence_barrier_init();
    }
    __syncthreads();
    const float Wf = (float)W, Hf = (float)H;
    const int tiles_per_frame = tiles_x * tiles_y;

    if (warp == 8) {
        // ================= producer warp =================
        int cur_b = -1;
        float Hm[9], px_min = 0.f, py_min = 0.f, kw = 0.f, kh = 0.f;
        for (int t = t_begin, i = 0; t < t_end; ++t, ++i) {
            const int stage = i % INV_STAGES;
            const uint32_t parity = (uint32_t)((i / INV_STAGES) & 1);
            const int b = t / tiles_per_frame, r = t - b * tiles_per_frame;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            if (b != cur_b) {
                cur_b = b;
                const vidc_frame_params* __restrict__ P = a.prm + b;
#pragma unroll
                for (int k = 0; k < 9; ++k) Hm[k] = __ldg(&P->H[k]);
                px_min = __ldg(&P->px_min); py_min = __ldg(&P->py_min); kw = __ldg(&P->kw); kh = __ldg(&P->kh);
            }
            const int cxp = min(tx * 32 + ((lane & 1) ? 31 : 0), W - 1);
            const int cyp = min(ty * 32 + ((lane & 2) ? 31 : 0), H - 1);
            float ix, iy;
            inverse_coords(Hm, px_min, py_min, kw, kh, a.cam, (float)cxp, (float)cyp, Wf, Hf, ix, iy);
            bool fin = fabsf(ix) < 1.0e8f && fabsf(iy) < 1.0e8f;
            float xmn = ix, xmx = ix, ymn = iy, ymx = iy;
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                xmn = fminf(xmn, __shfl_xor_sync(0xffffffffu, xmn, o)); xmx = fmaxf(xmx, __shfl_xor_sync(0xffffffffu, xmx, o));
                ymn = fminf(ymn, __shfl_xor_sync(0xffffffffu, ymn, o)); ymx = fmaxf(ymx, __shfl_xor_sync(0xffffffffu, ymx, o));
            }
            fin = __all_sync(0xffffffffu, fin);
            if (lane == 0) {
                int mode = TILE_FALLBACK, x_lo = 0, y_lo = 0, bh = 0, cls = 0;
                if (fin) {
                    x_lo = ((int)floorf(xmn) - 1) & ~3;
                    y_lo = (int)floorf(ymn) - 1;
                    const int x_hi = (int)floorf(xmx) + 2, y_hi = (int)floorf(ymx) + 2;
                    const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;
                    if (x_hi < 0 || x_lo >= W || y_hi < 0 || y_lo >= H) mode = TILE_EXTERIOR;
                    else if (need_w <= INV_BW && need_h <= INV_BH_MAX) {
                        cls = need_h <= 32 ? 0 : need_h <= 40 ? 1 : 2;
                        bh = inv_box_h(cls);
                        mode = TILE_STAGED;
                    }
                }
                mbar_wait(&empty_bar[stage], parity ^ 1u);       // slot released by all 8 consumer warps
                s_info[stage][0] = mode; s_info[stage][1] = x_lo; s_info[stage][2] = y_lo; s_info[stage][3] = bh;
                if (mode == TILE_STAGED) {
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(INV_BW * bh * 3 * 4));
                    tma_load_4d(ring + stage * INV_STAGE_FLOATS, &maps.m[cls], &full_bar[stage], x_lo, y_lo, 0, b);
                } else {
                    mbar_arrive(&full_bar[stage]);
                }
            }
            __syncwarp();
        }
        return;
    }

    // ================= consumer warps =================
    const int x_sh = a.x_sh, x_sc = a.x_sc, z_sh = a.z_sh, z_sc = a.z_sc;
    int cur_b = -1;
    float pr[32];
    for (int t = t_begin, i = 0; t < t_end; ++t, ++i) {
        const int stage = i % INV_STAGES;
        const uint32_t parity = (uint32_t)((i / INV_STAGES) & 1);
        const int b = t / tiles_per_frame, r = t - b * tiles_per_frame;
        const int ty = r / tiles_x, tx = r - ty * tiles_x;
        if (b != cur_b) { cur_b = b; load_params(a.prm + b, pr, 0, 8); }
        const float* Hm = pr;
        const float* R = pr + 9;
        const float px_min = pr[27], py_min = pr[28], kw = pr[29], kh = pr[30];
        const int X = tx * 32 + lane, Y0 = ty * 32 + warp * 4;
        const float Xf = (float)X;
        const float u0 = Hm[0] * Xf, v0 = Hm[3] * Xf, s0 = Hm[6] * Xf;
        const float* __restrict__ in = a.x + (long long)b * a.x_sn;
        float* __restrict__ o = a.z + ((long long)b * a.z_sn + (long long)Y0 * z_sh + X);
        const bool xlive = X < W;

        mbar_wait(&full_bar[stage], parity);
        const int mode = s_info[stage][0], x_lo = s_info[stage][1], y_lo = s_info[stage][2], bh = s_info[stage][3];
        const float* __restrict__ stg = ring + stage * INV_STAGE_FLOATS;
        const int plane = INV_BW * bh;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int Y = Y0 + j;
            const bool live = xlive && Y < H;
            const float Yf = (float)Y;
            const float u = fmaf(Hm[1], Yf, u0) + Hm[2];
            const float v = fmaf(Hm[4], Yf, v0) + Hm[5];
            const float s = fmaf(Hm[7], Yf, s0) + Hm[8];
            float tx_, ty_;
            div2_rn(u, v, s, tx_, ty_);
            const float cxp = kw * (tx_ - px_min);
            const float cyp = kh * (ty_ - py_min);
            const float gx = a.cam.inv_half_w * (cxp - a.cam.cx);
            const float gy = a.cam.inv_half_h * (cyp - a.cam.cy);
            Pos tp = make_pos(unnormalize(gx, Wf), unnormalize(gy, Hf), H, W);   // non-finite -> !touch, !interior
            tp.touch = tp.touch && live;
            const int rx = tp.x0 - x_lo, ry = tp.y0 - y_lo;
            // every tap of a lane inside the staged box <=> the lane may read shared memory blindly; lanes that touch
            // nothing at all (fully outside the image, non-finite, dead) read slot 0 and get weight-free zeros below
            const bool inbox = (unsigned)rx <= (unsigned)(INV_BW - 2) && (unsigned)ry <= (unsigned)(bh - 2);
            Px3 y = {0.0f, 0.0f, 0.0f};
            if (mode == TILE_STAGED && __all_sync(0xffffffffu, inbox || !tp.touch)) {
                if (tp.touch) {
                    const float* __restrict__ p = stg + (ry * INV_BW + rx);
                    y.a = bilerp(p[0], p[1], p[INV_BW], p[INV_BW + 1], tp);
                    y.b = bilerp(p[plane], p[plane + 1], p[plane + INV_BW], p[plane + INV_BW + 1], tp);
                    y.c = bilerp(p[2 * plane], p[2 * plane + 1], p[2 * plane + INV_BW], p[2 * plane + INV_BW + 1], tp);
                }
            } else {
                y = inv_sample_row(in, x_sh, x_sc, H, W, tp);
            }
            float z0 = fmaf(R[6], y.c, fmaf(R[3], y.b, fmaf(R[0], y.a, 0.0f)));
            float z1 = fmaf(R[7], y.c, fmaf(R[4], y.b, fmaf(R[1], y.a, 0.0f)));
            float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
            if (NORMALIZE) {
                const float n = clamp_min_eps(sqrtf((z0 * z0 + z1 * z1) + z2 * z2));
                div3_rn(z0, z1, z2, n);
            }
            if (live) {
                o[0] = z0; o[z_sc] = z1; o[2 * z_sc] = z2;
                if (a.valid) a.valid[((long long)b * H + Y) * W + X] = tp.touch ? 1 : 0;
            }
            o += z_sh;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }
}

}  // namespace vidc_k
