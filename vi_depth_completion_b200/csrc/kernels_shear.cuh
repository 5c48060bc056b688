// kernels_shear.cuh -- the forward and inverse warps with SHEARED row segments.
//
// The planar forward warp is bound by the L1 data pipe (profiles/r1_history.md): a canvas row of a rolled frame maps to
// a slanted source line, so the 32 lanes of a tap request touch one cache line per source row they cross (8 rows at
// 15 deg, 16 at 30 deg) and every one of the 16 requests of a row segment costs that many L1 wavefronts.
// Here lane l of a row segment does not take pixel (X0 + l, Y) but (X0 + l, Y0 + ((r + sh(l)) mod 32)), with
// sh(l) = round(l * slope) and slope = the canvas dY/dX along which the SOURCE y stays constant at the tile centre:
// the 32 lanes follow a source row, a tap request touches ~2 lines at any roll.  Every pixel of the 32x32 tile is still
// produced exactly once, by exactly the same arithmetic (only the lane -> pixel assignment changes, so the bits do
// not): 8 warps x 4 sheared rows cover each column's 32 rows.  The results are staged in a 32x32 pixel-interleaved shared tile (one
// 128-bit deposit per pixel) and leave as 128-bit row stores per plane; the validity mask and coverage are computed from the staged values.
//
// Compile-time geometry only (W a multiple of 32, contiguous planes, 16-byte aligned outputs); everything else takes
// warp_rgbd_fast_kernel.
#pragma once

namespace vidc_k {

// Resident CTAs per SM.  The sheared kernels keep no running output pointers, so their hot paths fit 32 registers (the
// few spills sit in the cold column-major path) and 8 CTAs = 64 warps are resident per SM, where the straight-row kernels
// run 5 CTAs at 48 registers.  Measured on the B200 the extra warps hide more of the gather latency than the smaller L1
// (8 x 17.5 KB of staging tiles) costs: forward 0.641 -> 0.621 ms, inverse 0.538 -> 0.510 ms on the bench workload,
// 0.695 -> 0.642 / 0.547 -> 0.477 ms on level frames (profiles/r1_history.md, occupancy sweep).
#ifndef VIDC_SHEAR_BLOCKS_FWD
#define VIDC_SHEAR_BLOCKS_FWD 8
#endif
#ifndef VIDC_SHEAR_BLOCKS_INV
#define VIDC_SHEAR_BLOCKS_INV 8
#endif

template <int GW, int GH, bool HAS_D>
__global__ void __launch_bounds__(256, VIDC_SHEAR_BLOCKS_FWD)
warp_rgbd_shear_kernel(const __grid_constant__ FwdArgs a) {
    static_assert(GW > 0 && GW % 32 == 0, "sheared tiles need a compile-time canvas whose width is a multiple of 32");
    static_assert(ROWS_PER_THREAD == 4 && PATCH_W == 32, "32x32 tile, 8 warps x 4 rows");
    constexpr int W = GW, H = GH;
    // pixel-interleaved staging tile: one 16-byte slot (r, g, b, d) per pixel, slot = column ^ (column >> 3).  The
    // swizzle keeps both sides conflict-free: a quarter-warp of the 128-bit deposits covers 8 consecutive columns
    // (any rows), a quarter-warp of the 128-bit read-out covers columns 4i + k, i = 0..7 of one row.
    __shared__ __align__(16) float4 tile[32][32];
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int X = blockIdx.x * TILE_W + lane, tileY0 = blockIdx.y * TILE_H;
    // params: Hinv = floats 18..26, px_min,py_min = 27,28, ikw,ikh = 31,32 -> float4 #4..#8 (floats 16..35)
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    if (pr[19] != 0.0f) {                                          // vidc_frame_params::fwd_col_major (CTA-uniform)
        warp_rgbd_col_major_tile<GW, GH, HAS_D>(a, pr);
        return;
    }
    const float* Hi = pr + 2;
    const float px_min = pr[11], py_min = pr[12], ikw = pr[15], ikh = pr[16];
    const float px = ikw * (float)X + px_min;
    const float u0 = Hi[0] * px, v0 = Hi[3] * px, s0 = Hi[6] * px;
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in_rgb = a.rgb + (long long)b * a.rgb_sn;
    const float* __restrict__ in_dep = HAS_D ? a.dep + (long long)b * a.dep_sn : nullptr;
    // shear of this tile: d(source y)/dX and /dY of the canvas -> source map at the tile centre (not part of any result)
    int sh_l;
    {
        const float pxc = ikw * (float)(blockIdx.x * TILE_W + TILE_W / 2) + px_min;
        const float pyc = ikh * (float)(tileY0 + TILE_H / 2) + py_min;
        const float vc = fmaf(Hi[4], pyc, Hi[3] * pxc) + Hi[5], sc = fmaf(Hi[7], pyc, Hi[6] * pxc) + Hi[8];
        float slope = -__fdividef(ikw * (Hi[3] * sc - vc * Hi[6]), ikh * (Hi[4] * sc - vc * Hi[7]));
        slope = fminf(fmaxf(slope, -4.0f), 4.0f);                  // NaN -> -4: any integer shear is a valid permutation
        sh_l = __float2int_rn(slope * (float)lane);
    }
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int Ys = (warp * ROWS_PER_THREAD + j + sh_l) & 31;
        const float py = ikh * (float)(tileY0 + Ys) + py_min;
        const float u = fmaf(Hi[1], py, u0) + Hi[2];
        const float v = fmaf(Hi[4], py, v0) + Hi[5];
        const float s = fmaf(Hi[7], py, s0) + Hi[8];
        float sx, sy;
        div2_rn(u, v, s, sx, sy);                                  // :146-147
        const float gx = a.cam.inv_half_w * (sx - a.cam.cx);
        const float gy = a.cam.inv_half_h * (sy - a.cam.cy);
        const float ix = unnormalize(gx, Wf), iy = unnormalize(gy, Hf);
        const Pos t = make_pos(ix, iy, H, W);
        const Px4 o = fwd_sample_row<HAS_D>(in_rgb, in_dep, W, W * H, H, W, a.mode_d, ix, iy, t);
        tile[Ys][lane ^ (lane >> 3)] = make_float4(o.r, o.g, o.b, o.d);
    }
    __syncthreads();
    // write-out: thread -> (row, 4 consecutive columns), 128-bit loads from the tile, 128-bit row stores
    const int tid = warp * 32 + lane, row = tid >> 3, c4 = (tid & 7) * 4;
    const int Yo = tileY0 + row, Xo = blockIdx.x * TILE_W + c4;
    unsigned int cnt = 0;
    if (H % 32 == 0 || Yo < H) {
        float4 px4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) px4[k] = tile[row][(c4 + k) ^ ((c4 + k) >> 3)];
        const float4 r = make_float4(px4[0].x, px4[1].x, px4[2].x, px4[3].x);
        const float4 g = make_float4(px4[0].y, px4[1].y, px4[2].y, px4[3].y);
        const float4 bl = make_float4(px4[0].z, px4[1].z, px4[2].z, px4[3].z);
        float* __restrict__ o_rgb = a.rgb_o + ((long long)b * a.rgbo_sn + Yo * W + Xo);
        *reinterpret_cast<float4*>(o_rgb) = r;
        *reinterpret_cast<float4*>(o_rgb + W * H) = g;
        *reinterpret_cast<float4*>(o_rgb + 2 * W * H) = bl;
        if (HAS_D)
            *reinterpret_cast<float4*>(a.dep_o + ((long long)b * a.depo_sn + Yo * W + Xo)) =
                make_float4(px4[0].w, px4[1].w, px4[2].w, px4[3].w);
        if (a.mask || a.coverage) {                                // surface_normal.py:151, four pixels at once
            const unsigned int m0 = (r.x + g.x) + bl.x > 0.01f, m1 = (r.y + g.y) + bl.y > 0.01f;
            const unsigned int m2 = (r.z + g.z) + bl.z > 0.01f, m3 = (r.w + g.w) + bl.w > 0.01f;
            if (a.mask) *reinterpret_cast<unsigned int*>(a.mask + (((long long)b * H + Yo) * W + Xo)) = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
            cnt = m0 + m1 + m2 + m3;
        }
    }
    if (a.coverage) {
        __shared__ unsigned int cta_cov;
        if (tid == 0) cta_cov = 0;
        __syncthreads();
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&cta_cov, cnt);
        __syncthreads();
        if (tid == 0 && cta_cov) atomicAdd(a.coverage + b, cta_cov);
    }
}

// ---- inverse: camera px -> canvas coords, 3 planes, R^T, renormalisation; same sheared rows and staging ----------
// The fourth component of the staging slot carries the optional validity flag.
template <int GW, int GH, bool NORMALIZE>
__global__ void __launch_bounds__(256, VIDC_SHEAR_BLOCKS_INV)
unwarp_normals_shear_kernel(const __grid_constant__ InvArgs a) {
    static_assert(GW > 0 && GW % 32 == 0, "sheared tiles need a compile-time canvas whose width is a multiple of 32");
    static_assert(ROWS_PER_THREAD == 4 && PATCH_W == 32, "32x32 tile, 8 warps x 4 rows");
    constexpr int W = GW, H = GH;
    __shared__ __align__(16) float4 tile[32][32];
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int X = blockIdx.x * TILE_W + lane, tileY0 = blockIdx.y * TILE_H;
    // H = floats 0..8, R = 9..17, px_min,py_min = 27,28, kw,kh = 29,30 -> float4 #0..#7 (floats 0..31)
    float pr[32];
    load_params(a.prm + b, pr, 0, 8);
    if (__ldg(&a.prm[b].inv_col_major) != 0.0f) {                   // CTA-uniform
        unwarp_normals_col_major_tile<GW, GH, NORMALIZE>(a, pr);
        return;
    }
    const float* Hm = pr;
    const float* R = pr + 9;
    const float px_min = pr[27], py_min = pr[28], kw = pr[29], kh = pr[30];
    const float Xf = (float)X;
    const float u0 = Hm[0] * Xf, v0 = Hm[3] * Xf, s0 = Hm[6] * Xf;
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in = a.x + (long long)b * a.x_sn;
    int sh_l;
    {
        const float xc = (float)(blockIdx.x * TILE_W + TILE_W / 2), yc = (float)(tileY0 + TILE_H / 2);
        const float vc = fmaf(Hm[4], yc, Hm[3] * xc) + Hm[5], sc = fmaf(Hm[7], yc, Hm[6] * xc) + Hm[8];
        float slope = -__fdividef(Hm[3] * sc - vc * Hm[6], Hm[4] * sc - vc * Hm[7]);
        slope = fminf(fmaxf(slope, -4.0f), 4.0f);
        sh_l = __float2int_rn(slope * (float)lane);
    }
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int Ys = (warp * ROWS_PER_THREAD + j + sh_l) & 31;
        const float Yf = (float)(tileY0 + Ys);
        const float s = fmaf(Hm[7], Yf, s0) + Hm[8];
        const float u = fmaf(Hm[1], Yf, u0) + Hm[2];
        const float v = fmaf(Hm[4], Yf, v0) + Hm[5];
        float tx, ty;
        div2_rn(u, v, s, tx, ty);                                  // :245
        const float cxp = kw * (tx - px_min);
        const float cyp = kh * (ty - py_min);
        const float gx = a.cam.inv_half_w * (cxp - a.cam.cx);
        const float gy = a.cam.inv_half_h * (cyp - a.cam.cy);
        const Pos t = make_pos(unnormalize(gx, Wf), unnormalize(gy, Hf), H, W);
        const Px3 y = inv_sample_row(in, W, W * H, H, W, t);
        // z = C_R_Cg.bmm(y), C_R_Cg = R^T: k-ascending FMA chain from a +0 accumulator (:253)
        float z0 = fmaf(R[6], y.c, fmaf(R[3], y.b, fmaf(R[0], y.a, 0.0f)));
        float z1 = fmaf(R[7], y.c, fmaf(R[4], y.b, fmaf(R[1], y.a, 0.0f)));
        float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
        if (NORMALIZE) normalize3_rn(z0, z1, z2);                  // surface_normal.py:170
        tile[Ys][lane ^ (lane >> 3)] = make_float4(z0, z1, z2, (a.valid && t.touch) ? 1.0f : 0.0f);
    }
    __syncthreads();
    const int tid = warp * 32 + lane, row = tid >> 3, c4 = (tid & 7) * 4;
    const int Yo = tileY0 + row, Xo = blockIdx.x * TILE_W + c4;
    if (H % 32 == 0 || Yo < H) {
        float4 px4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) px4[k] = tile[row][(c4 + k) ^ ((c4 + k) >> 3)];
        float* __restrict__ o = a.z + ((long long)b * a.z_sn + Yo * W + Xo);
        *reinterpret_cast<float4*>(o) = make_float4(px4[0].x, px4[1].x, px4[2].x, px4[3].x);
        *reinterpret_cast<float4*>(o + W * H) = make_float4(px4[0].y, px4[1].y, px4[2].y, px4[3].y);
        *reinterpret_cast<float4*>(o + 2 * W * H) = make_float4(px4[0].z, px4[1].z, px4[2].z, px4[3].z);
        if (a.valid)
            *reinterpret_cast<unsigned int*>(a.valid + (((long long)b * H + Yo) * W + Xo)) =
                (px4[0].w != 0.0f ? 1u : 0u) | (px4[1].w != 0.0f ? 1u << 8 : 0u) | (px4[2].w != 0.0f ? 1u << 16 : 0u) |
                (px4[3].w != 0.0f ? 1u << 24 : 0u);
    }
}

}  // namespace vidc_k
