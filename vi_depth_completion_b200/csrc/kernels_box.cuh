// kernels_box.cuh -- the inverse warp with the source footprint STAGED in shared memory.
//
// What the profiles of the sheared inverse kernel said (profiles/r1_shear_final_ncu_summary.txt): 83 % issue-active AND 83 % of
// the L1 data pipe -- 3 wavefronts per 32-lane tap request (a request touches 2-3 cache lines) on top of the staging tile's
// deposits and read-out -- so neither fewer instructions nor fewer wavefronts alone could reach 0.70 of the HBM roofline.
// This kernel removes both at once:
//
//   * The inverse warp of a 32x32 camera tile reads a canvas patch of about 31x31 pixels (bounding box of the mapped tile
//     corners, computed per frame and tile by frame_params_inv_boxes_kernel: vidc::inv_tile_boxes, frame_params.cuh).  The
//     CTA copies that box -- three planes -- into shared memory with coalesced 128-bit loads; everything outside the image
//     is staged as +0, which IS padding_mode='zeros', so border pixels need no predicates.
//   * The bilinear taps are then 12 LDS per pixel with immediate offsets off ONE 32-bit address.  The box rows have a pitch
//     of 65 or 63 floats (+-1 modulo the 32 banks, chosen per tile from the sign of the map's shear): the 32 taps of a
//     canvas row segment advance by |dx| + |dy| < 1 bank per lane and fall into distinct banks -- one wavefront per request
//     at any roll angle, so lanes simply run along canvas rows.
//   * With straight rows every warp stores whole 128-byte row segments per plane straight from registers: no output staging
//     tile, no second barrier, no write-out pass, no shear arithmetic.
//   * Frames whose projective denominator is provably well inside the window of the shared-reciprocal division
//     (vidc::inv_division_proven) skip the per-pixel window test.
//
// Every pixel is produced by exactly the arithmetic of the other kernels (device_common.cuh / kernels_fast.cuh), so the bits
// do not change; a pixel whose taps are not inside the staged box (box overflow at a projective pole, non-finite
// coordinates) takes the predicated global-memory path.  Contiguous planes, W % 4 == 0, 16-byte aligned frames.
#pragma once

namespace vidc_k {

#ifndef VIDC_BOX_BLOCKS
#define VIDC_BOX_BLOCKS 7
#endif
constexpr int BOX_PITCH_MAX = 65, BOX_PLANE = vidc::kBoxMaxH * BOX_PITCH_MAX;       // floats per staged plane

// ---- per-frame kernel: parameters (exactly frame_params_kernel) + the per-tile box table --------------------------------
__global__ void __launch_bounds__(320) frame_params_inv_boxes_kernel(vidc_camera cam, const float* __restrict__ Ig,
                                                                     const float* __restrict__ Ia, int B,
                                                                     vidc_frame_params* __restrict__ out, float* __restrict__ H_out,
                                                                     uint4* __restrict__ boxes, int tiles_x, int tiles_y) {
    __shared__ vidc_frame_params sp;
    __shared__ int s_proven;
    const int i = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        const float g[3] = {Ig[3 * i], Ig[3 * i + 1], Ig[3 * i + 2]};
        const float a[3] = {Ia[3 * i], Ia[3 * i + 1], Ia[3 * i + 2]};
        vidc_frame_params p;
        vidc::frame_params_from_gravity(cam, g, a, p);
        p.fwd_col_major = fabsf(p.Hinv[1] * p.ikh) > 4.0f * fabsf(p.Hinv[0] * p.ikw) ? 1.0f : 0.0f;
        p.inv_col_major = fabsf(p.H[1]) > 4.0f * fabsf(p.H[0]) ? 1.0f : 0.0f;
#pragma unroll
        for (int k = 0; k < 11; ++k) p.reserved[k] = 0.0f;
        sp = p;
        s_proven = vidc::inv_division_proven(p, cam) ? 1 : 0;
    }
    __syncthreads();
    float* __restrict__ o = reinterpret_cast<float*>(out + i);
    const float* spf = reinterpret_cast<const float*>(&sp);
    if (t < 48) o[t] = spf[t];
    if (H_out && t < 9) H_out[9 * i + t] = sp.H[t];
    const int nt = tiles_x * tiles_y;
    for (int k = t; k < nt; k += blockDim.x) {
        uint32_t e[4];
        vidc::inv_tile_boxes(sp, cam, k % tiles_x, k / tiles_x, s_proven != 0, e);
        boxes[(size_t)i * nt + k] = make_uint4(e[0], e[1], e[2], e[3]);
    }
}

// ---- staging: box -> shared memory ----------------------------------------------------------------------------------------
// 128-bit global loads, four 32-bit deposits each (the odd row pitch rules out 128-bit shared stores).  Lane -> (row, float4
// column): 4 rows x 8 columns for boxes up to 32 floats wide -- the deposits of a warp go to banks (+-r + 4 c4 + k) mod 32, all
// distinct -- and 2 rows x 16 columns for wider boxes (two-way conflicts, but no idle lanes).  Everything a lane needs is
// fixed before the loop; one pass of the 8 warps covers 32 (16) rows.  (Measured alternative, profiles/r2_history.md: 4-byte
// cp.async copies with all rows in flight at once shorten the staging phase but cost more LSU instructions: 0.59 vs 0.53 ms.)
template <int HWC>
__device__ __forceinline__ void stage_box(const float* __restrict__ in, int W, int H, int HW, float* __restrict__ sbox,
                                          int bx0, int by0, int bw, int bh, int pitch, int lane, int warp) {
    const int hw = HWC ? HWC : HW;
    const bool wide = bw > 32;
    const int c = wide ? (lane & 15) << 2 : (lane & 7) << 2;
    const int rstep = wide ? 16 : 32;
    int r = wide ? (warp << 1) + (lane >> 4) : (warp << 2) + (lane >> 3);
    if (c >= bw) return;
    const int gx = bx0 + c;
    const bool col_img = (unsigned)gx < (unsigned)W;              // W % 4 == 0: the float4 is all in or all out
    const float* __restrict__ g = in + ((by0 + r) * W + gx);
    float* __restrict__ d = sbox + (r * pitch + c);
    for (; r < bh; r += rstep, g += rstep * W, d += rstep * pitch) {
        float4 v0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), v1 = v0, v2 = v0;
        if (col_img && (unsigned)(by0 + r) < (unsigned)H) {
            v0 = __ldg(reinterpret_cast<const float4*>(g));
            v1 = __ldg(reinterpret_cast<const float4*>(g + hw));
            v2 = __ldg(reinterpret_cast<const float4*>(g + 2 * hw));
        }
        d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w;
        d[BOX_PLANE] = v1.x; d[BOX_PLANE + 1] = v1.y; d[BOX_PLANE + 2] = v1.z; d[BOX_PLANE + 3] = v1.w;
        d[2 * BOX_PLANE] = v2.x; d[2 * BOX_PLANE + 1] = v2.y; d[2 * BOX_PLANE + 2] = v2.z; d[2 * BOX_PLANE + 3] = v2.w;
    }
}

// u/s and v/s without the window test (the frame passed vidc::inv_division_proven) or with it
__device__ __forceinline__ void div2_sel(float u, float v, float s, bool proven, float& qu, float& qv) {
    const float r = rcp_refined(s);
    qu = div_with_rcp(u, s, r);
    qv = div_with_rcp(v, s, r);
    if (!proven) {                                                 // CTA-uniform
        const float as = fabsf(s);
        const float hi = fmaxf(fmaxf(fabsf(u), fabsf(v)), as * 0x1p40f);
        const float lo = fminf(fminf(fabsf(u), fabsf(v)), as * 0x1p-40f);
        if (!(lo >= 0x1p-80f && hi <= 0x1p80f)) {
            qu = ieee_div_slow(u, s);
            qv = ieee_div_slow(v, s);
        }
    }
}

#ifdef VIDC_BOX_TIMING
__device__ unsigned long long g_box_timing[8];      // development instrumentation: summed clock64 deltas of thread 0 of every CTA
#define VIDC_BT(i, t_prev) do { if (threadIdx.x == 0 && threadIdx.y == 0) { const long long now_ = clock64(); atomicAdd(&g_box_timing[i], (unsigned long long)(now_ - t_prev)); t_prev = now_; } } while (0)
#else
#define VIDC_BT(i, t_prev) do { } while (0)
#endif
struct BoxGeom { int x0, y0, wm1, hm1; };
__device__ __forceinline__ BoxGeom box_geom(unsigned int xy, unsigned int wh) {
    BoxGeom g;
    g.x0 = (int)(short)(xy & 0xffffu); g.y0 = (int)xy >> 16;
    g.wm1 = (int)(wh & 0xffu) - 1; g.hm1 = (int)((wh >> 8) & 0xffu) - 1;
    return g;
}

// L2 prefetch of one staged box: thread t takes (plane, row, 128-byte segment) number t -- 3 planes x <= 40 rows x 2 segments.
__device__ __forceinline__ void prefetch_box_l2(const float* __restrict__ in, int W, int H, int HW, const BoxGeom& g, int tid) {
    const int bh = g.hm1 + 1, bw = g.wm1 + 1;
    const int seg = tid & 1, rp = tid >> 1;                        // rp = plane * bh + row
    if (rp >= 3 * bh) return;
    const int plane = rp >= 2 * bh ? 2 : (rp >= bh ? 1 : 0);
    const int row = g.y0 + rp - plane * bh;
    int col = g.x0 + seg * 32;
    col = col < 0 ? 0 : col;
    if ((unsigned)row < (unsigned)H && col < W && seg * 32 < bw)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(in + ((long long)plane * HW + row * W + col)));
}

// One CTA = two vertically adjacent 32x32 camera tiles, one after the other through the same staging buffer: the per-CTA
// prologue (parameters into uniform registers, per-lane column terms) is paid once per 2048 pixels.
template <int GW, int GH, bool NORMALIZE, bool HAS_VALID>
__global__ void __launch_bounds__(256, VIDC_BOX_BLOCKS)
unwarp_normals_box_kernel(const __grid_constant__ InvArgs a, const uint4* __restrict__ boxes, int tiles_y, int3 pf) {
    static_assert(GW % 4 == 0, "rows are staged as float4");
    static_assert(ROWS_PER_THREAD == 4 && TILE_W == 32 && TILE_H == 32, "32x32 tile, 8 warps x 4 rows");
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H, HW = W * H;
    __shared__ __align__(16) float sbox[3 * BOX_PLANE];
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W;
    // H = floats 0..8, R = 9..17, px_min,py_min = 27,28, kw,kh = 29,30 -> float4 #0..#7 (floats 0..31)
    float pr[32];
    load_params(a.prm + b, pr, 0, 8);
    const float* Hm = pr;
    const float* R = pr + 9;
    const float px_min = pr[27], py_min = pr[28], kw = pr[29], kh = pr[30];
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in = a.x + (long long)b * a.x_sn;
    // The staging loads of a tile are a DRAM round trip that nothing in this CTA can overlap.  Each CTA therefore asks L2 for
    // the boxes of the CTA that will run about one wave later (pf = that distance in grid coordinates, computed by the host):
    // by the time that CTA stages, its loads are L2 hits.  Purely a hint -- no result depends on it.
#ifndef VIDC_BOX_NO_PREFETCH
    {
        int px = (int)blockIdx.x + pf.x, py = (int)blockIdx.y + pf.y, pz = (int)blockIdx.z + pf.z;
        if (px >= (int)gridDim.x) { px -= gridDim.x; ++py; }
        if (py >= (int)gridDim.y) { py -= gridDim.y; ++pz; }
        if (pz < (int)gridDim.z) {
            const float* __restrict__ pin = a.x + (long long)pz * a.x_sn;
            const int tid = warp * 32 + lane;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int ty = 2 * py + half;
                if (ty >= tiles_y) break;
                const uint4 e = __ldg(boxes + (((size_t)pz * tiles_y + ty) * gridDim.x + px));
                const int nsub = (int)((e.y >> 16) & 3u);
                if (nsub >= 1) prefetch_box_l2(pin, W, H, HW, box_geom(e.x, e.y), tid);
                if (nsub == 2) prefetch_box_l2(pin, W, H, HW, box_geom(e.z, e.w), tid);
            }
        }
    }
#endif
    long long tprev = 0;
#ifdef VIDC_BOX_TIMING
    tprev = clock64();
    if (threadIdx.x == 0 && threadIdx.y == 0) atomicAdd(&g_box_timing[7], 1ull);
#endif
    (void)tprev;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int ty = 2 * blockIdx.y + half;
        if (ty >= tiles_y) break;                                  // CTA-uniform (odd number of tile rows)
        const int tileY0 = ty * TILE_H;
        const uint4 e = __ldg(boxes + (((size_t)b * tiles_y + ty) * gridDim.x + blockIdx.x));
        const int nsub = (int)((e.y >> 16) & 3u);
        const int pitch = ((e.y >> 20) & 1u) ? 65 : 63;
        const bool proven = ((e.y >> 21) & 1u) != 0u;
        BoxGeom bg = box_geom(e.x, e.y);
        VIDC_BT(0 + 3 * half, tprev);                              // prologue / previous tile's tail + entry load
        if (half) __syncthreads();                                 // the previous tile's taps are done
        if (nsub) stage_box<GW * GH>(in, W, H, HW, sbox, bg.x0, bg.y0, bg.wm1 + 1, bg.hm1 + 1, pitch, lane, warp);
        __syncthreads();
        VIDC_BT(1 + 3 * half, tprev);                              // staging incl. barrier
        // per-lane column terms only now: nothing but the box geometry is live across the staging loop, whose three 128-bit
        // loads per row group must all be in flight together
        const int X = tileX0 + lane;
        const float Xf = (float)X;
        const float u0 = Hm[0] * Xf, v0 = Hm[3] * Xf, s0 = Hm[6] * Xf;
        const bool xlive = GW ? true : (X < W);
        float* __restrict__ o = a.z + ((long long)b * a.z_sn + (tileY0 + warp) * W + X);
        unsigned char* __restrict__ o_valid = HAS_VALID ? a.valid + (((long long)b * H + tileY0 + warp) * W + X) : nullptr;
#pragma unroll
        for (int j = 0; j < ROWS_PER_THREAD; ++j) {
            if (j == 2 && nsub == 2) {                             // CTA-uniform: the lower half's box replaces the upper one
                __syncthreads();
                bg = box_geom(e.z, e.w);
                stage_box<GW * GH>(in, W, H, HW, sbox, bg.x0, bg.y0, bg.wm1 + 1, bg.hm1 + 1, pitch, lane, warp);
                __syncthreads();
            }
            const int Y = tileY0 + warp + 8 * j;
            const float Yf = (float)Y;
            const float s = fmaf(Hm[7], Yf, s0) + Hm[8];
            const float u = fmaf(Hm[1], Yf, u0) + Hm[2];
            const float v = fmaf(Hm[4], Yf, v0) + Hm[5];
            float tx, ty2;
            div2_sel(u, v, s, proven, tx, ty2);                    // :245
            const float cxp = kw * (tx - px_min);
            const float cyp = kh * (ty2 - py_min);
            const float gx = a.cam.inv_half_w * (cxp - a.cam.cx);
            const float gy = a.cam.inv_half_h * (cyp - a.cam.cy);
            const Pos t = make_pos(unnormalize(gx, Wf), unnormalize(gy, Hf), H, W);
            const int ax = t.x0 - bg.x0, ay = t.y0 - bg.y0;
            const bool inbox = t.fin && (unsigned)ax < (unsigned)bg.wm1 && (unsigned)ay < (unsigned)bg.hm1;
            Px3 y = {0.0f, 0.0f, 0.0f};
            bool touch = false;
            if (__all_sync(0xffffffffu, inbox)) {
                const float* __restrict__ p0 = sbox + (ay * pitch + ax);
                const float* __restrict__ p1 = p0 + pitch;
                y.a = bilerp(p0[0], p0[1], p1[0], p1[1], t);
                y.b = bilerp(p0[BOX_PLANE], p0[BOX_PLANE + 1], p1[BOX_PLANE], p1[BOX_PLANE + 1], t);
                y.c = bilerp(p0[2 * BOX_PLANE], p0[2 * BOX_PLANE + 1], p1[2 * BOX_PLANE], p1[2 * BOX_PLANE + 1], t);
                if (HAS_VALID) touch = (unsigned)(t.x0 + 1) <= (unsigned)W && (unsigned)(t.y0 + 1) <= (unsigned)H;
            } else {                                               // box overflow / pole / non-finite: straight from global memory
                Pos tb = t;
                tb.touch = t.fin && (unsigned)(t.x0 + 1) <= (unsigned)W && (unsigned)(t.y0 + 1) <= (unsigned)H;
                touch = tb.touch;
                if (__any_sync(0xffffffffu, tb.touch)) {
                    y.a = sample_border(in, W, H, W, tb);
                    y.b = sample_border(in + HW, W, H, W, tb);
                    y.c = sample_border(in + 2 * HW, W, H, W, tb);
                }
            }
            // z = C_R_Cg.bmm(y), C_R_Cg = R^T: k-ascending FMA chain from a +0 accumulator (:253)
            float z0 = fmaf(R[6], y.c, fmaf(R[3], y.b, fmaf(R[0], y.a, 0.0f)));
            float z1 = fmaf(R[7], y.c, fmaf(R[4], y.b, fmaf(R[1], y.a, 0.0f)));
            float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
            if (NORMALIZE) normalize3_rn(z0, z1, z2);              // surface_normal.py:170
            if (xlive && ((GW && GH % 32 == 0) || Y < H)) {
                o[8 * j * W] = z0; o[8 * j * W + HW] = z1; o[8 * j * W + 2 * HW] = z2;
                if (HAS_VALID) o_valid[8 * j * W] = touch ? 1 : 0;
            }
        }
        VIDC_BT(2 + 3 * half, tprev);                              // compute + stores (thread 0's own)
    }
}

}  // namespace vidc_k
