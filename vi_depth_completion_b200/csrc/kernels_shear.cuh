// kernels_shear.cuh -- the forward and inverse warps with SHEARED segments: lanes follow SOURCE rows.
//
// The planar warps are bound by the L1 data pipe and by gather latency (profiles/r1_history.md): a canvas row of a
// rolled frame maps to a slanted source line, so the 32 lanes of a tap request touch one cache line per source row they
// cross (8 rows at 15 deg, 16 at 30 deg) and every one of the 12-16 requests of a segment costs that many L1 wavefronts.
// Here the 32 lanes of a segment do not take 32 pixels of one canvas row.  Per 32x32 tile the kernel looks at the
// direction in which the SOURCE y stays constant (from the derivatives of the canvas -> source map at the tile centre):
//
//   * mostly along canvas X (|roll| < 45 deg): lane l takes pixel (X0 + l, Y0 + ((r + sh(l)) mod 32)),
//   * mostly along canvas Y (|roll| > 45 deg): lane l takes pixel (X0 + ((r + sh(l)) mod 32), Y0 + l),
//
// with sh(l) = round(l * slope), |slope| <= 1: the lanes follow a source row, a tap request touches ~2 rows of lines at
// ANY roll (the straight-row kernels need a separate column-major tile path near +-90 deg and lose 20-40 % between
// 30 and 75 deg).  Every pixel of the tile is still produced exactly once -- 8 warps x 4 segments cover each column's
// (row's) 32 pixels -- by exactly the same arithmetic: only the lane -> pixel assignment changes, so the bits do not.
// The results are staged in shared memory and leave in one of two ways (chosen per CTA, same bits):
//   * TMA write-out (default, lanes along X): planar (plane, row, column) tile, bulk tensor stores issued by one thread
//     (StoreMaps below);
//   * LSU write-out (lanes along Y, validity output, VIDC_TMA_STORE=0): a 32x32 pixel-interleaved tile (one 128-bit deposit
//     per pixel, XOR-swizzled so that deposits and read-out are bank-conflict-free in both orientations), 128-bit row stores
//     per plane, validity mask and coverage count from the staged values.
//
// Without running output pointers the hot paths fit 32 registers: 8 CTAs = 64 warps are resident per SM where the
// straight-row kernels run 5 CTAs at 48 registers, which hides more of the gather latency than the smaller L1
// (8 x 16 KB of staging tiles) costs (profiles/r1_history.md, occupancy sweep).
//
// Contiguous planes whose width is a multiple of 32 (compile-time 640x480 and 320x240, or runtime geometry with input and
// canvas of the same size), 16-byte aligned outputs; everything else takes the straight-row kernels of kernels_fast.cuh.
#pragma once

namespace vidc_k {

// Segment (before the shear) that warp w processes in its j-th iteration.  Interleaved (w + 8 j): the eight warps of a
// CTA work on eight ADJACENT segments at any time, so the source lines two neighbouring segments share are reused while
// they are still in L1; blocked (4 w + j) leaves that reuse to the next iteration, after 63 other warps have gone through L1.
#ifndef VIDC_SEG_INTERLEAVED
#define VIDC_SEG_INTERLEAVED 1
#endif
#if VIDC_SEG_INTERLEAVED
#define VIDC_SEG(w, j) ((w) + 8 * (j))
#else
#define VIDC_SEG(w, j) ((w) * ROWS_PER_THREAD + (j))
#endif
// runtime-geometry instantiations (GW = GH = 0) carry W, H and W * H in registers
#ifndef VIDC_SHEAR_BLOCKS_RT
#define VIDC_SHEAR_BLOCKS_RT 6
#endif
#ifndef VIDC_SHEAR_BLOCKS_FWD
#define VIDC_SHEAR_BLOCKS_FWD 8
#endif
#ifndef VIDC_SHEAR_BLOCKS_INV
#define VIDC_SHEAR_BLOCKS_INV 8
#endif

// Staging-tile slot of pixel (row, col).  Deposits of a quarter-warp cover 8 consecutive columns (lanes along X, rows
// nearly equal) or 8 consecutive rows (lanes along Y, columns nearly equal); the read-out of a quarter-warp covers
// columns 4i + k, i = 0..7 of one row.  col ^ (col >> 3) spreads the read-out over the eight 16-byte bank groups, the
// extra (row & 7) term spreads the column-wise deposits.
template <bool ALONG_Y>
__device__ __forceinline__ int shear_slot(int row, int col) {
    return ALONG_Y ? (col ^ (col >> 3) ^ (row & 7)) : (col ^ (col >> 3));
}
// Orientation and per-lane shear of a tile from d(source y)/dX = ax and d(source y)/dY = ay (not part of any result).
__device__ __forceinline__ bool shear_of_tile(float ax, float ay, int lane, int& sh_l) {
    const bool along_y = fabsf(ax) > fabsf(ay);
    float slope = along_y ? -__fdividef(ay, ax) : -__fdividef(ax, ay);     // |slope| <= 1 in both orientations
    slope = (fabsf(slope) <= 1.0f) ? slope : 0.0f;                         // NaN / inf (degenerate frames): no shear
    sh_l = __float2int_rn(slope * (float)lane);
    return along_y;
}

// Exterior-tile bitmap (frame_params_tiles_kernel, kernels_params.cuh): bit t of vidc_frame_params::reserved is set when canvas
// tile t certainly lies outside the source footprint.  Parameters from any other producer carry zeros there.
__device__ __forceinline__ bool tile_marked_exterior(const vidc_frame_params* __restrict__ prm) {
    const unsigned int tile = blockIdx.y * gridDim.x + blockIdx.x;
    if (tile >= 320u) return false;
    const unsigned int word = __float_as_uint(__ldg(&prm->reserved[tile >> 5]));
    return (word >> (tile & 31u)) & 1u;
}

// L2 prefetch for the CTA that will run about half a wave later: the forward warp is bound by the latency of L1 misses that go
// all the way to DRAM (L2 hit rate 24 %, long-scoreboard 17.5 per issue: profiles/r1_shear_final_ncu_summary.txt); asking L2 for
// the rows of that tile's source box now turns them into L2 hits.  The box comes from the per-frame kernel
// (vidc::fwd_tile_src_box); thread -> (row, 128-byte segment); a hint only, no result depends on it.
// returns the frame index and the element offset (inside a plane) of this thread's 128-byte segment, or false
__device__ __forceinline__ bool prefetch_target(const uint4* __restrict__ boxes, int pf_x, int pf_y, int pf_z, int W, int& pz, long long& off) {
    if (!boxes) return false;
    int px = (int)blockIdx.x + pf_x, py = (int)blockIdx.y + pf_y;
    pz = (int)blockIdx.z + pf_z;
    if (px >= (int)gridDim.x) { px -= gridDim.x; ++py; }
    if (py >= (int)gridDim.y) { py -= gridDim.y; ++pz; }
    if (pz >= (int)gridDim.z) return false;
    const uint4 e = __ldg(boxes + (((size_t)pz * gridDim.y + py) * gridDim.x + px));
    if (e.z == 0u) return false;
    const int tid = threadIdx.y * 32 + threadIdx.x, seg = tid & 3, r = tid >> 2;           // <= 64 rows x 4 segments of 32 px
    if (r >= (int)e.w || seg * 32 >= (int)e.z + (int)(e.x & 31u)) return false;
    if ((int)((e.x & ~31u) + seg * 32) >= W) return false;
    off = (long long)(e.y + r) * W + ((e.x & ~31u) + seg * 32);
    return true;
}
template <bool HAS_D>
__device__ __forceinline__ void prefetch_source_box_l2(const FwdArgs& a, int W, int H) {
    int pz;
    long long off;
    if (!prefetch_target(a.src_boxes, a.pf_x, a.pf_y, a.pf_z, W, pz, off)) return;
    const float* __restrict__ rgb = a.rgb + (long long)pz * a.rgb_sn + off;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rgb));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rgb + W * H));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rgb + 2 * W * H));
    if (HAS_D) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dep + (long long)pz * a.dep_sn + off));
}

// ---- TMA write-out (TS = true; VIDC_TMA_STORE=0 turns it off) ---------------------------------------------------------------------
// Tiles whose lanes run along X deposit into a PLANAR staging tile (plane, row, column: the layout of a (32, 32, planes)
// tensor box; a lane's column is its bank, so deposits are conflict-free at any shear) and the tile leaves as bulk tensor
// stores issued by ONE thread (cp.async.bulk.tensor, SASS UTMASTG) instead of four LDS.128 + four STG.128 + a mask store per
// thread; the validity mask is staged as bytes and stored the same way; rows below a canvas whose height is not a multiple
// of 32 are clipped by the TMA unit.  Measured (profiles/r2_history.md): forward 0.515 -> 0.493 ms, inverse 0.492 -> 0.444 ms.
// Tiles whose lanes run along Y keep the interleaved tile and the LSU write-out (column-wise deposits into a planar tile
// would be 32-way bank conflicts, 4-way with the TMA's 128-byte swizzle -- no better than what they replace).
struct StoreMaps {
    CUtensorMap img;                         // (W, H, C, B) fp32 planes, box (32, 32, C, 1): RGB, normals, or the C planes of warp_forward
    CUtensorMap dep;                         // (W, H, B) fp32, box (32, 32, 1)
    CUtensorMap mask;                        // (W, H, B) uint8, box (32, 32, 1)
};
// one thread: the staged planes -> global memory; returns when the TMA unit has read the tile
__device__ __forceinline__ void tile_store_issue(const StoreMaps& maps, const void* img, const void* dep, const void* mask, int x0, int y0, int b) {
    tma_store_4d(&maps.img, img, x0, y0, 0, b);
    if (dep) tma_store_3d(&maps.dep, dep, x0, y0, b);
    if (mask) tma_store_3d(&maps.mask, mask, x0, y0, b);
    tma_store_commit_and_wait_read();
}

// ---- forward: RGB (3 planes) + optional depth, mask, coverage -------------------------------------------------------
// PLANAR (lanes along X only): deposits go to the planar tile `tp` and the mask bytes to `mt`; returns the lane's count of
// valid pixels (coverage).  Otherwise the interleaved float4 tile, mask and coverage from the staged values in the write-out.
template <int GW, int GH, bool HAS_D, bool ALONG_Y, bool PLANAR = false>
__device__ __forceinline__ unsigned int warp_rgbd_shear_segments(const FwdArgs& a, const float* pr, float4 (*tile)[32], int sh_l,
                                                                 unsigned char (*mt)[32] = nullptr) {
    static_assert(!(PLANAR && ALONG_Y), "planar deposits are for tiles whose lanes run along X");
    float (*tp)[32][32] = reinterpret_cast<float (*)[32][32]>(&tile[0][0]);
    unsigned int cnt = 0;
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    const float* Hi = pr + 2;
    const float px_min = pr[11], py_min = pr[12], ikw = pr[15], ikh = pr[16];
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in_rgb = a.rgb + (long long)b * a.rgb_sn;
    const float* __restrict__ in_dep = HAS_D ? a.dep + (long long)b * a.dep_sn : nullptr;
    // the coordinate a lane keeps over its four segments: X (lanes along X) or Y (lanes along Y)
    const float p_fix = ALONG_Y ? ikh * (float)(tileY0 + lane) + py_min : ikw * (float)(tileX0 + lane) + px_min;
    const float u0 = Hi[0] * p_fix, v0 = Hi[3] * p_fix, s0 = Hi[6] * p_fix;        // used when the fixed one is X
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int S = (VIDC_SEG(warp, j) + sh_l) & 31;
        float u, v, s;
        if (ALONG_Y) {
            const float px = ikw * (float)(tileX0 + S) + px_min;
            u = fmaf(Hi[1], p_fix, Hi[0] * px) + Hi[2];
            v = fmaf(Hi[4], p_fix, Hi[3] * px) + Hi[5];
            s = fmaf(Hi[7], p_fix, Hi[6] * px) + Hi[8];
        } else {
            const float py = ikh * (float)(tileY0 + S) + py_min;
            u = fmaf(Hi[1], py, u0) + Hi[2];
            v = fmaf(Hi[4], py, v0) + Hi[5];
            s = fmaf(Hi[7], py, s0) + Hi[8];
        }
        float sx, sy;
        div2_rn(u, v, s, sx, sy);                                  // :146-147
        const float gx = a.cam.inv_half_w * (sx - a.cam.cx);
        const float gy = a.cam.inv_half_h * (sy - a.cam.cy);
        const float ix = unnormalize(gx, Wf), iy = unnormalize(gy, Hf);
        const Pos t = make_pos(ix, iy, H, W);
        const Px4 o = fwd_sample_row<HAS_D>(in_rgb, in_dep, W, W * H, H, W, a.mode_d, ix, iy, t);
        if (PLANAR) {
            tp[0][S][lane] = o.r;
            tp[1][S][lane] = o.g;
            tp[2][S][lane] = o.b;
            if (HAS_D) tp[3][S][lane] = o.d;
            else if (a.dep_o) tp[3][S][lane] = 0.0f;               // sparse-depth route: the zero fill of the depth plane rides along
            if (a.mask || a.coverage) {                            // surface_normal.py:151
                const unsigned int m = (o.r + o.g) + o.b > 0.01f;
                mt[S][lane] = (unsigned char)m;
                if ((GW && GH % 32 == 0) || tileY0 + S < H) cnt += m;
            }
        } else {
            const int row = ALONG_Y ? lane : S, col = ALONG_Y ? S : lane;
            tile[row][shear_slot<ALONG_Y>(row, col)] = make_float4(o.r, o.g, o.b, o.d);
        }
    }
    return cnt;
}

// write-out: thread -> (row, 4 consecutive columns), 128-bit loads from the tile, 128-bit row stores per plane
template <int GW, int GH, bool HAS_D, bool ALONG_Y>
__device__ __forceinline__ void warp_rgbd_shear_write_out(const FwdArgs& a, const float4 (*tile)[32]) {
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane, row = tid >> 3, c4 = (tid & 7) * 4;
    const int Yo = blockIdx.y * TILE_H + row, Xo = blockIdx.x * TILE_W + c4;
    unsigned int cnt = 0;
    if ((GW && GH % 32 == 0) || Yo < H) {
        float4 px4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) px4[k] = tile[row][shear_slot<ALONG_Y>(row, c4 + k)];
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
        else if (a.dep_o)                                          // sparse-depth path: the plane is zero-filled here, the points
            *reinterpret_cast<float4*>(a.dep_o + ((long long)b * a.depo_sn + Yo * W + Xo)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // are warped by warp_sparse_depth_kernel
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

template <int GW, int GH, bool HAS_D, bool TS>
__global__ void __launch_bounds__(256, GW ? VIDC_SHEAR_BLOCKS_FWD : VIDC_SHEAR_BLOCKS_RT)
warp_rgbd_shear_kernel(const __grid_constant__ FwdArgs a, const __grid_constant__ StoreMaps maps) {
    static_assert(GW % 32 == 0, "sheared tiles need a canvas whose width is a multiple of 32 (GW = 0: runtime geometry)");
    static_assert(ROWS_PER_THREAD == 4 && PATCH_W == 32 && TILE_W == 32 && TILE_H == 32, "32x32 tile, 8 warps x 4 segments");
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    __shared__ __align__(128) float4 tile[32][32];
    __shared__ __align__(128) unsigned char mtile[TS ? 32 : 1][32];
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    prefetch_source_box_l2<HAS_D>(a, W, H);
    if (tile_marked_exterior(a.prm + b)) {                         // CTA-uniform: nothing of the source lands here, all zeros
        const int tid = warp * 32 + lane, row = tid >> 3, c4 = (tid & 7) * 4;
        const int Yo = tileY0 + row, Xo = tileX0 + c4;
        if ((GW && GH % 32 == 0) || Yo < H) {
            const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float* __restrict__ o_rgb = a.rgb_o + ((long long)b * a.rgbo_sn + Yo * W + Xo);
            *reinterpret_cast<float4*>(o_rgb) = z4;
            *reinterpret_cast<float4*>(o_rgb + W * H) = z4;
            *reinterpret_cast<float4*>(o_rgb + 2 * W * H) = z4;
            if (HAS_D || a.dep_o) *reinterpret_cast<float4*>(a.dep_o + ((long long)b * a.depo_sn + Yo * W + Xo)) = z4;
            if (a.mask) *reinterpret_cast<unsigned int*>(a.mask + (((long long)b * H + Yo) * W + Xo)) = 0u;
        }
        return;
    }
    // params: Hinv = floats 18..26, px_min,py_min = 27,28, ikw,ikh = 31,32 -> float4 #4..#8 (floats 16..35)
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    int sh_l;
    bool along_y;
    {
        const float* Hi = pr + 2;
        const float ikw = pr[15], ikh = pr[16];
        const float pxc = ikw * (float)(tileX0 + TILE_W / 2) + pr[11], pyc = ikh * (float)(tileY0 + TILE_H / 2) + pr[12];
        const float vc = fmaf(Hi[4], pyc, Hi[3] * pxc) + Hi[5], sc = fmaf(Hi[7], pyc, Hi[6] * pxc) + Hi[8];
        along_y = shear_of_tile(ikw * (Hi[3] * sc - vc * Hi[6]), ikh * (Hi[4] * sc - vc * Hi[7]), lane, sh_l);
    }
    if (along_y) {                                                 // CTA-uniform; the common orientation stays straight-line
        warp_rgbd_shear_segments<GW, GH, HAS_D, true>(a, pr, tile, sh_l);
        __syncthreads();
        warp_rgbd_shear_write_out<GW, GH, HAS_D, true>(a, tile);
        return;
    }
    if (TS) {
        unsigned int cnt = warp_rgbd_shear_segments<GW, GH, HAS_D, false, true>(a, pr, tile, sh_l, mtile);
        fence_async_smem();
        __syncthreads();
        const float* tp = reinterpret_cast<const float*>(&tile[0][0]);
        if (warp == 0 && lane == 0) tile_store_issue(maps, tp, (HAS_D || a.dep_o) ? tp + 3 * 1024 : nullptr, a.mask ? &mtile[0][0] : nullptr, tileX0, tileY0, b);
        if (a.coverage) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) atomicAdd(a.coverage + b, cnt);
        }
        return;
    }
    warp_rgbd_shear_segments<GW, GH, HAS_D, false>(a, pr, tile, sh_l);
    __syncthreads();
    warp_rgbd_shear_write_out<GW, GH, HAS_D, false>(a, tile);
}

// ---- forward, C planes in one interpolation mode: the reference-shaped call warp_with_gravity_center_aligned ------------
// (RGB: C = 3; depth through the 3-D path of :110-112: C = 1).  Same segments, staging and write-out; no mask.
struct PlanesArgs {
    const vidc_frame_params* prm; CamConst cam;
    const float* x; long long x_sn;         // input planes, contiguous W x H each
    float* y; long long y_sn;               // canvas planes, contiguous W x H each
    int mode;
    const uint4* src_boxes; int pf_x, pf_y, pf_z;      // L2 prefetch hints, as in FwdArgs
};

template <int GW, int GH, int C, bool ALONG_Y, bool PLANAR = false>
__device__ __forceinline__ void warp_planes_shear_segments(const PlanesArgs& a, const float* pr, float4 (*tile)[32], int sh_l) {
    static_assert(!(PLANAR && ALONG_Y), "planar deposits are for tiles whose lanes run along X");
    float (*tp)[32][32] = reinterpret_cast<float (*)[32][32]>(&tile[0][0]);
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    const float* Hi = pr + 2;
    const float px_min = pr[11], py_min = pr[12], ikw = pr[15], ikh = pr[16];
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in = a.x + (long long)b * a.x_sn;
    const float p_fix = ALONG_Y ? ikh * (float)(tileY0 + lane) + py_min : ikw * (float)(tileX0 + lane) + px_min;
    const float u0 = Hi[0] * p_fix, v0 = Hi[3] * p_fix, s0 = Hi[6] * p_fix;
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int S = (VIDC_SEG(warp, j) + sh_l) & 31;
        float u, v, s;
        if (ALONG_Y) {
            const float px = ikw * (float)(tileX0 + S) + px_min;
            u = fmaf(Hi[1], p_fix, Hi[0] * px) + Hi[2];
            v = fmaf(Hi[4], p_fix, Hi[3] * px) + Hi[5];
            s = fmaf(Hi[7], p_fix, Hi[6] * px) + Hi[8];
        } else {
            const float py = ikh * (float)(tileY0 + S) + py_min;
            u = fmaf(Hi[1], py, u0) + Hi[2];
            v = fmaf(Hi[4], py, v0) + Hi[5];
            s = fmaf(Hi[7], py, s0) + Hi[8];
        }
        float sx, sy;
        div2_rn(u, v, s, sx, sy);                                  // :146-147
        const float gx = a.cam.inv_half_w * (sx - a.cam.cx);
        const float gy = a.cam.inv_half_h * (sy - a.cam.cy);
        const float ix = unnormalize(gx, Wf), iy = unnormalize(gy, Hf);
        const Pos t = make_pos(ix, iy, H, W);
        float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (__any_sync(0xffffffffu, t.touch)) {                    // exterior segments: zeros
            if (a.mode != VIDC_BILINEAR) {
#pragma unroll
                for (int c = 0; c < C; ++c) o[c] = sample_nearest_pos(in + c * (W * H), ix, iy, H, W, W, t.touch);
            } else if (__all_sync(0xffffffffu, t.interior)) {
                const int off = t.y0 * W + t.x0;
#pragma unroll
                for (int c = 0; c < C; ++c) o[c] = sample_interior(in + c * (W * H), off, W, t);
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) o[c] = sample_border(in + c * (W * H), W, H, W, t);
            }
        }
        if (PLANAR) {
#pragma unroll
            for (int c = 0; c < C; ++c) tp[c][S][lane] = o[c];
        } else {
            const int row = ALONG_Y ? lane : S, col = ALONG_Y ? S : lane;
            tile[row][shear_slot<ALONG_Y>(row, col)] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <int GW, int GH, int C, bool ALONG_Y>
__device__ __forceinline__ void warp_planes_shear_write_out(const PlanesArgs& a, const float4 (*tile)[32]) {
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, tid = threadIdx.y * 32 + threadIdx.x, row = tid >> 3, c4 = (tid & 7) * 4;
    const int Yo = blockIdx.y * TILE_H + row, Xo = blockIdx.x * TILE_W + c4;
    if ((GW && GH % 32 == 0) || Yo < H) {
        float4 px4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) px4[k] = tile[row][shear_slot<ALONG_Y>(row, c4 + k)];
        float* __restrict__ o = a.y + ((long long)b * a.y_sn + Yo * W + Xo);
        *reinterpret_cast<float4*>(o) = make_float4(px4[0].x, px4[1].x, px4[2].x, px4[3].x);
        if (C > 1) *reinterpret_cast<float4*>(o + W * H) = make_float4(px4[0].y, px4[1].y, px4[2].y, px4[3].y);
        if (C > 2) *reinterpret_cast<float4*>(o + 2 * W * H) = make_float4(px4[0].z, px4[1].z, px4[2].z, px4[3].z);
        if (C > 3) *reinterpret_cast<float4*>(o + 3 * W * H) = make_float4(px4[0].w, px4[1].w, px4[2].w, px4[3].w);
    }
}

template <int GW, int GH, int C, bool TS>
__global__ void __launch_bounds__(256, GW ? VIDC_SHEAR_BLOCKS_FWD : VIDC_SHEAR_BLOCKS_RT)
warp_planes_shear_kernel(const __grid_constant__ PlanesArgs a, const __grid_constant__ StoreMaps maps) {
    static_assert(GW % 32 == 0 && C >= 1 && C <= 4, "canvas width a multiple of 32, 1-4 planes");
    static_assert(ROWS_PER_THREAD == 4 && PATCH_W == 32 && TILE_W == 32 && TILE_H == 32, "32x32 tile, 8 warps x 4 segments");
    __shared__ __align__(128) float4 tile[32][32];
    const int b = blockIdx.z, lane = threadIdx.x;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    if (C >= 3) {                                                  // L2 prefetch of a later tile's source box (see warp_rgbd_shear_kernel);
                                                                   // measured: RGB 0.412 -> 0.397 ms, but one plane 0.263 -> 0.284 ms
        const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
        int pz;
        long long off;
        if (prefetch_target(a.src_boxes, a.pf_x, a.pf_y, a.pf_z, W, pz, off)) {
            const float* __restrict__ x = a.x + (long long)pz * a.x_sn + off;
#pragma unroll
            for (int c = 0; c < C; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + c * (W * H)));
        }
    }
    if (tile_marked_exterior(a.prm + b)) {                         // CTA-uniform: all zeros
        const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
        const int tid = threadIdx.y * 32 + lane, row = tid >> 3, c4 = (tid & 7) * 4;
        const int Yo = tileY0 + row, Xo = tileX0 + c4;
        if ((GW && GH % 32 == 0) || Yo < H) {
            const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float* __restrict__ o = a.y + ((long long)b * a.y_sn + Yo * W + Xo);
#pragma unroll
            for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(o + c * (W * H)) = z4;
        }
        return;
    }
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    int sh_l;
    bool along_y;
    {
        const float* Hi = pr + 2;
        const float ikw = pr[15], ikh = pr[16];
        const float pxc = ikw * (float)(tileX0 + TILE_W / 2) + pr[11], pyc = ikh * (float)(tileY0 + TILE_H / 2) + pr[12];
        const float vc = fmaf(Hi[4], pyc, Hi[3] * pxc) + Hi[5], sc = fmaf(Hi[7], pyc, Hi[6] * pxc) + Hi[8];
        along_y = shear_of_tile(ikw * (Hi[3] * sc - vc * Hi[6]), ikh * (Hi[4] * sc - vc * Hi[7]), lane, sh_l);
    }
    if (along_y) {                                                 // CTA-uniform
        warp_planes_shear_segments<GW, GH, C, true>(a, pr, tile, sh_l);
        __syncthreads();
        warp_planes_shear_write_out<GW, GH, C, true>(a, tile);
        return;
    }
    if (TS) {
        warp_planes_shear_segments<GW, GH, C, false, true>(a, pr, tile, sh_l);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.y == 0 && lane == 0) tile_store_issue(maps, &tile[0][0], nullptr, nullptr, tileX0, tileY0, b);
        return;
    }
    warp_planes_shear_segments<GW, GH, C, false>(a, pr, tile, sh_l);
    __syncthreads();
    warp_planes_shear_write_out<GW, GH, C, false>(a, tile);
}

// ---- packed fp32x2 forms for the inverse warp (VIDC_SHEAR_PACKED) --------------------------------------------------------------
// The inverse kernel is bound by issue slots (81 % issue-active with the TMA write-out).  Blackwell's FFMA2 / FMUL2 / FADD2 do
// two independent fp32 operations per lane and instruction, each rounded exactly like its scalar form, so pairing the (u, v)
// halves of the coordinate chain, two of the three planes of the interpolation and of the rotation, and the divisions that
// share a reciprocal removes instructions without moving a bit -- with the one hazard kernels_fast.cuh documents (ptxas fuses
// a packed multiply that feeds a packed add into FFMA2 even under -fmad=false): such multiplies stay scalar.
// Measured (profiles/r2_history.md): 0.4479 -> 0.4411 ms for the inverse (the register pairing costs ~85 moves per kernel, about
// half of what the packed operations save); packing only the coordinate chain, or the chain and the rotation, spills at 32
// registers and is slower (0.475 / 0.488 ms).  -DVIDC_SHEAR_PACKED=0 restores the scalar form (same bits).
#ifndef VIDC_SHEAR_PACKED
#define VIDC_SHEAR_PACKED 1
#endif
#ifndef VIDC_SHEAR_PACKED_SAMPLE
#define VIDC_SHEAR_PACKED_SAMPLE 1
#endif
#ifndef VIDC_SHEAR_PACKED_ROT
#define VIDC_SHEAR_PACKED_ROT 1
#endif
__device__ __forceinline__ float2 div2p_sel(float2 uv, float s, bool proven) {       // packed div2_sel
    const float r = rcp_refined(s);
    float2 q = mul2(uv, bc(r));
    const float2 rem = fma2(bc(-s), q, uv);
    q = fma2(rem, bc(r), q);
    if (!proven) {                                                 // CTA-uniform
        const float as = fabsf(s);
        const float hi = fmaxf(fmaxf(fabsf(uv.x), fabsf(uv.y)), as * 0x1p40f);
        const float lo = fminf(fminf(fabsf(uv.x), fabsf(uv.y)), as * 0x1p-40f);
        if (!(lo >= 0x1p-80f && hi <= 0x1p80f)) {
            q.x = ieee_div_slow(uv.x, s);
            q.y = ieee_div_slow(uv.y, s);
        }
    }
    return q;
}
__device__ __forceinline__ void normalize3p_rn(float2& z01, float& z2) {             // packed normalize3_rn
    const float2 sq = mul2(z01, z01);
    const float ss = (sq.x + sq.y) + z2 * z2;
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(ss));
    const float s = ss * rs, h = rs * 0.5f;
    const float n = fmaf(fmaf(-s, s, ss), h, s);
    const float r = rcp_refined(n);
    float2 q = mul2(z01, bc(r));
    const float2 rem = fma2(bc(-n), q, z01);
    q = fma2(rem, bc(r), q);
    const float q2 = div_with_rcp(z2, n, r);
    const unsigned k0 = __float_as_uint(z01.x) * 2u - 1u, k1 = __float_as_uint(z01.y) * 2u - 1u, k2 = __float_as_uint(z2) * 2u - 1u;
    const bool comps_ok = min(min(k0, k1), k2) >= 0x42ffffffu;
    const bool ss_ok = (__float_as_uint(ss) - 0x21800000u) <= 0x3c000000u;
    if (comps_ok && ss_ok) {
        z01 = q; z2 = q2;
    } else {
        const float ns = ieee_norm_slow(ss);
        z01.x = ieee_div_slow(z01.x, ns); z01.y = ieee_div_slow(z01.y, ns); z2 = ieee_div_slow(z2, ns);
    }
}
__device__ __forceinline__ Px3 inv_sample_row_lazy_p(const float* __restrict__ in, int x_sh, int x_sc, int H, int W, const Pos& t0) {
    Px3 o = {0.0f, 0.0f, 0.0f};
    if (__all_sync(0xffffffffu, t0.interior)) {
        o = inv_sample_interior_p(in, x_sh, x_sc, t0);
    } else {
        Pos t = t0;
        t.touch = t0.fin && (unsigned)(t0.x0 + 1) <= (unsigned)W && (unsigned)(t0.y0 + 1) <= (unsigned)H;
        if (__any_sync(0xffffffffu, t.touch)) {
            o.a = sample_border(in, x_sh, H, W, t);
            o.b = sample_border(in + x_sc, x_sh, H, W, t);
            o.c = sample_border(in + 2 * x_sc, x_sh, H, W, t);
        }
    }
    return o;
}

// ---- inverse: camera px -> canvas coords, 3 planes, R^T, renormalisation ------------------------------------------------
// The fourth component of the staging slot carries the optional validity flag.
template <int GW, int GH, bool NORMALIZE, bool HAS_VALID, bool ALONG_Y, bool PLANAR = false>
__device__ __forceinline__ void unwarp_normals_shear_segments(const InvArgs& a, const float* pr, float4 (*tile)[32], int sh_l, bool proven) {
    static_assert(!(PLANAR && (ALONG_Y || HAS_VALID)), "planar deposits: lanes along X, no validity output");
    float (*tp)[32][32] = reinterpret_cast<float (*)[32][32]>(&tile[0][0]);
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    const float* Hm = pr;
    const float* R = pr + 9;
    const float px_min = pr[27], py_min = pr[28], kw = pr[29], kh = pr[30];
    const float Wf = (float)W, Hf = (float)H;
    const float* __restrict__ in = a.x + (long long)b * a.x_sn;
    const float c_fix = ALONG_Y ? (float)(tileY0 + lane) : (float)(tileX0 + lane);
    const float u0 = Hm[0] * c_fix, v0 = Hm[3] * c_fix, s0 = Hm[6] * c_fix;        // used when the fixed one is X
#pragma unroll
    for (int j = 0; j < ROWS_PER_THREAD; ++j) {
        const int S = (VIDC_SEG(warp, j) + sh_l) & 31;
#if VIDC_SHEAR_PACKED
        float s;
        float2 uv;
        if (ALONG_Y) {
            const float Xf = (float)(tileX0 + S);
            s = fmaf(Hm[7], c_fix, Hm[6] * Xf) + Hm[8];
            uv = add2(fma2(f2(Hm[1], Hm[4]), bc(c_fix), f2(Hm[0] * Xf, Hm[3] * Xf)), f2(Hm[2], Hm[5]));
        } else {
            const float Yf = (float)(tileY0 + S);
            s = fmaf(Hm[7], Yf, s0) + Hm[8];
            uv = add2(fma2(f2(Hm[1], Hm[4]), bc(Yf), f2(u0, v0)), f2(Hm[2], Hm[5]));
        }
        const float2 txy = div2p_sel(uv, s, proven);               // :245
        const float2 tm = sub2(txy, f2(px_min, py_min));
        const float2 cm = sub2(f2(kw * tm.x, kh * tm.y), f2(a.cam.cx, a.cam.cy));              // scalar multiplies: see above
        const float2 g1 = add2(f2(a.cam.inv_half_w * cm.x, a.cam.inv_half_h * cm.y), bc(1.0f));
        const float2 ixy = mul2(fma2(g1, f2(Wf, Hf), bc(-1.0f)), bc(0.5f));                    // ATen unnormalise
        const Pos t = make_pos_p(ixy, H, W);
        const Px3 y = HAS_VALID ? inv_sample_row(in, W, W * H, H, W, t)
                                : (VIDC_SHEAR_PACKED_SAMPLE ? inv_sample_row_lazy_p(in, W, W * H, H, W, t) : inv_sample_row_lazy(in, W, W * H, H, W, t));
#if VIDC_SHEAR_PACKED_ROT
        float2 z01 = fma2(f2(R[6], R[7]), bc(y.c), fma2(f2(R[3], R[4]), bc(y.b), fma2(f2(R[0], R[1]), bc(y.a), bc(0.0f))));   // :253
        float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
        if (NORMALIZE) normalize3p_rn(z01, z2);                    // surface_normal.py:170
        const float z0 = z01.x, z1 = z01.y;
#else
        float z0 = fmaf(R[6], y.c, fmaf(R[3], y.b, fmaf(R[0], y.a, 0.0f)));
        float z1 = fmaf(R[7], y.c, fmaf(R[4], y.b, fmaf(R[1], y.a, 0.0f)));
        float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
        if (NORMALIZE) normalize3_rn(z0, z1, z2);
#endif
#else
        float u, v, s;
        if (ALONG_Y) {
            const float Xf = (float)(tileX0 + S);
            s = fmaf(Hm[7], c_fix, Hm[6] * Xf) + Hm[8];
            u = fmaf(Hm[1], c_fix, Hm[0] * Xf) + Hm[2];
            v = fmaf(Hm[4], c_fix, Hm[3] * Xf) + Hm[5];
        } else {
            const float Yf = (float)(tileY0 + S);
            s = fmaf(Hm[7], Yf, s0) + Hm[8];
            u = fmaf(Hm[1], Yf, u0) + Hm[2];
            v = fmaf(Hm[4], Yf, v0) + Hm[5];
        }
        float tx, ty;
        div2_sel(u, v, s, proven, tx, ty);                         // :245 (window test only for frames that did not pass the proof)
        const float cxp = kw * (tx - px_min);
        const float cyp = kh * (ty - py_min);
        const float gx = a.cam.inv_half_w * (cxp - a.cam.cx);
        const float gy = a.cam.inv_half_h * (cyp - a.cam.cy);
        const Pos t = make_pos(unnormalize(gx, Wf), unnormalize(gy, Hf), H, W);
        const Px3 y = HAS_VALID ? inv_sample_row(in, W, W * H, H, W, t) : inv_sample_row_lazy(in, W, W * H, H, W, t);
        // z = C_R_Cg.bmm(y), C_R_Cg = R^T: k-ascending FMA chain from a +0 accumulator (:253)
        float z0 = fmaf(R[6], y.c, fmaf(R[3], y.b, fmaf(R[0], y.a, 0.0f)));
        float z1 = fmaf(R[7], y.c, fmaf(R[4], y.b, fmaf(R[1], y.a, 0.0f)));
        float z2 = fmaf(R[8], y.c, fmaf(R[5], y.b, fmaf(R[2], y.a, 0.0f)));
        if (NORMALIZE) normalize3_rn(z0, z1, z2);                  // surface_normal.py:170
#endif
        if (PLANAR) {
            tp[0][S][lane] = z0;
            tp[1][S][lane] = z1;
            tp[2][S][lane] = z2;
        } else {
            const int row = ALONG_Y ? lane : S, col = ALONG_Y ? S : lane;
            tile[row][shear_slot<ALONG_Y>(row, col)] = make_float4(z0, z1, z2, (HAS_VALID && t.touch) ? 1.0f : 0.0f);
        }
    }
}

// Two passes of (row, 2 consecutive columns) per thread, 64-bit row stores per plane.  (Measured against one pass of four
// columns with 128-bit stores: this form keeps the whole kernel inside 32 registers without spills.)
template <int GW, int GH, bool HAS_VALID, bool ALONG_Y>
__device__ __forceinline__ void unwarp_normals_shear_write_out(const InvArgs& a, const float4 (*tile)[32]) {
    const int W = GW ? GW : a.cam.W, H = GW ? GH : a.cam.H;
    const int b = blockIdx.z, tid = threadIdx.y * 32 + threadIdx.x, c2 = (tid & 15) * 2;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int row = (tid >> 4) + 16 * pass;
        const int Yo = blockIdx.y * TILE_H + row, Xo = blockIdx.x * TILE_W + c2;
        if ((GW && GH % 32 == 0) || Yo < H) {
            const float4 za = tile[row][shear_slot<ALONG_Y>(row, c2)], zb = tile[row][shear_slot<ALONG_Y>(row, c2 + 1)];
            float* __restrict__ o = a.z + ((long long)b * a.z_sn + Yo * W + Xo);
            *reinterpret_cast<float2*>(o) = make_float2(za.x, zb.x);
            *reinterpret_cast<float2*>(o + W * H) = make_float2(za.y, zb.y);
            *reinterpret_cast<float2*>(o + 2 * W * H) = make_float2(za.z, zb.z);
            if (HAS_VALID)
                *reinterpret_cast<unsigned short*>(a.valid + (((long long)b * H + Yo) * W + Xo)) =
                    (unsigned short)((za.w != 0.0f ? 1u : 0u) | (zb.w != 0.0f ? 0x100u : 0u));
        }
    }
}

// HAS_VALID = false (no validity output requested) drops the `touch` test from the all-interior fast path
// TS (TMA write-out, see StoreMaps above) is for calls without a validity output
template <int GW, int GH, bool NORMALIZE, bool HAS_VALID, bool TS>
__global__ void __launch_bounds__(256, GW ? VIDC_SHEAR_BLOCKS_INV : VIDC_SHEAR_BLOCKS_RT)
unwarp_normals_shear_kernel(const __grid_constant__ InvArgs a, const __grid_constant__ StoreMaps maps) {
    static_assert(GW % 32 == 0, "sheared tiles need a canvas whose width is a multiple of 32 (GW = 0: runtime geometry)");
    static_assert(ROWS_PER_THREAD == 4 && PATCH_W == 32 && TILE_W == 32 && TILE_H == 32, "32x32 tile, 8 warps x 4 segments");
    static_assert(!(TS && HAS_VALID), "the TMA write-out has no validity plane");
    __shared__ __align__(128) float4 tile[32][32];
    const int b = blockIdx.z, lane = threadIdx.x, warp = threadIdx.y;
    const int tileX0 = blockIdx.x * TILE_W, tileY0 = blockIdx.y * TILE_H;
    // H = floats 0..8, R = 9..17, px_min,py_min = 27,28, kw,kh = 29,30 -> float4 #0..#7 (floats 0..31)
    float pr[32];
    load_params(a.prm + b, pr, 0, 8);
    int sh_l;
    bool along_y;
    {
        const float* Hm = pr;
        const float xc = (float)(tileX0 + TILE_W / 2), yc = (float)(tileY0 + TILE_H / 2);
        const float vc = fmaf(Hm[4], yc, Hm[3] * xc) + Hm[5], sc = fmaf(Hm[7], yc, Hm[6] * xc) + Hm[8];
        along_y = shear_of_tile(Hm[3] * sc - vc * Hm[6], Hm[4] * sc - vc * Hm[7], lane, sh_l);
    }
    const bool proven = __ldg(&a.prm[b].reserved[10]) != 0.0f;     // CTA-uniform (vidc::inv_division_proven)
    if (along_y) {                                                 // CTA-uniform; the common orientation stays straight-line
        unwarp_normals_shear_segments<GW, GH, NORMALIZE, HAS_VALID, true>(a, pr, tile, sh_l, proven);
        __syncthreads();
        unwarp_normals_shear_write_out<GW, GH, HAS_VALID, true>(a, tile);
        return;
    }
    if (TS) {
        unwarp_normals_shear_segments<GW, GH, NORMALIZE, false, false, true>(a, pr, tile, sh_l, proven);
        fence_async_smem();
        __syncthreads();
        if (warp == 0 && lane == 0) tile_store_issue(maps, &tile[0][0], nullptr, nullptr, tileX0, tileY0, b);
        return;
    }
    unwarp_normals_shear_segments<GW, GH, NORMALIZE, HAS_VALID, false>(a, pr, tile, sh_l, proven);
    __syncthreads();
    unwarp_normals_shear_write_out<GW, GH, HAS_VALID, false>(a, tile);
}

}  // namespace vidc_k
