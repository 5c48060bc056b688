// kernels_sparse.cuh -- SURVEY.md section 8 row f2, second half: the forward warp of a SPARSE depth image done analytically.
//
// The reference rasterises ~150 KLT points per frame into a (1,H,W) image that is 99.8 % zeros (dataset.py:496-510) and then
// resamples that image like any other (warping_2dof_alignment.py:108-156, the 3-D input path of :110-112).  Resampling a
// mostly-zero plane costs the forward kernel a quarter of its gathers and 4 B/px of HBM reads.  Here the points themselves
// are warped: a few CTAs per frame, each of which
//   1. rasterises the frame's tracks exactly as the loader does (fp64 pixel arithmetic, truncation, LAST point on a pixel
//      wins) into a shared-memory hash pixel -> depth instead of an image,
//   2. for every occupied pixel finds the canvas pixels whose 2x2 bilinear footprint (or nearest tap) can contain it -- the
//      canvas-side bounding box of the +-1 px window around the pixel, from its four mapped corners, plus a pixel of margin --
//   3. and evaluates those canvas pixels with exactly the arithmetic of the dense kernels (forward_coords, bilinear_taps, the
//      nw, ne, sw, se FMA chain; absent taps read as +0 from the hash): the value written is the one the dense resample
//      produces, bit for bit, whichever of the (up to four) contributing points triggered it.
// The canvas plane itself is zero-filled by the RGB kernel's write-out (warp_rgbd_shear_kernel<.., HAS_D = false> with
// dep_o set).  A point whose window straddles a projective pole (no bounded pre-image) is evaluated against every canvas
// pixel by the whole CTA; correctness never depends on the bounding box being tight, only on it being conservative.
#pragma once

namespace vidc_k {

constexpr int SPARSE_SLOTS = 4096, SPARSE_MAX_POINTS = 2048;

struct SparseArgs {
    const vidc_frame_params* prm; CamConst cam;
    const double* tracks; const int* counts; int N, cols;
    double fc0, fc1, cc0, cc1;
    float* dep_o; long long depo_sn;
    int mode;
};

__device__ __forceinline__ unsigned int sparse_hash(int key) { return ((unsigned int)key * 2654435761u) >> 20; }   // 12 bits

__device__ __forceinline__ float sparse_lookup(const int* __restrict__ keys, const float* __restrict__ vals, int key) {
    unsigned int s = sparse_hash(key);
    while (true) {
        const int k = keys[s];
        if (k == key) return vals[s];
        if (k < 0) return 0.0f;                                   // an empty slot ends the probe sequence: the pixel holds +0
        s = (s + 1u) & (SPARSE_SLOTS - 1);
    }
}

// One canvas pixel against one occupied source pixel `key`: writes the dense kernels' value if the pixel is among its taps.
__device__ __forceinline__ void sparse_eval(const SparseArgs& a, const float* pr, const int* keys, const float* vals, int key, float val,
                                            int X, int Y, float* __restrict__ out) {
    const int W = a.cam.W, H = a.cam.H;
    const float* Hi = pr + 2;
    float ix, iy;
    forward_coords(Hi, pr[11], pr[12], pr[15], pr[16], a.cam, (float)X, (float)Y, (float)W, (float)H, ix, iy);
    if (a.mode == VIDC_BILINEAR) {
        const Taps t = bilinear_taps(ix, iy, H, W, W, 1);
        const bool hit = (t.b_nw && t.o_nw == key) || (t.b_ne && t.o_ne == key) || (t.b_sw && t.o_sw == key) || (t.b_se && t.o_se == key);
        if (!hit) return;
        const float v_nw = t.b_nw ? sparse_lookup(keys, vals, t.o_nw) : 0.0f;
        const float v_ne = t.b_ne ? sparse_lookup(keys, vals, t.o_ne) : 0.0f;
        const float v_sw = t.b_sw ? sparse_lookup(keys, vals, t.o_sw) : 0.0f;
        const float v_se = t.b_se ? sparse_lookup(keys, vals, t.o_se) : 0.0f;
        float acc = v_nw * t.w_nw;                                // sample_bilinear's chain (device_common.cuh)
        acc = fmaf(v_ne, t.w_ne, acc);
        acc = fmaf(v_sw, t.w_sw, acc);
        acc = fmaf(v_se, t.w_se, acc);
        out[Y * W + X] = acc;
    } else {
        const int xn = (int)rintf(ix), yn = (int)rintf(iy);       // sample_nearest
        if ((unsigned)xn < (unsigned)W && (unsigned)yn < (unsigned)H && yn * W + xn == key) out[Y * W + X] = val;
    }
}

__global__ void __launch_bounds__(256) warp_sparse_depth_kernel(const __grid_constant__ SparseArgs a) {
    __shared__ int keys[SPARSE_SLOTS];
    __shared__ int idxs[SPARSE_SLOTS];                            // winning point index while rasterising, then the depth in place
    float* vals = reinterpret_cast<float*>(idxs);
    __shared__ unsigned int slow_mask[SPARSE_SLOTS / 32];        // occupied pixels without a bounded pre-image
    __shared__ int n_slow;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.cam.W, H = a.cam.H;
    for (int s = tid; s < SPARSE_SLOTS; s += 256) { keys[s] = -1; idxs[s] = -1; }
    if (tid < SPARSE_SLOTS / 32) slow_mask[tid] = 0u;
    if (tid == 0) n_slow = 0;
    __syncthreads();
    // 1. rasterise (dataset.py:496-510): fp64, int() truncation, the last point on a pixel wins
    int n = a.counts ? a.counts[b] : a.N;
    n = n < a.N ? n : a.N;
    for (int i = tid; i < n; i += 256) {
        const double* t = a.tracks + ((long long)b * a.N + i) * a.cols;
        const double u = t[1] / t[3], v = t[2] / t[3];
        const double px = a.fc0 * u + a.cc0, py = a.fc1 * v + a.cc1;
        if (!(fabs(px) < 2.0e9) || !(fabs(py) < 2.0e9)) continue;
        const int col = (int)px, row = (int)py;
        if (row < 0 || row >= H || col < 0 || col >= W) continue;
        const int key = row * W + col;
        unsigned int s = sparse_hash(key);
        while (true) {
            const int prev = atomicCAS(&keys[s], -1, key);
            if (prev == -1 || prev == key) { atomicMax(&idxs[s], i); break; }
            s = (s + 1u) & (SPARSE_SLOTS - 1);
        }
    }
    __syncthreads();
    for (int s = tid; s < SPARSE_SLOTS; s += 256)
        if (keys[s] >= 0) {
            const int i = idxs[s];
            vals[s] = (float)a.tracks[((long long)b * a.N + i) * a.cols + 3];                          // klt_tracks[i, 3]  :510
        }
    __syncthreads();
    // Hinv = floats 18..26, px_min,py_min = 27,28, kw,kh = 29,30, ikw,ikh = 31,32 -> float4 #4..#8 (floats 16..35); H = floats 0..8
    float pr[20];
    load_params(a.prm + b, pr, 4, 5);
    float Hm[12];
    load_params(a.prm + b, Hm, 0, 3);
    const float kw = pr[13], kh = pr[14], px_min = pr[11], py_min = pr[12];
    float* __restrict__ out = a.dep_o + (long long)b * a.depo_sn;
    // 2./3. every occupied pixel: one warp each
    // (gridDim.y CTAs share a frame: each builds the whole table -- cheap -- and takes the pixels with key % gridDim.y ==
    // blockIdx.y.  The split must be a function of the KEY: the slot a key lands in depends on the insertion order, which
    // differs from CTA to CTA.)
    for (int base = warp * 32; base < SPARSE_SLOTS; base += 8 * 32) {
        const int k_l = keys[base + lane];
        unsigned int live = __ballot_sync(0xffffffffu, k_l >= 0 && (unsigned)k_l % gridDim.y == blockIdx.y);
        while (live) {
            const int s = base + (__ffs(live) - 1);
            live &= live - 1;
            const int key = keys[s];
            const float val = vals[s];
            const int py = key / W, px = key - py * W;
            // corners of the +-1 px window in source index space -> camera pixels -> canvas (lanes 0..3; approximate, margin below)
            const float six = (float)px + ((lane & 1) ? 1.0f : -1.0f), siy = (float)py + ((lane & 2) ? 1.0f : -1.0f);
            const float sx = six - 0.5f * (float)(W - 1) + a.cam.cx, sy = siy - 0.5f * (float)(H - 1) + a.cam.cy;
            const float t0 = Hm[6] * sx, t1 = Hm[7] * sy;
            const float sden = t0 + t1 + Hm[8];
            const float cX = kw * ((Hm[0] * sx + Hm[1] * sy + Hm[2]) / sden - px_min);
            const float cY = kh * ((Hm[3] * sx + Hm[4] * sy + Hm[5]) / sden - py_min);
            bool ok = fabsf(sden) > 1e-3f * (fabsf(t0) + fabsf(t1) + fabsf(Hm[8])) && fabsf(cX) < 1e8f && fabsf(cY) < 1e8f;
            float xlo = cX, xhi = cX, ylo = cY, yhi = cY, slo = sden, shi = sden;
#pragma unroll
            for (int m = 1; m <= 2; m <<= 1) {
                xlo = fminf(xlo, __shfl_xor_sync(0xffffffffu, xlo, m)); xhi = fmaxf(xhi, __shfl_xor_sync(0xffffffffu, xhi, m));
                ylo = fminf(ylo, __shfl_xor_sync(0xffffffffu, ylo, m)); yhi = fmaxf(yhi, __shfl_xor_sync(0xffffffffu, yhi, m));
                slo = fminf(slo, __shfl_xor_sync(0xffffffffu, slo, m)); shi = fmaxf(shi, __shfl_xor_sync(0xffffffffu, shi, m));
                const int other_ok = __shfl_xor_sync(0xffffffffu, (int)ok, m);     // unconditionally: every lane takes part
                ok = ok && other_ok != 0;
            }
            ok = __shfl_sync(0xffffffffu, (int)(ok && (slo > 0.0f || shi < 0.0f)), 0);
            xlo = __shfl_sync(0xffffffffu, xlo, 0); xhi = __shfl_sync(0xffffffffu, xhi, 0);
            ylo = __shfl_sync(0xffffffffu, ylo, 0); yhi = __shfl_sync(0xffffffffu, yhi, 0);
            if (!ok) {                                            // pole inside the window: no bounded pre-image
                if (lane == 0) { atomicOr(&slow_mask[s >> 5], 1u << (s & 31)); n_slow = 1; }
                continue;
            }
            int X0 = (int)floorf(xlo - 1.0f), X1 = (int)ceilf(xhi + 1.0f), Y0 = (int)floorf(ylo - 1.0f), Y1 = (int)ceilf(yhi + 1.0f);
            X0 = X0 < 0 ? 0 : X0; Y0 = Y0 < 0 ? 0 : Y0; X1 = X1 > W - 1 ? W - 1 : X1; Y1 = Y1 > H - 1 ? H - 1 : Y1;
            if (X1 < X0 || Y1 < Y0) continue;                     // the pixel's pre-image lies outside the canvas
            const int wc = X1 - X0 + 1, nc = wc * (Y1 - Y0 + 1);
            for (int c = lane; c < nc; c += 32) sparse_eval(a, pr, keys, vals, key, val, X0 + c % wc, Y0 + c / wc, out);
        }
    }
    __syncthreads();
    // Points without a bounded pre-image (frames whose horizon crosses the image): ONE pass of this CTA over the canvas for all of
    // them together -- a canvas pixel is evaluated like in the dense kernels and written when one of its taps is such a point.
    if (n_slow) {
        for (int p = tid; p < W * H; p += 256) {
            const int X = p % W, Y = p / W;
            float ix, iy;
            forward_coords(pr + 2, pr[11], pr[12], pr[15], pr[16], a.cam, (float)X, (float)Y, (float)W, (float)H, ix, iy);
            int cand[4]; bool inb[4];
            if (a.mode == VIDC_BILINEAR) {
                const Taps t = bilinear_taps(ix, iy, H, W, W, 1);
                cand[0] = t.o_nw; cand[1] = t.o_ne; cand[2] = t.o_sw; cand[3] = t.o_se;
                inb[0] = t.b_nw; inb[1] = t.b_ne; inb[2] = t.b_sw; inb[3] = t.b_se;
            } else {
                const int xn = (int)rintf(ix), yn = (int)rintf(iy);
                cand[0] = yn * W + xn; inb[0] = (unsigned)xn < (unsigned)W && (unsigned)yn < (unsigned)H;
                inb[1] = inb[2] = inb[3] = false; cand[1] = cand[2] = cand[3] = 0;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!inb[k]) continue;
                unsigned int s = sparse_hash(cand[k]);            // is this tap an occupied pixel flagged slow (by THIS CTA)?
                while (keys[s] >= 0 && keys[s] != cand[k]) s = (s + 1u) & (SPARSE_SLOTS - 1);
                if (keys[s] == cand[k] && ((slow_mask[s >> 5] >> (s & 31)) & 1u)) {
                    sparse_eval(a, pr, keys, vals, cand[k], vals[s], X, Y, out);
                    break;
                }
            }
        }
    }
}

}  // namespace vidc_k
