// Host build of the product's exact-arithmetic header (vi_depth_completion_b200/csrc/exact_math.cuh)
// so the CPU test-suite can sweep it against libm / torch without a GPU.
#include "../../vi_depth_completion_b200/csrc/exact_math.cuh"
#include <stddef.h>
extern "C" {
void host_glibc_atan2f(const float* y, const float* x, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = vidc::glibc_atan2f(y[i], x[i]);
}
void host_mkl_cosf_ha(const float* x, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = vidc::mkl_cosf_ha(x[i]);
}
// exhaustive self-check against the C library: all y-bit-patterns in [ylo, yhi) step ystep, for a fixed x
size_t host_atan2f_sweep_vs_libm(uint32_t ylo, uint32_t yhi, uint32_t ystep, float x) {
    size_t bad = 0;
    for (uint64_t b = ylo; b < yhi; b += ystep) {
        const float y = vidc::u2f((uint32_t)b);
        const float a = vidc::glibc_atan2f(y, x), r = atan2f(y, x);
        if (vidc::f2u(a) != vidc::f2u(r) && !(a != a && r != r)) ++bad;
    }
    return bad;
}
}

// Host build of the product's per-frame parameter code (csrc/frame_params.cuh): the same source the
// params kernel compiles for the device, so goldens can pin it without a GPU.
#include "../../vi_depth_completion_b200/csrc/frame_params.cuh"
extern "C" void host_frame_params(const vidc_camera* cam, const float* Ig, const float* Ia, int B, vidc_frame_params* out) {
    for (int i = 0; i < B; ++i) {
        memset(&out[i], 0, sizeof(vidc_frame_params));
        vidc::frame_params_from_gravity(*cam, Ig + 3 * i, Ia + 3 * i, out[i]);
    }
}

extern "C" void host_condition_gravity(const float* raw, int B, int rule, float* Ig, float* Ia) {
    for (int i = 0; i < B; ++i) vidc::condition_gravity(raw + 3 * i, rule, Ig + 3 * i, Ia + 3 * i);
}
extern "C" void host_mkl_sinf_ha(const float* x, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = vidc::mkl_sinf_ha(x[i]);
}

// exterior-tile bitmap of the forward kernels (csrc/frame_params.cuh: tile_certainly_exterior), one byte per 32x32 tile
extern "C" void host_exterior_tiles(const vidc_camera* cam, const float* Ig, const float* Ia, int B, unsigned char* out) {
    const int tiles_x = (cam->W + 31) / 32, tiles_y = (cam->H + 31) / 32;
    for (int i = 0; i < B; ++i) {
        vidc_frame_params p;
        memset(&p, 0, sizeof p);
        vidc::frame_params_from_gravity(*cam, Ig + 3 * i, Ia + 3 * i, p);
        for (int t = 0; t < tiles_x * tiles_y; ++t)
            out[(size_t)i * tiles_x * tiles_y + t] = vidc::tile_certainly_exterior(p, *cam, t % tiles_x, t / tiles_x) ? 1 : 0;
    }
}

// round 2: the per-frame proof for the inverse warp's shared-reciprocal division, and the per-tile tables of the kernels
// (prefetch / staging hints) -- the same __host__ __device__ source the per-frame kernels compile
extern "C" void host_inv_division_proven(const vidc_camera* cam, const float* Ig, const float* Ia, int B, unsigned char* out,
                                         vidc_frame_params* params_out) {
    for (int i = 0; i < B; ++i) {
        vidc_frame_params p;
        memset(&p, 0, sizeof p);
        vidc::frame_params_from_gravity(*cam, Ig + 3 * i, Ia + 3 * i, p);
        out[i] = vidc::inv_division_proven(p, *cam) ? 1 : 0;
        if (params_out) params_out[i] = p;
    }
}
extern "C" void host_tile_tables(const vidc_camera* cam, const float* Ig, const float* Ia, int B, uint32_t* inv_boxes, uint32_t* fwd_boxes) {
    const int tiles_x = (cam->W + 31) / 32, tiles_y = (cam->H + 31) / 32, nt = tiles_x * tiles_y;
    for (int i = 0; i < B; ++i) {
        vidc_frame_params p;
        memset(&p, 0, sizeof p);
        vidc::frame_params_from_gravity(*cam, Ig + 3 * i, Ia + 3 * i, p);
        const bool proven = vidc::inv_division_proven(p, *cam);
        for (int t = 0; t < nt; ++t) {
            vidc::inv_tile_boxes(p, *cam, t % tiles_x, t / tiles_x, proven, inv_boxes + ((size_t)i * nt + t) * 4);
            vidc::fwd_tile_src_box(p, *cam, t % tiles_x, t / tiles_x, fwd_boxes + ((size_t)i * nt + t) * 4);
        }
    }
}

// ToTensor's x / 255 as the device kernels evaluate it (csrc/frame_params.cuh: u8_to_unit)
extern "C" void host_u8_to_unit(const unsigned char* in, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = vidc::u8_to_unit(in[i]);
}
