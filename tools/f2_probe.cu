// Probe: do the packed fp32x2 intrinsics (FFMA2 / FADD2 / FMUL2) with broadcast / negated operands return, per half,
// exactly what the scalar IEEE operations return?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__global__ void k(const float* in, int n, unsigned* bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = in[6 * i], b = in[6 * i + 1], c = in[6 * i + 2], d = in[6 * i + 3], e = in[6 * i + 4], f = in[6 * i + 5];
    unsigned m = 0;
    { float2 r = __ffma2_rn(f2(a, b), bc(c), f2(d, e)); if (r.x != fmaf(a, c, d) || r.y != fmaf(b, c, e)) m |= 1; }
    { float2 r = __fadd2_rn(f2(a, b), f2(-c, -d)); if (r.x != a - c || r.y != b - d) m |= 2; }
    { float2 r = __fmul2_rn(f2(a, b), f2(e, f)); if (r.x != a * e || r.y != b * f) m |= 4; }
    { float2 q = __fmul2_rn(f2(a, b), bc(f)); float2 rem = __ffma2_rn(bc(-e), q, f2(a, b)); float2 r = __ffma2_rn(rem, bc(f), q);
      float q0 = a * f, r0 = fmaf(-e, q0, a), s0 = fmaf(r0, f, q0); float q1 = b * f, r1 = fmaf(-e, q1, b), s1 = fmaf(r1, f, q1);
      if (r.x != s0 || r.y != s1) m |= 8; }
    { float2 g = f2(a, b); float2 r = __fmul2_rn(__ffma2_rn(__fadd2_rn(g, bc(1.0f)), f2(640.f, 480.f), bc(-1.0f)), bc(0.5f));
      if (r.x != fmaf(a + 1.0f, 640.f, -1.0f) * 0.5f || r.y != fmaf(b + 1.0f, 480.f, -1.0f) * 0.5f) m |= 16; }
    { float2 z = __ffma2_rn(f2(a, b), bc(c), __ffma2_rn(f2(d, e), bc(f), __fmul2_rn(f2(b, a), bc(d))));
      if (z.x != fmaf(a, c, fmaf(d, f, b * d)) || z.y != fmaf(b, c, fmaf(e, f, a * d))) m |= 32; }
    if (m) atomicOr(bad, m);
}
int main() {
    const int n = 1 << 22;
    std::vector<float> h(6 * (size_t)n);
    srand(1);
    for (auto& v : h) v = (float)rand() / RAND_MAX * 4.f - 2.f;
    float* d; unsigned* bad; cudaMalloc(&d, h.size() * 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(d, n, bad);
    unsigned hb = 0; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("%s, mismatch mask 0x%x (bit per test: fma-bcast, sub, mul, div-steps, unnormalise, rotation)\n", cudaGetErrorString(cudaDeviceSynchronize()), hb);
    return 0;
}
