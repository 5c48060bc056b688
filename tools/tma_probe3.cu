// Probe 3: where does cuTensorMapEncodeTiled come from, and what does it write?
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
typedef CUresult (*ENC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void dump(const char* name, const CUtensorMap& m) {
    printf("%s:", name);
    for (int i = 0; i < 16; ++i) printf(" %016llx", (unsigned long long)m.opaque[i]);
    printf("\n");
}
int main() {
    float* d; cudaMalloc(&d, 320 * 240 * 3 * 2 * 4);
    printf("device ptr %p\n", (void*)d);
    cuuint64_t dims[2] = {320, 240 * 6}; cuuint64_t str[1] = {1280};
    cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
    void* p1 = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p1, cudaEnableDefault, &q);
    printf("cudaGetDriverEntryPoint: err %d q %d ptr %p\n", (int)e, (int)q, p1);
    void* h = dlopen("libcuda.so.1", RTLD_NOW);
    void* p2 = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
    printf("dlsym: %p\n", p2);
    void* p3 = nullptr;
    typedef CUresult (*GPA)(const char*, void**, int, cuuint64_t, CUdriverProcAddressQueryResult*);
    GPA gpa = h ? (GPA)dlsym(h, "cuGetProcAddress_v2") : nullptr;
    CUdriverProcAddressQueryResult qq;
    if (gpa) { CUresult r = gpa("cuTensorMapEncodeTiled", &p3, 12090, 0, &qq); printf("cuGetProcAddress_v2(12090): r %d ptr %p\n", (int)r, p3); }
    void* ps[3] = {p1, p2, p3};
    const char* names[3] = {"runtime-entry", "dlsym", "getproc"};
    for (int k = 0; k < 3; ++k) {
        if (!ps[k]) continue;
        CUtensorMap m; memset(&m, 0xAB, sizeof m);
        CUresult r = ((ENC)ps[k])(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("%s -> %d\n", names[k], (int)r);
        dump(names[k], m);
    }
    return 0;
}
