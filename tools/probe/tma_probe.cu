// tma_probe.cu -- development probe: which forms of a 3-D tiled TMA load work on this device
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __noinline__ void tma3_noinline(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    tma3(dst, tm, c0, c1, c2, bar);
}

struct args {
    CUtensorMap tm;
    const CUtensorMap* tmg;
    float* out;
    int mode, box, c0, c1;
};

__global__ void probe(const __grid_constant__ args a) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(a.box * 4) : "memory");
        if (a.mode == 0) tma3(s32(sm), &a.tm, a.c0, a.c1, 0, s32(&bar));
        else if (a.mode == 1) tma3_noinline(s32(sm), &a.tm, a.c0, a.c1, 0, s32(&bar));
        else if (a.mode == 2) tma3(s32(sm), a.tmg, a.c0, a.c1, 0, s32(&bar));
        else tma3_noinline(s32(sm), a.tmg, a.c0, a.c1, 0, s32(&bar));
    }
    uint32_t ok = 0, tries = 0;
    while (!ok && ++tries < (1u << 20))
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < a.box; i += blockDim.x) a.out[i] = ok ? sm[i] : -777.0f;
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int n = 512, members = 2;
    std::vector<float> h((size_t)n * n * members);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003);
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 1024 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    printf("encode fn %p q %d\n", fn, (int)q);
    for (int box : {256, 32}) {
        args a;
        cuuint64_t dims[3] = {n, n, members};
        cuuint64_t str[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
        cuuint32_t bx[3] = {(cuuint32_t)box, 1, 1}, es[3] = {1, 1, 1};
        CUresult rc = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %d encode rc %d\n", box, (int)rc);
        CUtensorMap* tmg;
        cudaMalloc(&tmg, 128);
        cudaMemcpy(tmg, &a.tm, 128, cudaMemcpyHostToDevice);
        a.tmg = tmg; a.out = out; a.box = box;
        for (int mode = 0; mode < 4; mode++) {
            if (only >= 0 && mode != only) continue;
            for (int c0 : {0, 5, -3}) {
                a.mode = mode; a.c0 = c0; a.c1 = 7;
                probe<<<1, 128, 2048>>>(a);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> r(box);
                if (e == cudaSuccess) cudaMemcpy(r.data(), out, box * 4, cudaMemcpyDeviceToHost);
                bool good = e == cudaSuccess;
                for (int i = 0; good && i < box; i++) {
                    const int y = c0 + i;
                    const float want = (y < 0 || y >= n) ? 0.0f : h[(size_t)7 * n + y];
                    good = r[i] == want;
                }
                printf("box %3d mode %d (%s, %s) c0 %2d: %s %s\n", box, mode, mode < 2 ? "param" : "global",
                       (mode & 1) ? "noinline" : "inline", c0, cudaGetErrorString(e), good ? "OK" : "MISMATCH");
                if (e != cudaSuccess) return 1;
            }
        }
    }
    return 0;
}
