// FP32 multiply-add issue rate of one SM sub-partition (development probe):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma_rate tools/probe/ffma_rate.cu && /tmp/ffma_rate
// Prints warp instructions per clock and SM for FFMA (3 register operands), FFMA2 (fma.rn.f32x2) and
// an FFMA + integer mix, for 2, 4 and 8 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b, long long* clk) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
    unsigned int u[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = __fmaf_rn(x[i], a, x[(i + 5) & 15]);
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long xa, xb, aa, ra;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(xa) : "f"(x[i]), "f"(x[i + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(xb) : "f"(x[(i + 6) & 15]), "f"(x[(i + 7) & 15]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a), "f"(b));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ra) : "l"(xa), "l"(aa), "l"(xb));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(ra));
                }
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = __fmaf_rn(x[i], a, x[(i + 5) & 15]);
#pragma unroll
                for (int i = 0; i < 4; i++) u[i] = (u[i] + u[(i + 1) & 3]) ^ (u[(i + 2) & 3] >> 3);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u[0] + u[1] + u[2] + u[3]);
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 3; mode++)
        for (int threads : {256, 512, 1024}) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, threads>>>(out, iters, 1.0001f, 0.9999f, clk);
                if (mode == 1) k<1><<<148, threads>>>(out, iters, 1.0001f, 0.9999f, clk);
                if (mode == 2) k<2><<<148, threads>>>(out, iters, 1.0001f, 0.9999f, clk);
            }
            long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            const double per_warp = mode == 1 ? 32.0 * iters : (mode == 0 ? 64.0 * iters : 64.0 * iters);
            const double warps = threads / 32.0;
            printf("mode %d (%s) threads %4d: %lld clocks, %.2f fp warp-instr/clk/SM (%.2f per scheduler)%s\n", mode,
                   mode == 0 ? "FFMA" : mode == 1 ? "FFMA2" : "FFMA+int", threads, c, per_warp * warps / c, per_warp * warps / c / 4,
                   mode == 1 ? "  [each = 2 fma per lane]" : mode == 2 ? "  [+ 12 int ops per 64 fma]" : "");
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
