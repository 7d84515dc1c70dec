// Issue-rate microbenchmark: scalar FFMA vs packed fma.rn.f32x2, register vs shared/constant operands.
// Decides how the tap contractions of the DP step are written.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_taps[64];
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int iters) {
    __shared__ float st[64];
    if (threadIdx.x < 64) st[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = in[threadIdx.x + i];
    float x0 = in[threadIdx.x + 20], x1 = in[threadIdx.x + 21];
    float b[8];
    for (int i = 0; i < 8; ++i) b[i] = in[threadIdx.x + 30 + i];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // scalar FFMA, all register operands
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x0, x1);
        } else if (MODE == 1) {     // packed FFMA2
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, av, xv, yv;
                    asm("mov.b64 %0, {%1,%2};" : "=l"(av) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x0), "f"(x0));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(yv) : "f"(x1), "f"(x1));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(av), "l"(xv), "l"(yv));
                    asm("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
                }
        } else if (MODE == 2) {     // scalar FFMA with a constant-bank tap operand (uniform)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(x0, c_taps[(r * 16 + i) & 63], a[i]);
        } else if (MODE == 4 || MODE == 5) {     // packed FFMA2 interleaved with scalar FFMA (1 : 1 or 2 : 1 instructions): do the two share one pipe?
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, av, xv, yv;
                    asm("mov.b64 %0, {%1,%2};" : "=l"(av) : "f"(a[i]), "f"(a[i + 1]));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(xv) : "f"(x0), "f"(x0));
                    asm("mov.b64 %0, {%1,%2};" : "=l"(yv) : "f"(x1), "f"(x1));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(av), "l"(xv), "l"(yv));
                    asm("mov.b64 {%0,%1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
                    if (MODE == 4 || (i & 2) == 0) b[i >> 1] = fmaf(b[i >> 1], x0, x1);
                }
        } else if (MODE == 3) {     // scalar FFMA with a shared-memory broadcast tap operand
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(x0, st[(r * 16 + i) & 63], a[i]);
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    for (int i = 0; i < 8; ++i) s += b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, float *out, float *in) {
    int iters = 4096, grid = 148 * 8;
    k<MODE><<<grid, 256>>>(out, in, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)grid * 256 * iters * (MODE == 4 ? 96 : MODE == 5 ? 80 : 64);
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (2*FMA)  err=%s\n", name, ms, 2 * fma / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run<0>("FFMA reg,reg,reg", out, in);
    run<1>("FFMA2 (f32x2)", out, in);
    run<2>("FFMA reg,const,reg", out, in);
    run<3>("FFMA reg,smem-bcast,reg", out, in);
    run<4>("FFMA2 + FFMA 1:1", out, in);
    run<5>("FFMA2 + FFMA 2:1", out, in);
    return 0;
}
