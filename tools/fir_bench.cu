// Microbenchmark of the FIR-like inner loop of dp_fast.cu in isolation: which formulation of the 2x2 complex MAC
// sustains the highest rate on B200?  Window in shared memory (conflict-free padded float4), taps as shared-memory
// broadcasts, 4 consecutive symbols per thread, 256 threads x 2 CTAs per SM, 148 x 2 CTAs.
//   V0: 4-mult, FFMA2 with a scalar-broadcast tap operand           (dp_fast.cu before the Gauss change)
//   V1: 3-mult (Gauss), FFMA2 with pair x pair operands             (dp_fast.cu now)
//   V2: 4-mult, scalar FFMA
//   V3: 3-mult (Gauss), scalar FFMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fir_bench fir_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NT = 256, R = 4, NLAG = 24, WLEN = NT * R + 64;

template <int V>
__global__ void __launch_bounds__(NT, 2) k(float *out, const float *in, int iters) {
    __shared__ float4 win[WLEN + WLEN / 4 + 4];
    __shared__ float4 taps[NLAG * 3];
    for (int i = threadIdx.x; i < WLEN + WLEN / 4 + 4; i += NT) win[i] = make_float4(in[i & 1023], in[(i + 1) & 1023], in[(i + 2) & 1023], in[(i + 3) & 1023]);
    for (int i = threadIdx.x; i < NLAG * 3; i += NT) taps[i] = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
    __syncthreads();
    float acc[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    const int i0 = R * threadIdx.x;
    auto LD = [&](int j) { return win[j + (j >> 2)]; };
    for (int it = 0; it < iters; ++it) {
        float4 w[R + 1];
        w[0] = LD(i0); w[1] = LD(i0 + 1); w[2] = LD(i0 + 2);
#pragma unroll 4
        for (int a = 0; a < NLAG; ++a) {
            w[3] = LD(i0 + a + 3);
            if (V == 0) {
                const float4 t0 = taps[2 * a], t1 = taps[2 * a + 1];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 x0 = make_float2(w[r].x, w[r].y), x1 = make_float2(w[r].z, w[r].w);
                    float2 *A = reinterpret_cast<float2 *>(acc[r]);
                    A[0] = __ffma2_rn(make_float2(t0.x, t0.x), x0, A[0]);
                    A[1] = __ffma2_rn(make_float2(t0.y, t0.y), x0, A[1]);
                    A[0] = __ffma2_rn(make_float2(t0.z, t0.z), x1, A[0]);
                    A[1] = __ffma2_rn(make_float2(t0.w, t0.w), x1, A[1]);
                    A[2] = __ffma2_rn(make_float2(t1.x, t1.x), x0, A[2]);
                    A[3] = __ffma2_rn(make_float2(t1.y, t1.y), x0, A[3]);
                    A[2] = __ffma2_rn(make_float2(t1.z, t1.z), x1, A[2]);
                    A[3] = __ffma2_rn(make_float2(t1.w, t1.w), x1, A[3]);
                }
            } else if (V == 1) {
                const float4 T0 = taps[3 * a], T1 = taps[3 * a + 1], T2 = taps[3 * a + 2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 x0 = make_float2(w[r].x, w[r].y), x1 = make_float2(w[r].z, w[r].w);
                    const float2 s = make_float2(w[r].x + w[r].y, w[r].z + w[r].w);       // (recomputed per use here: worst case)
                    float2 *A = reinterpret_cast<float2 *>(acc[r]);
                    A[0] = __ffma2_rn(make_float2(T0.x, T0.y), s, A[0]);
                    A[1] = __ffma2_rn(make_float2(T0.z, T0.w), s, A[1]);
                    A[2] = __ffma2_rn(x0, make_float2(T1.x, T1.y), A[2]);
                    A[2] = __ffma2_rn(x1, make_float2(T1.z, T1.w), A[2]);
                    A[3] = __ffma2_rn(x0, make_float2(T2.x, T2.y), A[3]);
                    A[3] = __ffma2_rn(x1, make_float2(T2.z, T2.w), A[3]);
                }
            } else if (V == 2) {
                const float4 t0 = taps[2 * a], t1 = taps[2 * a + 1];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 x = w[r];
                    acc[r][0] = fmaf(t0.x, x.x, acc[r][0]); acc[r][0] = fmaf(-t0.y, x.y, acc[r][0]);
                    acc[r][0] = fmaf(t0.z, x.z, acc[r][0]); acc[r][0] = fmaf(-t0.w, x.w, acc[r][0]);
                    acc[r][1] = fmaf(t0.x, x.y, acc[r][1]); acc[r][1] = fmaf(t0.y, x.x, acc[r][1]);
                    acc[r][1] = fmaf(t0.z, x.w, acc[r][1]); acc[r][1] = fmaf(t0.w, x.z, acc[r][1]);
                    acc[r][2] = fmaf(t1.x, x.x, acc[r][2]); acc[r][2] = fmaf(-t1.y, x.y, acc[r][2]);
                    acc[r][2] = fmaf(t1.z, x.z, acc[r][2]); acc[r][2] = fmaf(-t1.w, x.w, acc[r][2]);
                    acc[r][3] = fmaf(t1.x, x.y, acc[r][3]); acc[r][3] = fmaf(t1.y, x.x, acc[r][3]);
                    acc[r][3] = fmaf(t1.z, x.w, acc[r][3]); acc[r][3] = fmaf(t1.w, x.z, acc[r][3]);
                }
            } else {
                const float4 T0 = taps[3 * a], T1 = taps[3 * a + 1], T2 = taps[3 * a + 2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 x = w[r];
                    const float s0 = x.x + x.y, s1 = x.z + x.w;
                    acc[r][0] = fmaf(T0.x, s0, acc[r][0]); acc[r][0] = fmaf(T0.y, s1, acc[r][0]);
                    acc[r][1] = fmaf(T0.z, s0, acc[r][1]); acc[r][1] = fmaf(T0.w, s1, acc[r][1]);
                    acc[r][2] = fmaf(x.x, T1.x, acc[r][2]); acc[r][2] = fmaf(x.z, T1.z, acc[r][2]);
                    acc[r][3] = fmaf(x.y, T1.y, acc[r][3]); acc[r][3] = fmaf(x.w, T1.w, acc[r][3]);
                    acc[r][4] = fmaf(x.x, T2.x, acc[r][4]); acc[r][4] = fmaf(x.z, T2.z, acc[r][4]);
                    acc[r][5] = fmaf(x.y, T2.y, acc[r][5]); acc[r][5] = fmaf(x.w, T2.w, acc[r][5]);
                }
            }
            w[0] = w[1]; w[1] = w[2]; w[2] = w[3];
        }
    }
    float s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += acc[r][c];
    out[blockIdx.x * NT + threadIdx.x] = s;
}

template <int V>
void run(const char *name, float *out, float *in) {
    const int iters = 2000, grid = 148 * 2;
    k<V><<<grid, NT>>>(out, in, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<grid, NT>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // complex 2x2 MACs: per thread per lag step: 4 symbols x 2 outputs x 2 inputs = 16 complex MACs
    const double cmac = (double)grid * NT * iters * NLAG * 16;
    const double cyc_per_lagstep_per_warp = ms * 1e-3 * 1.965e9 / ((double)iters * NLAG) / 4.0;   // 16 warps per SM = 4 per SMSP
    printf("%-44s %8.3f ms  %7.2f G complex-MAC/s  %6.2f SMSP cycles per warp lag-step (at 1.965 GHz)  err=%s\n", name, ms,
           cmac / ms * 1e-6, cyc_per_lagstep_per_warp, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 2 * NT * 4); cudaMalloc(&in, 8192 * 4);
    cudaMemset(in, 0, 8192 * 4);
    run<0>("V0 4-mult FFMA2 (scalar-bcast tap)", out, in);
    run<1>("V1 3-mult FFMA2 (pair x pair)", out, in);
    run<2>("V2 4-mult scalar FFMA", out, in);
    run<3>("V3 3-mult scalar FFMA", out, in);
    return 0;
}
