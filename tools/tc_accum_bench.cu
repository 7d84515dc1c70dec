// Experiment: how does tcgen05.mma (kind::tf32, fp32 accumulator in TMEM) round when it adds a K = 8 product group to the accumulator?
// D starts at 1.0 (one MMA), then `n` MMAs each add t = +-1.75 ulp(1.0) (7 * 2^-26, exact in tf32):
//   round-to-nearest would give 1 + 2 n ulp, truncation 1 + n ulp (exact: 1 + 1.75 n ulp); negative t tells toward-zero from toward -inf.
// Second test: the 8 products of ONE MMA each carry 1.75 ulp (exact sum 14 ulp): are they summed exactly before the accumulate?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_accum_bench tools/tc_accum_bench.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
// core-matrix layout (K-major, no swizzle) of a [rows][8] matrix: element (r, k) at (r/8)*64 + (k/4)*32 + (r%8)*4 + (k%4)
__device__ __forceinline__ int cm(int r, int k) { return (r / 8) * 64 + (k / 4) * 32 + (r % 8) * 4 + (k % 4); }

__global__ void __launch_bounds__(128) k_acc(float *out, int test, int n, float t) {
    __shared__ __align__(1024) float A1[128 * 8], Aall[128 * 8], B1[32 * 8], Bt[32 * 8];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < 128 * 8; i += 128) { A1[i] = 0.f; Aall[i] = 0.f; }
    for (int i = tid; i < 32 * 8; i += 128) { B1[i] = 0.f; Bt[i] = 0.f; }
    __syncthreads();
    for (int r = tid; r < 128; r += 128) { A1[cm(r, 0)] = 1.f; for (int k = 0; k < 8; ++k) Aall[cm(r, k)] = 1.f; }
    if (tid < 32) { B1[cm(tid, 0)] = 1.f; for (int k = 0; k < 8; ++k) Bt[cm(tid, k)] = t; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        mma_tf32(tmem, make_desc(smem_u32(A1), 128, 256), make_desc(smem_u32(B1), 128, 256), idesc, 0);          // D = 1
        if (test == 0)
            for (int i = 0; i < n; ++i) mma_tf32(tmem, make_desc(smem_u32(A1), 128, 256), make_desc(smem_u32(Bt), 128, 256), idesc, 1);   // + t, n times
        else
            for (int i = 0; i < n; ++i) mma_tf32(tmem, make_desc(smem_u32(Aall), 128, 256), make_desc(smem_u32(Bt), 128, 256), idesc, 1); // + 8 t in one MMA
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tmem + ((uint32_t)(32 * wid) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (tid == 0) out[0] = __uint_as_float(v);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    float *d, h;
    cudaMalloc(&d, 4);
    const float ulp = ldexpf(1.f, -23);
    const float ts[4] = {7.f * ldexpf(1.f, -26), -7.f * ldexpf(1.f, -26), 5.f * ldexpf(1.f, -26), 3.f * ldexpf(1.f, -27)};   // 1.75, -1.75 (ulp of 0.5..1 is half), 1.25, 0.375 ulp
    for (int test = 0; test < 2; ++test)
        for (int ti = 0; ti < 4; ++ti)
            for (int n : {1, 16}) {
                k_acc<<<1, 128>>>(d, test, n, ts[ti]);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
                cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
                const double exact = 1.0 + (double)ts[ti] * n * (test ? 8 : 1);
                printf("%s: t = %+.3f ulp, n = %2d: D = 1 %+.3f ulp, exact 1 %+.3f ulp\n", test ? "8 products per MMA" : "1 product per MMA ", ts[ti] / ulp, n,
                       ((double)h - 1.0) / ulp, (exact - 1.0) / ulp);
            }
    return 0;
}
