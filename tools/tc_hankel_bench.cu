// Experiment: can tcgen05 read a HANKEL operand -- overlapping rows of one flat shared-memory array -- through a K-major swizzled
// descriptor?  This is what a sliding-window contraction (the butterfly FIR, the channel convolution) needs to run on the tensor
// cores without materialising an im2col matrix:
//     A[m][k] = X[(RB/4) m + k]          (row pitch RB bytes = one swizzle row, K runs on into the following rows)
// The array is written with the swizzle as a function of the ABSOLUTE byte address (16-byte chunk index XOR address bits 7..),
// K-step s starts the descriptor 32 s bytes further, i.e. after RB/32 steps on the NEXT row.  Whether the hardware swizzles by
// absolute address (then every shift works) or by row counter + base_offset (then only whole 128-byte shifts can work) is not
// documented here, so all variants are measured: RB in {32, 64, 128} x base_offset in {0, (start >> 7) & 7}.
// B is K-major without swizzle (validated in tc_corr_bench.cu test 3).  Second part: cycles per MMA for N = 32 / 64 / 128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_hankel_bench tools/tc_hankel_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_k(int M, int N) {   // tf32 x tf32 -> f32, both operands K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

constexpr int NX = 8192;               // floats of the flat array (32 KB)
constexpr int KT = 96;                 // K extent of the test (12 MMA steps): crosses 12 / 6 / 3 rows for RB = 32 / 64 / 128
constexpr int NN = 32;                 // N of the validation

// out: [128][NN] results; timing: out_t[3] cycles per MMA for N = 32, 64, 128 (A swizzle RB)
__global__ void __launch_bounds__(128) k_hankel(const float *x, const float *b, float *out, long long *out_t, int RB, int bo_mode, int timing) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *X = sm;                                   // 32 KB
    float *Bm = reinterpret_cast<float *>(sm + NX * 4);      // [KT/8 steps][NN/8][2][8][4] = KT * NN floats (timing: up to N = 128 -> 8 * 128 floats per step reused)
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, wid = tid >> 5;
    const uint32_t mask = RB == 128 ? 7u : RB == 64 ? 3u : 1u;
    const uint32_t ltype = RB == 128 ? 2u : RB == 64 ? 4u : 6u;
    for (int i = tid; i < NX; i += 128) {
        uint32_t L = 4 * i;
        L ^= ((L >> 7) & mask) << 4;
        *reinterpret_cast<float *>(X + L) = x[i];
    }
    for (int i = tid; i < KT * 128; i += 128) {              // b: [128][KT] row-major (n, k)
        const int n = i / KT, k = i % KT, s = k / 8, kk = k % 8;
        // step tile of N rows: core matrix (n/8, kk/4) of 128 B: n-groups at SBO = 256 B, k halves at LBO = 128 B
        Bm[s * (128 * 8) + (n / 8) * 64 + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4)] = b[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    uint32_t parity = 0;
    if (!timing) {
        if (tid == 0) {
            for (int s = 0; s < KT / 8; ++s) {
                const uint32_t start = smem_u32(X) + 32 * s;
                const uint64_t da = make_desc(start, 16, 8 * RB, ltype, bo_mode ? (start >> 7) & 7 : 0);
                const uint64_t db = make_desc(smem_u32(Bm) + s * (128 * 8 * 4), 128, 256, 0, 0);
                mma_tf32(tmem, da, db, make_idesc_k(128, NN), s > 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(&bar, parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * wid) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)tid * NN + j] = __uint_as_float(v[j]);
    } else {
        // cycles per MMA: `reps` x 12 back-to-back K steps into the same accumulator, for N = 32, 64, 128
        for (int t = 0; t < 3; ++t) {
            const int N = 32 << t, reps = 16;
            long long c0 = 0;
            if (tid == 0) {
                c0 = clock64();
                for (int r = 0; r < reps; ++r)
                    for (int s = 0; s < KT / 8; ++s) {
                        const uint32_t start = smem_u32(X) + 32 * s;
                        const uint64_t da = make_desc(start, 16, 8 * RB, ltype, bo_mode ? (start >> 7) & 7 : 0);
                        const uint64_t db = make_desc(smem_u32(Bm) + s * (128 * 8 * 4), 128, 256, 0, 0);
                        mma_tf32(tmem, da, db, make_idesc_k(128, N), 1);
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            mbar_wait(&bar, parity);
            parity ^= 1;
            if (tid == 0) out_t[t] = (clock64() - c0) / (reps * (KT / 8));
            __syncthreads();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static float trunc_tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xffffe000u; memcpy(&x, &b, 4); return x; }

int main() {
    std::vector<float> x(NX), b(128 * KT), out(128 * NN);
    srand(2);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &v : x) v = rnd();
    for (auto &v : b) v = rnd();
    float *dx, *db, *dout;
    long long *dt, ht[3];
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, out.size() * 4); cudaMalloc(&dt, 3 * sizeof(long long));
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = NX * 4 + (KT / 8) * 128 * 8 * 4 + 1024;
    cudaFuncSetAttribute(k_hankel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int rbs[3] = {32, 64, 128};
    for (int ri = 0; ri < 3; ++ri)
        for (int bo = 0; bo < 2; ++bo) {
            const int RB = rbs[ri];
            cudaMemset(dout, 0, out.size() * 4);
            k_hankel<<<1, 128, smem>>>(dx, db, dout, dt, RB, bo, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("RB %d bo %d: CUDA error %s\n", RB, bo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            // error of the full K = 96 product and of prefixes (which K step breaks first?)
            double err = 0;
            int bad_rows = 0;
            for (int m = 0; m < 128; ++m) {
                double rowerr = 0;
                for (int n = 0; n < NN; ++n) {
                    double r = 0;
                    for (int k = 0; k < KT; ++k) r += (double)trunc_tf32(x[(RB / 4) * m + k]) * trunc_tf32(b[n * KT + k]);
                    rowerr = fmax(rowerr, fabs(out[m * NN + n] - r));
                }
                err = fmax(err, rowerr);
                bad_rows += rowerr > 1e-4;
            }
            printf("hankel RB %3d base_offset %s: max |D - ref(truncated)| %.3e, rows off by > 1e-4: %d / 128\n", RB, bo ? "(start>>7)&7" : "0", err, bad_rows);
        }
    for (int ri = 0; ri < 3; ++ri) {
        k_hankel<<<1, 128, smem>>>(dx, db, dout, dt, rbs[ri], 1, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("timing RB %d: CUDA error %s\n", rbs[ri], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ht, dt, sizeof(ht), cudaMemcpyDeviceToHost);
        printf("timing RB %3d: cycles per 128 x N x 8 tf32 MMA (A, B from shared memory): N = 32: %lld, N = 64: %lld, N = 128: %lld\n", rbs[ri], ht[0], ht[1], ht[2]);
    }
    return 0;
}
