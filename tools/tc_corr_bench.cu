// Experiment (outside north_star's "no tensor cores" design): the tap-gradient correlations as block outer
// products on tcgen05.  For 16-symbol blocks k,  D[(row, j), (comp, n)] = sum_k A_row[16 k + j] * W_comp[16 k + n]  is a plain GEMM
// whose operands are the SoA rows exactly as they lie in memory (MN-major, 64-byte rows, SWIZZLE_64B); the lag-b correlation is
// the sum of a diagonal of D.  This program validates the descriptor encodings on a B200 before the kernel is built around them:
//   test 0: A = 8 rows x 16 (M = 128), B = 8 groups x 16 (N = 128) from separate buffers (canonical), K = 32 blocks (4 MMAs)
//   test 1: B's second N half = the SAME buffers one k-row (64 B) further (window groups alias, two N = 64 MMAs)
//   test 2: A in SWIZZLE_128B (4 rows x 32 samples: interleaved rx phases), B as in test 1
// and reports the error against references with truncated / rounded / exact fp32 inputs (what does kind::tf32 do to fp32 bits?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_corr_bench tools/tc_corr_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                                  // version = 1 (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {   // tf32 x tf32 -> f32, both operands MN-major
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ uint32_t swz64(uint32_t off) { return off ^ (((off >> 7) & 3u) << 4); }
__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

constexpr int NB = 32;                 // 16-symbol blocks per tile (512 symbols)
constexpr int ROWB = NB * 64;          // bytes of one SW64 row buffer (2048)
// smem: A64 [8 rows][ROWB] | A128 [4 rows][2 ROWB] | W [4 comps][ROWB + 64] | W2 [4 comps][ROWB] (second halves as separate buffers, test 0)
__global__ void __launch_bounds__(128) k_test(const float *arows, const float *xrows, const float *wrows, float *out, int test) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *A64 = sm, *A128 = A64 + 8 * ROWB, *W = A128 + 4 * 2 * ROWB, *W2 = W + 4 * 4096;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, wid = tid >> 5;
    // arows: [8][512] floats; xrows: [4][1024]; wrows: [4][512 + 16]
    for (int i = tid; i < 8 * 512; i += 128) {
        const int r = i / 512, u = i % 512;
        *reinterpret_cast<float *>(A64 + r * ROWB + swz64(4 * u)) = arows[i];
    }
    for (int i = tid; i < 4 * 1024; i += 128) {
        const int r = i / 1024, s = i % 1024;
        *reinterpret_cast<float *>(A128 + r * 2 * ROWB + swz128(4 * s)) = xrows[i];
    }
    for (int i = tid; i < 4 * 528; i += 128) {
        const int r = i / 528, u = i % 528;
        *reinterpret_cast<float *>(W + r * 4096 + swz64(4 * u)) = wrows[i];          // 4096-byte comp stride (ROWB + one more row, 1024-aligned)
        if (u >= 16) *reinterpret_cast<float *>(W2 + r * ROWB + swz64(4 * (u - 16))) = wrows[i];
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
    if (test == 4) {                                         // TMEM store / load round trip
        const uint32_t taddr = tmem + ((uint32_t)(32 * wid) << 16);
        const uint32_t val = 1000u * tid;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + 5), "r"(val) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr + 5));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[tid] = (float)r;
        out[128 + tid] = (float)tmem;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
        return;
    }
    if (test == 5 || test == 6) {
        // MN-major tf32: the only legal layout is SWIZZLE_128B_BASE32B: rows of 128 B (32 elements along MN), 4 k-rows per atom (512 B),
        // 32-byte chunks XORed with the k-row index.  A[m = 32 r + i][k] = xrows[r][32 k + i]  (4 rows x 1024 floats, k-row pitch 128 B)
        // test 5: B = A-like from wrows viewed as [4][..] with pitch 32 floats; test 6: B's k-row pitch is 16 elements (W[k][n] = w[16 k + n], materialised)
        __syncthreads();
        unsigned char *MA = A128, *MB = A64;
        for (int i = tid; i < 4 * 1024; i += 128) {
            const int r = i / 1024, sidx = i % 1024;
            uint32_t off = 4 * sidx;
            off ^= ((off >> 7) & 3u) << 5;
            *reinterpret_cast<float *>(MA + r * 4096 + off) = xrows[i];
        }
        for (int i = tid; i < 4 * 16 * 32; i += 128) {       // 16 k-rows x 32 columns per comp
            const int r = i / 512, k = (i % 512) / 32, n = i % 32;
            uint32_t off = 128 * k + 4 * n;
            off ^= ((off >> 7) & 3u) << 5;
            *reinterpret_cast<float *>(MB + r * 2048 + off) = test == 5 ? wrows[r * 528 + 32 * k + n] : wrows[r * 528 + 16 * k + n];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            for (int kk = 0; kk < 2; ++kk) {                 // K = 16 k-rows = 2 MMAs of 8
                const uint64_t da = make_desc(smem_u32(MA) + kk * 1024, 4096, 512, 1), db = make_desc(smem_u32(MB) + kk * 1024, 2048, 512, 1);
                mma_tf32(tmem, da, db, make_idesc(128, 128), kk > 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
    } else
    if (test == 3) {                                         // K-major, no swizzle: core matrices of 8 rows x 16 B; A = arows as [m = 128][k = 8]: a[m/16][m%16 + 16 kk]... simple: A[m][k] = arows[m * 8 + k]
        __syncthreads();
        float *A0 = reinterpret_cast<float *>(A64), *B0 = reinterpret_cast<float *>(A128);
        for (int i = tid; i < 128 * 8; i += 128) {
            const int m = i / 8, k = i % 8;
            // core matrix (m/8, k/4): 128 B; within: row m%8 at 16 B, element k%4.  m-groups at SBO = 256 B (two k core matrices side by side), k-groups at LBO = 128 B
            A0[(m / 8) * 64 + (k / 4) * 32 + (m % 8) * 4 + (k % 4)] = arows[i];
            B0[(m / 8) * 64 + (k / 4) * 32 + (m % 8) * 4 + (k % 4)] = wrows[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint64_t da = make_desc(smem_u32(A0), 128, 256, 0), db = make_desc(smem_u32(B0), 128, 256, 0);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            mma_tf32(tmem, da, db, idesc, 0);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
    } else
    if (tid == 0) {
        for (int kk = 0; kk < NB / 8; ++kk) {
            const uint32_t acc = kk > 0;
            if (test == 0) {
                const uint64_t da = make_desc(smem_u32(A64) + kk * 512, ROWB, 512, 4);
                // B groups: n/16 = 2 comp + half; half 0 in W (stride 4096), half 1 in W2 (stride ROWB): not one uniform LBO -> two MMAs as well
                const uint64_t db0 = make_desc(smem_u32(W) + kk * 512, 4096, 512, 4);
                const uint64_t db1 = make_desc(smem_u32(W2) + kk * 512, ROWB, 512, 4);
                mma_tf32(tmem, da, db0, make_idesc(128, 64), acc);
                mma_tf32(tmem + 64, da, db1, make_idesc(128, 64), acc);
            } else {
                const uint64_t da = test == 1 ? make_desc(smem_u32(A64) + kk * 512, ROWB, 512, 4) : make_desc(smem_u32(A128) + kk * 1024, 2 * ROWB, 1024, 2);
                const uint64_t db0 = make_desc(smem_u32(W) + kk * 512, 4096, 512, 4);
                const uint64_t db1 = make_desc(smem_u32(W) + kk * 512 + 64, 4096, 512, 4);      // one k-row further: window positions 16..31
                mma_tf32(tmem, da, db0, make_idesc(128, 64), acc);
                mma_tf32(tmem + 64, da, db1, make_idesc(128, 64), acc);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(&bar)),
            "r"(0)
            : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * wid) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static float trunc_tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xffffe000u; memcpy(&x, &b, 4); return x; }
static float round_tf32(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&x, &b, 4); return x; }

int main() {
    std::vector<float> a(8 * 512), x(4 * 1024), w(4 * 528), out(128 * 128);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &v : a) v = rnd();
    for (auto &v : x) v = rnd();
    for (auto &v : w) v = rnd();
    float *da, *dx, *dw, *dout;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&dout, out.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 8 * ROWB + 4 * 2 * ROWB + 4 * 4096 + 4 * ROWB + 1024;
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int test = 0; test < 7; ++test) {
        cudaMemset(dout, 0, out.size() * 4);
        k_test<<<1, 128, smem>>>(da, dx, dw, dout, test);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d: CUDA error %s\n", test, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        if (test == 4) { printf("test 4 (tmem st/ld): out[1] %g out[77] %g (expect 1000, 77000), tmem base %g\n", out[1], out[77], out[128]); continue; }
        if (test >= 5) {
            double e[3] = {0, 0, 0};
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
                double r[3] = {0, 0, 0};
                for (int k = 0; k < 16; ++k) {
                    const float av = x[(m / 32) * 1024 + 32 * k + m % 32], wv = w[(n / 32) * 528 + (test == 5 ? 32 : 16) * k + n % 32];
                    r[0] += (double)trunc_tf32(av) * trunc_tf32(wv); r[1] += (double)round_tf32(av) * round_tf32(wv); r[2] += (double)av * wv;
                }
                for (int t = 0; t < 3; ++t) e[t] = fmax(e[t], fabs(out[m * 128 + n] - r[t]));
            }
            printf("test %d (MN-major 128B_BASE32B): max err vs truncated %.3e rounded %.3e exact %.3e, D[0][0..3] = %g %g %g %g\n", test, e[0], e[1], e[2], out[0], out[1], out[2], out[3]);
            continue;
        }
        if (test == 3) {
            double e = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { double r = 0; for (int k = 0; k < 8; ++k) r += (double)trunc_tf32(a[m * 8 + k]) * trunc_tf32(w[n * 8 + k]); e = fmax(e, fabs(out[m * 128 + n] - r)); }
            printf("test 3 (K-major, no swizzle): max err vs truncated %.3e, D[0][0..3] = %g %g %g %g\n", e, out[0], out[1], out[2], out[3]);
            continue;
        }
        // reference: D[m][n] = sum_k A[m][k] W[n][k];  tests 0/1: m = 16 r + j -> a[r][16 k + j];  test 2: m = 32 r + s -> x[r][32 k + s]
        //            n = 64 half + 16 comp + j' -> w[comp][16 k + 16 half + j']
        double err[3] = {0, 0, 0}, mag = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                const int half = n / 64, comp = (n % 64) / 16, jp = n % 16;
                double s[3] = {0, 0, 0};
                for (int k = 0; k < NB; ++k) {
                    const float av = test < 2 ? a[(m / 16) * 512 + 16 * k + m % 16] : x[(m / 32) * 1024 + 32 * k + m % 32];
                    const float wv = w[comp * 528 + 16 * k + 16 * half + jp];
                    s[0] += (double)trunc_tf32(av) * trunc_tf32(wv);
                    s[1] += (double)round_tf32(av) * round_tf32(wv);
                    s[2] += (double)av * wv;
                }
                for (int t = 0; t < 3; ++t) err[t] = fmax(err[t], fabs(out[m * 128 + n] - s[t]));
                mag = fmax(mag, fabs(s[2]));
            }
        printf("test %d: max |D - ref|: truncated inputs %.3e, rounded inputs %.3e, exact inputs %.3e   (max |D| %.3f)  D[0][0..3] = %g %g %g %g\n", test,
               err[0], err[1], err[2], mag, out[0], out[1], out[2], out[3]);
    }
    return 0;
}
