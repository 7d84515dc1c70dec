// Evaluation kernels: shift search and SER estimators (reference: optical_DP_channel/shared_funcs.py:188-338).
// All are single-pass HBM scans (q: 128 B/symbol at 64-QAM, out: 16 B, tx: 8 B as float16) with integer
// error COUNTS returned next to the float SER so that "bit-exact SER counts" is checkable.
#include "common.cuh"

namespace vaeq {

constexpr int EV_NT = 256;
constexpr int EV_CHUNKS = 32;          // partial sums per shift (find_shift scratch = n_shift*8*EV_CHUNKS doubles)

__device__ __forceinline__ float tx_level(uint16_t bits, float scale) {
    // round(scale*tx.float()+scale), two fp32 roundings then round-half-even like torch.round (sf:198, sf:239)
    return rintf(__fadd_rn(__fmul_rn(scale, half_bits_to_float(bits)), scale));
}

// ---------------------------------------------------------------------------------------------
// find_shift / find_shift_symb_full
// ---------------------------------------------------------------------------------------------
// grid (n_shift, EV_CHUNKS): block (i, c) accumulates, over its slice of t, the 8 sums
//   S[comp][b][a] = sum_t tx[a][comp][t] * E[b][(t - (i - half)) mod N]          (sf:300-304, circular roll)
template <bool FROM_Q>
__global__ void __launch_bounds__(EV_NT) k_shift_corr(const float *q, int64_t ld_q, const float *out, int64_t ld_out,
                                                      const uint16_t *tx, int64_t ld_tx, const float *amp, int n_lev,
                                                      int N, int n_shift, double *part) {
    __shared__ double red[8 * 32];
    const int i = blockIdx.x, chunk = blockIdx.y, half = n_shift / 2;
    const int64_t per = ((int64_t)N + gridDim.y - 1) / gridDim.y;
    const int64_t t_lo = chunk * per, t_hi = min((int64_t)N, t_lo + per);
    float a_l[VAEQ_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < VAEQ_MAX_LEVELS; ++l) a_l[l] = (FROM_Q && l < n_lev) ? amp[l] : 0.f;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t t = t_lo + threadIdx.x; t < t_hi; t += EV_NT) {
        int64_t src = (t - (i - half)) % N;
        if (src < 0) src += N;
        float E[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (FROM_Q) {
                float e = 0.f;                                            // E_q[x_I] = sum_l a_l q_I[l]  (sf:297)
                for (int l = 0; l < n_lev; ++l) e += a_l[l] * q[(int64_t)(b * 2 * n_lev + l) * ld_q + src];
                E[b] = e;
            } else {
                E[b] = out[(int64_t)(b * 2) * ld_out + src];             // rx[:,0,:]  (sf:321)
            }
        }
#pragma unroll
        for (int comp = 0; comp < 2; ++comp)
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const float x = half_bits_to_float(tx[(int64_t)(a * 2 + comp) * ld_tx + t]);
                acc[comp * 4 + 0 * 2 + a] += (double)(x * E[0]);
                acc[comp * 4 + 1 * 2 + a] += (double)(x * E[1]);
            }
    }
    block_sum<8>(acc, red);
    if (threadIdx.x == 0) {
        double *dst = part + ((int64_t)i * gridDim.y + chunk) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = acc[k];
    }
}

// one thread: reduce chunks, then torch.max / argmax logic of sf:303-314 (first index wins ties)
__global__ void k_shift_decide(const double *part, int n_shift, int chunks, float *corr_out, int16_t *shift_out, int *r_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int half = n_shift / 2;
    float cmax[8];
    int cind[8];
    for (int k = 0; k < 8; ++k) {                            // k = comp*4 + b*2 + a
        cmax[k] = -1.f;
        cind[k] = 0;
        for (int i = 0; i < n_shift; ++i) {
            double s = 0.0;
            for (int c = 0; c < chunks; ++c) s += part[((int64_t)i * chunks + c) * 8 + k];
            const float v = fabsf((float)s);
            if (corr_out) corr_out[k * n_shift + i] = v;
            if (v > cmax[k]) {
                cmax[k] = v;
                cind[k] = i;
            }
        }
    }
    float best[4];
    int which[4];
    for (int ba = 0; ba < 4; ++ba) {                         // max over the tx component, first wins
        which[ba] = (cmax[4 + ba] > cmax[ba]) ? 1 : 0;
        best[ba] = which[ba] ? cmax[4 + ba] : cmax[ba];
    }
    // ba index = b*2 + a
    const bool straight = (best[0] + best[3]) >= (best[1] + best[2]);
    if (straight) {
        shift_out[0] = (int16_t)(half - cind[which[0] * 4 + 0]);
        shift_out[1] = (int16_t)(half - cind[which[3] * 4 + 3]);
        *r_out = 0;
    } else {
        shift_out[0] = (int16_t)(half - cind[which[1] * 4 + 1]);
        shift_out[1] = (int16_t)(half - cind[which[2] * 4 + 2]);
        *r_out = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// SER from posteriors: hard decision argmax(q), 4 rotations x IQ flip (sf:188-222)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void count_hypotheses(int (&cnt)[8], float DI, float DQ, float S, const float (&hI)[4],
                                                 const float (&hQ)[4]) {
    const float DQf = S - DQ;                                 // IQ-flipped reference data (sf:199)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        cnt[r] += (DI != hI[r]) || (DQ != hQ[r]);
        cnt[4 + r] += (DI != hI[r]) || (DQf != hQ[r]);
    }
}

template <int NL>
__global__ void __launch_bounds__(EV_NT) k_ser_iqflip(const float *q, int64_t ld_q, const uint16_t *tx, int64_t ld_tx,
                                                      int N, int *counts) {
    __shared__ int red[16 * 32];
    const float S = (float)(NL - 1), scale = (float)((NL - 1) / 2.0);
    int cnt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt[k] = 0;
    for (int64_t t = (int64_t)blockIdx.x * EV_NT + threadIdx.x; t < N; t += (int64_t)gridDim.x * EV_NT) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            int d[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {                     // torch.argmax: first maximal index
                float best = q[(int64_t)(p * 2 * NL + c * NL) * ld_q + t];
                int bi = 0;
#pragma unroll
                for (int l = 1; l < NL; ++l) {
                    const float v = q[(int64_t)(p * 2 * NL + c * NL + l) * ld_q + t];
                    if (v > best) {
                        best = v;
                        bi = l;
                    }
                }
                d[c] = bi;
            }
            const float dI = (float)d[0], dQ = (float)d[1];
            const float hI[4] = {dI, S - dI, S - dQ, dQ};     // 0, pi, pi/2, 3pi/2  (sf:201-219)
            const float hQ[4] = {dQ, S - dQ, dI, S - dI};
            const float DI = tx_level(tx[(int64_t)(p * 2 + 0) * ld_tx + t], scale);
            const float DQ = tx_level(tx[(int64_t)(p * 2 + 1) * ld_tx + t], scale);
            int c8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            count_hypotheses(c8, DI, DQ, S, hI, hQ);
#pragma unroll
            for (int r = 0; r < 4; ++r) {                     // counts[flip][pol][rot]
                cnt[(0 * 2 + p) * 4 + r] += c8[r];
                cnt[(1 * 2 + p) * 4 + r] += c8[4 + r];
            }
        }
    }
    // block reduction of 16 integer counters, then one atomic per counter (integer: order independent)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int v = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k * 32 + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        int v = 0;
        for (int w = 0; w < EV_NT / 32; ++w) v += red[threadIdx.x * 32 + w];
        if (v) atomicAdd(&counts[threadIdx.x], v);
    }
}

__global__ void k_ser_min(const int *counts, int N, float *ser_out) {
    if (threadIdx.x < 2) {
        const int p = threadIdx.x;
        float best = 3.0e38f;
        for (int f = 0; f < 2; ++f)
            for (int r = 0; r < 4; ++r) best = fminf(best, __fdiv_rn((float)counts[(f * 2 + p) * 4 + r], (float)N));
        ser_out[p] = best;                                    // torch.amin over (flip, rotation)  (sf:221, sf:264)
    }
}

// ---------------------------------------------------------------------------------------------
// SER from the constellation with PCS-aware thresholds (sf:225-287)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EV_NT) k_constell_norms(const float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx,
                                                          int N, double *sums) {
    __shared__ double red[2 * 32];
    double acc[2] = {0.0, 0.0};
    for (int64_t t = (int64_t)blockIdx.x * EV_NT + threadIdx.x; t < N; t += (int64_t)gridDim.x * EV_NT) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const float a = half_bits_to_float(tx[(int64_t)(p * 2) * ld_tx + t]), b = half_bits_to_float(tx[(int64_t)(p * 2 + 1) * ld_tx + t]);
            const float x = rx[(int64_t)(p * 2) * ld_rx + t], y = rx[(int64_t)(p * 2 + 1) * ld_rx + t];
            acc[0] += (double)sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
            acc[1] += (double)sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
        }
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], acc[0]);
        atomicAdd(&sums[1], acc[1]);
    }
}

template <int NL>
__global__ void __launch_bounds__(EV_NT) k_ser_constell(float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx,
                                                        const float *amp, const float *var, float nu_sc, int N,
                                                        const double *sums, int *counts) {
    __shared__ int red[16 * 32];
    __shared__ float lo[NL], hi[NL];
    if (threadIdx.x < NL) {
        const int l = threadIdx.x;
        // d = (1 + 2 nu_sc var[0]) (a_l + a_{l+1}) / 2, padded with -inf / +inf  (sf:234-236)
        const float f = __fadd_rn(1.f, __fmul_rn(2.f * nu_sc, var[0]));
        lo[l] = (l == 0) ? -INFINITY : __fmul_rn(f, __fadd_rn(amp[l - 1], amp[l])) * 0.5f;
        hi[l] = (l == NL - 1) ? INFINITY : __fmul_rn(f, __fadd_rn(amp[l], amp[l + 1])) * 0.5f;
    }
    __syncthreads();
    const double cnt2 = 2.0 * (double)N;
    const float g = __fdiv_rn((float)(sums[0] / cnt2), (float)(sums[1] / cnt2));      // sf:242
    const float S = (float)(NL - 1), scale = (float)((NL - 1) / 2.0);
    int cnt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt[k] = 0;
    for (int64_t t = (int64_t)blockIdx.x * EV_NT + threadIdx.x; t < N; t += (int64_t)gridDim.x * EV_NT) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float *pI = rx + (int64_t)(p * 2) * ld_rx + t, *pQ = rx + (int64_t)(p * 2 + 1) * ld_rx + t;
            const float yI = __fmul_rn(*pI, g), yQ = __fmul_rn(*pQ, g);
            *pI = yI;                                         // in-place rescale, visible to the caller (sf:242)
            *pQ = yQ;
            const float DI = tx_level(tx[(int64_t)(p * 2) * ld_tx + t], scale), DQ = tx_level(tx[(int64_t)(p * 2 + 1) * ld_tx + t], scale);
            const float DQf = S - DQ;
            const int iI = min(max((int)DI, 0), NL - 1), iQ = min(max((int)DQ, 0), NL - 1), iQf = min(max((int)DQf, 0), NL - 1);
            const float hI[4] = {yI, -yI, -yQ, yQ};           // sf:245-262
            const float hQ[4] = {yQ, -yQ, yI, -yI};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool okI = (lo[iI] <= hI[r]) && (hI[r] < hi[iI]);
                const bool okQ = (lo[iQ] <= hQ[r]) && (hQ[r] < hi[iQ]);
                const bool okQf = (lo[iQf] <= hQ[r]) && (hQ[r] < hi[iQf]);
                cnt[(0 * 2 + p) * 4 + r] += !(okI && okQ);
                cnt[(1 * 2 + p) * 4 + r] += !(okI && okQf);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int v = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k * 32 + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        int v = 0;
        for (int w = 0; w < EV_NT / 32; ++w) v += red[threadIdx.x * 32 + w];
        if (v) atomicAdd(&counts[threadIdx.x], v);
    }
}

// ---------------------------------------------------------------------------------------------
// extension (not in the reference): GMI-style achievable rate from the posteriors
// ---------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(EV_NT) k_gmi(const float *q, int64_t ld_q, const uint16_t *tx, int64_t ld_tx, int N, double *sums) {
    __shared__ double red[2 * 32];
    const float scale = (float)((NL - 1) / 2.0);
    double acc[2] = {0.0, 0.0};
    for (int64_t t = (int64_t)blockIdx.x * EV_NT + threadIdx.x; t < N; t += (int64_t)gridDim.x * EV_NT) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int iI = min(max((int)tx_level(tx[(int64_t)(p * 2) * ld_tx + t], scale), 0), NL - 1);
            const int iQ = min(max((int)tx_level(tx[(int64_t)(p * 2 + 1) * ld_tx + t], scale), 0), NL - 1);
            const float qi = q[(int64_t)(p * 2 * NL + iI) * ld_q + t], qq = q[(int64_t)(p * 2 * NL + NL + iQ) * ld_q + t];
            acc[p] += log2((double)fmaxf(qi, 1e-30f)) + log2((double)fmaxf(qq, 1e-30f));
        }
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[0], acc[0]);
        atomicAdd(&sums[1], acc[1]);
    }
}
__global__ void k_gmi_fin(const double *sums, const float *P, int n_lev, int N, float *gmi_out) {
    if (threadIdx.x < 2) {
        double H = 0.0;
        for (int l = 0; l < n_lev; ++l) H -= 2.0 * (double)P[l] * log2((double)P[l]);
        gmi_out[threadIdx.x] = (float)(H + sums[threadIdx.x] / (double)N);
    }
}

static int ev_grid(int N) { return max(1, min((N + EV_NT - 1) / EV_NT, sm_count() * 8)); }

}  // namespace vaeq

using namespace vaeq;

extern "C" int vaeq_find_shift(const float *q, int64_t ld_q, const float *out, int64_t ld_out, const uint16_t *tx,
                               int64_t ld_tx, const float *amp, int32_t n_lev, int32_t N, int32_t n_shift, float *corr_out,
                               int16_t *shift_out, int32_t *r_out, void *scratch, void *stream) {
    VAEQ_CHECK_ARG((q != nullptr) != (out != nullptr), "exactly one of q / out must be given");
    VAEQ_CHECK_ARG(tx && shift_out && r_out && scratch && N > 0 && n_shift > 0 && n_shift <= 64, "bad find_shift arguments");
    VAEQ_CHECK_ARG(!q || (amp && n_lev >= 1 && n_lev <= VAEQ_MAX_LEVELS), "bad amp / n_lev");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (int)min((int64_t)EV_CHUNKS, ((int64_t)N + 4095) / 4096);
    dim3 grid(n_shift, chunks);
    double *part = static_cast<double *>(scratch);
    if (q) k_shift_corr<true><<<grid, EV_NT, 0, st>>>(q, ld_q, nullptr, 0, tx, ld_tx, amp, n_lev, N, n_shift, part);
    else k_shift_corr<false><<<grid, EV_NT, 0, st>>>(nullptr, 0, out, ld_out, tx, ld_tx, nullptr, 0, N, n_shift, part);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_shift_corr");
    k_shift_decide<<<1, 32, 0, st>>>(part, n_shift, chunks, corr_out, shift_out, r_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_shift_decide");
    return VAEQ_OK;
}

extern "C" int vaeq_ser_iqflip(const float *q, int64_t ld_q, const uint16_t *tx, int64_t ld_tx, int32_t n_lev, int32_t N,
                               int32_t *counts_out, float *ser_out, void *stream) {
    VAEQ_CHECK_ARG(q && tx && counts_out && ser_out && N > 0, "bad ser_iqflip arguments");
    VAEQ_CHECK_ARG(n_lev == 2 || n_lev == 4 || n_lev == 8, "n_lev=%d must be 2, 4 or 8", n_lev);
    cudaStream_t st = (cudaStream_t)stream;
    VAEQ_CUDA(cudaMemsetAsync(counts_out, 0, 16 * sizeof(int), st));
    const int grid = ev_grid(N);
    if (n_lev == 2) k_ser_iqflip<2><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, counts_out);
    else if (n_lev == 4) k_ser_iqflip<4><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, counts_out);
    else k_ser_iqflip<8><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, counts_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_iqflip");
    k_ser_min<<<1, 32, 0, st>>>(counts_out, N, ser_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_min");
    return VAEQ_OK;
}

extern "C" int vaeq_ser_constell(float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx, const float *amp,
                                 const float *var, float nu_sc, int32_t n_lev, int32_t N, int32_t *counts_out,
                                 float *ser_out, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(rx && tx && amp && var && counts_out && ser_out && scratch && N > 0, "bad ser_constell arguments");
    VAEQ_CHECK_ARG(n_lev == 2 || n_lev == 4 || n_lev == 8, "n_lev=%d must be 2, 4 or 8", n_lev);
    cudaStream_t st = (cudaStream_t)stream;
    double *sums = static_cast<double *>(scratch);
    VAEQ_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    VAEQ_CUDA(cudaMemsetAsync(counts_out, 0, 16 * sizeof(int), st));
    const int grid = ev_grid(N);
    k_constell_norms<<<grid, EV_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, N, sums);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_constell_norms");
    if (n_lev == 2) k_ser_constell<2><<<grid, EV_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, var, nu_sc, N, sums, counts_out);
    else if (n_lev == 4) k_ser_constell<4><<<grid, EV_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, var, nu_sc, N, sums, counts_out);
    else k_ser_constell<8><<<grid, EV_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, var, nu_sc, N, sums, counts_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_constell");
    k_ser_min<<<1, 32, 0, st>>>(counts_out, N, ser_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_min");
    return VAEQ_OK;
}

extern "C" int vaeq_gmi(const float *q, int64_t ld_q, const uint16_t *tx, int64_t ld_tx, const float *P, int32_t n_lev,
                        int32_t N, float *gmi_out, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(q && tx && P && gmi_out && scratch && N > 0, "bad gmi arguments");
    VAEQ_CHECK_ARG(n_lev == 2 || n_lev == 4 || n_lev == 8, "n_lev=%d must be 2, 4 or 8", n_lev);
    cudaStream_t st = (cudaStream_t)stream;
    double *sums = static_cast<double *>(scratch);
    VAEQ_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    const int grid = ev_grid(N);
    if (n_lev == 2) k_gmi<2><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, sums);
    else if (n_lev == 4) k_gmi<4><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, sums);
    else k_gmi<8><<<grid, EV_NT, 0, st>>>(q, ld_q, tx, ld_tx, N, sums);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_gmi");
    k_gmi_fin<<<1, 32, 0, st>>>(sums, P, n_lev, N, gmi_out);
    ktime_begin(VAEQ_K_EVAL, st); ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_gmi_fin");
    return VAEQ_OK;
}
