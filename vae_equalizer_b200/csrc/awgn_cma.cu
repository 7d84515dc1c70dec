// Single-polarisation CMA baseline of the AWGN module (AWGN_channel/func_CMA_MQAM_shaping.py, `cm:` below):
//   CMA :142-168 (complex FIR h[0] + j h[1], tap update after every symbol, NO power normalisation -- unlike the DP version sf:350),
//   SER_CMA :63-93 (in-place rescale, nearest-level decisions, 4 rotations), find_shift_symb :127-140 (non-circular correlation of
//   the first 1000 symbols).  CPE :170-196 (no unwrap) is cma.cu's pipeline with unwrap = 0.
// The output index k = i//sps - mh is negative for the first symbols and wraps to the END of out / e like in the DP version (cm:155).
#include "common.cuh"

namespace vaeq {

// one warp per run; lane l owns taps l and l + 32
template <int NSLOT>
__global__ void __launch_bounds__(32) k_cma1_sample(const float *Rx, float *h, float *out, float *e, int N, int M, int sps, float R,
                                                    float lr, int train) {
    const int lane = threadIdx.x, mh = M / 2, Nsym = N / sps;
    const float *y = Rx + (int64_t)blockIdx.x * 2 * N;
    float *hh = h + (int64_t)blockIdx.x * 2 * M, *o = out + (int64_t)blockIdx.x * 2 * Nsym, *er = e + (int64_t)blockIdx.x * Nsym;
    float h0[NSLOT], h1[NSLOT];
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
        const int k = lane + 32 * sl;
        h0[sl] = k < M ? hh[k] : 0.f;
        h1[sl] = k < M ? hh[M + k] : 0.f;
    }
    const float lr2 = 2.f * lr;
    const int nsym_loop = (N + sps - 1) / sps;
    float n0[NSLOT], n1[NSLOT];                                  // the next symbol's window, loaded one iteration ahead
    auto load_window = [&](int ks) {
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            const int k = lane + 32 * sl, s = ks * sps - mh + k;  // zero padding of mh samples per side (cm:149-150)
            const bool ok = (k < M) && (s >= 0) && (s < N);
            n0[sl] = ok ? y[s] : 0.f;
            n1[sl] = ok ? y[N + s] : 0.f;
        }
    };
    load_window(0);
    for (int ks = 0; ks < nsym_loop; ++ks) {
        float y0[NSLOT], y1[NSLOT];
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            y0[sl] = n0[sl];
            y1[sl] = n1[sl];
        }
        if (ks + 1 < nsym_loop) load_window(ks + 1);
        float a = 0.f, b = 0.f, c = 0.f, d = 0.f;                // y0.h0, y1.h1, y0.h1, y1.h0  (four matmuls, cm:157-158)
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            a = fmaf(y0[sl], h0[sl], a);
            b = fmaf(y1[sl], h1[sl], b);
            c = fmaf(y0[sl], h1[sl], c);
            d = fmaf(y1[sl], h0[sl], d);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        c = warp_sum(c);
        d = warp_sum(d);
        const float o0 = __fsub_rn(a, b), o1 = __fadd_rn(c, d);
        const float err = __fsub_rn(__fsub_rn(R, __fmul_rn(o0, o0)), __fmul_rn(o1, o1));      // cm:160
        int k = (mh + ks * sps) / sps - mh;                      // cm:155
        if (k < 0) k += Nsym;
        if (lane == 0 && k >= 0 && k < Nsym) {
            o[k] = o0;
            o[Nsym + k] = o1;
            er[k] = err;
        }
        if (train) {
            const float f = lr2 * err;
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl) {                 // cm:163-164
                h0[sl] += f * (o0 * y0[sl] + o1 * y1[sl]);
                h1[sl] += f * (o1 * y0[sl] - o0 * y1[sl]);
            }
        }
    }
    if (train) {
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            const int k = lane + 32 * sl;
            if (k < M) {
                hh[k] = h0[sl];
                hh[M + k] = h1[sl];
            }
        }
    }
}

// ---- SER_CMA (cm:63-93) --------------------------------------------------------------------------------------------------
constexpr int SC1_NT = 256;
__global__ void __launch_bounds__(SC1_NT) k_ser_cma_norms(const float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx, int N, double *part) {
    __shared__ double red[2 * 32];
    double acc[2] = {0.0, 0.0};
    for (int t = blockIdx.x * SC1_NT + threadIdx.x; t < N; t += gridDim.x * SC1_NT) {
        const float a = half_bits_to_float(tx[t]), b = half_bits_to_float(tx[ld_tx + t]), x = rx[t], y = rx[ld_rx + t];
        acc[0] += (double)sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
        acc[1] += (double)sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {                                      // per-CTA partials, summed in fixed order by the consumer
        part[2 * blockIdx.x] = acc[0];
        part[2 * blockIdx.x + 1] = acc[1];
    }
}

template <int NL>
__global__ void __launch_bounds__(SC1_NT) k_ser_cma(float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx, const float *amp, int N,
                                                    const double *part, int nparts, int *counts) {
    __shared__ double bc[2];
    __shared__ int red[4 * 32];
    __shared__ float a[NL];
    if (threadIdx.x < NL) a[threadIdx.x] = amp[threadIdx.x];
    if (threadIdx.x < 32) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = threadIdx.x; i < nparts; i += 32) {
            s0 += part[2 * i];
            s1 += part[2 * i + 1];
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (threadIdx.x == 0) {
            bc[0] = s0;
            bc[1] = s1;
        }
    }
    __syncthreads();
    const float g = __fdiv_rn((float)(bc[0] / (double)N), (float)(bc[1] / (double)N));          // cm:73
    const float scale = (float)((NL - 1) / 2.0);
    const int S = NL - 1;
    int cnt[4] = {0, 0, 0, 0};
    for (int t = blockIdx.x * SC1_NT + threadIdx.x; t < N; t += gridDim.x * SC1_NT) {
        const float sI = __fmul_rn(rx[t], g), sQ = __fmul_rn(rx[ld_rx + t], g);
        rx[t] = sI;                                              // in-place rescale, visible to the caller (cm:73)
        rx[ld_rx + t] = sQ;
        int dI = 0, dQ = 0;
        float bI = fabsf(__fsub_rn(sI, a[0])), bQ = fabsf(__fsub_rn(sQ, a[0]));
#pragma unroll
        for (int l = 1; l < NL; ++l) {                           // torch.argmin: first minimum (cm:76)
            const float vI = fabsf(__fsub_rn(sI, a[l])), vQ = fabsf(__fsub_rn(sQ, a[l]));
            if (vI < bI) {
                bI = vI;
                dI = l;
            }
            if (vQ < bQ) {
                bQ = vQ;
                dQ = l;
            }
        }
        const int xI = (int)rintf(__fadd_rn(__fmul_rn(scale, half_bits_to_float(tx[t])), scale));                // cm:72, round half to even
        const int xQ = (int)rintf(__fadd_rn(__fmul_rn(scale, half_bits_to_float(tx[ld_tx + t])), scale));
        cnt[0] += (xI != dI) || (xQ != dQ);                      // 0
        cnt[1] += (xI != S - dI) || (xQ != S - dQ);              // pi          (cm:80)
        cnt[2] += (xI != S - dQ) || (xQ != dI);                  // "pi/4"      (cm:84-86)
        cnt[3] += (xI != dQ) || (xQ != S - dI);                  // "3pi/4"     (cm:89-91)
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k * 32 + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int v = 0;
        for (int w = 0; w < SC1_NT / 32; ++w) v += red[threadIdx.x * 32 + w];
        if (v) atomicAdd(&counts[threadIdx.x], v);
    }
}

__global__ void k_ser_cma_min(const int *counts, int N, float *ser) {
    if (threadIdx.x == 0) {
        float best = 1.f;                                        // SER = torch.ones(4) (cm:69)
        for (int r = 0; r < 4; ++r) best = fminf(best, __fdiv_rn((float)counts[r], (float)N));
        ser[0] = best;
    }
}

// ---- find_shift_symb (cm:127-140): corr_c[i] = sum_{t < win - half} tx_c[half + t] rx_I[i + t], i < n_shift, win = 1000 -----------
__global__ void __launch_bounds__(1024) k_find_shift_symb(const float *rx, int n_rx, const uint16_t *tx, int64_t ld_tx, int n_shift, int win,
                                                          float *corr, int *shift) {
    __shared__ double acc[2][64];
    const int half = n_shift / 2, len = win - half, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int job = wid; job < 2 * n_shift; job += nw) {          // a warp per (component, shift)
        const int c = job / n_shift, i = job - c * n_shift;
        double s = 0.0;
        for (int t = lane; t < len; t += 32) s += (double)half_bits_to_float(tx[(int64_t)c * ld_tx + half + t]) * (double)rx[i + t];
        s = warp_sum(s);
        if (lane == 0) acc[c][i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best[2] = {-1.f, -1.f};
        int arg[2] = {0, 0};
        for (int c = 0; c < 2; ++c)
            for (int i = 0; i < n_shift; ++i) {
                const float v = fabsf((float)acc[c][i]);
                corr[c * n_shift + i] = (float)acc[c][i];
                if (v > best[c]) {                               // torch.argmax: first maximum
                    best[c] = v;
                    arg[c] = i;
                }
            }
        int pick = arg[0];
        if (!(best[0] >= 0.02f * (float)n_rx) && best[1] >= best[0]) pick = arg[1];      // cm:133-140
        shift[0] = pick - half;
    }
}

}  // namespace vaeq

using namespace vaeq;

extern "C" int vaeq_cma_awgn(const float *Rx, int32_t N, float R, float *h, int32_t M, float lr, int32_t sps, int32_t train, float *out,
                             float *e, int32_t n_runs, void *stream) {
    VAEQ_CHECK_ARG(Rx && h && out && e && N > 0 && n_runs > 0, "bad cma_awgn arguments");
    VAEQ_CHECK_ARG(M >= 1 && M <= VAEQ_MAX_TAPS && (M & 1), "M=%d must be odd and <= %d", M, VAEQ_MAX_TAPS);
    VAEQ_CHECK_ARG(sps >= 1 && N % sps == 0, "N=%d must be a multiple of sps=%d", N, sps);
    cudaStream_t st = (cudaStream_t)stream;
    ktime_begin(VAEQ_K_CMA, st);
    if (M <= 32) k_cma1_sample<1><<<n_runs, 32, 0, st>>>(Rx, h, out, e, N, M, sps, R, lr, train);
    else k_cma1_sample<2><<<n_runs, 32, 0, st>>>(Rx, h, out, e, N, M, sps, R, lr, train);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cma1_sample");
    return VAEQ_OK;
}

extern "C" int vaeq_ser_cma(float *rx, int64_t ld_rx, const uint16_t *tx, int64_t ld_tx, const float *amp, int32_t n_lev, int32_t N,
                            int32_t *counts_out, float *ser_out, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(rx && tx && amp && counts_out && ser_out && scratch && N > 0, "bad ser_cma arguments");
    VAEQ_CHECK_ARG(n_lev == 2 || n_lev == 4 || n_lev == 8, "n_lev=%d must be 2, 4 or 8", n_lev);
    cudaStream_t st = (cudaStream_t)stream;
    double *part = static_cast<double *>(scratch);
    VAEQ_CUDA(cudaMemsetAsync(counts_out, 0, 4 * sizeof(int), st));
    const int grid = max(1, min(min((N + SC1_NT - 1) / SC1_NT, sm_count() * 8), (int)(VAEQ_EVAL_SCRATCH_BYTES / (2 * sizeof(double)))));
    ktime_begin(VAEQ_K_EVAL, st);
    k_ser_cma_norms<<<grid, SC1_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, N, part);
    ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_cma_norms");
    ktime_begin(VAEQ_K_EVAL, st);
    switch (n_lev) {
        case 2: k_ser_cma<2><<<grid, SC1_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, N, part, grid, counts_out); break;
        case 4: k_ser_cma<4><<<grid, SC1_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, N, part, grid, counts_out); break;
        default: k_ser_cma<8><<<grid, SC1_NT, 0, st>>>(rx, ld_rx, tx, ld_tx, amp, N, part, grid, counts_out); break;
    }
    ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_cma");
    ktime_begin(VAEQ_K_EVAL, st);
    k_ser_cma_min<<<1, 32, 0, st>>>(counts_out, N, ser_out);
    ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_ser_cma_min");
    return VAEQ_OK;
}

extern "C" int vaeq_find_shift_symb(const float *rx, int32_t n_rx, const uint16_t *tx, int64_t ld_tx, int32_t n_tx, int32_t n_shift,
                                    float *corr_out, int32_t *shift_out, void *stream) {
    VAEQ_CHECK_ARG(rx && tx && corr_out && shift_out, "NULL pointer");
    VAEQ_CHECK_ARG(n_shift > 0 && n_shift <= 64, "n_shift=%d must be in [1, 64]", n_shift);
    const int win = 1000;                                        // cm:128-131: the first 1000 symbols
    VAEQ_CHECK_ARG(n_tx >= win && n_rx >= win + n_shift / 2, "find_shift_symb needs at least %d symbols (got rx %d, tx %d)", win + n_shift / 2, n_rx, n_tx);
    cudaStream_t st = (cudaStream_t)stream;
    ktime_begin(VAEQ_K_EVAL, st);
    k_find_shift_symb<<<1, 1024, 0, st>>>(rx, n_rx, tx, ld_tx, n_shift, win, corr_out, shift_out);
    ktime_end(VAEQ_K_EVAL, st);
    VAEQ_LAUNCH_CHECK("k_find_shift_symb");
    return VAEQ_OK;
}
