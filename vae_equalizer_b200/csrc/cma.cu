// CMA / CMAbatch / CMAflex baselines and Viterbi-Viterbi CPE for B200.
// Reference: optical_DP_channel/shared_funcs.py  CMA :341-379, CMAbatch :381-434, CMAflex :436-488, CPE :140-186.
//
// Data flow per run:  k_cma_power (mean power incl. the zero pads, sf:349-350) -> k_cma_scale (y = Rx / power)
//   -> k_cma_sample  : one WARP per run, lanes own taps, per-symbol recurrence (inherently sequential, sf:355-378)
//   or k_cma_block   : one CTA per run; taps are constant between two updates, so the symbols of a segment are
//                      equalized in parallel and the update  h += 2 lr sum_k e_k * inc_k  (sf:424-433, :478-487) is
//                      a correlation of g_k = e_k * out_k with the input window -- the dW contraction of the VAE step.
// Index quirk reproduced: the reference writes symbol ks to out[k], e[k], buf[k] with k = (mh + ks*sps)//sps - mh
// (sf:357), which is NEGATIVE for the first symbols and wraps to the end of the arrays (torch indexing).
#include <algorithm>
#include "common.cuh"

namespace vaeq {

constexpr int CMA_NT = 256;

struct CmaRun {
    const float *Rx;   // (2,2,N) of this run
    float *y;          // scaled copy (2,2,N)
    float *h;          // (2,2,2,M)
    float *out;        // (2,2,Nsym)
    float *e;          // (Nsym,2)
};

__device__ __forceinline__ CmaRun cma_run_ptrs(const float *Rx, float *ys, float *h, float *out, float *e, int run, int N,
                                               int M, int Nsym) {
    CmaRun r;
    r.Rx = Rx + (int64_t)run * 4 * N;
    r.y = ys + (int64_t)run * 4 * N;
    r.h = h + (int64_t)run * 8 * M;
    r.out = out + (int64_t)run * 4 * Nsym;
    r.e = e + (int64_t)run * 2 * Nsym;
    return r;
}

// mean(y_I^2 + y_Q^2) over both pols and the PADDED length N + 2*mh  (sf:350): one CTA per run
__global__ void __launch_bounds__(CMA_NT) k_cma_power(const float *Rx, int N, int mh, float *pw) {
    __shared__ double red[32];
    const float *x = Rx + (int64_t)blockIdx.x * 4 * N;
    double acc[1] = {0.0};
    for (int i = threadIdx.x; i < 2 * N; i += CMA_NT) {
        const int p = i / N, s = i - p * N;
        const float a = x[(int64_t)(2 * p) * N + s], b = x[(int64_t)(2 * p + 1) * N + s];
        acc[0] += (double)__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) pw[blockIdx.x] = (float)(acc[0] / (2.0 * (double)(N + 2 * mh)));
}

__global__ void k_cma_scale(const float *Rx, int N, const float *pw, float *ys, int n_runs) {
    const int64_t total = (int64_t)n_runs * 4 * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        ys[i] = __fdiv_rn(Rx[i], pw[i / (4 * (int64_t)N)]);                       // y /= mean power (sf:350)
}

__device__ __forceinline__ int cma_out_index(int ks, int sps, int mh, int Nsym) {
    const int k = (mh + ks * sps) / sps - mh;                                    // sf:357
    return k;                                                                    // may be negative
}

// ---------------------------------------------------------------------------------------------
// CMA: per-symbol update.  One warp per run; lane l owns taps k = l and k = l + 32.
// ---------------------------------------------------------------------------------------------
// NSLOT = 1 for M <= 32 (every reference setting uses M = 25): no second, all-zero tap slot to multiply, update and load for
template <int NSLOT>
__global__ void __launch_bounds__(32) k_cma_sample(const float *Rx, float *ys, float *h, float *out, float *e, int N, int M,
                                                   int sps, float R, float lr, int train) {
    const int lane = threadIdx.x, mh = M / 2, Nsym = N / sps;
    const CmaRun r = cma_run_ptrs(Rx, ys, h, out, e, blockIdx.x, N, M, Nsym);
    float hr[NSLOT][2][2], hi[NSLOT][2][2];                                              // [tap slot][o][i]
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
        const int k = lane + 32 * sl;
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                hr[sl][o][i] = k < M ? r.h[((o * 2 + i) * 2 + 0) * M + k] : 0.f;
                hi[sl][o][i] = k < M ? r.h[((o * 2 + i) * 2 + 1) * M + k] : 0.f;
            }
    }
    const float lr2 = 2.f * lr;
    const int nsym_loop = (N + sps - 1) / sps;
    // the window of symbol ks + 1 does not depend on the tap recurrence: it is loaded one iteration ahead, so that the
    // L2 latency of the loads is not part of the per-symbol dependency chain (one warp per run has nothing else to hide it)
    float nI[NSLOT][2], nQ[NSLOT][2];                                                   // [slot][in pol]
    const float *yrow[4] = {r.y, r.y + N, r.y + 2 * (int64_t)N, r.y + 3 * (int64_t)N};   // row bases once: the loop is one warp's dependent chain,
    auto load_window = [&](int ks) {                                                  // every instruction in it counts
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            const int k = lane + 32 * sl, s = ks * sps - mh + k;
            const bool ok = (k < M) && ((unsigned)s < (unsigned)N);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                nI[sl][i] = ok ? yrow[2 * i][s] : 0.f;
                nQ[sl][i] = ok ? yrow[2 * i + 1][s] : 0.f;
            }
        }
    };
    const int koff = mh / sps - mh;                                                   // cma_out_index(ks) = ks + koff: (mh + ks sps) / sps = ks + mh / sps
    load_window(0);
    for (int ks = 0; ks < nsym_loop; ++ks) {
        float yI[NSLOT][2], yQ[NSLOT][2];
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                yI[sl][i] = nI[sl][i];
                yQ[sl][i] = nQ[sl][i];
            }
        if (ks + 1 < nsym_loop) load_window(ks + 1);
        float oI[2] = {0.f, 0.f}, oQ[2] = {0.f, 0.f};
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int i = 0; i < 2; ++i) {                                    // sf:360-364
                    oI[o] += yI[sl][i] * hr[sl][o][i] - yQ[sl][i] * hi[sl][o][i];
                    oQ[o] += yI[sl][i] * hi[sl][o][i] + yQ[sl][i] * hr[sl][o][i];
                }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            oI[o] = warp_sum(oI[o]);
            oQ[o] = warp_sum(oQ[o]);
        }
        float err[2];
#pragma unroll
        for (int o = 0; o < 2; ++o) err[o] = R - oI[o] * oI[o] - oQ[o] * oQ[o];  // sf:366-367
        int k = ks + koff;                                                       // sf:357, may be negative: wraps to the end
        if (k < 0) k += Nsym;
        if (lane == 0 && k >= 0 && k < Nsym) {
            r.out[0 * Nsym + k] = oI[0];
            r.out[1 * Nsym + k] = oQ[0];
            r.out[2 * Nsym + k] = oI[1];
            r.out[3 * Nsym + k] = oQ[1];
            r.e[2 * k + 0] = err[0];
            r.e[2 * k + 1] = err[1];
        }
        if (train) {
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const float f = lr2 * err[o];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {                                // sf:370-378
                        hr[sl][o][i] += f * (oI[o] * yI[sl][i] + oQ[o] * yQ[sl][i]);
                        hi[sl][o][i] += f * (oQ[o] * yI[sl][i] - oI[o] * yQ[sl][i]);
                    }
                }
        }
    }
    if (train) {
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
            const int k = lane + 32 * sl;
            if (k < M) {
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        r.h[((o * 2 + i) * 2 + 0) * M + k] = hr[sl][o][i];
                        r.h[((o * 2 + i) * 2 + 1) * M + k] = hi[sl][o][i];
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CMAbatch / CMAflex: one CTA per run, segment-parallel.
// ---------------------------------------------------------------------------------------------
// STAGED: the sample window of a chunk of symbols (equalize stage) and the window, outputs and errors of an update are staged in
// shared memory first, so the sequential per-thread loops (25 taps per symbol; batchlen symbols per tap gradient) run on shared-memory
// latency instead of a global round trip per step.  Same operations in the same order: results are bit-identical to the
// unstaged form, which remains for windows that do not fit (batchlen > CMA_WCAP).
constexpr int CMA_WCAP = 1024;                                                   // largest update window (symbols) held in shared memory
template <bool STAGED>
__global__ void __launch_bounds__(CMA_NT) k_cma_block(const float *Rx, float *ys, float *h, float *out, float *e, int N, int M,
                                                      int sps, float R, float lr, int train, int mode, int batchlen,
                                                      int symb_step, int ycap) {
    extern __shared__ float hs[];                                                // (2,2,2,M) taps of this run | staging buffers
    float *ybuf = hs + 8 * M;                                                    // [4][ycap] samples
    float *obuf = ybuf + 4 * ycap;                                               // [4][batchlen] equalizer outputs of the update window
    float *ebuf = obuf + (STAGED ? 4 * batchlen : 0);                            // [batchlen][2] errors
    const int tid = threadIdx.x, mh = M / 2, Nsym = N / sps;
    const CmaRun r = cma_run_ptrs(Rx, ys, h, out, e, blockIdx.x, N, M, Nsym);
    for (int i = tid; i < 8 * M; i += CMA_NT) hs[i] = r.h[i];
    __syncthreads();
    const int nsym_loop = (N + sps - 1) / sps;
    const int k_first = cma_out_index(0, sps, mh, Nsym), k_last = cma_out_index(nsym_loop - 1, sps, mh, Nsym);
    const int off = -k_first;                                                    // ks = k + off
    const float lr2 = 2.f * lr;
    int k_cur = k_first;
    while (k_cur <= k_last) {
        // next firing symbol index >= max(k_cur, 1)
        int k_fire;
        if (!train) {
            k_fire = k_last + 1;
        } else if (mode == VAEQ_CMA_BATCH) {                                     // k % batchlen == 0 and k != 0  (sf:424)
            const int lo = max(k_cur, 1);
            k_fire = ((lo + batchlen - 1) / batchlen) * batchlen;
        } else {                                                                 // k % symb_step == 0 and k >= batchlen (sf:478)
            const int lo = max(k_cur, batchlen);
            k_fire = ((lo + symb_step - 1) / symb_step) * symb_step;
        }
        const int k_end = min(k_fire, k_last);                                   // inclusive
        // ---- equalize symbols k_cur..k_end with the current taps, CMA_NT symbols at a time ------------------------------
        for (int kc = k_cur; kc <= k_end; kc += CMA_NT) {
            const int kc_hi = min(kc + CMA_NT - 1, k_end);
            const int s_base = (kc + off) * sps - mh, ns = (kc_hi - kc) * sps + M;
            if (STAGED) {
                for (int j = tid; j < 4 * ns; j += CMA_NT) {
                    const int row = j / ns, jj = j - row * ns, sidx = s_base + jj;
                    ybuf[row * ycap + jj] = (sidx >= 0 && sidx < N) ? r.y[(int64_t)row * N + sidx] : 0.f;
                }
                __syncthreads();
            }
            const int k = kc + tid;
            if (k <= kc_hi) {
                const int ks = k + off;
                float oI[2] = {0.f, 0.f}, oQ[2] = {0.f, 0.f};
                const int sm0 = ks * sps - mh, m_lo = max(0, -sm0), m_hi = min(M, N - sm0);      // taps whose sample lies inside [0, N): one range
#pragma unroll 2
                for (int m = m_lo; m < m_hi; ++m) {
                    const int s = sm0 + m;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float a = STAGED ? ybuf[(2 * i) * ycap + (s - s_base)] : r.y[(int64_t)(2 * i) * N + s];
                        const float b = STAGED ? ybuf[(2 * i + 1) * ycap + (s - s_base)] : r.y[(int64_t)(2 * i + 1) * N + s];
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const float wr = hs[((o * 2 + i) * 2 + 0) * M + m], wi = hs[((o * 2 + i) * 2 + 1) * M + m];
                            oI[o] += a * wr - b * wi;
                            oQ[o] += a * wi + b * wr;
                        }
                    }
                }
                const int kk = k < 0 ? k + Nsym : k;
                if (kk >= 0 && kk < Nsym) {
                    r.out[0 * Nsym + kk] = oI[0];
                    r.out[1 * Nsym + kk] = oQ[0];
                    r.out[2 * Nsym + kk] = oI[1];
                    r.out[3 * Nsym + kk] = oQ[1];
                    r.e[2 * kk + 0] = R - oI[0] * oI[0] - oQ[0] * oQ[0];
                    r.e[2 * kk + 1] = R - oI[1] * oI[1] - oQ[1] * oQ[1];
                }
            }
            __syncthreads();                                                     // outputs visible to the update; ybuf free again
        }
        // ---- tap update from the window [k_fire - batchlen, k_fire) ------------------------------
        if (train && k_fire <= k_last) {
            const int kap0 = k_fire - batchlen, sw_base = (kap0 + off) * sps - mh, nsw = (batchlen - 1) * sps + M;
            if (STAGED) {
                for (int j = tid; j < 4 * nsw; j += CMA_NT) {
                    const int row = j / nsw, jj = j - row * nsw, sidx = sw_base + jj;
                    ybuf[row * ycap + jj] = (sidx >= 0 && sidx < N) ? r.y[(int64_t)row * N + sidx] : 0.f;
                }
                for (int j = tid; j < 4 * batchlen; j += CMA_NT) {
                    const int row = j / batchlen, kap = kap0 + (j - row * batchlen);
                    obuf[j] = (kap >= 0 && kap < Nsym) ? r.out[row * Nsym + kap] : 0.f;
                }
                for (int j = tid; j < 2 * batchlen; j += CMA_NT) {
                    const int kap = kap0 + (j >> 1);
                    ebuf[j] = (kap >= 0 && kap < Nsym) ? r.e[2 * kap + (j & 1)] : 0.f;
                }
                __syncthreads();
            }
            if (STAGED) {
                // item = (o, i, m), BOTH components c of the tap increment at once: they share all five shared-memory operands of a symbol.
                // The valid symbols of the window form ONE range (kap >= 0 and 0 <= s < N, s = (kap + off) sps - mh + m ascending in kap): bounds
                // hoisted, pointers stepped: 5 shared loads + 6 FP operations per symbol (the generic loop below was 35 instructions per symbol and
                // component, 68 % of the kernel's instruction stream); same expressions in the same order per accumulator: bit-identical sums
                for (int it = tid; it < 4 * M; it += CMA_NT) {
                    const int m = it % M, oi = it / M, i = oi & 1, o = oi >> 1;
                    float acc0 = 0.f, acc1 = 0.f;
                    const int t0s = mh - m - off * sps;                              // s >= 0  <=>  kap * sps >= t0s
                    const int lo_s = t0s <= 0 ? 0 : (t0s + sps - 1) / sps;
                    const int hi_s = (N - 1 + t0s) >= 0 ? (N - 1 + t0s) / sps : -1;  // s <= N - 1  <=>  kap * sps <= N - 1 + t0s
                    const int ka = max(max(kap0, 0), lo_s), kb = min(k_fire, hi_s + 1);
                    if (ka < kb) {
                        const int w0 = ka - kap0, s0 = (ka + off) * sps - mh + m - sw_base;
                        const float *ya = ybuf + (2 * i) * ycap + s0, *yb = ybuf + (2 * i + 1) * ycap + s0;
                        const float *pI = obuf + (2 * o) * batchlen + w0, *pQ = obuf + (2 * o + 1) * batchlen + w0, *pe = ebuf + 2 * w0 + o;
                        const int cnt = kb - ka;
#pragma unroll 4
                        for (int n = 0; n < cnt; ++n) {
                            const float a = ya[n * sps], b = yb[n * sps], vI = pI[n], vQ = pQ[n], ev = pe[2 * n];
                            const float inc0 = (vI * a + vQ * b);                    // sf:414-422
                            const float inc1 = (vQ * a - vI * b);
                            acc0 += ev * inc0;
                            acc1 += ev * inc1;
                        }
                    }
                    hs[((o * 2 + i) * 2 + 0) * M + m] += lr2 * acc0;                 // sf:425-433
                    hs[((o * 2 + i) * 2 + 1) * M + m] += lr2 * acc1;
                }
            } else {
            for (int idx = tid; idx < 8 * M; idx += CMA_NT) {
                const int m = idx % M, oic = idx / M, c = oic & 1, oi = oic >> 1, i = oi & 1, o = oi >> 1;
                float acc = 0.f;
                {
                for (int kap = kap0; kap < k_fire; ++kap) {
                    const int ks = kap + off, s = ks * sps - mh + m;
                    if (kap < 0 || s < 0 || s >= N) continue;                    // kap < 0 would read torch.empty garbage in the reference
                    const int w = kap - kap0;
                    const float a = STAGED ? ybuf[(2 * i) * ycap + (s - sw_base)] : r.y[(int64_t)(2 * i) * N + s];
                    const float b = STAGED ? ybuf[(2 * i + 1) * ycap + (s - sw_base)] : r.y[(int64_t)(2 * i + 1) * N + s];
                    const float vI = STAGED ? obuf[(2 * o) * batchlen + w] : r.out[(2 * o) * Nsym + kap];
                    const float vQ = STAGED ? obuf[(2 * o + 1) * batchlen + w] : r.out[(2 * o + 1) * Nsym + kap];
                    const float inc = c ? (vQ * a - vI * b) : (vI * a + vQ * b);     // sf:414-422
                    acc += (STAGED ? ebuf[2 * w + o] : r.e[2 * kap + o]) * inc;
                }
                }
                hs[idx] += lr2 * acc;                                            // sf:425-433
            }
            }
            __syncthreads();
        }
        k_cur = k_end + 1;
    }
    if (train)
        for (int i = tid; i < 8 * M; i += CMA_NT) r.h[i] = hs[i];
}

// ---------------------------------------------------------------------------------------------
// CPE (sf:140-186)
// ---------------------------------------------------------------------------------------------
constexpr int CPE_MA = 501;

// npol = 2 n_runs: the (n_runs,2,2,N) batch is npol independent (I, Q) row pairs
__global__ void k_cpe_pow4(const float *y, int N, int npol, float *p4) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npol * (int64_t)N; t += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(t / N), s = (int)(t - (int64_t)p * N);
        const float a = y[(int64_t)(2 * p) * N + s], b = y[(int64_t)(2 * p + 1) * N + s];
        const float a2 = a * a, b2 = b * b;
        p4[(int64_t)(2 * p) * N + s] = __fadd_rn(__fsub_rn(__fmul_rn(a2, a2), __fmul_rn(__fmul_rn(6.f, a2), b2)), __fmul_rn(b2, b2));   // sf:152
        p4[(int64_t)(2 * p + 1) * N + s] = __fmul_rn(4.f, __fsub_rn(__fmul_rn(__fmul_rn(a2, a), b), __fmul_rn(__fmul_rn(a, b2), b)));     // sf:153
    }
}

// phi[p][n] = atan2(ma_im, -ma_re)/4 with a zero-padded 501-tap moving average (sf:158-163).
// Every output keeps the reference's summation order (taps in ascending sample order, one fused multiply-add per tap), so the
// phases are bit-identical to the one-output-per-thread loop this replaces; a thread owns CPE_OPT consecutive outputs whose windows
// overlap, so one staged sample feeds CPE_OPT independent accumulator pairs (8x fewer loads, 16 independent chains instead of 2:
// 1.83 -> 0.3 ms for 592 runs x 9980 symbols, profiles/r01d_sweep_and_datagen.txt).  grid (ceil(N / CPE_TILE), npol).
constexpr int CPE_OPT = 8, CPE_PT = 128, CPE_TILE = CPE_OPT * CPE_PT, CPE_HALF = CPE_MA / 2, CPE_SPAN = CPE_TILE + 2 * CPE_HALF;
__device__ __forceinline__ int cpe_pad(int i) { return i + (i >> 3); }       // lanes read 8 apart: +1 per 8 keeps them on distinct banks
__global__ void __launch_bounds__(CPE_PT) k_cpe_phase(const float *p4, int N, int npol, float *phi) {
    __shared__ float s_re[CPE_SPAN + CPE_SPAN / 8 + 1], s_im[CPE_SPAN + CPE_SPAN / 8 + 1];
    const float w = (float)(1.0 / CPE_MA);
    const int p = blockIdx.y, blk0 = blockIdx.x * CPE_TILE, tid = threadIdx.x;
    const float *re = p4 + (int64_t)(2 * p) * N, *im = p4 + (int64_t)(2 * p + 1) * N;
    for (int i = tid; i < CPE_SPAN; i += CPE_PT) {
        const int g = blk0 - CPE_HALF + i;
        const bool ok = g >= 0 && g < N;
        s_re[cpe_pad(i)] = ok ? re[g] : 0.f;
        s_im[cpe_pad(i)] = ok ? im[g] : 0.f;
    }
    __syncthreads();
    float sr[CPE_OPT], si[CPE_OPT];
#pragma unroll
    for (int k = 0; k < CPE_OPT; ++k) sr[k] = si[k] = 0.f;
    const int i0 = CPE_OPT * tid;                                 // tile index of the first sample of output 0's window
    const bool interior = blk0 - CPE_HALF >= 0 && blk0 + CPE_TILE + CPE_HALF <= N;
    // output k (sample o = blk0 + i0 + k) sums tile indices i0 + k ... i0 + k + 2 CPE_HALF, i.e. steps j = k ... k + 2 CPE_HALF
    auto step = [&](int j, bool check) {
        const int i = i0 + j, g = blk0 - CPE_HALF + i;
        const float xr = s_re[cpe_pad(i)], xi = s_im[cpe_pad(i)];
        const bool in_seq = !check || (g >= 0 && g < N);          // the reference sums only over existing samples (lo / hi clipping)
#pragma unroll
        for (int k = 0; k < CPE_OPT; ++k) {
            if (in_seq && j >= k && j <= k + 2 * CPE_HALF) {
                sr[k] = fmaf(xr, w, sr[k]);
                si[k] = fmaf(xi, w, si[k]);
            }
        }
    };
    if (interior) {
#pragma unroll
        for (int j = 0; j < CPE_OPT - 1; ++j) step(j, false);
#pragma unroll 4
        for (int j = CPE_OPT - 1; j <= 2 * CPE_HALF; ++j) {       // every output's window is open: no predicates
            const int i = i0 + j;
            const float xr = s_re[cpe_pad(i)], xi = s_im[cpe_pad(i)];
#pragma unroll
            for (int k = 0; k < CPE_OPT; ++k) {
                sr[k] = fmaf(xr, w, sr[k]);
                si[k] = fmaf(xi, w, si[k]);
            }
        }
#pragma unroll
        for (int j = 2 * CPE_HALF + 1; j < 2 * CPE_HALF + CPE_OPT; ++j) step(j, false);
    } else {
        for (int j = 0; j < 2 * CPE_HALF + CPE_OPT; ++j) step(j, true);
    }
#pragma unroll
    for (int k = 0; k < CPE_OPT; ++k) {
        const int o = blk0 + i0 + k;
        if (o < N) phi[(int64_t)p * N + o] = atan2f(si[k], -sr[k]) * 0.25f;
    }
}

// one CTA per pol: running counts of +/- jumps of the WRAPPED phase (sf:164-169), then derotation (sf:182-185)
__global__ void __launch_bounds__(1024) k_cpe_unwrap_rotate(const float *y, const float *phi, int N, float *out, int unwrap) {
    __shared__ int wsum[2][32];
    __shared__ int carry[2];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float *ph = phi + (int64_t)p * N;
    const float pi = 3.14159274101257324f, pi2 = pi * 0.5f, pi4 = pi * 0.25f;
    if (tid < 2) carry[tid] = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int n = base + tid;
        // flag at n: jump between n-1 and n  (phi[i+1:] is shifted when diff[i] crosses +-pi/4)
        int fp = 0, fn = 0;
        if (unwrap && n < N && n >= 1) {                                         // unwrap == 0: the AWGN module's CPE (no unwrapping)
            const float d = __fsub_rn(ph[n], ph[n - 1]);
            fp = d > pi4;
            fn = d < -pi4;
        }
        int sp = fp, sn = fn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, sp, o), b = __shfl_up_sync(0xffffffffu, sn, o);
            if (lane >= o) {
                sp += a;
                sn += b;
            }
        }
        if (lane == 31) {
            wsum[0][wid] = sp;
            wsum[1][wid] = sn;
        }
        __syncthreads();
        int bp = carry[0], bn = carry[1];
        for (int w = 0; w < wid; ++w) {
            bp += wsum[0][w];
            bn += wsum[1][w];
        }
        const int np = bp + sp, nn = bn + sn;                                    // inclusive counts up to n
        if (n < N) {
            float x = ph[n];
            for (int i = 0; i < np; ++i) x = __fsub_rn(x, pi2);                  // same rounding sequence as the loops at sf:166-169
            for (int i = 0; i < nn; ++i) x = __fadd_rn(x, pi2);
            const float c = cosf(x), s = sinf(x);
            const float a = y[(int64_t)(2 * p) * N + n], b = y[(int64_t)(2 * p + 1) * N + n];
            out[(int64_t)(2 * p) * N + n] = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s));
            out[(int64_t)(2 * p + 1) * N + n] = __fadd_rn(__fmul_rn(b, c), __fmul_rn(a, s));
        }
        __syncthreads();
        if (tid == 1023) {
            carry[0] = np;
            carry[1] = nn;
        }
        __syncthreads();
    }
}

}  // namespace vaeq

using namespace vaeq;

extern "C" size_t vaeq_cma_scratch_bytes(int32_t N, int32_t M, int32_t n_runs) {
    (void)M;
    if (N <= 0 || n_runs <= 0) return 0;
    return align_up((size_t)n_runs * 4 * N * sizeof(float), 256) + align_up((size_t)n_runs * sizeof(float), 256);
}

extern "C" int vaeq_cma(int32_t mode, const float *Rx, int32_t N, float R, float *h, int32_t M, float lr, int32_t batchlen,
                        int32_t symb_step, int32_t sps, int32_t train, float *out, float *e, int32_t n_runs, void *scratch,
                        void *stream) {
    VAEQ_CHECK_ARG(Rx && h && out && e && scratch && N > 0 && n_runs > 0, "bad cma arguments");
    VAEQ_CHECK_ARG(mode >= VAEQ_CMA_SAMPLE && mode <= VAEQ_CMA_FLEX, "bad cma mode %d", mode);
    VAEQ_CHECK_ARG(M >= 1 && M <= VAEQ_MAX_TAPS && (M & 1), "M=%d must be odd and <= %d", M, VAEQ_MAX_TAPS);
    VAEQ_CHECK_ARG(sps >= 1 && N % sps == 0, "N=%d must be a multiple of sps=%d", N, sps);
    VAEQ_CHECK_ARG(mode == VAEQ_CMA_SAMPLE || batchlen > 0, "batchlen must be positive");
    VAEQ_CHECK_ARG(mode != VAEQ_CMA_FLEX || symb_step > 0, "symb_step must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    float *ys = static_cast<float *>(scratch);
    float *pw = reinterpret_cast<float *>(static_cast<char *>(scratch) + align_up((size_t)n_runs * 4 * N * sizeof(float), 256));
    ktime_begin(VAEQ_K_CMA, st);
    k_cma_power<<<n_runs, CMA_NT, 0, st>>>(Rx, N, M / 2, pw);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cma_power");
    const int64_t total = (int64_t)n_runs * 4 * N;
    ktime_begin(VAEQ_K_CMA, st);
    k_cma_scale<<<(int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8), 256, 0, st>>>(Rx, N, pw, ys, n_runs);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cma_scale");
    if (mode == VAEQ_CMA_SAMPLE) {
        ktime_begin(VAEQ_K_CMA, st);
        if (M <= 32) k_cma_sample<1><<<n_runs, 32, 0, st>>>(Rx, ys, h, out, e, N, M, sps, R, lr, train);
        else k_cma_sample<2><<<n_runs, 32, 0, st>>>(Rx, ys, h, out, e, N, M, sps, R, lr, train);
        ktime_end(VAEQ_K_CMA, st);
        VAEQ_LAUNCH_CHECK("k_cma_sample");
    } else {
        const bool staged = batchlen <= CMA_WCAP;
        const int ycap = (std::max(CMA_NT, batchlen) - 1) * sps + M;            // samples per row of the staged window
        const size_t smem = (size_t)(8 * M + (staged ? 4 * ycap + 6 * batchlen : 0)) * sizeof(float);
        ktime_begin(VAEQ_K_CMA, st);
        if (staged) {
            static SmemAttrCache set_smem;
            if (int rc = ensure_dyn_smem(k_cma_block<true>, smem, set_smem)) return rc;
            k_cma_block<true><<<n_runs, CMA_NT, smem, st>>>(Rx, ys, h, out, e, N, M, sps, R, lr, train, mode, batchlen, symb_step, ycap);
        } else {
            k_cma_block<false><<<n_runs, CMA_NT, smem, st>>>(Rx, ys, h, out, e, N, M, sps, R, lr, train, mode, batchlen, symb_step, 0);
        }
        ktime_end(VAEQ_K_CMA, st);
        VAEQ_LAUNCH_CHECK("k_cma_block");
    }
    return VAEQ_OK;
}

extern "C" size_t vaeq_cpe_runs_scratch_bytes(int32_t N, int32_t n_runs) {
    if (N <= 0 || n_runs <= 0) return 0;
    return align_up((size_t)n_runs * 4 * N * sizeof(float), 256) + align_up((size_t)n_runs * 2 * N * sizeof(float), 256);
}
extern "C" size_t vaeq_cpe_scratch_bytes(int32_t N) { return vaeq_cpe_runs_scratch_bytes(N, 1); }

// npol independent complex sequences y (npol,2,N): 4th power, 501-tap moving average, phase (/4), optional unwrap, derotation
static int cpe_launch(const float *y, int N, int npol, int unwrap, float *y_corr, void *scratch, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    float *p4 = static_cast<float *>(scratch);
    float *phi = reinterpret_cast<float *>(static_cast<char *>(scratch) + align_up((size_t)npol * 2 * N * sizeof(float), 256));
    const int grid = (int)std::min<int64_t>((npol * (int64_t)N + 255) / 256, (int64_t)sm_count() * 16);
    ktime_begin(VAEQ_K_CMA, st);
    k_cpe_pow4<<<grid, 256, 0, st>>>(y, N, npol, p4);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cpe_pow4");
    ktime_begin(VAEQ_K_CMA, st);
    k_cpe_phase<<<dim3((N + CPE_TILE - 1) / CPE_TILE, npol), CPE_PT, 0, st>>>(p4, N, npol, phi);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cpe_phase");
    ktime_begin(VAEQ_K_CMA, st);
    k_cpe_unwrap_rotate<<<npol, 1024, 0, st>>>(y, phi, N, y_corr, unwrap);
    ktime_end(VAEQ_K_CMA, st);
    VAEQ_LAUNCH_CHECK("k_cpe_unwrap_rotate");
    return VAEQ_OK;
}

extern "C" int vaeq_cpe_runs(const float *y, int32_t N, int32_t n_runs, float *y_corr, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(y && y_corr && scratch && N > 1 && n_runs > 0 && n_runs <= 32767, "bad cpe arguments");
    return cpe_launch(y, N, 2 * n_runs, 1, y_corr, scratch, stream);
}

// CPE of the AWGN module (AWGN_channel/func_CMA_MQAM_shaping.py:170-196): single polarisation y (n_runs,2,N), NO unwrapping
extern "C" int vaeq_cpe_awgn(const float *y, int32_t N, int32_t n_runs, float *y_corr, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(y && y_corr && scratch && N > 1 && n_runs > 0 && n_runs <= 65535, "bad cpe arguments");
    return cpe_launch(y, N, n_runs, 0, y_corr, scratch, stream);
}

extern "C" int vaeq_cpe(const float *y, int32_t N, float *y_corr, void *scratch, void *stream) {
    return vaeq_cpe_runs(y, N, 1, y_corr, scratch, stream);
}
