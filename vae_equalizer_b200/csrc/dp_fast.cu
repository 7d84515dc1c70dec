// DP VAE step, fast path for B200 (sm_100a): register-blocked tap contractions.
//
// Same math as dp_step.cu (the generic kernels remain the fallback and the in-library cross-check), organised
// for FFMA issue rate, which -- not HBM -- bounds this step at M_est = 25 (DESIGN.md "Roofline"):
//   * a thread owns FT_R = 4 CONSECUTIVE symbols, so a K-lag contraction slides a register window over
//     K + 3 shared-memory float4 loads for 64 K FFMA (the generic kernel does 1 LDS per 2-4 FFMA);
//   * rx is split into its even / odd sample phases (sps = 2), each stored as float4 {I0,Q0,I1,Q1} per
//     position, so every lag of the stride-2 FIR is a unit-stride access; the float4 index is padded
//     (i + i/4) so that the "lane owns 4 consecutive positions" pattern is bank-conflict free with
//     compile-time offsets;
//   * scratch that crosses kernels (residual e, moments, dL/dout) is SoA rows like q, so all global traffic
//     is float4 per thread, 512 B per warp instruction;
//   * point-wise math uses ex2/lg2/rcp approximations, a reciprocal multiply for 1/(2 var) and the analytic
//     log q = (z_min - z_l) - log2 s for the entropy, ~100 instructions per component instead of ~450;
//   * the backward pass is two kernels (dE_q + softmin backward + dh, then dW) so that 2-3 CTAs fit per SM.
// Reference lines: twoXtwoFIR.forward sf:500-527, loss_function_shaping sf:92-137 (sf = optical_DP_channel/shared_funcs.py).
#include "dp_math.cuh"
#include "dp_kernels.cuh"

namespace vaeq {

constexpr int FT_NT = 256;                 // threads per CTA
constexpr int FT_R = 4;                    // consecutive symbols per thread
constexpr int FT_TE = FT_NT * FT_R;        // symbols per tile incl. halo (1024)
constexpr int FT_HP = 8;                   // halo per side in symbols (>= MH/2, multiple of 4)
constexpr int FT_T = FT_TE - 2 * FT_HP;    // owned symbols per tile (1008)
constexpr int FT_XOFF = 8;                 // extra margin of the x phase arrays (FIR reaches MH/2 further)
constexpr int FT_XN = FT_TE + 2 * FT_XOFF; // logical length of xe / xo
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__host__ __device__ constexpr int fdiv4(int c) { return c >= 0 ? c / 4 : -((3 - c) / 4); }
// padded float4 index of logical position 4*l + c (c compile-time, may be negative)
__host__ __device__ constexpr int poff(int c) { return c + fdiv4(c); }
__device__ __forceinline__ int pidx(int i) { return i + (i >> 2); }
constexpr int FT_XS = FT_XN + FT_XN / 4 + 4;   // padded lengths (float4)
constexpr int FT_ES = FT_TE + FT_TE / 4 + 4;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct FastConst {
    float amp[VAEQ_MAX_LEVELS], a2[VAEQ_MAX_LEVELS];
    float nua2l[VAEQ_MAX_LEVELS];          // nu_sc a^2 log2(e)
    float lgP[VAEQ_MAX_LEVELS];            // log2 P_l
    float c2[2];                           // log2(e) / (2 var_p)
    float inv_var[2];
};

__device__ __forceinline__ void load_fast_const(FastConst *c, const DpK &p, int n_lev) {
    const int t = threadIdx.x;
    if (t < VAEQ_MAX_LEVELS) {
        const float a = t < n_lev ? p.amp[t] : 0.f;
        c->amp[t] = a;
        c->a2[t] = a * a;
        c->nua2l[t] = p.nu_sc * (a * a) * LOG2E;
        c->lgP[t] = t < n_lev ? log2f(p.P[t]) : 0.f;
    }
    if (t < 2) {
        c->c2[t] = LOG2E / (2.f * p.var[t]);
        c->inv_var[t] = 1.f / p.var[t];
    }
}

// soft demapper for one component: q, first two moments and the entropy term  sum_l -q_l ln(q_l / P_l)
template <int NL>
__device__ __forceinline__ void demap_fast(float y, float c2, const FastConst &c, float (&q)[NL], float &m1, float &m2,
                                           float &ent) {
    float z[NL];
    float zmin = 3.0e38f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const float d = y - c.amp[l];
        z[l] = fmaf(d * c2, d, c.nua2l[l]);                 // ((y-a)^2/(2 var) + nu_sc a^2) * log2 e   (sf:521)
        zmin = fminf(zmin, z[l]);
    }
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        z[l] = zmin - z[l];                                 // log2 of the unnormalised posterior
        q[l] = ex2_approx(z[l]);
        s += q[l];
    }
    const float r = rcp_approx(s), lgs = lg2_approx(s);
    m1 = 0.f;
    m2 = 0.f;
    float e = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        q[l] *= r;
        m1 = fmaf(c.amp[l], q[l], m1);
        m2 = fmaf(c.a2[l], q[l], m2);
        e = fmaf(q[l], (z[l] - lgs) - c.lgP[l], e);          // q log2(q/P); the 1e-12 of sf:132 only matters where q/P < 1e-9
    }
    ent = -LN2 * e;
}

// softmin + moments + entropy backward for one component (closed form in oracle/closed_form.py)
template <int NL>
__device__ __forceinline__ float demap_backward_fast(float y, float inv_var, const FastConst &c, const float (&q)[NL],
                                                     float g1, float g2, float ent_w) {
    float gq[NL];
    float dot = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float g = fmaf(c.amp[l], g1, c.a2[l] * g2);
        // d/dq [ q ln(q/P + eps) ] = ln(q/P + eps) + (q/P)/(q/P + eps); the second term is a constant 1 wherever
        // q matters and a per-symbol constant cancels in (gq - dot) because sum_l q_l = 1
        g = fmaf(ent_w, lg2_approx(fmaxf(q[l], 1e-37f)) - c.lgP[l], g);
        gq[l] = g;
        dot = fmaf(q[l], g, dot);
    }
    float gy = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) gy = fmaf(q[l] * (dot - gq[l]), y - c.amp[l], gy);
    return gy * inv_var;
}

// ---------------------------------------------------------------------------------------------
// FIR-like contraction: 4 consecutive outputs per thread, NLAG lags, 2x2 complex taps.
//   acc[r][2o+c] += sum_i tap(o,i) * win[4l + r + a + C0]_i     (complex product, taps = {t00r,t00i,t01r,t01i | t10..t11})
// win is a padded float4 array, pb = 5*l its base for this thread, taps: 2 float4 per lag (smem broadcast).
// ---------------------------------------------------------------------------------------------
template <int NLAG, int C0>
__device__ __forceinline__ void fir4(const float4 *__restrict__ win, int pb, const float4 *__restrict__ taps,
                                     float (&acc)[FT_R][4]) {
    float4 w[4];
#pragma unroll
    for (int r = 0; r < 3; ++r) w[r] = win[pb + poff(C0 + r)];
#pragma unroll
    for (int a = 0; a < NLAG; ++a) {
        w[(a + 3) & 3] = win[pb + poff(C0 + a + 3)];
        const float4 t0 = taps[2 * a], t1 = taps[2 * a + 1];
#pragma unroll
        for (int r = 0; r < FT_R; ++r) {
            const float4 x = w[(a + r) & 3];
            acc[r][0] = fmaf(t0.x, x.x, fmaf(-t0.y, x.y, fmaf(t0.z, x.z, fmaf(-t0.w, x.w, acc[r][0]))));
            acc[r][1] = fmaf(t0.x, x.y, fmaf(t0.y, x.x, fmaf(t0.z, x.w, fmaf(t0.w, x.z, acc[r][1]))));
            acc[r][2] = fmaf(t1.x, x.x, fmaf(-t1.y, x.y, fmaf(t1.z, x.z, fmaf(-t1.w, x.w, acc[r][2]))));
            acc[r][3] = fmaf(t1.x, x.y, fmaf(t1.y, x.x, fmaf(t1.z, x.w, fmaf(t1.w, x.z, acc[r][3]))));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tap-gradient correlation: acc[a][2(2o+i)+c] += sum_r g[r]_o * conj(win[4l + r + a + C0]_i)
// ---------------------------------------------------------------------------------------------
template <int A, int C0>
__device__ __forceinline__ void corr4(const float4 *__restrict__ win, int pb, const float4 (&g)[FT_R], float (&acc)[A][8]) {
#pragma unroll
    for (int cpos = 0; cpos < A + FT_R - 1; ++cpos) {
        const float4 x = win[pb + poff(C0 + cpos)];
#pragma unroll
        for (int r = 0; r < FT_R; ++r) {
            const int a = cpos - r;
            if (a >= 0 && a < A) {
                const float4 gg = g[r];
                acc[a][0] = fmaf(gg.x, x.x, fmaf(gg.y, x.y, acc[a][0]));      // (o0,i0) re
                acc[a][1] = fmaf(gg.y, x.x, fmaf(-gg.x, x.y, acc[a][1]));     //         im
                acc[a][2] = fmaf(gg.x, x.z, fmaf(gg.y, x.w, acc[a][2]));      // (o0,i1)
                acc[a][3] = fmaf(gg.y, x.z, fmaf(-gg.x, x.w, acc[a][3]));
                acc[a][4] = fmaf(gg.z, x.x, fmaf(gg.w, x.y, acc[a][4]));      // (o1,i0)
                acc[a][5] = fmaf(gg.w, x.x, fmaf(-gg.z, x.y, acc[a][5]));
                acc[a][6] = fmaf(gg.z, x.z, fmaf(gg.w, x.w, acc[a][6]));      // (o1,i1)
                acc[a][7] = fmaf(gg.w, x.z, fmaf(-gg.z, x.w, acc[a][7]));
            }
        }
    }
}

// load the even/odd phase arrays of rx for the tile starting at symbol t0 (logical position 0 <-> symbol t0-HP-XOFF)
__device__ __forceinline__ void load_x_phases(const DpK &p, int t0, float4 *xe, float4 *xo) {
    const int sym0 = t0 - FT_HP - FT_XOFF;
    for (int j = threadIdx.x; j < FT_XN / 2; j += FT_NT) {
        const int64_t s0 = 2 * (int64_t)sym0 + 4 * j;           // first of 4 consecutive samples
        float v[4][4];
        if (s0 >= 0 && s0 + 3 < p.L) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(p.rx + (int64_t)r * p.ld_rx + s0));
                v[r][0] = t.x; v[r][1] = t.y; v[r][2] = t.z; v[r][3] = t.w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t s = s0 + k;
                    v[r][k] = (s >= 0 && s < p.L) ? p.rx[(int64_t)r * p.ld_rx + s] : 0.f;
                }
        }
        xe[pidx(2 * j)] = make_float4(v[0][0], v[1][0], v[2][0], v[3][0]);
        xo[pidx(2 * j)] = make_float4(v[0][1], v[1][1], v[2][1], v[3][1]);
        xe[pidx(2 * j + 1)] = make_float4(v[0][2], v[1][2], v[2][2], v[3][2]);
        xo[pidx(2 * j + 1)] = make_float4(v[0][3], v[1][3], v[2][3], v[3][3]);
    }
}

// SoA row helpers: 4 consecutive symbols of one row as a float4
__device__ __forceinline__ float4 ld_row4(const float *base, int64_t ld, int row, int u) {
    return __ldg(reinterpret_cast<const float4 *>(base + (int64_t)row * ld + u));
}
__device__ __forceinline__ void st_row4(float *base, int64_t ld, int row, int u, float4 v) {
    *reinterpret_cast<float4 *>(base + (int64_t)row * ld + u) = v;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NL, int MH>
__global__ void __launch_bounds__(FT_NT, 2) k_dp_fwd_fast(DpK p) {
    static_assert(MH % 2 == 0 && MH / 2 <= FT_HP - 2, "fast path needs M_est = 1 (mod 4) and M_est <= 25");
    constexpr int M = 2 * MH + 1, HF = MH / 2, NE = MH + 1, NO = MH;
    extern __shared__ __align__(16) float4 smem4[];
    float4 *xe = smem4, *xo = xe + FT_XS, *m1s = xo + FT_XS;
    float4 *tapF = m1s + FT_ES;                              // FIR taps: [phase][lag][2]
    float4 *tapD = tapF + 2 * (NE + NO);                     // channel taps, reversed per phase
    FastConst *cst = reinterpret_cast<FastConst *>(tapD + 2 * (NE + NO));
    float *red = reinterpret_cast<float *>(cst + 1);
    const int tid = threadIdx.x;

    // tap tables (see file header of dp_step.cu for the W / h layouts)
    for (int idx = tid; idx < (NE + NO) * 8; idx += FT_NT) {
        const int e = idx & 7, la = idx >> 3;
        const int ph = la >= NE, a = ph ? la - NE : la;
        const int o = e >> 2, i = (e >> 1) & 1, c = e & 1;
        const int k = 2 * a + ph;
        reinterpret_cast<float *>(tapF)[idx] = p.W[(o * 4 + 2 * c + i) * M + k];
        const int j = ph ? (2 * MH - 1 - 2 * a) : (2 * MH - 2 * a);
        reinterpret_cast<float *>(tapD)[idx] = p.h[((o * 2 + i) * 2 + c) * M + j];
    }
    load_fast_const(cst, p, NL);
    __syncthreads();
    const FastConst &c = *cst;
    const float4 *tFe = tapF, *tFo = tapF + 2 * NE, *tDe = tapD, *tDo = tapD + 2 * NE;

    float accC[2] = {0.f, 0.f}, accEnt = 0.f, accV[2] = {0.f, 0.f};
    const int pb = 5 * tid;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int t0 = tile * FT_T;
        load_x_phases(p, t0, xe, xo);
        __syncthreads();

        const int i0 = FT_R * tid;                           // local symbol index of this thread's first symbol
        const int u0 = t0 - FT_HP + i0;
        const bool in_seq = (u0 >= 0) && (u0 < p.B);         // B % 4 == 0: all four symbols in or out together
        const bool owned = in_seq && (i0 >= FT_HP) && (i0 < FT_HP + FT_T);
        float4 mom[FT_R];
#pragma unroll
        for (int r = 0; r < FT_R; ++r) mom[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_seq) {
            float y[FT_R][4];
#pragma unroll
            for (int r = 0; r < FT_R; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) y[r][k] = 0.f;
            // x[2u + k - MH]: even k = 2a -> xe[u + a - HF], odd k = 2a+1 -> xo[u + a - HF]
            fir4<NE, FT_XOFF - HF>(xe, pb, tFe, y);
            fir4<NO, FT_XOFF - HF>(xo, pb, tFo, y);
            float vs[FT_R][2];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float qv[FT_R][NL];
                float m1v[FT_R];
#pragma unroll
                for (int r = 0; r < FT_R; ++r) {
                    float m2, ent;
                    demap_fast<NL>(y[r][cc], c.c2[cc >> 1], c, qv[r], m1v[r], m2, ent);
                    const int u = u0 + r;
                    if (owned && u >= MH && u < p.B - MH) accEnt += ent;                  // sf:132
                    const float v = m2 - m1v[r] * m1v[r];                                 // sf:113
                    if (cc & 1) vs[r][cc >> 1] += v; else vs[r][cc >> 1] = v;
                }
                if (owned) {
#pragma unroll
                    for (int l = 0; l < NL; ++l)
                        st_row4(p.q, p.ld_q, cc * NL + l, u0, make_float4(qv[0][l], qv[1][l], qv[2][l], qv[3][l]));
                    st_row4(p.out, p.ld_out, cc, u0, make_float4(y[0][cc], y[1][cc], y[2][cc], y[3][cc]));
                    st_row4(p.m1rows, p.B, cc, u0, make_float4(m1v[0], m1v[1], m1v[2], m1v[3]));
                    if (p.qk != nullptr) {
#pragma unroll
                        for (int r = 0; r < FT_R; ++r) {
                            const int u = u0 + r;
                            if (u >= p.keep_lo && u < p.keep_lo + p.keep_n) {
                                const int64_t col = p.keep_base + (u - p.keep_lo);
#pragma unroll
                                for (int l = 0; l < NL; ++l) p.qk[(int64_t)(cc * NL + l) * p.ld_qk + col] = qv[r][l];
                                p.outk[(int64_t)cc * p.ld_outk + col] = y[r][cc];
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < FT_R; ++r) reinterpret_cast<float *>(&mom[r])[cc] = m1v[r];
            }
            if (owned) {
#pragma unroll
                for (int r = 0; r < FT_R; ++r) {
                    const int u = u0 + r;
                    accV[0] += vs[r][0];
                    accV[1] += vs[r][1];
                    if (u < MH || u >= p.B - MH) {
                        const int slot = (u < MH) ? u : MH + (u - (p.B - MH));
                        p.edge_vs[slot] = vs[r][0];
                        p.edge_vs[2 * MH + slot] = vs[r][1];
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < FT_R; ++r) m1s[pb + r] = mom[r];
        __syncthreads();

        // ---- D = h * E_q for the owned samples, residual e = D - rx ------------------------------------
        if (owned) {
            float de[FT_R][4], dod[FT_R][4];
#pragma unroll
            for (int r = 0; r < FT_R; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) de[r][k] = dod[r][k] = 0.f;
            fir4<NE, -HF>(m1s, pb, tDe, de);                 // even samples: sum_a h[2MH-2a] E_q[u + a - HF]
            fir4<NO, -HF + 1>(m1s, pb, tDo, dod);            // odd samples:  sum_b h[2MH-1-2b] E_q[u + b - HF + 1]
            float ee[4][FT_R], eo[4][FT_R];
#pragma unroll
            for (int r = 0; r < FT_R; ++r) {
                const int u = u0 + r;
                const float4 xr0 = xe[pb + poff(FT_XOFF + r)], xr1 = xo[pb + poff(FT_XOFF + r)];
                const bool v0 = (2 * u >= MH) && (2 * u < p.L - MH), v1 = (2 * u + 1 >= MH) && (2 * u + 1 < p.L - MH);
                const float e0[4] = {de[r][0] - xr0.x, de[r][1] - xr0.y, de[r][2] - xr0.z, de[r][3] - xr0.w};
                const float e1[4] = {dod[r][0] - xr1.x, dod[r][1] - xr1.y, dod[r][2] - xr1.z, dod[r][3] - xr1.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ee[k][r] = v0 ? e0[k] : 0.f;
                    eo[k][r] = v1 ? e1[k] : 0.f;
                    accC[k >> 1] += ee[k][r] * ee[k][r] + eo[k][r] * eo[k][r];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                st_row4(p.erows, p.B, k, u0, make_float4(ee[k][0], ee[k][1], ee[k][2], ee[k][3]));
                st_row4(p.erows, p.B, 4 + k, u0, make_float4(eo[k][0], eo[k][1], eo[k][2], eo[k][3]));
            }
        }
        __syncthreads();
    }

    float v[5] = {accC[0], accC[1], accEnt, accV[0], accV[1]};
    block_sum<5>(v, red);
    if (tid == 0) {
        double *dst = p.part_fwd + (int64_t)blockIdx.x * 8;
#pragma unroll
        for (int i = 0; i < 5; ++i) dst[i] = (double)v[i];
    }
}

// reduce the per-thread tap-gradient accumulators of one role over the CTA and add them to gpart
template <int A>
__device__ __forceinline__ void reduce_role(float (&acc)[A][8], float *red /* 8 warps x A*8 */, int role, int n_real,
                                            float *dst_base, int M, int fam_is_h, int ph, int a0, int MH) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = warp_sum(acc[a][k]);
            if (lane == 0) red[wid * (A * 8) + a * 8 + k] = s;
        }
    __syncthreads();
    // warps w and w+4 share a role
    if (wid < 4 && wid == role) {
        for (int idx = lane; idx < n_real * 8; idx += 32) {
            const int a = idx >> 3, k = idx & 7, oi = k >> 1, c = k & 1, o = oi >> 1, i = oi & 1;
            const float s = red[wid * (A * 8) + idx] + red[(wid + 4) * (A * 8) + idx];
            const int lag = a0 + a;
            if (fam_is_h) {                                   // dh[chi=o][nu=i][c][j], lags mirrored like the D taps
                const int j = ph ? (2 * MH - 1 - 2 * lag) : (2 * MH - 2 * lag);
                dst_base[8 * M + ((o * 2 + i) * 2 + c) * M + j] = s;
            } else {                                          // dW[o][i | 2+i][k]
                const int kk = 2 * lag + ph;
                dst_base[(o * 4 + 2 * c + i) * M + kk] = s;
            }
        }
    }
}

// roles: 0 = even phase lags [0, A0e), 1 = even [A0e, NE), 2 = odd [0, A0o), 3 = odd [A0o, NO)
template <int MH>
struct Roles {
    static constexpr int NE = MH + 1, NO = MH;
    static constexpr int A0e = (NE + 1) / 2, A1e = NE - A0e, A0o = (NO + 1) / 2, A1o = NO - A0o;
    static constexpr int AMAX = A0e;
};

// ---------------------------------------------------------------------------------------------
// backward 1: dL/dE_q (FIR-like over gD), softmin backward -> dL/dout rows, dh partials
// ---------------------------------------------------------------------------------------------
template <int NL, int MH>
__global__ void __launch_bounds__(FT_NT, 2) k_dp_bwd1_fast(DpK p) {
    constexpr int M = 2 * MH + 1, HF = MH / 2, NE = MH + 1, NO = MH;
    using RL = Roles<MH>;
    extern __shared__ __align__(16) float4 smem4[];
    float4 *ge = smem4, *go = ge + FT_ES, *m1s = go + FT_ES;
    float4 *tapG = m1s + FT_ES;                              // conj(h) taps for dE_q: [phase][lag][2]
    FastConst *cst = reinterpret_cast<FastConst *>(tapG + 2 * (NE + NO));
    float *PSg = reinterpret_cast<float *>(cst + 1);         // (2, M+1)
    float *red = PSg + 2 * (M + 1) + 2;                      // 8 * AMAX * 8
    const int tid = threadIdx.x, wid = tid >> 5;
    const float kap0 = p.scal[DP_KAPPA_OFF], kap1 = p.scal[DP_KAPPA_OFF + 1];

    for (int idx = tid; idx < (NE + NO) * 8; idx += FT_NT) {
        const int e = idx & 7, la = idx >> 3;
        const int ph = la >= NE, a = ph ? la - NE : la;
        const int nu = e >> 2, chi = (e >> 1) & 1, c = e & 1;     // out index nu, in index chi
        const int j = 2 * a + ph;
        const float hv = p.h[((chi * 2 + nu) * 2 + c) * M + j];
        reinterpret_cast<float *>(tapG)[idx] = c ? -hv : hv;       // conj(h)
    }
    load_fast_const(cst, p, NL);
    if (tid < 2) {
        const int nu = tid;
        float a = 0.f;
        PSg[nu * (M + 1)] = 0.f;
        for (int j = 0; j < M; ++j) {
            const float h0r = p.h[((0 * 2 + nu) * 2 + 0) * M + j], h0i = p.h[((0 * 2 + nu) * 2 + 1) * M + j];
            const float h1r = p.h[((1 * 2 + nu) * 2 + 0) * M + j], h1i = p.h[((1 * 2 + nu) * 2 + 1) * M + j];
            a += kap0 * (h0r * h0r + h0i * h0i) + kap1 * (h1r * h1r + h1i * h1i);
            PSg[nu * (M + 1) + j + 1] = a;
        }
    }
    __syncthreads();
    const FastConst &c = *cst;
    const float4 *tGe = tapG, *tGo = tapG + 2 * NE;
    const int pb = 5 * tid;
    const int role = wid & 3;

    float acc[RL::AMAX][8];
#pragma unroll
    for (int a = 0; a < RL::AMAX; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int t0 = tile * FT_T;
        const int i0 = FT_R * tid, u0 = t0 - FT_HP + i0;
        const bool in_seq = (u0 >= 0) && (u0 < p.B);
        const bool owned = in_seq && (i0 >= FT_HP) && (i0 < FT_HP + FT_T);
        // ---- stage gD = 2 kappa e (tile + halo) and E_q (tile + halo) -----------------------------------
        {
            float4 er[8], mr[4];
            if (in_seq) {
#pragma unroll
                for (int k = 0; k < 8; ++k) er[k] = ld_row4(p.erows, p.B, k, u0);
#pragma unroll
                for (int k = 0; k < 4; ++k) mr[k] = ld_row4(p.m1rows, p.B, k, u0);
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) er[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) mr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float s0 = 2.f * kap0, s1 = 2.f * kap1;
#pragma unroll
            for (int r = 0; r < FT_R; ++r) {
                auto comp = [r](const float4 &v) { return reinterpret_cast<const float *>(&v)[r]; };
                ge[pb + r] = make_float4(s0 * comp(er[0]), s0 * comp(er[1]), s1 * comp(er[2]), s1 * comp(er[3]));
                go[pb + r] = make_float4(s0 * comp(er[4]), s0 * comp(er[5]), s1 * comp(er[6]), s1 * comp(er[7]));
                m1s[pb + r] = make_float4(comp(mr[0]), comp(mr[1]), comp(mr[2]), comp(mr[3]));
            }
        }
        __syncthreads();

        if (owned) {
            float gE[FT_R][4];
#pragma unroll
            for (int r = 0; r < FT_R; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) gE[r][k] = 0.f;
            // dL/dE_q(u) = sum_chi sum_j conj(h[chi][nu][j]) gD_chi(2u - MH + j): even j -> ge[u+a-HF], odd -> go[u+a-HF]
            fir4<NE, -HF>(ge, pb, tGe, gE);
            fir4<NO, -HF>(go, pb, tGo, gE);
            float gy[4][FT_R];
            float gV[2][FT_R], entw[FT_R];
#pragma unroll
            for (int r = 0; r < FT_R; ++r) {
                const int u = u0 + r;
                const int jlo = max(0, 2 * MH - 2 * u), jhi = min(M, p.L - 2 * u);
                gV[0][r] = PSg[jhi] - PSg[jlo];
                gV[1][r] = PSg[(M + 1) + jhi] - PSg[(M + 1) + jlo];
                entw[r] = (u >= MH && u < p.B - MH) ? LN2 : 0.f;
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float4 q4[NL];
#pragma unroll
                for (int l = 0; l < NL; ++l) q4[l] = ld_row4(p.q, p.ld_q, cc * NL + l, u0);
                const float4 y4 = ld_row4(p.out, p.ld_out, cc, u0);
#pragma unroll
                for (int r = 0; r < FT_R; ++r) {
                    float q[NL];
#pragma unroll
                    for (int l = 0; l < NL; ++l) q[l] = reinterpret_cast<const float *>(&q4[l])[r];
                    const float y = reinterpret_cast<const float *>(&y4)[r];
                    const float m1 = reinterpret_cast<const float *>(&m1s[pb + r])[cc];
                    const float g2 = gV[cc >> 1][r];
                    const float g1 = gE[r][cc] - 2.f * m1 * g2;
                    gy[cc][r] = demap_backward_fast<NL>(y, c.inv_var[cc >> 1], c, q, g1, g2, entw[r]);
                }
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) st_row4(p.gyrows, p.B, cc, u0, make_float4(gy[cc][0], gy[cc][1], gy[cc][2], gy[cc][3]));
        }

        // ---- dh partials: roles (phase, lag half); warp w handles symbol groups of parity w>>2 ----------
        {
            // this thread's 4 symbols are handled for ALL roles by the 8 warps in turn: loop over the 8 groups of 128 symbols
            for (int g = (wid >> 2); g < FT_NT / 32; g += 2) {
                const int l = g * 32 + (tid & 31);           // thread-slot whose 4 symbols we process
                const int li0 = FT_R * l, uu0 = t0 - FT_HP + li0;
                const bool own = (uu0 >= 0) && (uu0 < p.B) && (li0 >= FT_HP) && (li0 < FT_HP + FT_T);
                if (!own) continue;                          // halo slots carry no owned samples (and their windows leave the tile)
                const int pbl = 5 * l;
                float4 gd[FT_R];
                const float4 *src = (role < 2) ? ge : go;
#pragma unroll
                for (int r = 0; r < FT_R; ++r) gd[r] = src[pbl + r];
                // window over E_q[v + a' - HF] (even phase) or E_q[v + a'' - HF + 1] (odd phase)
                if (role == 0) corr4<RL::AMAX, -HF>(m1s, pbl, gd, acc);
                else if (role == 1) corr4<RL::AMAX, -HF + RL::A0e>(m1s, pbl, gd, acc);
                else if (role == 2) corr4<RL::AMAX, -HF + 1>(m1s, pbl, gd, acc);
                else corr4<RL::AMAX, -HF + 1 + RL::A0o>(m1s, pbl, gd, acc);
            }
        }
        __syncthreads();
    }
    const int n_real = role == 0 ? RL::A0e : role == 1 ? RL::A1e : role == 2 ? RL::A0o : RL::A1o;
    const int a0 = role == 0 ? 0 : role == 1 ? RL::A0e : role == 2 ? 0 : RL::A0o;
    reduce_role<RL::AMAX>(acc, red, role, n_real, p.gpart + (int64_t)blockIdx.x * 16 * M, M, 1, role >> 1, a0, MH);
}

// ---------------------------------------------------------------------------------------------
// backward 2: dW partials = correlation of dL/dout with the rx phases
// ---------------------------------------------------------------------------------------------
template <int MH>
__global__ void __launch_bounds__(FT_NT, 2) k_dp_bwd2_fast(DpK p) {
    constexpr int M = 2 * MH + 1, HF = MH / 2;
    using RL = Roles<MH>;
    extern __shared__ __align__(16) float4 smem4[];
    float4 *xe = smem4, *xo = xe + FT_XS, *gys = xo + FT_XS;
    float *red = reinterpret_cast<float *>(gys + FT_ES);
    const int tid = threadIdx.x, wid = tid >> 5, role = wid & 3;
    const int pb = 5 * tid;

    float acc[RL::AMAX][8];
#pragma unroll
    for (int a = 0; a < RL::AMAX; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int t0 = tile * FT_T;
        load_x_phases(p, t0, xe, xo);
        {
            const int i0 = FT_R * tid, u0 = t0 - FT_HP + i0;
            const bool owned = (u0 >= 0) && (u0 < p.B) && (i0 >= FT_HP) && (i0 < FT_HP + FT_T);
            float4 gr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) gr[k] = owned ? ld_row4(p.gyrows, p.B, k, u0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < FT_R; ++r) {
                auto comp = [r](const float4 &v) { return reinterpret_cast<const float *>(&v)[r]; };
                gys[pb + r] = make_float4(comp(gr[0]), comp(gr[1]), comp(gr[2]), comp(gr[3]));
            }
        }
        __syncthreads();
        for (int g = (wid >> 2); g < FT_NT / 32; g += 2) {
            const int l = g * 32 + (tid & 31);
            const int li0 = FT_R * l, uu0 = t0 - FT_HP + li0;
            if (!((uu0 >= 0) && (uu0 < p.B) && (li0 >= FT_HP) && (li0 < FT_HP + FT_T))) continue;
            const int pbl = 5 * l;
            float4 gd[FT_R];
#pragma unroll
            for (int r = 0; r < FT_R; ++r) gd[r] = gys[pbl + r];
            // dW[o][i][k = 2a + ph] = sum_u gy_o(u) conj(x_ph,i[u + a - HF])
            if (role == 0) corr4<RL::AMAX, FT_XOFF - HF>(xe, pbl, gd, acc);
            else if (role == 1) corr4<RL::AMAX, FT_XOFF - HF + RL::A0e>(xe, pbl, gd, acc);
            else if (role == 2) corr4<RL::AMAX, FT_XOFF - HF>(xo, pbl, gd, acc);
            else corr4<RL::AMAX, FT_XOFF - HF + RL::A0o>(xo, pbl, gd, acc);
        }
        __syncthreads();
    }
    const int n_real = role == 0 ? RL::A0e : role == 1 ? RL::A1e : role == 2 ? RL::A0o : RL::A1o;
    const int a0 = role == 0 ? 0 : role == 1 ? RL::A0e : role == 2 ? 0 : RL::A0o;
    reduce_role<RL::AMAX>(acc, red, role, n_real, p.gpart + (int64_t)blockIdx.x * 16 * M, M, 0, role >> 1, a0, MH);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int MH>
static size_t fast_smem_fwd() {
    return (size_t)(2 * FT_XS + FT_ES + 4 * (2 * MH + 1)) * sizeof(float4) + sizeof(FastConst) + 5 * 32 * sizeof(float) + 64;
}
template <int MH>
static size_t fast_smem_bwd1() {
    return (size_t)(3 * FT_ES + 2 * (2 * MH + 1)) * sizeof(float4) + sizeof(FastConst) +
           (2 * (2 * MH + 2) + 2 + 8 * Roles<MH>::AMAX * 8) * sizeof(float) + 64;
}
template <int MH>
static size_t fast_smem_bwd2() {
    return (size_t)(2 * FT_XS + FT_ES) * sizeof(float4) + 8 * Roles<MH>::AMAX * 8 * sizeof(float) + 64;
}

template <int NL, int MH>
static int dp_run_fast_t(DpK p, int mode, cudaStream_t st, int *grid_bwd_out) {
    static int gridF = 0, gridB = 0;
    const size_t sf = fast_smem_fwd<MH>(), s1 = fast_smem_bwd1<MH>(), s2 = fast_smem_bwd2<MH>();
    if (!gridF) {
        VAEQ_CUDA(cudaFuncSetAttribute(k_dp_fwd_fast<NL, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf));
        VAEQ_CUDA(cudaFuncSetAttribute(k_dp_bwd1_fast<NL, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
        VAEQ_CUDA(cudaFuncSetAttribute(k_dp_bwd2_fast<MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
        int a = 0, b = 0, c = 0;
        VAEQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dp_fwd_fast<NL, MH>, FT_NT, sf));
        VAEQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_dp_bwd1_fast<NL, MH>, FT_NT, s1));
        VAEQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_dp_bwd2_fast<MH>, FT_NT, s2));
        gridF = max(1, a) * sm_count();
        gridB = max(1, min(b, c)) * sm_count();
    }
    p.T = FT_T;
    p.ntiles = (p.B + FT_T - 1) / FT_T;
    const int gf = min(min(gridF, DP_GRID_CAP), p.ntiles), gb = min(min(gridB, DP_GRID_CAP), p.ntiles);
    ktime_begin(VAEQ_K_DP_FWD, st);
    k_dp_fwd_fast<NL, MH><<<gf, FT_NT, sf, st>>>(p);
    ktime_end(VAEQ_K_DP_FWD, st);
    VAEQ_LAUNCH_CHECK("k_dp_fwd_fast");
    dp_launch_fin(p, gf, st);
    if (mode == DP_MODE_FWD) return VAEQ_OK;
    ktime_begin(VAEQ_K_DP_BWD, st);
    k_dp_bwd1_fast<NL, MH><<<gb, FT_NT, s1, st>>>(p);
    ktime_end(VAEQ_K_DP_BWD, st);
    VAEQ_LAUNCH_CHECK("k_dp_bwd1_fast");
    ktime_begin(VAEQ_K_DP_BWD2, st);
    k_dp_bwd2_fast<MH><<<gb, FT_NT, s2, st>>>(p);
    ktime_end(VAEQ_K_DP_BWD2, st);
    VAEQ_LAUNCH_CHECK("k_dp_bwd2_fast");
    *grid_bwd_out = gb;
    return VAEQ_OK;
}

// returns 1 if the fast path ran (and *grid_bwd_out is the number of gradient partials), 0 if not applicable
int dp_try_fast(const DpK &p, int n_lev, int mode, cudaStream_t st, int *grid_bwd_out, int *rc) {
    const bool aligned = (p.B % 4 == 0) && (p.ld_rx % 4 == 0) && (p.ld_q % 4 == 0) && (p.ld_out % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.rx) | reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.out)) % 16 == 0);
    if (!aligned || p.B < 2 * FT_T) return 0;                // small batches: the generic kernels (one tile) are as good
    *rc = VAEQ_OK;
#define FAST_CASE(NL_, MH_)                                             \
    if (n_lev == NL_ && p.mh == MH_) {                                  \
        *rc = dp_run_fast_t<NL_, MH_>(p, mode, st, grid_bwd_out);       \
        return 1;                                                       \
    }
    FAST_CASE(8, 12)
    FAST_CASE(8, 6)
    FAST_CASE(8, 4)
    FAST_CASE(8, 2)
    FAST_CASE(4, 12)
    FAST_CASE(2, 12)
#undef FAST_CASE
    return 0;
}

}  // namespace vaeq
