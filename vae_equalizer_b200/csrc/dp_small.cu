// DP VAE step for REFERENCE-SIZE minibatches (batch_len <= 512; Eval_run_DP.py:38 uses 100): persistent frame kernel.
//
// At batch_len = 100 one training step is ~0.5 MFLOP: launched as separate kernels a frame of 100-990 sequential steps
// (func_VAELE_DP_MQAM_shaping.py:57-66, func_VAEflex_DP_MQAM_shaping.py:59-70) is pure launch + global-memory latency.
// Here ONE CTA owns one run for the whole frame:
//   * taps W, channel estimate h, Adam moments and every per-step intermediate (E_q, backward coefficients, residual e,
//     dL/dout, partial sums) live in SHARED MEMORY across all steps; global traffic per step is the rx window in and
//     the kept columns of q / out (+ loss, var_est) out;
//   * the step is a sequence of CTA phases separated by __syncthreads(): FIR over (symbol, output pol) items, soft
//     demapper over (symbol, component) items (ex2/lg2/rcp math of dp_math.cuh, which also yields the backward
//     coefficients S1,T2,S3 so the backward needs no q and no transcendental), channel convolution + residual over
//     (sample, rx pol) items, scalar ELBO assembly, dL/dE_q + dL/dout, the two tap-gradient correlations, Adam;
//   * blockIdx.x = run: independent runs (the sweep cells of Eval_run_DP.py:68-95) are batched in the same launch, one
//     CTA each, 3-4 CTAs per SM.
// Same math as dp_step.cu / dp_fast.cu (closed form restated in oracle/closed_form.py); any odd M_est <= 63.
#include "dp_math.cuh"
#include "dp_kernels.cuh"

namespace vaeq {

constexpr int SM_NT = 256;

// shared-memory plan (offsets in floats); doubles first (8-byte aligned), then float4 arrays, then floats
struct SmallLayout {
    int XA, XO, SA, SO;                     // phase-array lengths (float4) and zero margins of x and e
    int dsc, xph, m1s, eph, gys, Wt, hD, hG, part4, cst, ys, srow, vsc, PSg, Wf, hf, adam, gfin, Ssh, hsq, red, scal, total;
};
__host__ __device__ inline SmallLayout small_layout(int B, int M) {
    SmallLayout l;
    const int mh = M / 2;
    l.XO = (mh + 1) / 2 + 1;
    l.XA = B + 2 * l.XO;
    l.SO = (mh + 1) / 2 + 1;
    l.SA = B + 2 * l.SO;
    int off = 0;
    auto take = [&](int nfloats) {
        int o = off;
        off += (nfloats + 3) & ~3;          // keep 16-byte alignment throughout
        return o;
    };
    l.dsc = take(2 * 12);                   // 12 doubles: tot[5], Esh[2], bc1, spare
    l.xph = take(4 * 2 * l.XA);
    l.m1s = take(4 * B);
    l.eph = take(4 * 2 * l.SA);
    l.gys = take(4 * B);
    l.Wt = take(4 * 2 * M);
    l.hD = take(4 * 2 * M);
    l.hG = take(4 * 2 * M);
    l.part4 = take(4 * SM_NT);
    l.cst = take((int)(sizeof(FastConst) / 4));
    l.ys = take(4 * B);
    l.srow = take(12 * B);
    l.vsc = take(4 * B);
    l.PSg = take(2 * (M + 1));
    l.Wf = take(8 * M);
    l.hf = take(8 * M);
    l.adam = take(48 * M);
    l.gfin = take(16 * M);
    l.Ssh = take(2 * M);
    l.hsq = take(4 * M);
    l.red = take(5 * 32);
    l.scal = take(8);
    l.total = off;
    return l;
}

template <int NL>
__global__ void __launch_bounds__(SM_NT) k_dp_frame_fast(DpK p, DpRunsK rs, int n_steps, int stride_sym, int keep_lo_in_dst,
                                                         float lr_w, float lr_h, int amsgrad) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int B = p.B, L = 2 * p.B, M = p.M, mh = p.mh, Mh = 2 * p.mh, run = blockIdx.x;
    const SmallLayout lay = small_layout(B, M);
    const int XA = lay.XA, XO = lay.XO, SA = lay.SA, SO = lay.SO;
    double *dsc = reinterpret_cast<double *>(sm + lay.dsc);
    float4 *xph = reinterpret_cast<float4 *>(sm + lay.xph);      // [2][XA]  rx phases {I0,Q0,I1,Q1}: xph[ph][a+XO] = rx[:, 2a+ph]
    float4 *m1s = reinterpret_cast<float4 *>(sm + lay.m1s);      // [B]      E_q[x] {p0 I, p0 Q, p1 I, p1 Q}
    float4 *eph = reinterpret_cast<float4 *>(sm + lay.eph);      // [2][SA]  residual D - rx {chi0 re, im, chi1 re, im} by sample phase
    float4 *gys = reinterpret_cast<float4 *>(sm + lay.gys);      // [B]      dL/dout
    float4 *Wt = reinterpret_cast<float4 *>(sm + lay.Wt);        // [o][k]   {wr<-p0, wi<-p0, wr<-p1, wi<-p1}
    float4 *hD = reinterpret_cast<float4 *>(sm + lay.hD);        // [chi][j] {h_chi,0 re, im, h_chi,1 re, im}
    float4 *hG = reinterpret_cast<float4 *>(sm + lay.hG);        // [nu][j]  2 kappa_chi * {h_0,nu re, im, h_1,nu re, im}
    float4 *part4 = reinterpret_cast<float4 *>(sm + lay.part4);
    FastConst *cst = reinterpret_cast<FastConst *>(sm + lay.cst);
    float *ys = sm + lay.ys, *srow = sm + lay.srow, *vsc = sm + lay.vsc, *PSg = sm + lay.PSg, *Wf = sm + lay.Wf, *hf = sm + lay.hf;
    float *ad = sm + lay.adam, *gfin = sm + lay.gfin, *Ssh = sm + lay.Ssh, *hsq = sm + lay.hsq, *red = sm + lay.red, *scal = sm + lay.scal;

    // ---- this run's tensors (blockIdx.x = run) ---------------------------------------------------------------------
    const float *rx0 = p.rx + run * rs.rs_rx;
    float *Wg = p.W + run * rs.rs_W, *hg = p.h + run * rs.rs_h, *adg = p.adam + run * rs.rs_adam;
    float *qk = p.qk ? p.qk + run * rs.rs_qk : nullptr, *outk = p.qk ? p.outk + run * rs.rs_outk : nullptr;
    if (rs.lr_w) lr_w = rs.lr_w[run];
    if (rs.lr_h) lr_h = rs.lr_h[run];
    const float nu_sc = rs.nu_sc ? rs.nu_sc[run] : p.nu_sc;

    for (int i = tid; i < 8 * M; i += SM_NT) {
        Wf[i] = Wg[i];
        hf[i] = hg[i];
    }
    for (int i = tid; i < 48 * M; i += SM_NT) ad[i] = adg[i];
    for (int i = tid; i < 2 * SA; i += SM_NT) eph[i] = make_float4(0.f, 0.f, 0.f, 0.f);     // margins stay zero for the whole frame
    load_fast_const(cst, p.amp + run * rs.rs_amp, p.P + run * rs.rs_P, p.var + run * rs.rs_var, nu_sc, NL);
    const int step0 = *reinterpret_cast<const int *>(adg + 48 * M);
    const double width = (double)(L - Mh);
    __syncthreads();
    const FastConst &c = *cst;

    for (int m = 0; m < n_steps; ++m) {
        const float *rx = rx0 + (int64_t)m * stride_sym * 2;
        const int64_t keep_base = (int64_t)m * stride_sym + (keep_lo_in_dst ? p.keep_lo : 0);
        const bool last = m == n_steps - 1;

        // ---- P0: tap tables from the master copies, rx window -> phase arrays (zero padded), Adam bias corrections ------
        for (int idx = tid; idx < 2 * M; idx += SM_NT) {
            const int o = idx / M, k = idx - o * M;
            Wt[idx] = make_float4(Wf[(o * 4 + 0) * M + k], Wf[(o * 4 + 2) * M + k], Wf[(o * 4 + 1) * M + k], Wf[(o * 4 + 3) * M + k]);
            hD[idx] = make_float4(hf[((o * 2 + 0) * 2 + 0) * M + k], hf[((o * 2 + 0) * 2 + 1) * M + k],
                                  hf[((o * 2 + 1) * 2 + 0) * M + k], hf[((o * 2 + 1) * 2 + 1) * M + k]);
        }
        for (int i = tid; i < 4 * M; i += SM_NT) {
            const int cn = i / M, j = i - cn * M;
            const float hr = hf[(cn * 2 + 0) * M + j], hi = hf[(cn * 2 + 1) * M + j];
            hsq[i] = hr * hr + hi * hi;
        }
        for (int idx = tid; idx < 2 * XA; idx += SM_NT) {
            const int ph = idx >= XA, a = idx - ph * XA - XO, s = 2 * a + ph;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s >= 0 && s < L) v = make_float4(rx[s], rx[p.ld_rx + s], rx[2 * p.ld_rx + s], rx[3 * p.ld_rx + s]);
            xph[idx] = v;
        }
        if (tid == SM_NT - 1) {
            float bc2s;
            adam_bias(step0 + m + 1, &dsc[8], &bc2s);
            scal[6] = bc2s;
        }
        __syncthreads();

        // ---- P1: butterfly FIR (sf:500-518), one (symbol, output pol) item per thread --------------------------------
        for (int it = tid; it < 2 * B; it += SM_NT) {
            const int o = it >= B, u = it - o * B;
            float reA = 0.f, reB = 0.f, imA = 0.f, imB = 0.f;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const int k0 = (mh + ph) & 1;                    // taps k = k0, k0+2, ... read samples of phase ph
                const float4 *xb = xph + ph * XA + u + XO + ((k0 - mh - ph) >> 1);
                const float4 *wb = Wt + o * M + k0;
                const int n = (M - k0 + 1) >> 1;
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const float4 x = xb[i], w = wb[2 * i];
                    reA = fmaf(w.x, x.x, reA); reA = fmaf(-w.y, x.y, reA);
                    reB = fmaf(w.z, x.z, reB); reB = fmaf(-w.w, x.w, reB);
                    imA = fmaf(w.x, x.y, imA); imA = fmaf(w.y, x.x, imA);
                    imB = fmaf(w.z, x.w, imB); imB = fmaf(w.w, x.z, imB);
                }
            }
            ys[(2 * o) * B + u] = reA + reB;
            ys[(2 * o + 1) * B + u] = imA + imB;
        }
        __syncthreads();

        // ---- P2: soft demapper + moments + entropy + backward coefficients, one (symbol, component) item per thread ----
        float accEnt = 0.f, accV0 = 0.f, accV1 = 0.f;
        for (int it = tid; it < 4 * B; it += SM_NT) {
            const int cc = it / B, u = it - cc * B, pol = cc >> 1;
            const float y = ys[it];
            float q[NL], m1, m2, ent, S1, S2, S3;
            demap_fast<NL, true>(y, c.c2[pol], c.inv_var[pol], c, q, m1, m2, ent, S1, S2, S3);
            if (qk != nullptr && u >= p.keep_lo && u < p.keep_lo + p.keep_n) {       // VAELE_DP:61-62 / VAEflex_DP:64-65
                const int64_t col = keep_base + (u - p.keep_lo);
#pragma unroll
                for (int l = 0; l < NL; ++l) qk[(int64_t)(cc * NL + l) * p.ld_qk + col] = q[l];
                outk[(int64_t)cc * p.ld_outk + col] = y;
            }
            reinterpret_cast<float *>(m1s + u)[cc] = m1;
            srow[cc * B + u] = S1;
            srow[(4 + cc) * B + u] = fmaf(-2.f * m1, S1, S2);
            srow[(8 + cc) * B + u] = S3;
            const float v = m2 - m1 * m1;                                            // sf:113
            vsc[it] = v;
            if (pol) accV1 += v; else accV0 += v;
            if (u >= mh && u < B - mh) accEnt += ent;                                // sf:132: [mh:-mh] in SYMBOLS
        }
        __syncthreads();

        // ---- P3: D = h * E_q on the valid samples, residual e = D - rx (sf:115-134), one (sample, rx pol) item per thread --
        float accC0 = 0.f, accC1 = 0.f;
        for (int it = tid; it < 2 * L; it += SM_NT) {
            const int chi = it >= L, s = it - chi * L;
            float er = 0.f, ei = 0.f;
            if (s >= mh && s < L - mh) {                                             // sf:120 "valid", in SAMPLES
                float drA = 0.f, drB = 0.f, diA = 0.f, diB = 0.f;
                const int par = (s + mh) & 1;
                const float4 *hb = hD + chi * M + par;
                const float4 *mb = m1s + ((s + mh - par) >> 1);                       // E_q[(s + mh - j)/2], j = par + 2i
                const int n = (M - par + 1) >> 1;
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const float4 hh = hb[2 * i], mm = mb[-i];
                    drA = fmaf(hh.x, mm.x, drA); drA = fmaf(-hh.y, mm.y, drA);
                    drB = fmaf(hh.z, mm.z, drB); drB = fmaf(-hh.w, mm.w, drB);
                    diA = fmaf(hh.y, mm.x, diA); diA = fmaf(hh.x, mm.y, diA);
                    diB = fmaf(hh.w, mm.z, diB); diB = fmaf(hh.z, mm.w, diB);
                }
                const float4 x = xph[(s & 1) * XA + (s >> 1) + XO];
                er = (drA + drB) - (chi ? x.z : x.x);
                ei = (diA + diB) - (chi ? x.w : x.y);
            }
            reinterpret_cast<float2 *>(eph + (s & 1) * SA + (s >> 1) + SO)[chi] = make_float2(er, ei);
            const float e2 = er * er + ei * ei;
            if (chi) accC1 += e2; else accC0 += e2;
        }

        // ---- P4: ELBO scalars (sf:131-137): C, loss, var_est, kappa = (L-Mh)/C, S_nu(j) ------------------------------------
        {
            float v[5] = {accC0, accC1, accEnt, accV0, accV1};
            block_sum<5>(v, red);                                                    // has the barriers that publish eph
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < 5; ++i) dsc[i] = (double)v[i];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < 2 * M; idx += SM_NT) {                             // S_nu(j): Var sum over Mh <= 2u+j < L (sf:128)
            const int nu = idx / M, j = idx - nu * M;
            double s = dsc[3 + nu];
            const int u_lo = (Mh - j + 1) >> 1, u_hi = (L - 1 - j) >> 1;
            for (int u = 0; u < u_lo && u < B; ++u) s -= (double)(vsc[(2 * nu) * B + u] + vsc[(2 * nu + 1) * B + u]);
            for (int u = u_hi + 1; u < B; ++u) s -= (double)(vsc[(2 * nu) * B + u] + vsc[(2 * nu + 1) * B + u]);
            Ssh[idx] = (float)s;
        }
        __syncthreads();
        if (wid < 2) {                                                               // E_chi = sum |h|^2 S_nu(j)  (sf:129)
            double E = 0.0;
            for (int idx = lane; idx < 2 * M; idx += 32) E += (double)hsq[wid * 2 * M + idx] * (double)Ssh[idx];
            E = warp_sum(E);
            if (lane == 0) dsc[5 + wid] = E;
        }
        __syncthreads();
        if (tid < 2) {
            const int chi = tid;
            const double C = dsc[chi] + dsc[5 + chi];                                // sf:133-134
            const double term = width * log(C);                                      // sf:136
            scal[chi] = (float)(width / C);                                          // kappa
            const float ve = (float)(C / width);                                     // sf:137
            if (rs.var_steps) rs.var_steps[((int64_t)run * 2 + chi) * n_steps + m] = ve;
            if (last && rs.var_last) rs.var_last[2 * run + chi] = ve;
            const double term1 = __shfl_sync(0x3u, term, 1);
            if (chi == 0) {
                const float loss = (float)((-dsc[2] + term) + term1);
                if (rs.loss_steps) rs.loss_steps[(int64_t)run * n_steps + m] = loss;
                if (last && rs.loss_last) rs.loss_last[run] = loss;
            }
        }
        __syncthreads();
        const float kap0 = scal[0], kap1 = scal[1];
        for (int idx = tid; idx < 2 * M; idx += SM_NT) {                             // conj(h) taps of dL/dE_q, scaled by 2 kappa_chi
            const int nu = idx / M, j = idx - nu * M;
            hG[idx] = make_float4(2.f * kap0 * hf[((0 * 2 + nu) * 2 + 0) * M + j], 2.f * kap0 * hf[((0 * 2 + nu) * 2 + 1) * M + j],
                                  2.f * kap1 * hf[((1 * 2 + nu) * 2 + 0) * M + j], 2.f * kap1 * hf[((1 * 2 + nu) * 2 + 1) * M + j]);
        }
        if (tid >= 32 && tid < 34) {                                                 // dL/dVar prefix sums over the taps
            const int nu = tid - 32;
            float a = 0.f;
            PSg[nu * (M + 1)] = 0.f;
            for (int j = 0; j < M; ++j) {
                a += kap0 * hsq[(0 * 2 + nu) * M + j] + kap1 * hsq[(1 * 2 + nu) * M + j];
                PSg[nu * (M + 1) + j + 1] = a;
            }
        }
        __syncthreads();

        // ---- P5: dL/dE_q = conj(h) (*) gD, then dL/dout = dL/dE_q S1 + dL/dVar T2 + w S3; one (symbol, tx pol) item per thread --
        for (int it = tid; it < 2 * B; it += SM_NT) {
            const int nu = it >= B, u = it - nu * B;
            float grA = 0.f, grB = 0.f, giA = 0.f, giB = 0.f;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const int j0 = (mh + ph) & 1;                    // gD sample 2u - mh + j has phase ph for j = j0, j0+2, ...
                const float4 *eb = eph + ph * SA + u + SO + ((j0 - mh - ph) >> 1);
                const float4 *hb = hG + nu * M + j0;
                const int n = (M - j0 + 1) >> 1;
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const float4 e = eb[i], hh = hb[2 * i];
                    grA = fmaf(hh.x, e.x, grA); grA = fmaf(hh.y, e.y, grA);
                    grB = fmaf(hh.z, e.z, grB); grB = fmaf(hh.w, e.w, grB);
                    giA = fmaf(hh.x, e.y, giA); giA = fmaf(-hh.y, e.x, giA);
                    giB = fmaf(hh.z, e.w, giB); giB = fmaf(-hh.w, e.z, giB);
                }
            }
            const int jlo = max(0, Mh - 2 * u), jhi = min(M, L - 2 * u);
            const float gV = PSg[nu * (M + 1) + jhi] - PSg[nu * (M + 1) + jlo];
            const float entw = (u >= mh && u < B - mh) ? LN2 : 0.f;
            const int cc = 2 * nu;
            const float gI = fmaf(grA + grB, srow[cc * B + u], fmaf(gV, srow[(4 + cc) * B + u], entw * srow[(8 + cc) * B + u]));
            const float gQ = fmaf(giA + giB, srow[(cc + 1) * B + u], fmaf(gV, srow[(5 + cc) * B + u], entw * srow[(9 + cc) * B + u]));
            reinterpret_cast<float2 *>(gys + u)[nu] = make_float2(gI, gQ);
        }
        __syncthreads();

        // ---- P6: tap gradients.  item < 2M: dW[o=0,1][in][k] = sum_u gy_o(u) conj(x_in[2u+k-mh]);
        //          item >= 2M: dh[chi=0,1][nu][j] = 2 kappa_chi sum_v e_chi(2v+j-mh) conj(E_q,nu[v]).  The symbol range is cut
        //          into `parts` chunks so that all threads work; partials are summed in fixed order (deterministic).
        {
            const int items = 4 * M, parts = max(1, SM_NT / items), chunk = (B + parts - 1) / parts;
            if (tid < items * parts) {
                const int item = tid % items, part = tid / items;
                const int u0 = part * chunk, u1 = min(B, u0 + chunk);
                float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
                if (item < 2 * M) {
                    const int in = item / M, k = item - in * M;
                    const int t = k - mh, ph = t & 1;
                    const float4 *xb = xph + ph * XA + XO + ((t - ph) >> 1);
#pragma unroll 4
                    for (int u = u0; u < u1; ++u) {
                        const float4 x = xb[u], g = gys[u];
                        const float a = in ? x.z : x.x, b = in ? x.w : x.y;
                        a0r = fmaf(g.x, a, a0r); a0r = fmaf(g.y, b, a0r);
                        a0i = fmaf(g.y, a, a0i); a0i = fmaf(-g.x, b, a0i);
                        a1r = fmaf(g.z, a, a1r); a1r = fmaf(g.w, b, a1r);
                        a1i = fmaf(g.w, a, a1i); a1i = fmaf(-g.z, b, a1i);
                    }
                } else {
                    const int it2 = item - 2 * M, nu = it2 / M, j = it2 - nu * M;
                    const int t = j - mh, ph = t & 1;
                    const float4 *eb = eph + ph * SA + SO + ((t - ph) >> 1);
#pragma unroll 4
                    for (int v = u0; v < u1; ++v) {
                        const float4 e = eb[v], mm = m1s[v];
                        const float a = nu ? mm.z : mm.x, b = nu ? mm.w : mm.y;
                        a0r = fmaf(e.x, a, a0r); a0r = fmaf(e.y, b, a0r);
                        a0i = fmaf(e.y, a, a0i); a0i = fmaf(-e.x, b, a0i);
                        a1r = fmaf(e.z, a, a1r); a1r = fmaf(e.w, b, a1r);
                        a1i = fmaf(e.w, a, a1i); a1i = fmaf(-e.z, b, a1i);
                    }
                }
                part4[part * items + item] = make_float4(a0r, a0i, a1r, a1i);
            }
            __syncthreads();
            if (tid < items) {
                float4 s = part4[tid];
                for (int part = 1; part < parts; ++part) {
                    const float4 t4 = part4[part * items + tid];
                    s.x += t4.x; s.y += t4.y; s.z += t4.z; s.w += t4.w;
                }
                if (tid < 2 * M) {
                    const int in = tid / M, k = tid - in * M;
                    gfin[(0 * 4 + in) * M + k] = s.x;
                    gfin[(0 * 4 + 2 + in) * M + k] = s.y;
                    gfin[(1 * 4 + in) * M + k] = s.z;
                    gfin[(1 * 4 + 2 + in) * M + k] = s.w;
                } else {
                    const int it2 = tid - 2 * M, nu = it2 / M, j = it2 - nu * M;
                    gfin[8 * M + ((0 * 2 + nu) * 2 + 0) * M + j] = 2.f * kap0 * s.x;
                    gfin[8 * M + ((0 * 2 + nu) * 2 + 1) * M + j] = 2.f * kap0 * s.y;
                    gfin[8 * M + ((1 * 2 + nu) * 2 + 0) * M + j] = 2.f * kap1 * s.z;
                    gfin[8 * M + ((1 * 2 + nu) * 2 + 1) * M + j] = 2.f * kap1 * s.w;
                }
            }
            __syncthreads();
        }

        // ---- P7: E-term of dh, Adam on both parameter groups (VAELE_DP:28-31,66), one parameter per thread -----------------
        {
            const double bc1 = dsc[8];
            const float bc2s = scal[6];
            for (int i = tid; i < 16 * M; i += SM_NT) {
                float g = gfin[i];
                if (i >= 8 * M) {
                    const int r = i - 8 * M, j = r % M, cn = r / (2 * M), chi = cn >> 1, nu = cn & 1;
                    g = (float)((double)g + 2.0 * (double)(chi ? kap1 : kap0) * (double)hf[r] * (double)Ssh[nu * M + j]);
                }
                if (last) {
                    if (i < 8 * M) {
                        if (rs.gW_last) rs.gW_last[(int64_t)run * 8 * M + i] = g;
                    } else if (rs.gh_last) {
                        rs.gh_last[(int64_t)run * 8 * M + i - 8 * M] = g;
                    }
                }
                if (i < 8 * M) adam_apply(Wf, g, ad, ad + 8 * M, ad + 16 * M, i, lr_w, amsgrad != 0, bc1, bc2s);
                else adam_apply(hf, g, ad + 24 * M, ad + 32 * M, ad + 40 * M, i - 8 * M, lr_h, amsgrad != 0, bc1, bc2s);
            }
        }
        __syncthreads();
    }

    // ---- write the trained state back -------------------------------------------------------------------------------
    for (int i = tid; i < 8 * M; i += SM_NT) {
        Wg[i] = Wf[i];
        hg[i] = hf[i];
    }
    for (int i = tid; i < 48 * M; i += SM_NT) adg[i] = ad[i];
    if (tid == 0) *reinterpret_cast<int *>(adg + 48 * M) = step0 + n_steps;
}

size_t dp_small_smem(int B, int M) { return (size_t)small_layout(B, M).total * sizeof(float); }

int dp_small_launch(const DpK &p, const DpRunsK &rs, int n_lev, int n_runs, int n_steps, int stride_sym, int keep_lo_in_dst,
                    float lr_w, float lr_h, int amsgrad, cudaStream_t st) {
    const size_t smem = dp_small_smem(p.B, p.M);
    static size_t set_smem[3] = {0, 0, 0};
#define SMALL_CASE(NL_, IDX_)                                                                                              \
    {                                                                                                                      \
        if (smem > set_smem[IDX_]) {                                                                                       \
            VAEQ_CUDA(cudaFuncSetAttribute(k_dp_frame_fast<NL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            set_smem[IDX_] = smem;                                                                                         \
        }                                                                                                                  \
        ktime_begin(VAEQ_K_DP_FRAME, st);                                                                                  \
        k_dp_frame_fast<NL_><<<n_runs, SM_NT, smem, st>>>(p, rs, n_steps, stride_sym, keep_lo_in_dst, lr_w, lr_h, amsgrad); \
        ktime_end(VAEQ_K_DP_FRAME, st);                                                                                    \
    }
    if (n_lev == 2) SMALL_CASE(2, 0)
    else if (n_lev == 4) SMALL_CASE(4, 1)
    else SMALL_CASE(8, 2)
#undef SMALL_CASE
    VAEQ_LAUNCH_CHECK("k_dp_frame_fast");
    return VAEQ_OK;
}

}  // namespace vaeq
