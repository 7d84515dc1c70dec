// DP VAE step, backward pass as ONE warp-specialised kernel for B200 (sm_100a).
//
// The three backward kernels of dp_fast.cu (dL/dE_q + softmin backward -> dL/dout rows; dW; dh) read 192 B per symbol where 128 are
// needed (the residual rows e twice, dL/dout written and read back, rx a second time) and each of them runs its own
// stage -> barrier -> contract phases.  Here one persistent CTA per SM holds three warp families that work on DIFFERENT tiles at the same time:
//   A (4 warps)  tile t+1: dL/dE_q = conj(h) (*) gD (3-multiplication FIR over the residual window), dL/dout from the forward pass's
//                S1/T2/S3 rows; leaves the residual window and the dL/dout window in shared memory.  One A thread is also the
//                producer: every global read of the kernel is a TMA bulk copy (cp.async.bulk, completion on an mbarrier) of a row
//                segment exactly as it lies in HBM, issued one tile ahead;
//   B (4 warps)  tile t: dW = rx (*) dL/dout;     C (4 warps)  tile t: dh = E_q (*) residual.
// A has the highest warp ids (the issue arbiter prefers them): it finishes its tile early and waits, B and C never wait for it.
// Registers are per family (B and C each keep 56 tap accumulators, A none), so the "accumulators live across the point-wise
// stage" objection to a fused backward does not apply; families synchronise through mbarriers only (no CTA-wide barrier in the
// tile loop).  In B and C the per-symbol operand is the one that is used as its raw SoA rows (rx samples, E_q rows: a lane's four
// symbols are one or two float4 per row) and the SLIDING operand is the one an A thread produced for its own four symbols
// (dL/dout, residual), so nothing is transposed by anybody but its producer:
//     dW[o][i][2a+ph]   = sum_p x_ph,i(p)   conj-pair with gy_o(p + HF - a)
//     dh[chi][nu][j]    = sum_u E_q,nu(u)   conj-pair with e_ph,chi(u + b),   j = ph + 2 (b + HF)
// (every (symbol, lag) pair counted once, partitioned by the position of the per-symbol operand).  DRAM: e 32 + E_q 16 +
// S 48 + rx 32 = 128 B per symbol.  Reference: loss.backward() of loss_function_shaping sf:92-137 through twoXtwoFIR.forward sf:500-527.
#include "dp_fast.cuh"
#include "tma.cuh"

namespace vaeq {

constexpr int FB_NA = 4;                                    // warps of family A (= B = C)
constexpr int FB_NT = 32 * 3 * FB_NA;                       // 384 threads: B, C, A (168 registers each, one CTA per SM)
constexpr int FB_SL = 32 * FB_NA;                           // 128 slots of FT_R = 4 symbols
constexpr int FB_TE = FB_SL * FT_R;                         // 512 slot symbols per tile
constexpr int FB_HP = 8;                                    // slots' margin per side (>= MH/2 + 2, multiple of 4)
constexpr int FB_T = FB_TE - 2 * FB_HP;                     // 496 owned positions per tile
constexpr int FB_EX = 8;                                    // residual margin per side beyond the slots
constexpr int FB_EN = FB_TE + 2 * FB_EX;                    // 528 staged residual positions
constexpr int FB_ES = FB_EN + FB_EN / 4 + 4;                // padded float4 lengths
constexpr int FB_GS = FB_TE + FB_TE / 4 + 4;
constexpr int FB_NBAR = 16;
constexpr int FB_RB = 3;                                    // stages of the E_q / rx ring (consumed one tile later than the residual ring)

// ---- FIR-like contraction over a SWIZZLED window {re_0, re_1, im_0, im_1} (3-multiplication form, see dp_fast.cuh) -------------
//     P1[r][o] += (tr_{o,0}, tr_{o,1}) * (re_0 + im_0, re_1 + im_1)
//     P2[r][o] += (re_0, re_1) * (td_{o,0}, td_{o,1})         P3[r][o] += (im_0, im_1) * (ts_{o,0}, ts_{o,1})
// halves summed at the end, Re = P1 - P3, Im = P1 + P2.  Tap table per lag:
//     T0 = {tr00, tr01, tr10, tr11}, T1 = {td00, td01, ts00, ts01}, T2 = {td10, td11, ts10, ts11}     (index o,i)
struct FirAccS {
    float2 P1[FT_R][2], P2[FT_R][2], P3[FT_R][2];
};
struct WinElS {
    float2 xr, xi, s;
};
__device__ __forceinline__ void fir_step_s(const float4 T0, const float4 T1, const float4 T2, const WinElS &xa, const WinElS &xb,
                                           const WinElS &xc, const WinElS &xd, FirAccS &a) {
    const WinElS *xs[4] = {&xa, &xb, &xc, &xd};
#pragma unroll
    for (int r = 0; r < FT_R; ++r) {
        const WinElS &w = *xs[r];
        a.P1[r][0] = __ffma2_rn(make_float2(T0.x, T0.y), w.s, a.P1[r][0]);
        a.P1[r][1] = __ffma2_rn(make_float2(T0.z, T0.w), w.s, a.P1[r][1]);
        a.P2[r][0] = __ffma2_rn(w.xr, make_float2(T1.x, T1.y), a.P2[r][0]);
        a.P3[r][0] = __ffma2_rn(w.xi, make_float2(T1.z, T1.w), a.P3[r][0]);
        a.P2[r][1] = __ffma2_rn(w.xr, make_float2(T2.x, T2.y), a.P2[r][1]);
        a.P3[r][1] = __ffma2_rn(w.xi, make_float2(T2.z, T2.w), a.P3[r][1]);
    }
}
// window element c of the group whose first element is logical position 4 q + M4 (padded float4 index, compile-time offset)
template <int M4>
__device__ __forceinline__ void fir4s(const float4 *__restrict__ win, int q, const float4 *__restrict__ taps, int nlag, FirAccS &acc) {
    auto LD = [win](int q_, int c) {
        const float4 x = win[5 * q_ + c];
        WinElS w;
        w.xr = make_float2(x.x, x.y);
        w.xi = make_float2(x.z, x.w);
        w.s = __fadd2_rn(w.xr, w.xi);
        return w;
    };
    WinElS w0 = LD(q, poff(M4)), w1 = LD(q, poff(M4 + 1)), w2 = LD(q, poff(M4 + 2)), w3;
    int a = 0;
#pragma unroll 1
    for (; a + 4 <= nlag; a += 4) {
        w3 = LD(q, poff(M4 + 3));
        fir_step_s(taps[0], taps[1], taps[2], w0, w1, w2, w3, acc);
        w0 = LD(q, poff(M4 + 4));
        fir_step_s(taps[3], taps[4], taps[5], w1, w2, w3, w0, acc);
        w1 = LD(q, poff(M4 + 5));
        fir_step_s(taps[6], taps[7], taps[8], w2, w3, w0, w1, acc);
        w2 = LD(q, poff(M4 + 6));
        fir_step_s(taps[9], taps[10], taps[11], w3, w0, w1, w2, acc);
        q += 1;
        taps += 4 * FT_TAPV;
    }
    if (a < nlag) {                                             // nlag % 4 <= 1 here (13 even, 12 odd lags at M_est = 25; 7/6, 5/4, 3/2)
        w3 = LD(q, poff(M4 + 3));
        fir_step_s(taps[0], taps[1], taps[2], w0, w1, w2, w3, acc);
        if (a + 1 < nlag) {
            w0 = LD(q, poff(M4 + 4));
            fir_step_s(taps[3], taps[4], taps[5], w1, w2, w3, w0, acc);
            if (a + 2 < nlag) {
                w1 = LD(q, poff(M4 + 5));
                fir_step_s(taps[6], taps[7], taps[8], w2, w3, w0, w1, acc);
            }
        }
    }
}

// corr4 (dp_fast.cuh) with compile-time window offsets: the first window element is logical position 4 q + M4
template <int A, int M4>
__device__ __forceinline__ void corr4m(const float4 *__restrict__ win, int q, const float4 (&g)[FT_R], float2 (&acc2)[A][4]) {
    float ng[FT_R][2];
#pragma unroll
    for (int r = 0; r < FT_R; ++r) {
        ng[r][0] = -g[r].x;
        ng[r][1] = -g[r].z;
    }
#pragma unroll
    for (int cpos = 0; cpos < A + FT_R - 1; ++cpos) {
        const float4 x = win[5 * q + poff(M4 + cpos)];
        const float2 xr = make_float2(x.x, x.y), xi = make_float2(x.z, x.w);
#pragma unroll
        for (int r = 0; r < FT_R; ++r) {
            const int a = cpos - r;
            if (a >= 0 && a < A) {
                const float4 gg = g[r];
                acc2[a][0] = __ffma2_rn(make_float2(gg.x, gg.x), xr, acc2[a][0]);
                acc2[a][1] = __ffma2_rn(make_float2(gg.y, gg.y), xr, acc2[a][1]);
                acc2[a][2] = __ffma2_rn(make_float2(gg.z, gg.z), xr, acc2[a][2]);
                acc2[a][3] = __ffma2_rn(make_float2(gg.w, gg.w), xr, acc2[a][3]);
                acc2[a][0] = __ffma2_rn(make_float2(gg.y, gg.y), xi, acc2[a][0]);
                acc2[a][1] = __ffma2_rn(make_float2(ng[r][0], ng[r][0]), xi, acc2[a][1]);
                acc2[a][2] = __ffma2_rn(make_float2(gg.w, gg.w), xi, acc2[a][2]);
                acc2[a][3] = __ffma2_rn(make_float2(ng[r][1], ng[r][1]), xi, acc2[a][3]);
            }
        }
    }
}

template <int MH>
__global__ void __launch_bounds__(FB_NT, 1) k_dp_bwd_fused(DpK p) {
    constexpr int M = 2 * MH + 1, HF = MH / 2, NE = MH + 1, NO = MH, BASE = MH / 2, AMAX = BASE + 1;
    static_assert(MH % 2 == 0 && HF + 2 <= FB_HP && HF <= FB_EX - 2, "fused backward needs M_est = 1 (mod 4), M_est <= 25");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    float *eraw = reinterpret_cast<float *>(smem_raw + 8 * FB_NBAR);     // [2][8][FB_EN]
    float *sraw = eraw + 2 * 8 * FB_EN;                                  // [2][12][FB_TE]
    float *m1raw = sraw + 2 * 12 * FB_TE;                                // [FB_RB][4][FB_TE]
    float *xraw = m1raw + FB_RB * 4 * FB_TE;                             // [FB_RB][4][2 FB_TE]
    float4 *gew = reinterpret_cast<float4 *>(xraw + FB_RB * 4 * 2 * FB_TE);   // [2][phase][FB_ES]: residual window {chi0 re, chi1 re, chi0 im, chi1 im}
    float4 *gyw = gew + 2 * 2 * FB_ES;                                   // [2][FB_GS]: dL/dout window {p0 I, p1 I, p0 Q, p1 Q}
    float4 *tapG = gyw + 2 * FB_GS;                                      // 2 kappa_chi conj(h) taps for dE_q: [phase][lag][FT_TAPV]
    float *PSg = reinterpret_cast<float *>(tapG + FT_TAPV * (NE + NO));  // (2, M+1)
    float *red = PSg + 2 * (M + 1);                                      // [8 warps][AMAX * 8]
    uint64_t *fullA = bars, *fullB = bars + 2, *ge_ready = bars + 5, *gy_ready = bars + 7, *ge_free = bars + 9, *gy_free = bars + 11;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const float kap0 = p.scal[DP_KAPPA_OFF], kap1 = p.scal[DP_KAPPA_OFF + 1];

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(fullA + s, 1);
            mbar_init(ge_ready + s, 1);
            mbar_init(gy_ready + s, FB_NA);
            mbar_init(ge_free + s, FB_NA);
            mbar_init(gy_free + s, FB_NA);
        }
        for (int s = 0; s < FB_RB; ++s) mbar_init(fullB + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < (NE + NO) * 4 * FT_TAPV; idx += FB_NT) {
        const int la = idx / (4 * FT_TAPV), e = idx - la * (4 * FT_TAPV);
        const int ph = la >= NE, a = ph ? la - NE : la;
        int nu, chi, kind = -1;                                   // out index nu, in index chi
        if (e < 4) { nu = e >> 1; chi = e & 1; }
        else { const int q = e - 4; nu = q >> 2; kind = (q >> 1) & 1; chi = q & 1; }
        const int j = 2 * a + ph;
        const float sc = 2.f * (chi ? kap1 : kap0);               // gD = 2 kappa_chi e folded into the taps
        const float tr = sc * p.h[((chi * 2 + nu) * 2 + 0) * M + j], ti = -sc * p.h[((chi * 2 + nu) * 2 + 1) * M + j];
        reinterpret_cast<float *>(tapG)[idx] = kind < 0 ? tr : (kind ? tr + ti : ti - tr);
    }
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

    const int nt = p.ntiles;
    if (wid >= 2 * FB_NA) {
        // ================= family A: producer, dL/dE_q, dL/dout =================
        const int l = tid - 2 * FB_NA * 32;                      // slot
        // all global reads of tile `it`: residual + S rows into ring A (2 stages), E_q + rx rows into ring B (3 stages: B / C use them one tile later)
        auto produce = [&](int tile, int it) {                   // called by warp 0 of the family
            const int sa = it & 1, sb = it % FB_RB;
            const int64_t s0 = (int64_t)p.sym_lo + (int64_t)tile * FB_T - FB_HP;
            float *de = eraw + sa * 8 * FB_EN, *ds = sraw + sa * 12 * FB_TE, *dm = m1raw + sb * 4 * FB_TE, *dx = xraw + sb * 4 * 2 * FB_TE;
            rows_zero_fill(de, 8, s0 - FB_EX, FB_EN, 0, p.B, lane);
            rows_zero_fill(ds, 12, s0, FB_TE, 0, p.B, lane);
            rows_zero_fill(dm, 4, s0, FB_TE, 0, p.B, lane);
            rows_zero_fill(dx, 4, 2 * s0, 2 * FB_TE, 0, p.L, lane);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(fullA + sa, rows_bytes(8, s0 - FB_EX, FB_EN, 0, p.B) + rows_bytes(12, s0, FB_TE, 0, p.B));
                rows_issue(de, p.erows, p.B, 8, s0 - FB_EX, FB_EN, 0, p.B, fullA + sa);
                rows_issue(ds, p.srows, p.B, 12, s0, FB_TE, 0, p.B, fullA + sa);
                mbar_arrive_expect_tx(fullB + sb, rows_bytes(4, s0, FB_TE, 0, p.B) + rows_bytes(4, 2 * s0, 2 * FB_TE, 0, p.L));
                rows_issue(dm, p.m1rows, p.B, 4, s0, FB_TE, 0, p.B, fullB + sb);
                rows_issue(dx, p.rx, p.ld_rx, 4, 2 * s0, 2 * FB_TE, 0, p.L, fullB + sb);
            }
        };
        if (wid == 2 * FB_NA && (int)blockIdx.x < nt) produce(blockIdx.x, 0);
        int it = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < nt; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t pf = (it >> 1) & 1, pe = pf ^ 1;
            const int s0 = p.sym_lo + tile * FB_T - FB_HP, u0 = s0 + FT_R * l;
            float4 *ge = gew + (s * 2 + 0) * FB_ES, *go = gew + (s * 2 + 1) * FB_ES, *gy = gyw + s * FB_GS;
            mbar_wait(fullA + s, pf);
            mbar_wait(ge_free + s, pe);                          // C is done with tile it-2 (its windows and its ring-B stage)
            mbar_wait(gy_free + s, pe);                          // B likewise
            {   // residual rows -> window layout (own positions; the 4 groups beyond the slots by threads 0..3)
                const float4 *er4 = reinterpret_cast<const float4 *>(eraw + s * 8 * FB_EN);
#pragma unroll 1
                for (int g = l; g < FB_EN / 4; g += FB_SL) {
                    float4 er[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) er[k] = er4[k * (FB_EN / 4) + g];
#pragma unroll
                    for (int r = 0; r < FT_R; ++r) {
                        ge[5 * g + r] = make_float4(f4c(er[0], r), f4c(er[2], r), f4c(er[1], r), f4c(er[3], r));
                        go[5 * g + r] = make_float4(f4c(er[4], r), f4c(er[6], r), f4c(er[5], r), f4c(er[7], r));
                    }
                }
            }
            named_bar_sync(1, FB_SL);                            // every A warp has left tile it-1: ring A's other stage is free
            if (l == 0) mbar_arrive(ge_ready + s);
            if (wid == 2 * FB_NA && tile + (int)gridDim.x < nt) produce(tile + gridDim.x, it + 1);
            // dL/dE_q(u) = sum_chi sum_j conj(h[chi][nu][j]) gD_chi(2u - MH + j): even j -> ge[u + a - HF], odd j -> go[u + a - HF]
            FirAccS fa;
#pragma unroll
            for (int r = 0; r < FT_R; ++r)
#pragma unroll
                for (int o = 0; o < 2; ++o) fa.P1[r][o] = fa.P2[r][o] = fa.P3[r][o] = make_float2(0.f, 0.f);
            constexpr int MA = (FB_EX - HF) & 3, QA = (FB_EX - HF) >> 2;      // first window element of slot l: 4 (l + QA) + MA
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) fir4s<MA>(ph ? go : ge, l + QA, tapG + (ph ? FT_TAPV * NE : 0), ph ? NO : NE, fa);
            float gE[FT_R][4];
#pragma unroll
            for (int r = 0; r < FT_R; ++r)
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const float p1 = fa.P1[r][o].x + fa.P1[r][o].y, p2 = fa.P2[r][o].x + fa.P2[r][o].y, p3 = fa.P3[r][o].x + fa.P3[r][o].y;
                    gE[r][2 * o] = p1 - p3;
                    gE[r][2 * o + 1] = p1 + p2;
                }
            // dL/dout = dL/dE_q * S1 + dL/dVar * T2 + w * S3 with the coefficients the forward pass left in srows
            const float4 *sr4 = reinterpret_cast<const float4 *>(sraw + s * 12 * FB_TE);
            float gyv[4][FT_R];
#pragma unroll
            for (int pol = 0; pol < 2; ++pol) {
                float gV[FT_R], entw[FT_R];
#pragma unroll
                for (int r = 0; r < FT_R; ++r) {
                    const int u = u0 + r;
                    const int jlo = min(M, max(0, 2 * MH - 2 * u)), jhi = max(0, min(M, p.L - 2 * u));
                    gV[r] = PSg[pol * (M + 1) + jhi] - PSg[pol * (M + 1) + jlo];
                    entw[r] = (u >= MH && u < p.B - MH) ? LN2 : 0.f;
                }
#pragma unroll
                for (int cq = 0; cq < 2; ++cq) {
                    const int cc = 2 * pol + cq;
                    const float4 s1 = sr4[cc * FB_SL + l], t2 = sr4[(4 + cc) * FB_SL + l], s3 = sr4[(8 + cc) * FB_SL + l];
#pragma unroll
                    for (int r = 0; r < FT_R; ++r) gyv[cc][r] = fmaf(gE[r][cc], f4c(s1, r), fmaf(gV[r], f4c(t2, r), entw[r] * f4c(s3, r)));
                }
            }
#pragma unroll
            for (int r = 0; r < FT_R; ++r) gy[5 * l + r] = make_float4(gyv[0][r], gyv[2][r], gyv[1][r], gyv[3][r]);
            __syncwarp();
            if (lane == 0) mbar_arrive(gy_ready + s);
        }
    } else {
        // ================= families B (dW) and C (dh): role = (phase, lag half) =================
        const bool isB = wid < FB_NA;
        const int role = wid & 3, ph = role >> 1;
        const int a0 = (role & 1) ? (ph ? BASE : BASE + 1) : 0;  // first window offset of the role; even phase: 0..B-1 | B (shared) | B+1..2B, odd: 0..B-1 | B..2B-1
        // window position of offset a' = 0 for slot position 0:  B: dL/dout(p + a' - HF + ph), window index 0 <-> symbol s0;
        //                                                        C: e_ph(u + a' - HF), window index 0 <-> symbol s0 - FB_EX
        const int c0 = (isB ? -HF + ph : FB_EX - HF) + a0, c0x = (isB ? -HF : FB_EX - HF) + BASE;
        const int m4 = c0 & 3, q0 = c0 >> 2, q0x = c0x >> 2;     // c0x is a multiple of 4 (HF = BASE, FB_EX = 8)
        static_assert(FB_EX % 4 == 0, "shared-lag window offset");
        float2 acc2[AMAX][4];                                    // tap accumulators
#pragma unroll
        for (int a = 0; a < AMAX; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc2[a][k] = make_float2(0.f, 0.f);
        int it = 0;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < nt; tile += gridDim.x, ++it) {
            const int s = it & 1, sb = it % FB_RB;
            const uint32_t pf = (it >> 1) & 1, pfb = (it / FB_RB) & 1;
            const int s0 = p.sym_lo + tile * FB_T - FB_HP;
            const float4 *win = isB ? gyw + s * FB_GS : gew + (s * 2 + ph) * FB_ES;
            const float4 *winx = isB ? win : gew + (s * 2 + 0) * FB_ES;     // even-phase window of the shared lag
            mbar_wait(fullB + sb, pfb);
            mbar_wait((isB ? gy_ready : ge_ready) + s, pf);
#pragma unroll 1
            for (int g = 0; g < FB_SL / 32; ++g) {
                const int sl = g * 32 + lane, li0 = FT_R * sl, uu0 = s0 + li0;
                if (!(li0 >= FB_HP && li0 < FB_HP + FB_T && uu0 < p.sym_hi)) continue;     // margin slots own nothing
                float4 gd[FT_R], gx[FT_R];                       // per-symbol operand of this phase / of the even phase
                if (isB) {
                    const float4 *x4 = reinterpret_cast<const float4 *>(xraw + sb * 4 * 2 * FB_TE);
                    float4 va[4], vb[4];                         // samples {e0,o0,e1,o1}, {e2,o2,e3,o3} of the 4 rows
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        va[k] = x4[k * (2 * FB_TE / 4) + 2 * sl];
                        vb[k] = x4[k * (2 * FB_TE / 4) + 2 * sl + 1];
                    }
                    gx[0] = make_float4(va[0].x, va[1].x, va[2].x, va[3].x);
                    gx[1] = make_float4(va[0].z, va[1].z, va[2].z, va[3].z);
                    gx[2] = make_float4(vb[0].x, vb[1].x, vb[2].x, vb[3].x);
                    gx[3] = make_float4(vb[0].z, vb[1].z, vb[2].z, vb[3].z);
                    if (ph) {
                        gd[0] = make_float4(va[0].y, va[1].y, va[2].y, va[3].y);
                        gd[1] = make_float4(va[0].w, va[1].w, va[2].w, va[3].w);
                        gd[2] = make_float4(vb[0].y, vb[1].y, vb[2].y, vb[3].y);
                        gd[3] = make_float4(vb[0].w, vb[1].w, vb[2].w, vb[3].w);
                    } else {
#pragma unroll
                        for (int r = 0; r < FT_R; ++r) gd[r] = gx[r];
                    }
                } else {
                    const float4 *m4r = reinterpret_cast<const float4 *>(m1raw + sb * 4 * FB_TE);
                    const float4 m0 = m4r[sl], m1 = m4r[FB_SL + sl], m2 = m4r[2 * FB_SL + sl], m3 = m4r[3 * FB_SL + sl];
                    gd[0] = make_float4(m0.x, m1.x, m2.x, m3.x);
                    gd[1] = make_float4(m0.y, m1.y, m2.y, m3.y);
                    gd[2] = make_float4(m0.z, m1.z, m2.z, m3.z);
                    gd[3] = make_float4(m0.w, m1.w, m2.w, m3.w);
#pragma unroll
                    for (int r = 0; r < FT_R; ++r) gx[r] = gd[r];
                }
                float2(&accB)[BASE][4] = reinterpret_cast<float2(&)[BASE][4]>(acc2);
                switch (m4) {                                    // warp-uniform: compile-time padded offsets inside
                    case 0: corr4m<BASE, 0>(win, sl + q0, gd, accB); break;
                    case 1: corr4m<BASE, 1>(win, sl + q0, gd, accB); break;
                    case 2: corr4m<BASE, 2>(win, sl + q0, gd, accB); break;
                    default: corr4m<BASE, 3>(win, sl + q0, gd, accB); break;
                }
                if ((g & 3) == role) corr4m<1, 0>(winx, sl + q0x, gx, reinterpret_cast<float2(&)[1][4]>(acc2[BASE]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive((isB ? gy_free : ge_free) + s);
        }
        // warp reduction of the accumulators: acc2[a][2 o + c] is the pair over i of output (o, i), c = 0 real, 1 imaginary
#pragma unroll
        for (int a = 0; a < AMAX; ++a)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 v2 = acc2[a][2 * (k >> 2) + (k & 1)];
                const float sum = warp_sum(((k >> 1) & 1) ? v2.y : v2.x);
                if (lane == 0) red[wid * (AMAX * 8) + a * 8 + k] = sum;
            }
    }
    __syncthreads();
    if (wid < 2 * FB_NA) {                                      // publish this CTA's partial (fixed order)
        const bool isB = wid < FB_NA;
        const int fam0 = isB ? 0 : FB_NA, role = wid & 3, ph = role >> 1;
        const int a0 = (role & 1) ? (ph ? BASE : BASE + 1) : 0, n_real = role == 0 ? BASE + 1 : BASE;
        float *dst = p.gpart + (int64_t)blockIdx.x * 16 * M;
        for (int idx = lane; idx < n_real * 8; idx += 32) {
            const int a = idx >> 3, k = idx & 7, oc = k >> 2, ic = (k >> 1) & 1, cidx = k & 1;     // oc: per-symbol operand's index, ic: window's
            float sum = 0.f;
            if (a == BASE) {                                     // the shared lag (role 0 publishes it): one partial per role, fixed order
#pragma unroll
                for (int w = 0; w < 4; ++w) sum += red[(fam0 + w) * (AMAX * 8) + idx];
            } else {
                sum = red[(fam0 + role) * (AMAX * 8) + idx];
            }
            if (cidx) sum = -sum;                                // corr4 formed x conj(gy) / E_q conj(e): the gradients are the conjugates
            const int ap = a0 + a;                               // window offset a'
            if (isB) dst[(ic * 4 + 2 * cidx + oc) * M + (2 * MH - ph - 2 * ap)] = sum;                                      // o = gy pol = ic, i = rx pol = oc
            else dst[8 * M + ((ic * 2 + oc) * 2 + cidx) * M + (ph + 2 * ap)] = 2.f * p.scal[DP_KAPPA_OFF + ic] * sum;       // chi = ic, nu = oc; gD = 2 kappa_chi e
        }
    }
}

template <int MH>
static size_t fused_bwd_smem() {
    return 8 * FB_NBAR + (size_t)(2 * 8 * FB_EN + 2 * 12 * FB_TE + FB_RB * 4 * FB_TE + FB_RB * 4 * 2 * FB_TE) * sizeof(float) +
           (size_t)(2 * 2 * FB_ES + 2 * FB_GS + FT_TAPV * (2 * MH + 1)) * sizeof(float4) +
           (size_t)(2 * (2 * MH + 2) + 8 * (MH / 2 + 1) * 8) * sizeof(float) + 128;
}

template <int MH>
static int fused_bwd_launch_t(DpK p, cudaStream_t st, int *nparts) {
    static SmemAttrCache set;
    const size_t sm = fused_bwd_smem<MH>();
    if (int rc = ensure_dyn_smem(k_dp_bwd_fused<MH>, sm, set)) return rc;
    p.ntiles = (p.sym_hi - p.sym_lo + FB_T - 1) / FB_T;
    const int grid = min(sm_count(), p.ntiles);
    ktime_begin(VAEQ_K_DP_BWD, st);
    k_dp_bwd_fused<MH><<<grid, FB_NT, sm, st>>>(p);
    ktime_end(VAEQ_K_DP_BWD, st);
    VAEQ_LAUNCH_CHECK("k_dp_bwd_fused");
    *nparts = grid;
    return VAEQ_OK;
}

// backward pass of the fast path as one launch; returns 0 if M_est does not qualify (the caller then runs the three kernels)
int dp_bwd_fused_launch(const DpK &p, cudaStream_t st, int *nparts, int *rc) {
    switch (p.mh) {
        case 12: *rc = fused_bwd_launch_t<12>(p, st, nparts); return 1;
        case 6: *rc = fused_bwd_launch_t<6>(p, st, nparts); return 1;
        case 4: *rc = fused_bwd_launch_t<4>(p, st, nparts); return 1;
        case 2: *rc = fused_bwd_launch_t<2>(p, st, nparts); return 1;
        default: return 0;
    }
}

}  // namespace vaeq
