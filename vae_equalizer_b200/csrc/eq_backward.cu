// Backward of twoXtwoFIR.forward (sf:500-527) as a stand-alone operator: given ARBITRARY upstream gradients dL/dq and dL/dout it
// returns dL/dconv_w.weight.  The fused training step never comes here (it differentiates loss(W, h) directly, dp_fast.cu /
// dp_step.cu); this is what autograd calls when a caller derives something else from q or out (a clamp, a mask, another
// loss) before taking gradients, so that the drop-in nn.Module differentiates like the reference's.
//   softmin backward:  z_l = (y - a_l)^2 / (2 var) + nu_sc a_l^2,  q = softmin(z)  =>  dq_l/dy = -q_l (z'_l - sum_k q_k z'_k),
//                      z'_l = (y - a_l) / var   (the PCS term does not depend on y)
//   FIR backward:      out_I[o] = conv(x_in_I, w[o]), out_Q[o] = conv(x_in_Q, w[o]) with x_in_I = [xI_0, xI_1, -xQ_0, -xQ_1],
//                      x_in_Q = [xQ_0, xQ_1, xI_0, xI_1] (sf:504-509)  =>
//                      dW[o][c][k] = sum_n gyI_o[n] x_in_I[c][2n + k - mh] + gyQ_o[n] x_in_Q[c][2n + k - mh]   (zero padded)
// Off the hot path: plain kernels, per-CTA partials summed in a fixed order.
#include "common.cuh"

namespace vaeq {

constexpr int EB_NT = 256;
constexpr int EB_CHUNK = 1024;             // symbols per CTA of the tap-gradient kernel

template <int NL>
__global__ void __launch_bounds__(EB_NT) k_eq_gy(const float *q, int64_t ld_q, const float *out, int64_t ld_out, const float *gq,
                                                  int64_t ld_gq, const float *gout, int64_t ld_gout, const float *amp, const float *var,
                                                  int B, float *gy) {
    __shared__ float a[NL];
    if (threadIdx.x < NL) a[threadIdx.x] = amp[threadIdx.x];
    __syncthreads();
    for (int64_t idx = (int64_t)blockIdx.x * EB_NT + threadIdx.x; idx < 4 * (int64_t)B; idx += (int64_t)gridDim.x * EB_NT) {
        const int cc = (int)(idx / B), t = (int)(idx - (int64_t)cc * B);
        float g = gout ? gout[(int64_t)cc * ld_gout + t] : 0.f;
        if (gq) {
            const float y = out[(int64_t)cc * ld_out + t], iv = 1.f / var[cc >> 1];
            float ql[NL], zp[NL], zbar = 0.f;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                ql[l] = q[(int64_t)(cc * NL + l) * ld_q + t];
                zp[l] = (y - a[l]) * iv;
                zbar = fmaf(ql[l], zp[l], zbar);
            }
#pragma unroll
            for (int l = 0; l < NL; ++l) g -= gq[(int64_t)(cc * NL + l) * ld_gq + t] * ql[l] * (zp[l] - zbar);
        }
        gy[idx] = g;
    }
}

// thread j < 8M owns tap (o, c, k); a CTA sums over EB_CHUNK symbols staged in shared memory
__global__ void __launch_bounds__(EB_NT) k_eq_dw(const float *rx, int64_t ld_rx, int L, const float *gy, int B, int M, double *part) {
    extern __shared__ float sm[];
    const int mh = M / 2, n0 = blockIdx.x * EB_CHUNK, nn = min(EB_CHUNK, B - n0);
    const int xs0 = 2 * n0 - mh, xn = 2 * EB_CHUNK + M - 1;          // staged sample window [xs0, xs0 + xn)
    float *xw = sm, *gw = sm + 4 * xn;                                // xw[r][xn], gw[4][EB_CHUNK]
    for (int i = threadIdx.x; i < 4 * xn; i += EB_NT) {
        const int r = i / xn, s = xs0 + (i - r * xn);
        xw[i] = (s >= 0 && s < L) ? rx[(int64_t)r * ld_rx + s] : 0.f;
    }
    for (int i = threadIdx.x; i < 4 * EB_CHUNK; i += EB_NT) {
        const int cc = i / EB_CHUNK, t = i - cc * EB_CHUNK;
        gw[i] = t < nn ? gy[(int64_t)cc * B + n0 + t] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 8 * M; j += EB_NT) {
        const int o = j / (4 * M), c = (j / M) & 3, k = j % M;
        // rows of rx: (pol, I/Q) = 2*pol + iq.  x_in_I[c]: c<2 -> +xI_c, else -xQ_{c-2};  x_in_Q[c]: c<2 -> xQ_c, else xI_{c-2}
        const int pol = c & 1;
        const float *xI = xw + (2 * pol) * xn + k, *xQ = xw + (2 * pol + 1) * xn + k;
        const float *gI = gw + (2 * o) * EB_CHUNK, *gQ = gw + (2 * o + 1) * EB_CHUNK;
        double acc = 0.0;
        float a = 0.f;
        for (int t = 0; t < nn; ++t) {
            const float xi = xI[2 * t], xq = xQ[2 * t];
            a += c < 2 ? gI[t] * xi + gQ[t] * xq : gQ[t] * xi - gI[t] * xq;
            if ((t & 63) == 63) {
                acc += (double)a;
                a = 0.f;
            }
        }
        part[(int64_t)blockIdx.x * 8 * M + j] = acc + (double)a;
    }
}

__global__ void k_eq_dw_sum(const double *part, int nparts, int n, float *gW) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[(int64_t)i * n + j];
    gW[j] = (float)s;
}

}  // namespace vaeq

using namespace vaeq;

extern "C" size_t vaeq_eq_backward_scratch_bytes(int32_t B, int32_t M) {
    if (B <= 0 || M <= 0) return 0;
    const size_t chunks = ((size_t)B + EB_CHUNK - 1) / EB_CHUNK;
    return align_up((size_t)4 * B * sizeof(float), 256) + align_up(chunks * 8 * M * sizeof(double), 256);
}

extern "C" int vaeq_eq_backward(const float *rx, int64_t ld_rx, const float *q, int64_t ld_q, const float *out, int64_t ld_out,
                                const float *gq, int64_t ld_gq, const float *gout, int64_t ld_gout, const float *amp, const float *var,
                                int32_t n_lev, int32_t B, int32_t M, float *gW, void *scratch, void *stream) {
    VAEQ_CHECK_ARG(rx && amp && var && gW && scratch && (gq || gout), "NULL pointer (rx, amp, var, gW, scratch; one of gq / gout)");
    VAEQ_CHECK_ARG(!gq || (q && out), "dL/dq needs q and out");
    VAEQ_CHECK_ARG(n_lev == 2 || n_lev == 4 || n_lev == 8, "n_lev=%d must be 2, 4 or 8", n_lev);
    VAEQ_CHECK_ARG(B > 0 && M >= 1 && M <= VAEQ_MAX_TAPS && (M & 1), "bad B=%d / M_est=%d", B, M);
    cudaStream_t st = (cudaStream_t)stream;
    float *gy = static_cast<float *>(scratch);
    double *part = reinterpret_cast<double *>(static_cast<char *>(scratch) + align_up((size_t)4 * B * sizeof(float), 256));
    const int grid_gy = (int)std::min<int64_t>((4 * (int64_t)B + EB_NT - 1) / EB_NT, (int64_t)sm_count() * 8);
    ktime_begin(VAEQ_K_OTHER, st);
    switch (n_lev) {
        case 2: k_eq_gy<2><<<grid_gy, EB_NT, 0, st>>>(q, ld_q, out, ld_out, gq, ld_gq, gout, ld_gout, amp, var, B, gy); break;
        case 4: k_eq_gy<4><<<grid_gy, EB_NT, 0, st>>>(q, ld_q, out, ld_out, gq, ld_gq, gout, ld_gout, amp, var, B, gy); break;
        default: k_eq_gy<8><<<grid_gy, EB_NT, 0, st>>>(q, ld_q, out, ld_out, gq, ld_gq, gout, ld_gout, amp, var, B, gy); break;
    }
    ktime_end(VAEQ_K_OTHER, st);
    VAEQ_LAUNCH_CHECK("k_eq_gy");
    const int chunks = (B + EB_CHUNK - 1) / EB_CHUNK;
    const size_t smem = (size_t)(4 * (2 * EB_CHUNK + M - 1) + 4 * EB_CHUNK) * sizeof(float);
    static SmemAttrCache set_smem;
    if (int rc = ensure_dyn_smem(k_eq_dw, smem, set_smem)) return rc;
    ktime_begin(VAEQ_K_OTHER, st);
    k_eq_dw<<<chunks, EB_NT, smem, st>>>(rx, ld_rx, 2 * B, gy, B, M, part);
    ktime_end(VAEQ_K_OTHER, st);
    VAEQ_LAUNCH_CHECK("k_eq_dw");
    ktime_begin(VAEQ_K_OTHER, st);
    k_eq_dw_sum<<<(8 * M + 127) / 128, 128, 0, st>>>(part, chunks, 8 * M, gW);
    ktime_end(VAEQ_K_OTHER, st);
    VAEQ_LAUNCH_CHECK("k_eq_dw_sum");
    return VAEQ_OK;
}
