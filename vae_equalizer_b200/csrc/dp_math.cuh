// Point-wise math of the DP VAE step shared by every kernel variant (fp32, reference op order).
#pragma once
#include "common.cuh"

namespace vaeq {

// Per-run demapper constants staged once per CTA in shared memory.
struct DemapConst {
    float amp[VAEQ_MAX_LEVELS];   // a_l                       (sf:568)
    float a2[VAEQ_MAX_LEVELS];    // a_l^2
    float nua2[VAEQ_MAX_LEVELS];  // float(nu_sc) * a_l^2       (sf:520 PCS term)
    float P[VAEQ_MAX_LEVELS];     // prior pmf                  (sf:572)
    float var[2];                 // demapper variance per pol  (sf:581)
    float pad[2];
};

__device__ __forceinline__ void load_demap_const(DemapConst *c, const float *amp, const float *P, const float *var,
                                                 float nu_sc, int n_lev) {
    const int t = threadIdx.x;
    if (t < VAEQ_MAX_LEVELS) {
        float a = t < n_lev ? amp[t] : 0.f;
        c->amp[t] = a;
        c->a2[t] = a * a;
        c->nua2[t] = nu_sc * (a * a);
        c->P[t] = (t < n_lev && P != nullptr) ? P[t] : 1.f;
    }
    if (t < 2) c->var[t] = var[t];
}

// q_l = softmin_l( (y-a_l)^2/2/var + nu_sc a_l^2 )   (sf:521-523), max-subtracted like nn.Softmin.
// Also returns the first two posterior moments (sf:108-112).
template <int NL>
__device__ __forceinline__ void demap_component(float y, float var, const DemapConst &c, float (&q)[NL], float &m1,
                                                float &m2) {
    float z[NL];
    float zmin = 3.0e38f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float d = y - c.amp[l];
        z[l] = __fadd_rn(__fdiv_rn(__fmul_rn(__fmul_rn(d, d), 0.5f), var), c.nua2[l]);
        zmin = fminf(zmin, z[l]);
    }
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        q[l] = expf(zmin - z[l]);
        s += q[l];
    }
    m1 = 0.f;
    m2 = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        q[l] = __fdiv_rn(q[l], s);
        m1 = fmaf(c.amp[l], q[l], m1);
        m2 = fmaf(c.a2[l], q[l], m2);
    }
}

// entropy contribution  sum_l -q_l log(q_l/P_l + 1e-12)   (sf:132)
template <int NL>
__device__ __forceinline__ float entropy_component(const float (&q)[NL], const DemapConst &c) {
    float e = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) e -= q[l] * logf(__fdiv_rn(q[l], c.P[l]) + 1e-12f);
    return e;
}

// Back-propagate through moments + entropy + softmin for one I/Q component:
//   g1 = dL/dE_q[x] (already including -2 E_q[x] dL/dVar), g2 = dL/dE_q[x^2] = dL/dVar,
//   ent_on = 1 if this symbol is inside the entropy window.  Returns dL/dy.
template <int NL>
__device__ __forceinline__ float demap_backward(float y, float var, const DemapConst &c, const float (&q)[NL], float g1,
                                                float g2, bool ent_on) {
    float gq[NL];
    float dot = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float g = fmaf(c.amp[l], g1, c.a2[l] * g2);
        if (ent_on) {
            float u = __fdiv_rn(q[l], c.P[l]);
            float t = u + 1e-12f;
            g += logf(t) + __fdiv_rn(u, t);
        }
        gq[l] = g;
        dot = fmaf(q[l], g, dot);
    }
    float gy = 0.f;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float gz = -q[l] * (gq[l] - dot);
        gy = fmaf(gz, y - c.amp[l], gy);
    }
    return __fdiv_rn(gy, var);
}

}  // namespace vaeq
