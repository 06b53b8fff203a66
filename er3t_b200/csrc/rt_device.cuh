// rt_device.cuh -- device-side building blocks of the sm_100a photon-transport kernels:
// counter-based Philox4x32-10 streams, phase-function sampling/evaluation (Rayleigh, HG, tabulated),
// surface BRDF models (Lambertian, DSM/Cox-Munk, LSRT).  fp32 in flight, fp64 only in tallies.
//
// Input semantics follow the reference's solver contract (er3t/rtm/mca/mca_inp.py:19-364):
//   apf encoding        er3t/rtm/mca/mca_atm.py:101,262,276-277,301 ; er3t/rtm/mca/util.py:153
//   phase tables        er3t/rtm/mca/mca_sca.py:82-95
//   surface parameters  er3t/rtm/mca/mca_sfc.py:89-133 ; er3t/pre/sfc/sfc_gen.py:119-145
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RT_PI 3.14159265358979323846f
#define RT_2PI 6.28318530717958647692f
#define RT_INF 3.0e38f

// ------------------------------------------------------------------ Philox4x32-10
struct Philox4 {
    uint32_t k0, k1;       // key   = job seed
    uint32_t c0, c1;       // ctr.xy = global photon index
    uint32_t c2, c3;       // ctr.z = draw counter, ctr.w = stream id
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u01(uint32_t x) {   // (0,1), 24 bits
    return (float(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float4 rng4(Philox4& g) {
    const uint4 r = philox4x32_10(g.c0, g.c1, g.c2, g.c3, g.k0, g.k1);
    g.c2++;
    return make_float4(u01(r.x), u01(r.y), u01(r.z), u01(r.w));
}

// ------------------------------------------------------------------ small vector helpers
__device__ __forceinline__ float3 rotate_dir(const float3 d, float mu, float phi) {
    const float st = sqrtf(fmaxf(0.0f, 1.0f - mu * mu));
    float sp, cp;
    __sincosf(phi, &sp, &cp);
    float3 r;
    if (fabsf(d.z) > 0.99999f) {
        const float sg = d.z > 0 ? 1.0f : -1.0f;
        r.x = st * cp; r.y = st * sp * sg; r.z = mu * sg;
    } else {
        const float den = sqrtf(1.0f - d.z * d.z);
        const float inv = 1.0f / den;
        r.x = st * (d.x * d.z * cp - d.y * sp) * inv + d.x * mu;
        r.y = st * (d.y * d.z * cp + d.x * sp) * inv + d.y * mu;
        r.z = -st * cp * den + d.z * mu;
    }
    const float n = rsqrtf(r.x * r.x + r.y * r.y + r.z * r.z);
    r.x *= n; r.y *= n; r.z *= n;
    return r;
}

// ------------------------------------------------------------------ phase functions
// Normalisation everywhere: (1/2) int_{-1}^{1} P(mu) dmu = 1, local estimates use P / (4 pi).
struct PhaseTab {
    int npf, nang;
    const float* mu;    // [nang]        cos(scattering angle), DEcreasing (angle increasing)
    const float* p;     // [npf][nang]   normalised phase function, piecewise linear in mu
    const float* cdf;   // [npf][nang]   F[j] = Prob(angle <= ang[j]) ; F[0] = 0, F[nang-1] = 1
    // O(1) guide tables (accelerators only: the interval found is the one a full search finds, see tab_*1)
    const unsigned short* gs;   // [npf][RT_NGS]  sampling: largest j with F[j] <= k / RT_NGS
    const unsigned short* ge;   // [RT_NGE]       evaluation: an index at or before the interval of any mu whose
                                //                q = sqrt(2 - 2 mu) falls into bin k (bins uniform in q ~ angle)
};
#define RT_NGS 2048
#define RT_NGE 2048

__device__ __forceinline__ float hg_eval(float g, float mu) {
    const float d = 1.0f + g * g - 2.0f * g * mu;
    return (1.0f - g * g) * rsqrtf(d) / d;
}
__device__ __forceinline__ float hg_sample(float g, float xi) {
    if (fabsf(g) < 1e-4f) return 2.0f * xi - 1.0f;
    const float s = (1.0f - g * g) / (1.0f - g + 2.0f * g * xi);
    const float mu = (1.0f + g * g - s * s) / (2.0f * g);
    return fminf(1.0f, fmaxf(-1.0f, mu));
}
__device__ __forceinline__ float ray_eval(float mu) { return 0.75f * (1.0f + mu * mu); }
__device__ __forceinline__ float ray_sample(float xi) {
    const float u = 4.0f * xi - 2.0f;
    const float q = cbrtf(u + sqrtf(u * u + 1.0f));
    return fminf(1.0f, fmaxf(-1.0f, q - 1.0f / q));
}

// Tabulated phase functions.  MCARaTS looks its tables up in O(1) (Sca_ntg = 20000 equal-probability bins,
// er3t/rtm/mca/mca_inp.py:53); here a guide table gives the neighbourhood and a short scan finds the EXACT interval of
// the caller's angle grid (largest lo with m[lo] >= mu, resp. F[lo] <= xi -- what a binary search over the whole table
// returns), so the piecewise-linear function that is sampled and evaluated does not depend on the guide resolution.
// Expected scan length < 1 step on the 498-angle Mie grid (9 dependent loads for the binary search it replaces).
__device__ __noinline__ float tab_eval1(const PhaseTab& T, int it, float mu) {
    const float* m = T.mu;
    const int n = T.nang;
    if (mu >= __ldg(m)) return __ldg(T.p + size_t(it) * n);
    if (mu <= __ldg(m + n - 1)) return __ldg(T.p + size_t(it) * n + n - 1);
    const float q = sqrtf(fmaxf(0.0f, 2.0f - 2.0f * mu));
    int lo = int(__ldg(T.ge + min(RT_NGE - 1, int(q * (0.5f * RT_NGE)))));
    float m0 = __ldg(m + lo);
    while (lo > 0 && m0 < mu) { --lo; m0 = __ldg(m + lo); }              // guard (rounding of q); normally not taken
    float m1 = __ldg(m + lo + 1);
    while (lo + 1 < n - 1 && m1 >= mu) { ++lo; m0 = m1; m1 = __ldg(m + lo + 1); }
    const float p0 = __ldg(T.p + size_t(it) * n + lo), p1 = __ldg(T.p + size_t(it) * n + lo + 1);
    const float f = (m0 - mu) / (m0 - m1);
    return p0 + f * (p1 - p0);
}

__device__ __noinline__ float tab_sample1(const PhaseTab& T, int it, float xi) {
    const int n = T.nang;
    const float* F = T.cdf + size_t(it) * n;
    int lo = int(__ldg(T.gs + size_t(it) * RT_NGS + min(RT_NGS - 1, int(xi * float(RT_NGS)))));
    float F0 = __ldg(F + lo);
    while (lo > 0 && F0 > xi) { --lo; F0 = __ldg(F + lo); }              // guard; normally not taken
    float F1 = __ldg(F + lo + 1);
    while (lo + 1 < n - 1 && F1 <= xi) { ++lo; F0 = F1; F1 = __ldg(F + lo + 1); }
    const int hi = lo + 1;
    const float m0 = __ldg(T.mu + lo), m1 = __ldg(T.mu + hi);
    const float p0 = __ldg(T.p + size_t(it) * n + lo), p1 = __ldg(T.p + size_t(it) * n + hi);
    const float dm = m0 - m1;
    const float c = 2.0f * (xi - F0);
    const float s = (p1 - p0) / dm;
    const float disc = fmaxf(0.0f, p0 * p0 + 2.0f * s * c);
    const float den = p0 + sqrtf(disc);
    float t = den > 0.0f ? 2.0f * c / den : 0.0f;
    t = fminf(dm, fmaxf(0.0f, t));
    return m0 - t;
}

// apf decoding: <= -1 Rayleigh ; (-1, 1) Henyey-Greenstein g ; >= 1 real-valued 1-based table index
__device__ __forceinline__ float phase_eval(const PhaseTab& T, float apf, float mu) {
    if (apf <= -1.0f) return ray_eval(mu);
    if (apf < 1.0f) return hg_eval(apf, mu);
    if (T.npf == 0) return 1.0f;
    const float a = fminf(float(T.npf), fmaxf(1.0f, apf));
    int i = int(floorf(a));
    float f = a - float(i);
    if (i >= T.npf) { i = T.npf; f = 0.0f; }
    float v = tab_eval1(T, i - 1, mu);
    if (f > 0.0f) v = (1.0f - f) * v + f * tab_eval1(T, i, mu);
    return v;
}
__device__ __forceinline__ float phase_sample(const PhaseTab& T, float apf, float xi, float xi_tab) {
    if (apf <= -1.0f) return ray_sample(xi);
    if (apf < 1.0f) return hg_sample(apf, xi);
    if (T.npf == 0) return 2.0f * xi - 1.0f;
    const float a = fminf(float(T.npf), fmaxf(1.0f, apf));
    int i = int(floorf(a));
    float f = a - float(i);
    if (i >= T.npf) { i = T.npf; f = 0.0f; }
    if (f > 0.0f && xi_tab < f) i += 1;
    return tab_sample1(T, i - 1, xi);
}

// ------------------------------------------------------------------ surface BRDF models
// wi, wo: unit vectors pointing AWAY from the surface (z > 0); wi = -incident direction.
__device__ __forceinline__ float fresnel_unpol(float cosg, float nr, float ni) {
    // complex m^2, g = sqrt(m^2 - 1 + cos^2)
    const float ar = nr * nr - ni * ni, ai = 2.0f * nr * ni;            // m^2
    const float br = ar - 1.0f + cosg * cosg, bi = ai;                  // g^2
    const float mod = sqrtf(br * br + bi * bi);
    float gr = sqrtf(fmaxf(0.0f, 0.5f * (mod + br)));
    float gi = sqrtf(fmaxf(0.0f, 0.5f * (mod - br)));
    if (bi < 0.0f) gi = -gi;
    // rs = (c - g)/(c + g)
    const float n1r = cosg - gr, n1i = -gi, d1r = cosg + gr, d1i = gi;
    const float rs2 = (n1r * n1r + n1i * n1i) / (d1r * d1r + d1i * d1i);
    // rp = (m2 c - g)/(m2 c + g)
    const float n2r = ar * cosg - gr, n2i = ai * cosg - gi, d2r = ar * cosg + gr, d2i = ai * cosg + gi;
    const float rp2 = (n2r * n2r + n2i * n2i) / (d2r * d2r + d2i * d2i);
    return 0.5f * (rs2 + rp2);
}
__device__ __forceinline__ float cm_lambda(float mu, float sig2) {
    if (mu >= 0.999999f) return 0.0f;
    const float cot = mu * rsqrtf(fmaxf(1e-30f, 1.0f - mu * mu));
    const float nu = cot * rsqrtf(sig2);
    if (nu > 6.0f) return 0.0f;
    return 0.5f * (__expf(-nu * nu) / (1.7724538509f * nu) - erfcf(nu));
}
__device__ __forceinline__ float cm_shadow(float mui, float mur, float sig2) {
    return 1.0f / (1.0f + cm_lambda(mui, sig2) + cm_lambda(mur, sig2));
}
__device__ __forceinline__ float dsm_spec_brdf(const float* p, const float3 wi, const float3 wo) {
    const float sig2 = fmaxf(1e-6f, p[4]);
    float3 h = make_float3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z);
    const float hn2 = h.x * h.x + h.y * h.y + h.z * h.z;
    if (hn2 <= 0.0f || h.z <= 0.0f) return 0.0f;
    const float inv = rsqrtf(hn2);
    h.x *= inv; h.y *= inv; h.z *= inv;
    const float cosg = wi.x * h.x + wi.y * h.y + wi.z * h.z;
    if (cosg <= 0.0f) return 0.0f;
    const float cn = h.z, cn2 = cn * cn;
    const float tan2 = (1.0f - cn2) / cn2;
    const float P = expf(-tan2 / sig2) / (RT_PI * sig2);
    const float F = fresnel_unpol(cosg, p[2], p[3]);
    return F * P * cm_shadow(wi.z, wo.z, sig2) / (4.0f * wi.z * wo.z * cn2 * cn2);
}
__device__ __forceinline__ float lsrt_kernel_sum(const float* p, const float3 wi, const float3 wo) {
    const float ci = fminf(1.0f, fmaxf(1e-6f, wi.z)), cr = fminf(1.0f, fmaxf(1e-6f, wo.z));
    const float si = sqrtf(fmaxf(0.0f, 1.0f - ci * ci)), sr = sqrtf(fmaxf(0.0f, 1.0f - cr * cr));
    float cphi = 1.0f, sphi = 0.0f;
    if (si > 1e-6f && sr > 1e-6f) {
        cphi = (wi.x * wo.x + wi.y * wo.y) / (si * sr);
        cphi = fminf(1.0f, fmaxf(-1.0f, cphi));
        sphi = sqrtf(fmaxf(0.0f, 1.0f - cphi * cphi));
    }
    const float cxi = fminf(1.0f, fmaxf(-1.0f, ci * cr + si * sr * cphi));
    const float xi = acosf(cxi);
    const float sxi = sinf(xi);
    const float kvol = ((0.5f * RT_PI - xi) * cxi + sxi) / (ci + cr) - 0.25f * RT_PI;
    const float ti = si / ci, tr = sr / cr;
    const float seci = 1.0f / ci, secr = 1.0f / cr;
    const float D2 = fmaxf(0.0f, ti * ti + tr * tr - 2.0f * ti * tr * cphi);
    const float q = ti * tr * sphi;
    float cost = 2.0f * sqrtf(D2 + q * q) / (seci + secr);
    cost = fminf(1.0f, fmaxf(-1.0f, cost));
    const float t = acosf(cost);
    const float O = (t - sinf(t) * cost) * (seci + secr) / RT_PI;
    const float kgeo = O - seci - secr + 0.5f * (1.0f + cxi) * seci * secr;
    const float v = p[0] + p[1] * kgeo + p[2] * kvol;
    return v > 0.0f ? v : 0.0f;
}
__device__ __noinline__ float brdf_eval(int type, float p0, float p1, float p2, float p3, float p4, const float3 wi, const float3 wo) {
    if (wi.z <= 0.0f || wo.z <= 0.0f) return 0.0f;
    const float p[5] = {p0, p1, p2, p3, p4};
    if (type == 2) return p[1] * p[0] * (1.0f / RT_PI) + (1.0f - p[1]) * dsm_spec_brdf(p, wi, wo);
    if (type == 4) return lsrt_kernel_sum(p, wi, wo) * (1.0f / RT_PI);
    return p[0] * (1.0f / RT_PI);
}
