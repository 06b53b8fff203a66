// oracle_mc.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// fp64 CPU restatement of the photon-transport path that er3t delegates to the external MCARaTS
// binary (er3t/rtm/mca/mca_run.py:110-113,179-181).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this file's shared object.
//
// PARITY UNPINNED against MCARaTS itself: its Fortran source is not in the reference tree and cannot
// be built offline, and the reference's tests hold no golden vectors for transport.
// What this oracle restates is the *published algorithm family* (Iwabuchi 2006, JAS 63:2324: forward
// Monte Carlo, local-estimate radiance, path-integrated gas absorption) under the parameter semantics
// the reference documents in er3t/rtm/mca/mca_inp.py:19-364 and actually sets in
// er3t/rtm/mca/mcarats.py:234-414.  It is pinned instead by tests/test_oracle_*.py against
//   (1) a deterministic adding-doubling plane-parallel solver (oracle/adding_doubling.py),
//   (2) analytic identities (direct beam, R+T+A=1, Lambertian no-atmosphere, single scattering).
//
// Design choice that makes it an INDEPENDENT check of the CUDA path: free paths are sampled by exact
// voxel-by-voxel traversal (no majorants / null collisions), everything is fp64, and the random
// stream is consumed in a different order.  Agreement with the GPU is therefore statistical.
//
// Parameter semantics (file:line of the reference that defines each input):
//   grid / 1-D profiles      er3t/rtm/mca/mca_atm.py:74-102,105-139
//   3-D fields, iz3l, nz3    er3t/rtm/mca/mca_atm.py:231-337  (x fastest, mca_atm.py:383-388)
//   apf encoding             mca_atm.py:101,262,276-277,301 ; er3t/rtm/mca/util.py:153
//   phase tables             er3t/rtm/mca/mca_sca.py:82-95 ; er3t/pre/pha/pha_hg.py:24-25
//   surface types/params     er3t/rtm/mca/mca_sfc.py:89-133 ; er3t/pre/sfc/sfc_gen.py:119-145
//   source / sensor angles   er3t/rtm/mca/mcarats.py:285-307,374-383,527-549
//   output order/layout      er3t/rtm/mca/mca_out.py:350-352 (direct-down, down, up), :473 (rad)
//   roulette                 Pho_wmin/wfac, er3t/rtm/mca/mca_inp.py:196-198
//
// build: g++ -O3 -march=x86-64-v3 -fopenmp -shared -fPIC -I../include oracle_mc.cpp -o _build/liboracle_mc.so

#include "b200rt.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double PI = 3.14159265358979323846;
constexpr double DEG = PI / 180.0;

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have = 0;
    static inline void round1(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = 0xD2511F53ull * c[0];
        const uint64_t p1 = 0xCD9E8D57ull * c[2];
        const uint32_t hi0 = uint32_t(p0 >> 32), lo0 = uint32_t(p0);
        const uint32_t hi1 = uint32_t(p1 >> 32), lo1 = uint32_t(p1);
        const uint32_t n0 = hi1 ^ c[1] ^ k[0];
        const uint32_t n1 = lo1;
        const uint32_t n2 = hi0 ^ c[3] ^ k[1];
        const uint32_t n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void block(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t o[4]) {
        uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
        uint32_t k[2] = {key_in[0], key_in[1]};
        for (int r = 0; r < 10; ++r) {
            round1(c, k);
            k[0] += 0x9E3779B9u;
            k[1] += 0xBB67AE85u;
        }
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[3];
    }
    void init(uint64_t seed, uint64_t photon, uint32_t stream) {
        key[0] = uint32_t(seed); key[1] = uint32_t(seed >> 32);
        ctr[0] = uint32_t(photon); ctr[1] = uint32_t(photon >> 32);
        ctr[2] = 0; ctr[3] = stream;
        have = 0;
    }
    inline uint32_t next_u32() {
        if (have == 0) { block(ctr, key, out); ctr[2]++; have = 4; }
        return out[4 - (have--)];
    }
    inline double uni() {  // (0,1)
        return (double(next_u32()) + 0.5) * (1.0 / 4294967296.0);
    }
};

// ---------------------------------------------------------------- phase functions
struct PhaseTable {           // piecewise-linear in mu, normalised to (1/2) int_{-1}^{1} P dmu = 1
    int n = 0;
    std::vector<double> mu;   // increasing
    std::vector<double> p;
    std::vector<double> cdf;  // cdf[0]=0 .. cdf[n-1]=1 (probability of mu' <= mu)
};

void build_table(const double* ang_deg, const double* pha, int nang, PhaseTable& t) {
    t.n = nang;
    t.mu.resize(nang); t.p.resize(nang); t.cdf.resize(nang);
    for (int i = 0; i < nang; ++i) {            // reverse: angle increasing -> mu decreasing
        int j = nang - 1 - i;
        double m = std::cos(ang_deg[j] * DEG);
        if (j == 0 && std::fabs(ang_deg[j]) < 1e-9) m = 1.0;
        if (j == nang - 1 && std::fabs(ang_deg[j] - 180.0) < 1e-9) m = -1.0;
        t.mu[i] = m;
        t.p[i] = pha[j] > 0.0 ? pha[j] : 0.0;
    }
    double area = 0.0;
    t.cdf[0] = 0.0;
    for (int i = 1; i < nang; ++i) {
        area += 0.5 * (t.p[i] + t.p[i - 1]) * (t.mu[i] - t.mu[i - 1]);
        t.cdf[i] = area;
    }
    const double norm = area > 0 ? area : 1.0;   // int P dmu; want 2
    for (int i = 0; i < nang; ++i) { t.p[i] *= 2.0 / norm; t.cdf[i] /= norm; }
    t.cdf[nang - 1] = 1.0;
}

inline double table_eval(const PhaseTable& t, double mu) {
    if (mu <= t.mu[0]) return t.p[0];
    if (mu >= t.mu[t.n - 1]) return t.p[t.n - 1];
    int lo = int(std::upper_bound(t.mu.begin(), t.mu.end(), mu) - t.mu.begin()) - 1;
    if (lo > t.n - 2) lo = t.n - 2;
    double f = (mu - t.mu[lo]) / (t.mu[lo + 1] - t.mu[lo]);
    return t.p[lo] + f * (t.p[lo + 1] - t.p[lo]);
}

inline double table_sample(const PhaseTable& t, double xi) {
    int lo = int(std::upper_bound(t.cdf.begin(), t.cdf.end(), xi) - t.cdf.begin()) - 1;
    if (lo < 0) lo = 0;
    if (lo > t.n - 2) lo = t.n - 2;
    const double dmu = t.mu[lo + 1] - t.mu[lo];
    const double c = 2.0 * (xi - t.cdf[lo]);            // area (in units where int P dmu = 2)
    const double p0 = t.p[lo], p1 = t.p[lo + 1];
    const double s = (p1 - p0) / dmu;
    const double disc = p0 * p0 + 2.0 * s * c;
    const double den = p0 + std::sqrt(disc > 0 ? disc : 0.0);
    double tt = den > 0 ? 2.0 * c / den : 0.0;
    if (tt < 0) tt = 0; if (tt > dmu) tt = dmu;
    return t.mu[lo] + tt;
}

inline double hg_eval(double g, double mu) {
    const double d = 1.0 + g * g - 2.0 * g * mu;
    return (1.0 - g * g) / (d * std::sqrt(d));
}
inline double hg_sample(double g, double xi) {
    if (std::fabs(g) < 1e-6) return 2.0 * xi - 1.0;
    const double s = (1.0 - g * g) / (1.0 - g + 2.0 * g * xi);
    double mu = (1.0 + g * g - s * s) / (2.0 * g);
    return std::min(1.0, std::max(-1.0, mu));
}
inline double ray_eval(double mu) { return 0.75 * (1.0 + mu * mu); }
inline double ray_sample(double xi) {
    const double u = 4.0 * xi - 2.0;
    const double q = std::cbrt(u + std::sqrt(u * u + 1.0));
    return std::min(1.0, std::max(-1.0, q - 1.0 / q));
}

struct Phase {
    std::vector<PhaseTable> tab;
    // apf decoding: <= -1 Rayleigh ; (-1,1) HG ; >= 1 real-valued 1-based index
    double eval(double apf, double mu) const {
        if (apf <= -1.0) return ray_eval(mu);
        if (apf < 1.0) return hg_eval(apf, mu);
        const int np = int(tab.size());
        if (np == 0) return 1.0;
        double a = std::min(double(np), std::max(1.0, apf));
        int i = int(std::floor(a));
        double f = a - i;
        if (i >= np) { i = np; f = 0; }
        double v = table_eval(tab[i - 1], mu);
        if (f > 0) v = (1.0 - f) * v + f * table_eval(tab[i], mu);
        return v;
    }
    double sample(double apf, double xi, double xi_tab) const {
        if (apf <= -1.0) return ray_sample(xi);
        if (apf < 1.0) return hg_sample(apf, xi);
        const int np = int(tab.size());
        if (np == 0) return 2.0 * xi - 1.0;
        double a = std::min(double(np), std::max(1.0, apf));
        int i = int(std::floor(a));
        double f = a - i;
        if (i >= np) { i = np; f = 0; }
        if (f > 0 && xi_tab < f) i += 1;
        return table_sample(tab[i - 1], xi);
    }
};

// ---------------------------------------------------------------- vectors
struct V3 { double x, y, z; };
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

inline V3 rotate_dir(const V3& d, double mu, double phi) {
    const double st = std::sqrt(std::max(0.0, 1.0 - mu * mu));
    const double cp = std::cos(phi), sp = std::sin(phi);
    V3 r;
    if (std::fabs(d.z) > 0.99999) {
        const double sg = d.z > 0 ? 1.0 : -1.0;
        r.x = st * cp; r.y = st * sp * sg; r.z = mu * sg;
    } else {
        const double den = std::sqrt(1.0 - d.z * d.z);
        r.x = st * (d.x * d.z * cp - d.y * sp) / den + d.x * mu;
        r.y = st * (d.y * d.z * cp + d.x * sp) / den + d.y * mu;
        r.z = -st * cp * den + d.z * mu;
    }
    const double n = std::sqrt(dot(r, r));
    r.x /= n; r.y /= n; r.z /= n;
    return r;
}

// ---------------------------------------------------------------- surface BRDF models
// DSM (type 2): param = (diffuse_alb, diffuse_frac, refrac_r, refrac_i, slope variance sigma^2)
//   er3t/pre/sfc/sfc_gen.py:137-141 ; er3t/pre/sfc/util.py:93-148 (Cox & Munk isotropic, Koepke whitecaps)
// LSRT (type 4): param = (f_iso, f_geo, f_vol), er3t/pre/sfc/sfc_gen.py:123-125
inline double fresnel_unpol(double cosg, double nr, double ni) {
    using cd = std::complex<double>;
    const cd m(nr, ni);
    const cd m2 = m * m;
    const cd g = std::sqrt(m2 - 1.0 + cosg * cosg);
    const cd rs = (cosg - g) / (cosg + g);
    const cd rp = (m2 * cosg - g) / (m2 * cosg + g);
    return 0.5 * (std::norm(rs) + std::norm(rp));
}
inline double cm_lambda(double mu, double sig2) {
    if (mu >= 1.0) return 0.0;
    const double cot = mu / std::sqrt(std::max(1e-300, 1.0 - mu * mu));
    const double nu = cot / std::sqrt(sig2);
    return 0.5 * (std::exp(-nu * nu) / (std::sqrt(PI) * nu) - std::erfc(nu));
}
inline double cm_shadow(double mui, double mur, double sig2) {
    return 1.0 / (1.0 + cm_lambda(mui, sig2) + cm_lambda(mur, sig2));
}
// specular (glint) part of DSM, wi and wo both pointing AWAY from the surface (z > 0)
inline double dsm_spec_brdf(const float* p, const V3& wi, const V3& wo) {
    const double sig2 = std::max(1e-6, double(p[4]));
    V3 h{wi.x + wo.x, wi.y + wo.y, wi.z + wo.z};
    const double hn = std::sqrt(dot(h, h));
    if (hn <= 0 || h.z <= 0) return 0.0;
    h.x /= hn; h.y /= hn; h.z /= hn;
    const double cosg = dot(wi, h);
    if (cosg <= 0) return 0.0;
    const double cn = h.z;
    const double tan2 = (1.0 - cn * cn) / (cn * cn);
    const double P = std::exp(-tan2 / sig2) / (PI * sig2);
    const double F = fresnel_unpol(cosg, p[2], p[3]);
    return F * P * cm_shadow(wi.z, wo.z, sig2) / (4.0 * wi.z * wo.z * cn * cn * cn * cn);
}
inline double lsrt_kernel_sum(const float* p, const V3& wi, const V3& wo) {
    const double ci = std::min(1.0, std::max(1e-6, wi.z)), cr = std::min(1.0, std::max(1e-6, wo.z));
    const double si = std::sqrt(std::max(0.0, 1.0 - ci * ci)), sr = std::sqrt(std::max(0.0, 1.0 - cr * cr));
    double cphi = 1.0, sphi = 0.0;
    if (si > 1e-12 && sr > 1e-12) {
        cphi = (wi.x * wo.x + wi.y * wo.y) / (si * sr);
        cphi = std::min(1.0, std::max(-1.0, cphi));
        sphi = std::sqrt(std::max(0.0, 1.0 - cphi * cphi));
    }
    const double cxi = std::min(1.0, std::max(-1.0, ci * cr + si * sr * cphi));
    const double xi = std::acos(cxi);
    const double sxi = std::sin(xi);
    const double kvol = ((PI / 2 - xi) * cxi + sxi) / (ci + cr) - PI / 4;
    const double ti = si / ci, tr = sr / cr;                    // b/r = 1
    const double seci = 1.0 / ci, secr = 1.0 / cr;
    const double D2 = std::max(0.0, ti * ti + tr * tr - 2.0 * ti * tr * cphi);
    double cost = 2.0 * std::sqrt(D2 + (ti * tr * sphi) * (ti * tr * sphi)) / (seci + secr);  // h/b = 2
    cost = std::min(1.0, std::max(-1.0, cost));
    const double t = std::acos(cost);
    const double O = (t - std::sin(t) * cost) * (seci + secr) / PI;
    const double kgeo = O - seci - secr + 0.5 * (1.0 + cxi) * seci * secr;
    const double v = p[0] + p[1] * kgeo + p[2] * kvol;
    return v > 0 ? v : 0.0;
}
// full BRDF [1/sr]
inline double brdf_eval(int type, const float* p, const V3& wi, const V3& wo) {
    if (wi.z <= 0 || wo.z <= 0) return 0.0;
    switch (type) {
        case B200RT_SFC_LAMBERT: return p[0] / PI;
        case B200RT_SFC_DSM: return p[1] * p[0] / PI + (1.0 - p[1]) * dsm_spec_brdf(p, wi, wo);
        case B200RT_SFC_LSRT: return lsrt_kernel_sum(p, wi, wo) / PI;
        default: return p[0] / PI;
    }
}

// ---------------------------------------------------------------- scene
struct Sensor {
    V3 s; double zloc, zref; int nxr, nyr;
    // all-sky camera (Rad_mrkind = 1): `s` = viewing axis, (ex, ey) complete the camera frame (Z-Y-Z rotation by phi, the, psi)
    int kind = 2;
    V3 cpos{0, 0, 0}, ex{1, 0, 0}, ey{0, 1, 0};
    double cos_half_fov = -1, u_half = PI / 2, v_half = PI / 2, ap2 = 0;
};

struct Scene {
    int nx, ny, nz, iz0 /*0-based first 3-D layer*/, nz3, np1d, np3d;
    double dx, dy, Lx, Ly;
    std::vector<double> z;                      // nz+1
    std::vector<double> e1, o1, a1;             // [np1d][nz]
    std::vector<double> e1tot;                  // [nz]
    const float *e3, *o3, *a3, *abs3;           // caller-owned (host)
    Phase phase;
    int sfc_nx, sfc_ny;
    const int32_t* sfc_type; const float* sfc_param;
    V3 src; double src_cos_half; double mu0;
    std::vector<Sensor> sens;
    int solver, target;
    double wmin, wfac; int iso_ss, iso_max;
    inline bool in3d(int iz) const { return nz3 > 0 && iz >= iz0 && iz < iz0 + nz3; }
    int layout = 0;   // b200rt_scene.layout3d
    inline size_t vox(int iz, int iy, int ix) const { return (size_t(iz - iz0) * ny + iy) * nx + ix; }
    // index of component k of voxel (iz, iy, ix) in the caller's 3-D arrays
    inline size_t idx3(int k, int iz, int iy, int ix) const {
        if (layout == 0) return size_t(k) * nz3 * ny * nx + vox(iz, iy, ix);
        return ((size_t(ix) * ny + iy) * nz3 + (iz - iz0)) * np3d_in + k;
    }
    int np3d_in = 0;
    inline double ext3tot(int iz, int iy, int ix) const {
        double s = 0;
        for (int k = 0; k < np3d; ++k) s += double(e3[idx3(k, iz, iy, ix)]);
        return s;
    }
};

inline V3 dir_from_angles(double the_deg, double phi_deg) {
    const double t = the_deg * DEG, p = phi_deg * DEG;
    V3 d{std::sin(t) * std::cos(p), std::sin(t) * std::sin(p), std::cos(t)};
    if (std::fabs(d.x) < 1e-15) d.x = 0; if (std::fabs(d.y) < 1e-15) d.y = 0;
    return d;
}

struct Tally {
    double* flux; double* rad; double* heat;   // may be null
    bool atomic;
    size_t nlev_stride, var_stride, slab_stride_flux;
    inline void add(double* p, double v) const {
        if (atomic) {
#pragma omp atomic
            *p += v;
        } else *p += v;
    }
};

struct Counters {
    uint64_t photons = 0, n_cell = 0, n_coll = 0, n_sfc = 0, n_le = 0, n_le_visit = 0, n_tally = 0, n_kill = 0;
    double w_toa = 0, w_sfc = 0, w_atm = 0, w_rr = 0;
};

struct Photon {
    double x, y, z; V3 d; double w; int ix, iy, iz; int order; bool direct; bool frozen;
};

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline double wrap(double x, double L) { x = std::fmod(x, L); if (x < 0) x += L; if (x >= L) x = 0; return x; }

// Optical depth (extinction + absorption) from point (x,y,z in layer iz, column ix,iy) along unit
// vector s until z reaches ztarget.  Exact voxel traversal; IPA/frozen: the column never changes.
double tau_to_level(const Scene& S, const double* abs1d, double x, double y, double z, int ix, int iy, int iz,
                    const V3& s, double ztarget, bool frozen, uint64_t& visits) {
    double tau = 0.0;
    const bool up = s.z > 0;
    while (true) {
        const double zb = up ? std::min(S.z[iz + 1], ztarget) : std::max(S.z[iz], ztarget);
        double dz = (zb - z) / s.z;
        if (dz < 0) dz = 0;
        const double sa = abs1d ? abs1d[iz] : 0.0;
        if (!S.in3d(iz) || frozen) {
            double se = S.e1tot[iz];
            double sa3 = 0;
            if (S.in3d(iz)) { se += S.ext3tot(iz, iy, ix); if (S.abs3) sa3 = S.abs3[S.vox(iz, iy, ix)]; }
            tau += (se + sa + sa3) * dz;
            ++visits;
            if (!frozen) { x = wrap(x + s.x * dz, S.Lx); y = wrap(y + s.y * dz, S.Ly); }
        } else {
            // walk columns inside this layer
            double rem = dz;
            while (true) {
                double tx = 1e300, ty = 1e300;
                if (s.x > 0) tx = ((ix + 1) * S.dx - x) / s.x; else if (s.x < 0) tx = (ix * S.dx - x) / s.x;
                if (s.y > 0) ty = ((iy + 1) * S.dy - y) / s.y; else if (s.y < 0) ty = (iy * S.dy - y) / s.y;
                if (tx < 0) tx = 0; if (ty < 0) ty = 0;
                double step = std::min(rem, std::min(tx, ty));
                double sig = S.e1tot[iz] + S.ext3tot(iz, iy, ix) + sa + (S.abs3 ? S.abs3[S.vox(iz, iy, ix)] : 0.0);
                tau += sig * step;
                ++visits;
                if (step >= rem) { x += s.x * rem; y += s.y * rem; break; }
                rem -= step;
                x += s.x * step; y += s.y * step;
                if (tx <= ty) {
                    if (s.x > 0) { ix++; x = ix * S.dx; if (ix == S.nx) { ix = 0; x = 0; } }
                    else { x = ix * S.dx; ix--; if (ix < 0) { ix = S.nx - 1; x = S.Lx; } }
                } else {
                    if (s.y > 0) { iy++; y = iy * S.dy; if (iy == S.ny) { iy = 0; y = 0; } }
                    else { y = iy * S.dy; iy--; if (iy < 0) { iy = S.ny - 1; y = S.Ly; } }
                }
            }
        }
        z = zb;
        if (up) { if (zb >= ztarget || iz + 1 >= S.nz) break; iz++; }
        else { if (zb <= ztarget || iz == 0) break; iz--; }
        if (S.in3d(iz) && !frozen) {   // entering the 3-D block (or moving inside it): refresh column
            ix = clampi(int(std::floor(x / S.dx)), 0, S.nx - 1);
            iy = clampi(int(std::floor(y / S.dy)), 0, S.ny - 1);
            // keep face ownership consistent with direction of travel
            if (s.x < 0 && x == ix * S.dx && x > 0) ix--;
            if (s.y < 0 && y == iy * S.dy && y > 0) iy--;
        }
    }
    return tau;
}

struct JobCtx {
    const double* abs1d; const double* fscale; double rscale; int slab; double norm;
};

// local estimate toward every sensor.  `pfun(mu)`: angular density [1/sr] of the event toward direction s
template <class F>
void local_estimate(const Scene& S, const JobCtx& J, const Tally& T, const Photon& p, double wgt, F&& dens,
                    Counters& C) {
    for (size_t k = 0; k < S.sens.size(); ++k) {
        const Sensor& se = S.sens[k];
        if (se.kind == 1) {
            // ---- all-sky camera: I_pix += w f exp(-tau) / (R^2 dOmega_pix) x (power per photon); nearest periodic image;
            //      polar mapping U = theta cos(az), V = theta sin(az); dOmega = (sin(theta) / theta) dU dV
            if (p.frozen) continue;
            double dx = se.cpos.x - p.x, dy = se.cpos.y - p.y;
            dx -= S.Lx * std::nearbyint(dx / S.Lx); dy -= S.Ly * std::nearbyint(dy / S.Ly);
            const double dz = se.cpos.z - p.z;
            const double R2 = dx * dx + dy * dy + dz * dz;
            if (!(R2 > 1e-6) || std::fabs(dz) < 1e-3) continue;
            const double R = std::sqrt(R2);
            const V3 sd{dx / R, dy / R, dz / R};
            if (std::fabs(sd.z) < 1e-4) continue;
            const V3 v{-sd.x, -sd.y, -sd.z};
            const double cq = dot(v, se.s);
            if (cq < se.cos_half_fov) continue;
            const double theta = std::acos(std::min(1.0, cq));
            const double vx = dot(v, se.ex), vy = dot(v, se.ey), rho = std::sqrt(vx * vx + vy * vy);
            const double U = rho > 0 ? theta * vx / rho : 0.0, V = rho > 0 ? theta * vy / rho : 0.0;
            if (std::fabs(U) >= se.u_half || std::fabs(V) >= se.v_half) continue;
            const int px = clampi(int((U + se.u_half) / (2 * se.u_half) * se.nxr), 0, se.nxr - 1);
            const int py = clampi(int((V + se.v_half) / (2 * se.v_half) * se.nyr), 0, se.nyr - 1);
            const double f = dens(sd);
            if (f <= 0) continue;
            int iz = p.iz;
            if (sd.z > 0 && p.z >= S.z[iz + 1] && iz + 1 < S.nz) iz++;
            if (sd.z < 0 && p.z <= S.z[iz] && iz > 0) iz--;
            int ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1), iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1);
            if (S.in3d(p.iz)) { ix = p.ix; iy = p.iy; }
            const double tau = tau_to_level(S, J.abs1d, p.x, p.y, p.z, ix, iy, iz, sd, se.cpos.z, false, C.n_le_visit);
            ++C.n_le;
            const double sinc = theta > 1e-6 ? std::sin(theta) / theta : 1.0;
            const double domega = sinc * (2 * se.u_half / se.nxr) * (2 * se.v_half / se.nyr);
            const double contrib = wgt * f * std::exp(-tau) / (std::max(R2, se.ap2) * domega);
            size_t off = 0;
            for (size_t q = 0; q < k; ++q) off += size_t(S.sens[q].nxr) * S.sens[q].nyr;
            size_t slabsz = 0;
            for (auto& q : S.sens) slabsz += size_t(q.nxr) * q.nyr;
            T.add(&T.rad[size_t(J.slab) * slabsz + off + size_t(py) * se.nxr + px], contrib * J.rscale * J.norm * S.Lx * S.Ly);
            ++C.n_tally;
            continue;
        }
        const double ztop = S.z[S.nz], zbot = S.z[0];
        const double zt = std::min(ztop, std::max(zbot, se.zloc));
        const double dzs = (zt - p.z) / se.s.z;
        if (!(dzs > 0)) { if (!(dzs == 0 && p.z == zt)) continue; }
        const double f = dens(se.s);
        if (f <= 0) continue;
        int iz = p.iz;
        // a point exactly on a level belongs to the layer the ray enters
        if (se.s.z > 0 && p.z >= S.z[iz + 1] && iz + 1 < S.nz) iz++;
        if (se.s.z < 0 && p.z <= S.z[iz] && iz > 0) iz--;
        int ix = p.ix, iy = p.iy;
        if (!p.frozen) {
            ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1);
            iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1);
            if (S.in3d(p.iz)) { ix = p.ix; iy = p.iy; }
        }
        const double tau = tau_to_level(S, J.abs1d, p.x, p.y, p.z, ix, iy, iz, se.s, zt, p.frozen, C.n_le_visit);
        ++C.n_le;
        const double contrib = wgt * f * std::exp(-tau) / std::fabs(se.s.z);
        // pixel registration: line of sight meets z = zref
        double xr = p.x, yr = p.y;
        if (!p.frozen) {
            const double t = (se.zref - p.z) / se.s.z;
            xr = wrap(p.x + se.s.x * t, S.Lx); yr = wrap(p.y + se.s.y * t, S.Ly);
        }
        int px = clampi(int(std::floor(xr / S.Lx * se.nxr)), 0, se.nxr - 1);
        int py = clampi(int(std::floor(yr / S.Ly * se.nyr)), 0, se.nyr - 1);
        if (p.frozen) {   // IPA: the entry column's pixel
            px = clampi(int((double(p.ix) + 0.5) / S.nx * se.nxr), 0, se.nxr - 1);
            py = clampi(int((double(p.iy) + 0.5) / S.ny * se.nyr), 0, se.nyr - 1);
        }
        size_t off = 0;
        for (size_t q = 0; q < k; ++q) off += size_t(S.sens[q].nxr) * S.sens[q].nyr;
        size_t slabsz = 0;
        for (auto& q : S.sens) slabsz += size_t(q.nxr) * q.nyr;
        T.add(&T.rad[size_t(J.slab) * slabsz + off + size_t(py) * se.nxr + px],
              contrib * J.rscale * J.norm * double(se.nxr) * se.nyr);
        ++C.n_tally;
    }
}

inline void flux_tally(const Scene& S, const JobCtx& J, const Tally& T, const Photon& p, int var, int lev, double w,
                       Counters& C) {
    if (!T.flux) return;
    int ix, iy;
    if (p.frozen || S.in3d(p.iz)) { ix = p.ix; iy = p.iy; }
    else { ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1); iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1); }
    const size_t nxy = size_t(S.nx) * S.ny, nlev = size_t(S.nz) + 1;
    const double sc = (J.fscale ? J.fscale[lev] : 1.0) * J.norm * double(nxy);
    T.add(&T.flux[((size_t(J.slab) * 3 + var) * nlev + lev) * nxy + size_t(iy) * S.nx + ix], w * sc);
    ++C.n_tally;
}

void trace_photon(const Scene& S, const JobCtx& J, const Tally& T, Philox& R, Counters& C) {
    Photon p;
    p.x = R.uni() * S.Lx; p.y = R.uni() * S.Ly; p.z = S.z[S.nz];
    {   // direction uniform in the solar cone (Src_qmax = FULL cone angle, mcarats.py:378)
        const double c = 1.0 - R.uni() * (1.0 - S.src_cos_half);
        const double ph = 2.0 * PI * R.uni();
        p.d = (S.src_cos_half < 1.0) ? rotate_dir(S.src, c, ph) : S.src;
    }
    p.w = 1.0; p.order = 0; p.direct = true; p.iz = S.nz - 1;
    p.ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1);
    p.iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1);
    p.frozen = (S.solver == B200RT_SOLVER_IPA);
    ++C.photons;
    flux_tally(S, J, T, p, 0, S.nz, p.w, C);
    flux_tally(S, J, T, p, 1, S.nz, p.w, C);
    const size_t nxy = size_t(S.nx) * S.ny;

    while (true) {
        double tau = -std::log(R.uni());
        bool collided = false;
        // ---- fly until collision, surface or TOA
        while (true) {
            const int iz = p.iz;
            const bool is3 = S.in3d(iz);
            double sig = S.e1tot[iz], sa = J.abs1d ? J.abs1d[iz] : 0.0;
            if (is3) { sig += S.ext3tot(iz, p.iy, p.ix); if (S.abs3) sa += S.abs3[S.vox(iz, p.iy, p.ix)]; }
            ++C.n_cell;
            double tz = 1e300, tx = 1e300, ty = 1e300;
            if (p.d.z > 0) tz = (S.z[iz + 1] - p.z) / p.d.z; else if (p.d.z < 0) tz = (S.z[iz] - p.z) / p.d.z;
            if (is3 && !p.frozen) {
                if (p.d.x > 0) tx = ((p.ix + 1) * S.dx - p.x) / p.d.x; else if (p.d.x < 0) tx = (p.ix * S.dx - p.x) / p.d.x;
                if (p.d.y > 0) ty = ((p.iy + 1) * S.dy - p.y) / p.d.y; else if (p.d.y < 0) ty = (p.iy * S.dy - p.y) / p.d.y;
            }
            if (tz < 0) tz = 0; if (tx < 0) tx = 0; if (ty < 0) ty = 0;
            const double dexit = std::min(tz, std::min(tx, ty));
            const double dcol = sig > 0 ? tau / sig : 1e300;
            const double dmove = std::min(dexit, dcol);
            // path-integrated absorption
            if (sa > 0) {
                const double wn = p.w * std::exp(-sa * dmove);
                const double dep = p.w - wn;
                C.w_atm += dep;
                if (T.heat) {
                    int hx = p.ix, hy = p.iy;
                    if (!is3 && !p.frozen) { hx = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1); hy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1); }
                    T.add(&T.heat[(size_t(J.slab) * S.nz + iz) * nxy + size_t(hy) * S.nx + hx], dep * J.norm * double(nxy) * (J.fscale ? J.fscale[iz] : 1.0));
                }
                p.w = wn;
            }
            if (dcol <= dexit) {
                p.x += p.d.x * dcol; p.y += p.d.y * dcol; p.z += p.d.z * dcol;
                if (!is3 && !p.frozen) { p.x = wrap(p.x, S.Lx); p.y = wrap(p.y, S.Ly); }
                if (p.frozen) { p.x = wrap(p.x, S.Lx); p.y = wrap(p.y, S.Ly); }
                collided = true;
                break;
            }
            tau -= sig * dexit;
            p.x += p.d.x * dexit; p.y += p.d.y * dexit; p.z += p.d.z * dexit;
            if (!is3 || p.frozen) { p.x = wrap(p.x, S.Lx); p.y = wrap(p.y, S.Ly); }
            if (tz <= tx && tz <= ty) {
                if (p.d.z > 0) {
                    p.z = S.z[iz + 1];
                    flux_tally(S, J, T, p, 2, iz + 1, p.w, C);
                    if (iz + 1 >= S.nz) { C.w_toa += p.w; return; }
                    p.iz = iz + 1;
                } else {
                    p.z = S.z[iz];
                    if (p.direct) flux_tally(S, J, T, p, 0, iz, p.w, C);
                    flux_tally(S, J, T, p, 1, iz, p.w, C);
                    if (iz == 0) break;   // surface
                    p.iz = iz - 1;
                }
                if (S.in3d(p.iz) && !is3 && !p.frozen) {   // entering the 3-D block
                    p.ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1);
                    p.iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1);
                }
            } else if (tx <= ty) {
                if (p.d.x > 0) { p.ix++; p.x = p.ix * S.dx; if (p.ix == S.nx) { p.ix = 0; p.x = 0; } }
                else { p.x = p.ix * S.dx; p.ix--; if (p.ix < 0) { p.ix = S.nx - 1; p.x = S.Lx; } }
            } else {
                if (p.d.y > 0) { p.iy++; p.y = p.iy * S.dy; if (p.iy == S.ny) { p.iy = 0; p.y = 0; } }
                else { p.y = p.iy * S.dy; p.iy--; if (p.iy < 0) { p.iy = S.ny - 1; p.y = S.Ly; } }
            }
        }

        if (collided) {
            const int iz = p.iz;
            const bool is3 = S.in3d(iz);
            // choose component with probability ext_k / sum ext
            double sig = S.e1tot[iz];
            if (is3) sig += S.ext3tot(iz, p.iy, p.ix);
            double u = R.uni() * sig;
            double omg = 1.0, apf = 0.0;
            bool found = false;
            if (is3) {
                for (int k = 0; k < S.np3d && !found; ++k) {
                    const size_t q = S.idx3(k, iz, p.iy, p.ix);
                    const double e = S.e3[q];
                    if (u < e) { omg = S.o3[q]; apf = S.a3[q]; found = true; }
                    else u -= e;
                }
            }
            for (int k = 0; k < S.np1d && !found; ++k) {
                const double e = S.e1[size_t(k) * S.nz + iz];
                if (u < e || k == S.np1d - 1) { omg = S.o1[size_t(k) * S.nz + iz]; apf = S.a1[size_t(k) * S.nz + iz]; found = true; }
                else u -= e;
            }
            ++C.n_coll;
            omg = std::min(1.0, std::max(0.0, omg));
            const double wn = p.w * omg;
            C.w_atm += p.w - wn;
            if (T.heat && p.w > wn) {
                int hx = p.ix, hy = p.iy;
                if (!is3 && !p.frozen) { hx = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1); hy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1); }
                T.add(&T.heat[(size_t(J.slab) * S.nz + iz) * nxy + size_t(hy) * S.nx + hx], (p.w - wn) * J.norm * double(nxy) * (J.fscale ? J.fscale[iz] : 1.0));
            }
            p.w = wn;
            p.order++; p.direct = false;
            if (p.w <= 0) return;
            if (S.solver == B200RT_SOLVER_PARTIAL_3D && p.order >= S.iso_ss && !p.frozen) {
                if (!is3) { p.ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1); p.iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1); }
                p.frozen = true;
            }
            if (T.rad) {
                const V3 din = p.d;
                local_estimate(S, J, T, p, p.w, [&](const V3& s) { return S.phase.eval(apf, dot(din, s)) / (4.0 * PI); }, C);
            }
            const double xi1 = R.uni(), xi2 = R.uni(), xi3 = R.uni();
            const double mu = S.phase.sample(apf, xi1, xi3);
            p.d = rotate_dir(p.d, mu, 2.0 * PI * xi2);
            if (p.order >= S.iso_max) { C.w_rr -= p.w; return; }
        } else {
            // ---- surface at z = z[0]
            ++C.n_sfc;
            int sx = clampi(int(std::floor(p.x / S.Lx * S.sfc_nx)), 0, S.sfc_nx - 1);
            int sy = clampi(int(std::floor(p.y / S.Ly * S.sfc_ny)), 0, S.sfc_ny - 1);
            if (p.frozen) {
                sx = clampi(int((double(p.ix) + 0.5) / S.nx * S.sfc_nx), 0, S.sfc_nx - 1);
                sy = clampi(int((double(p.iy) + 0.5) / S.ny * S.sfc_ny), 0, S.sfc_ny - 1);
            }
            const size_t sn = size_t(S.sfc_nx) * S.sfc_ny, si = size_t(sy) * S.sfc_nx + sx;
            const int type = S.sfc_type[si];
            float prm[5];
            for (int q = 0; q < 5; ++q) prm[q] = S.sfc_param[q * sn + si];
            const V3 wi{-p.d.x, -p.d.y, -p.d.z};
            const double win = p.w;
            if (T.rad) {
                local_estimate(S, J, T, p, win, [&](const V3& s) { return s.z > 0 ? brdf_eval(type, prm, wi, s) * s.z : 0.0; }, C);
            }
            // sample reflected direction
            V3 wo; double fac = 0.0;
            const double xi1 = R.uni(), xi2 = R.uni(), xi3 = R.uni();
            bool diffuse = true;
            if (type == B200RT_SFC_DSM && xi3 >= prm[1]) diffuse = false;
            if (diffuse) {
                const double ct = std::sqrt(xi1), st = std::sqrt(1.0 - xi1), ph = 2.0 * PI * xi2;
                wo = V3{st * std::cos(ph), st * std::sin(ph), ct};
                if (wo.z < 1e-9) wo.z = 1e-9;
                if (type == B200RT_SFC_LSRT) fac = lsrt_kernel_sum(prm, wi, wo);
                else fac = prm[0];      // Lambertian albedo (type 1) or whitecap albedo (DSM)
            } else {
                // facet slope from the isotropic Gaussian, mirror reflection
                const double sig2 = std::max(1e-6, double(prm[4]));
                const double r = std::sqrt(-sig2 * std::log(1.0 - xi1 * (1.0 - 1e-16)));
                const double ph = 2.0 * PI * xi2;
                const double zx = r * std::cos(ph), zy = r * std::sin(ph);
                const double nn = 1.0 / std::sqrt(1.0 + zx * zx + zy * zy);
                const V3 n{-zx * nn, -zy * nn, nn};
                const double cosg = dot(wi, n);
                if (cosg > 0) {
                    wo = V3{2.0 * cosg * n.x - wi.x, 2.0 * cosg * n.y - wi.y, 2.0 * cosg * n.z - wi.z};
                    if (wo.z > 0) fac = fresnel_unpol(cosg, prm[2], prm[3]) * cosg / (wi.z * n.z) * cm_shadow(wi.z, wo.z, sig2);
                }
            }
            const double wn = win * fac;
            C.w_sfc += win - wn;
            p.w = wn;
            if (!(p.w > 0)) return;
            p.d = wo; p.direct = false; p.order++;
            p.iz = 0; p.z = S.z[0];
            if (S.in3d(0) && !p.frozen) { /* column unchanged */ }
            flux_tally(S, J, T, p, 2, 0, p.w, C);
            if (S.solver == B200RT_SOLVER_PARTIAL_3D && p.order >= S.iso_ss && !p.frozen) {
                if (!S.in3d(0)) { p.ix = clampi(int(std::floor(p.x / S.dx)), 0, S.nx - 1); p.iy = clampi(int(std::floor(p.y / S.dy)), 0, S.ny - 1); }
                p.frozen = true;
            }
        }
        // ---- Russian roulette (Pho_wmin / Pho_wfac)
        if (S.wmin > 0 && p.w < S.wmin) {
            const double wt = S.wfac > 0 ? S.wfac : 1.0;
            if (R.uni() * wt < p.w) { C.w_rr += wt - p.w; p.w = wt; }
            else { C.w_rr -= p.w; ++C.n_kill; return; }
        }
        if (p.w < 1e-30) { C.w_rr -= p.w; return; }
    }
}

}  // namespace

extern "C" {

int oracle_version(void) { return B200RT_VERSION; }

void oracle_philox_fill(uint64_t seed, uint64_t first, uint32_t c2, uint32_t c3, uint32_t* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t idx = first + uint64_t(i);
        uint32_t ctr[4] = {uint32_t(idx), uint32_t(idx >> 32), c2, c3};
        uint32_t key[2] = {uint32_t(seed), uint32_t(seed >> 32)};
        Philox::block(ctr, key, out + 4 * i);
    }
}

static int build_phase(const b200rt_scene* sc, Phase& ph) {
    ph.tab.resize(sc->npf);
    for (int i = 0; i < sc->npf; ++i) build_table(sc->ang, sc->pha + size_t(i) * sc->nang, sc->nang, ph.tab[i]);
    return 0;
}

void oracle_phase_eval(const b200rt_scene* sc, double apf, const double* mu, double* out, int64_t n) {
    Phase ph; build_phase(sc, ph);
    for (int64_t i = 0; i < n; ++i) out[i] = ph.eval(apf, mu[i]);
}
void oracle_phase_sample(const b200rt_scene* sc, double apf, const double* xi, double* out, int64_t n) {
    Phase ph; build_phase(sc, ph);
    for (int64_t i = 0; i < n; ++i) out[i] = ph.sample(apf, xi[i], 0.999999);
}
void oracle_brdf_eval(int32_t type, const float* p5, const double* din, const double* dout, double* f, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        V3 wi{-din[3 * i], -din[3 * i + 1], -din[3 * i + 2]};
        V3 wo{dout[3 * i], dout[3 * i + 1], dout[3 * i + 2]};
        f[i] = brdf_eval(type, p5, wi, wo);
    }
}

// All pointers are HOST pointers.  flux/rad/heat may be NULL when the target does not include them;
// they are accumulated into (caller zeroes them).
int oracle_run(const b200rt_scene* sc, const b200rt_options* opt, const b200rt_job* jobs, int njob,
               double* flux, double* rad, double* heat, b200rt_stats* stats, int nthreads) {
    Scene S;
    S.nx = sc->nx; S.ny = sc->ny; S.nz = sc->nz; S.nz3 = sc->nz3; S.iz0 = sc->iz3l - 1;
    S.np1d = sc->np1d; S.np3d = sc->nz3 > 0 ? sc->np3d : 0;
    S.dx = sc->dx; S.dy = sc->dy; S.Lx = sc->nx * sc->dx; S.Ly = sc->ny * sc->dy;
    if (S.nz < 1 || S.nx < 1 || S.ny < 1 || S.np1d < 1) return B200RT_ERR_ARG;
    if (S.nz3 > 0 && (S.iz0 < 0 || S.iz0 + S.nz3 > S.nz)) return B200RT_ERR_ARG;
    S.z.assign(sc->zgrd, sc->zgrd + S.nz + 1);
    S.e1.assign(sc->ext1d, sc->ext1d + size_t(S.np1d) * S.nz);
    S.o1.assign(sc->omg1d, sc->omg1d + size_t(S.np1d) * S.nz);
    S.a1.assign(sc->apf1d, sc->apf1d + size_t(S.np1d) * S.nz);
    S.e1tot.assign(S.nz, 0.0);
    for (int k = 0; k < S.np1d; ++k) for (int i = 0; i < S.nz; ++i) S.e1tot[i] += S.e1[size_t(k) * S.nz + i];
    S.e3 = sc->ext3d; S.o3 = sc->omg3d; S.a3 = sc->apf3d; S.abs3 = sc->abs3d;
    S.layout = sc->layout3d; S.np3d_in = sc->np3d;
    build_phase(sc, S.phase);
    S.sfc_nx = sc->sfc_nx; S.sfc_ny = sc->sfc_ny; S.sfc_type = sc->sfc_type; S.sfc_param = sc->sfc_param;
    S.src = dir_from_angles(sc->src_the, sc->src_phi);
    if (!(S.src.z < 0)) return B200RT_ERR_ARG;
    S.mu0 = -S.src.z;
    S.src_cos_half = std::cos(0.5 * sc->src_qmax * DEG);
    S.solver = opt->solver; S.target = opt->target;
    S.wmin = opt->wmin; S.wfac = opt->wfac;
    S.iso_ss = opt->iso_ss > 0 ? opt->iso_ss : 1;
    S.iso_max = opt->iso_max > 0 ? opt->iso_max : 1000000;
    for (int k = 0; k < sc->nrad; ++k) {
        const b200rt_sensor& q = sc->sensors[k];
        const V3 view = dir_from_angles(q.the, q.phi);
        Sensor se; se.s = V3{-view.x, -view.y, -view.z};
        if (se.s.x == 0) se.s.x = 0; if (se.s.y == 0) se.s.y = 0;   // no negative zeros
        se.kind = q.kind;
        if (q.kind == 1) {
            const double t = q.the * DEG, f = q.phi * DEG, ps = q.psi * DEG;
            const double ct = std::cos(t), st = std::sin(t), cf = std::cos(f), sf = std::sin(f), cp = std::cos(ps), sp = std::sin(ps);
            se.s = view;
            se.ex = V3{cp * ct * cf - sp * sf, cp * ct * sf + sp * cf, -cp * st};
            se.ey = V3{-sp * ct * cf - cp * sf, -sp * ct * sf + cp * cf, sp * st};
            se.cpos = V3{q.xpos * S.Lx, q.ypos * S.Ly, std::min(S.z[S.nz], std::max(S.z[0], q.zloc))};
            se.cos_half_fov = std::cos(0.5 * std::min(q.qmax, 360.0) * DEG);
            se.u_half = 0.5 * q.umax * DEG; se.v_half = 0.5 * q.vmax * DEG; se.ap2 = q.apsize * q.apsize;
        } else if (std::fabs(se.s.z) < 1e-6) return B200RT_ERR_ARG;
        se.zloc = q.zloc; se.zref = q.zref; se.nxr = q.nxr; se.nyr = q.nyr;
        S.sens.push_back(se);
    }
    const bool want_flux = (opt->target & B200RT_TARGET_FLUX) && flux;
    const bool want_rad = (opt->target & B200RT_TARGET_RADIANCE) && rad && !S.sens.empty();
    const bool want_heat = (opt->target & B200RT_TARGET_HEATING) && heat;

    const size_t nxy = size_t(S.nx) * S.ny;
    const size_t nflux = want_flux ? size_t(opt->nslab) * 3 * (S.nz + 1) * nxy : 0;
    size_t radslab = 0; for (auto& q : S.sens) radslab += size_t(q.nxr) * q.nyr;
    const size_t nradn = want_rad ? size_t(opt->nslab) * radslab : 0;
    const size_t nheat = want_heat ? size_t(opt->nslab) * S.nz * nxy : 0;
    const size_t ntot = nflux + nradn + nheat;
    const bool priv = ntot <= (size_t(1) << 16);   // small tallies: per-thread copies, no atomics

#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    const int nth = omp_get_max_threads();
#else
    const int nth = 1;
#endif
    std::vector<Counters> cnt(nth);
    std::vector<std::vector<double>> pf(priv ? nth : 0), pr(priv ? nth : 0), phh(priv ? nth : 0);
    const int world = opt->shard_world > 0 ? opt->shard_world : 1;
    const int rank = opt->shard_rank;

    for (int j = 0; j < njob; ++j) {
        const b200rt_job& jb = jobs[j];
        if (jb.slab < 0 || jb.slab >= opt->nslab) return B200RT_ERR_ARG;
        JobCtx J; J.abs1d = jb.abs1d; J.fscale = jb.flx_scale; J.rscale = jb.rad_scale; J.slab = jb.slab;
        J.norm = jb.nphot > 0 ? S.mu0 * sc->src_flx / double(jb.nphot) : 0.0;
#pragma omp parallel
        {
#ifdef _OPENMP
            const int tid = omp_get_thread_num();
#else
            const int tid = 0;
#endif
            Tally T;
            T.atomic = !priv;
            if (priv) {
                if (pf[tid].size() != nflux) pf[tid].assign(nflux, 0.0);
                if (pr[tid].size() != nradn) pr[tid].assign(nradn, 0.0);
                if (phh[tid].size() != nheat) phh[tid].assign(nheat, 0.0);
                T.flux = want_flux ? pf[tid].data() : nullptr;
                T.rad = want_rad ? pr[tid].data() : nullptr;
                T.heat = want_heat ? phh[tid].data() : nullptr;
            } else {
                T.flux = want_flux ? flux : nullptr; T.rad = want_rad ? rad : nullptr; T.heat = want_heat ? heat : nullptr;
            }
            Counters& C = cnt[tid];
#pragma omp for schedule(dynamic, 4096)
            for (int64_t i = rank; i < jb.nphot; i += world) {
                Philox R; R.init(jb.seed, uint64_t(i), 0x0ACC1Eu);
                trace_photon(S, J, T, R, C);
            }
        }
    }
    if (priv) {
        for (int t = 0; t < nth; ++t) {
            for (size_t i = 0; i < pf[t].size() && want_flux; ++i) flux[i] += pf[t][i];
            for (size_t i = 0; i < pr[t].size() && want_rad; ++i) rad[i] += pr[t][i];
            for (size_t i = 0; i < phh[t].size() && want_heat; ++i) heat[i] += phh[t][i];
        }
    }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (auto& c : cnt) {
            stats->photons += c.photons; stats->n_cell += c.n_cell; stats->n_tent += c.n_coll; stats->n_coll += c.n_coll;
            stats->n_sfc += c.n_sfc; stats->n_le += c.n_le; stats->n_le_visit += c.n_le_visit; stats->n_tally += c.n_tally;
            stats->n_roulette_kill += c.n_kill;
            stats->w_toa_up += c.w_toa; stats->w_sfc_abs += c.w_sfc; stats->w_atm_abs += c.w_atm; stats->w_roulette += c.w_rr;
        }
    }
    return B200RT_OK;
}

}  // extern "C"
