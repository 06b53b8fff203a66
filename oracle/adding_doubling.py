"""
adding_doubling.py -- TEST INFRASTRUCTURE.  Deterministic (non Monte Carlo) plane-parallel solver used to
pin oracle_mc.cpp and the CUDA path, because the reference ships no golden vectors for the transport
itself ("validate against plane-parallel analytic/benchmark cases instead").

Scalar, azimuthally averaged (m = 0) adding-doubling (van de Hulst 1963; Hansen & Travis 1974) on a
Gauss-Radau quadrature of (0, 1] that contains mu = 1, so that nadir radiance (azimuth independent)
and all fluxes are obtained exactly from the m = 0 Fourier term.  The solar direction must be one of the
quadrature nodes (`radau_nodes(n)`; tests choose SZA = arccos(node)).

`solve_views` adds what SURVEY.md 8c(2) asks for beyond that: TABULATED phase functions (piecewise linear in cos(Theta)
on the table's own angle grid -- the definition the transport kernels sample and evaluate, include/b200rt.h) and OBLIQUE
views.  The azimuth dependence is carried by a cosine Fourier series I = sum_m I^m(mu) cos m(phi - phi0); every mode obeys
the same adding-doubling equations with the phase matrices p^m(mu, mu') = (1/2pi) int P(cos Theta) cos(m phi) dphi.  View
zenith angles enter as extra quadrature nodes of weight ZERO (they receive radiance, they do not feed back).  The modes
m >= nmode, which only first-order scattering of a peaked phase function populates, are restored in closed form: exact
single scattering minus its own truncated Fourier series (the TMS idea of Nakajima & Tanaka 1988).

Layer model identical to the transport contract (include/b200rt.h): each layer holds several scattering
components (ext_k, omega_k, apf_k) plus a pure absorption coefficient; Lambertian surface below.
The role the reference gives to such a check: examples/00_er3t_bmk.py:470-579 (MCARaTS vs libRadtran).
"""

import numpy as np
from numpy.polynomial import legendre as L

__all__ = ['radau_nodes', 'solve', 'solve_views', 'solve_beam', 'gauss_nodes', 'hg', 'rayleigh', 'table', 'single_scatter_toa']


def radau_nodes(n):
    """n-point Gauss-Radau rule on (0, 1] including the node mu = 1.  Returns (mu, w), mu increasing."""
    # left Radau on [-1, 1] (x = -1 fixed): roots of P_{n-1}(x) + P_n(x)
    c = np.zeros(n + 1)
    c[n - 1] = 1.0
    c[n] = 1.0
    x = np.sort(L.legroots(c).real)
    x[0] = -1.0
    pn1 = L.legval(x, np.eye(n)[n - 1])          # P_{n-1}(x)
    w = (1.0 - x) / (n * n * pn1 * pn1)
    w[0] = 2.0 / (n * n)
    # flip so that +1 is included, map [-1,1] -> [0,1]
    xr = -x[::-1]
    wr = w[::-1]
    mu = 0.5 * (xr + 1.0)
    return mu, 0.5 * wr


def hg(g):
    def f(c):
        d = 1.0 + g * g - 2.0 * g * c
        return (1.0 - g * g) / (d * np.sqrt(d))
    return f


def rayleigh():
    return lambda c: 0.75 * (1.0 + c * c)


def _azimuth_mean(pfun, mu_i, mu_j, nphi=1440):
    """pbar(mu_i, mu_j) = (1/2pi) int P(mu_i mu_j + sqrt(1-mu_i^2) sqrt(1-mu_j^2) cos phi) dphi."""
    phi = (np.arange(nphi) + 0.5) * (2.0 * np.pi / nphi)
    si = np.sqrt(np.clip(1.0 - mu_i ** 2, 0, None))[:, None, None]
    sj = np.sqrt(np.clip(1.0 - mu_j ** 2, 0, None))[None, :, None]
    c = mu_i[:, None, None] * mu_j[None, :, None] + si * sj * np.cos(phi)[None, None, :]
    return pfun(np.clip(c, -1.0, 1.0)).mean(axis=-1)


def _layer_rt(tau, omega, ppp, ppm, mu, w, dt_max=1.0e-6):
    """R, T of a homogeneous layer by doubling from an infinitesimal layer."""
    n = mu.size
    eye = np.eye(n)
    if tau <= 0.0:
        return np.zeros((n, n)), eye.copy()
    ndbl = max(0, int(np.ceil(np.log2(tau / dt_max))))
    dt = tau / (2.0 ** ndbl)
    minv = 1.0 / mu[:, None]
    R = 0.5 * omega * dt * minv * ppm * w[None, :]
    T = eye - dt * np.diag(1.0 / mu) + 0.5 * omega * dt * minv * ppp * w[None, :]
    for _ in range(ndbl):
        Q = np.linalg.solve(eye - R @ R, eye)
        TQ = T @ Q
        R, T = R + TQ @ R @ T, TQ @ T
    return R, T


def _add(top, bot):
    """Combine (R, T, R*, T*) of `top` over `bot`."""
    Ra, Ta, Ras, Tas = top
    Rb, Tb, Rbs, Tbs = bot
    n = Ra.shape[0]
    eye = np.eye(n)
    A = np.linalg.solve(eye - Rb @ Ras, eye)     # (I - Rb Ra*)^-1
    B = np.linalg.solve(eye - Ras @ Rb, eye)     # (I - Ra* Rb)^-1
    R = Ra + Tas @ A @ Rb @ Ta
    T = Tb @ B @ Ta
    Rs = Rbs + Tb @ B @ Ras @ Tbs
    Ts = Tas @ A @ Tbs
    return R, T, Rs, Ts


def solve(layers, albedo, mu0_index, nstream=48):
    """
    layers : list from TOP to BOTTOM of dict(dz=..., comps=[(ext, omega, pfun), ...], absorb=kabs)
             (ext, kabs in 1/m, dz in m; pfun(cosTheta) normalised to (1/2) int P dmu = 1)
    albedo : Lambertian surface albedo
    mu0_index : index into radau_nodes(nstream)[0] of the solar direction
    returns dict with (per unit flux density normal to the beam, like Src_flx = 1):
        mu0, f_up[levels], f_down[levels], f_down_direct[levels]  (levels from BOTTOM (surface) to TOP,
        the ordering of mca_out_ng), rad_nadir_toa
    """
    mu, w = radau_nodes(nstream)
    n = mu.size
    eye = np.eye(n)
    k0 = mu0_index
    mu0 = mu[k0]

    pcache = {}

    def phase_mats(pfun):
        key = id(pfun)
        if key not in pcache:
            ppp = _azimuth_mean(pfun, mu, mu)
            ppm = _azimuth_mean(pfun, mu, -mu)
            # renormalise so that every incident direction conserves energy: (1/2) sum_j w_j (p++ + p+-) = 1
            s = 0.5 * ((ppp + ppm) * w[:, None]).sum(axis=0)   # sum over outgoing i for incident j
            ppp = ppp / s[None, :]
            ppm = ppm / s[None, :]
            pcache[key] = (ppp, ppm)
        return pcache[key]

    stacks = []
    taus = []
    for lay in layers:
        ext = sum(c[0] for c in lay['comps'])
        sca = sum(c[0] * c[1] for c in lay['comps'])
        kext = ext + lay.get('absorb', 0.0)
        tau = kext * lay['dz']
        taus.append(tau)
        if sca > 0:
            ppp = sum(c[0] * c[1] * phase_mats(c[2])[0] for c in lay['comps']) / sca
            ppm = sum(c[0] * c[1] * phase_mats(c[2])[1] for c in lay['comps']) / sca
            om = sca / kext
        else:
            ppp = ppm = np.zeros((n, n))
            om = 0.0
        R, T = _layer_rt(tau, om, ppp, ppm, mu, w)
        stacks.append((R, T, R, T))

    Rs = 2.0 * albedo * np.outer(np.ones(n), mu * w)
    surf = (Rs, np.zeros((n, n)), Rs, np.zeros((n, n)))

    nl = len(layers)
    # cumulative from the top: top_part[i] = layers[0..i-1]; from the bottom: bot_part[i] = layers[i..] + surface
    ident = (np.zeros((n, n)), eye.copy(), np.zeros((n, n)), eye.copy())
    top_part = [ident]
    for i in range(nl):
        top_part.append(_add(top_part[-1], stacks[i]))
    bot_part = [None] * (nl + 1)
    bot_part[nl] = surf
    for i in range(nl - 1, -1, -1):
        bot_part[i] = _add(stacks[i], bot_part[i + 1])

    inc = np.zeros(n)
    inc[k0] = 1.0 / (2.0 * np.pi * w[k0])

    f_up = np.zeros(nl + 1)
    f_dn = np.zeros(nl + 1)
    f_dir = np.zeros(nl + 1)
    rad_up = np.zeros((nl + 1, n))
    tau_cum = np.concatenate([[0.0], np.cumsum(taus)])
    for i in range(nl + 1):            # interface i counted from the top (0 = TOA)
        Ra, Ta, Ras, Tas = top_part[i]
        Rb = bot_part[i][0]
        idn = np.linalg.solve(eye - Ras @ Rb, Ta @ inc)
        iup = Rb @ idn
        lev = nl - i                   # bottom-up index
        f_dn[lev] = 2.0 * np.pi * np.sum(w * mu * idn)
        f_up[lev] = 2.0 * np.pi * np.sum(w * mu * iup)
        f_dir[lev] = mu0 * np.exp(-tau_cum[i] / mu0)
        rad_up[lev] = iup
    return {'mu0': mu0, 'mu': mu, 'w': w, 'f_up': f_up, 'f_down': f_dn, 'f_down_direct': f_dir,
            'rad_nadir_toa': rad_up[nl][-1], 'rad_up': rad_up}


# ---------------------------------------------------------------------------------------------------------------------
# tabulated phase functions, oblique views (azimuthal Fourier modes)
# ---------------------------------------------------------------------------------------------------------------------
class table:
    """Tabulated phase function as the transport kernels define it: piecewise LINEAR in mu = cos(angle) between the
    table's angles, negative entries clipped, normalised to (1/2) int P dmu = 1 by the exact (trapezoid) integral."""

    nphi = 16384          # azimuth quadrature points (a Mie forward peak is ~0.5 deg wide)

    def __init__(self, ang_deg, pha):
        ang = np.asarray(ang_deg, dtype=np.float64)
        p = np.clip(np.asarray(pha, dtype=np.float64), 0.0, None)
        mu = np.cos(np.deg2rad(ang))
        if abs(ang[0]) < 1e-9:
            mu[0] = 1.0
        if abs(ang[-1] - 180.0) < 1e-9:
            mu[-1] = -1.0
        area = np.sum(0.5 * (p[1:] + p[:-1]) * (mu[:-1] - mu[1:]))
        self.mu = mu[::-1].copy()                 # increasing
        self.p = (p * 2.0 / area)[::-1].copy()

    def __call__(self, c):
        return np.interp(c, self.mu, self.p)


def _nphi_of(pfun):
    return int(getattr(pfun, 'nphi', 2048))


def phase_fourier(pfun, mu_out, mu_in, nmode, nphi=None):
    """p^m[m, i, j] = (1/2pi) int_0^2pi P(mu_out_i mu_in_j + s_i s_j cos phi) cos(m phi) dphi for m < nmode
    (signed cosines: negative = downward).  Midpoint rule on [0, pi] (the integrand is even in phi)."""
    nphi = _nphi_of(pfun) if nphi is None else int(nphi)
    nh = nphi // 2
    phi = (np.arange(nh) + 0.5) * (np.pi / nh)
    cphi = np.cos(phi)
    cm = np.cos(np.outer(phi, np.arange(nmode))) / nh           # (nh, nmode): mean over [0, pi] == mean over [0, 2pi]
    mu_out = np.asarray(mu_out, dtype=np.float64)
    mu_in = np.asarray(mu_in, dtype=np.float64)
    so = np.sqrt(np.clip(1.0 - mu_out ** 2, 0.0, None))
    si = np.sqrt(np.clip(1.0 - mu_in ** 2, 0.0, None))
    out = np.empty((nmode, mu_out.size, mu_in.size))
    for i in range(mu_out.size):
        c = mu_out[i] * mu_in[:, None] + (so[i] * si)[:, None] * cphi[None, :]
        out[:, i, :] = (pfun(np.clip(c, -1.0, 1.0)) @ cm).T
    return out


def single_scatter_toa(layers, mu0, mu_v, pvals):
    """First-order scattered radiance leaving TOA toward mu_v (per unit flux density normal to the beam).
    pvals[l] = sum_k ext_k omega_k P_k(Theta) of layer l (top to bottom) for this geometry."""
    a = 1.0 / mu0 + 1.0 / mu_v
    t = 0.0
    rad = 0.0
    for lay, sp in zip(layers, pvals):
        kext = sum(c[0] for c in lay['comps']) + lay.get('absorb', 0.0)
        tau = kext * lay['dz']
        if kext > 0.0:
            rad += sp / (4.0 * np.pi * mu_v) * np.exp(-t * a) * (-np.expm1(-tau * a)) / (kext * a)
        t += tau
    return rad


def solve_views(layers, albedo, mu0_index, nstream=48, views=(), nmode=32, progress=None):
    """
    Like `solve`, plus the TOA radiance toward oblique directions.

    views : sequence of (vza_deg, dphi_deg): zenith angle of the VIEW (0 = nadir-looking) and the azimuth of the photon
            direction of travel toward the sensor relative to the azimuth of the solar direction of travel
            (dphi = 0: the sensor receives forward-scattered light).
    returns the dict of `solve` plus 'rad_views' (one value per view), 'rad_views_ms_only' (Fourier part alone) and
    'ss_tail' (closed-form first-order scattering in the modes m >= nmode).
    """
    mu_q, w_q = radau_nodes(nstream)
    vmu = [float(np.cos(np.deg2rad(v[0]))) for v in views]
    extra = sorted(set(m for m in vmu if np.min(np.abs(mu_q - m)) > 1e-12))
    mu = np.concatenate([mu_q, extra])
    w = np.concatenate([w_q, np.zeros(len(extra))])
    n = mu.size
    eye = np.eye(n)
    k0 = mu0_index
    mu0 = mu[k0]
    iview = [int(np.argmin(np.abs(mu - m))) for m in vmu]

    pf = {}

    def mats(pfun):
        key = id(pfun)
        if key not in pf:
            ppp = phase_fourier(pfun, mu, mu, nmode)
            ppm = phase_fourier(pfun, mu, -mu, nmode)
            s = 0.5 * ((ppp[0] + ppm[0]) * w[:, None]).sum(axis=0)      # energy per incident direction (m = 0)
            pf[key] = (ppp / s[None, None, :], ppm / s[None, None, :])
        return pf[key]

    res = None
    nl = len(layers)
    rad_m = np.zeros((nmode, len(views)))
    taus = []
    for lay in layers:
        taus.append((sum(c[0] for c in lay['comps']) + lay.get('absorb', 0.0)) * lay['dz'])
    for m in range(nmode):
        stacks = []
        for lay, tau in zip(layers, taus):
            sca = sum(c[0] * c[1] for c in lay['comps'])
            kext = tau / lay['dz']
            if sca > 0:
                ppp = sum(c[0] * c[1] * mats(c[2])[0][m] for c in lay['comps']) / sca
                ppm = sum(c[0] * c[1] * mats(c[2])[1][m] for c in lay['comps']) / sca
                om = sca / kext
            else:
                ppp = ppm = np.zeros((n, n))
                om = 0.0
            R, T = _layer_rt(tau, om, ppp, ppm, mu, w, dt_max=min(1.0e-6, 1.0e-3 * mu.min()))
            stacks.append((R, T, R, T))
        Rs = 2.0 * albedo * np.outer(np.ones(n), mu * w) if m == 0 else np.zeros((n, n))
        surf = (Rs, np.zeros((n, n)), Rs, np.zeros((n, n)))
        inc = np.zeros(n)
        inc[k0] = (1.0 if m == 0 else 2.0) / (2.0 * np.pi * w[k0])
        if m == 0:
            ident = (np.zeros((n, n)), eye.copy(), np.zeros((n, n)), eye.copy())
            top_part = [ident]
            for i in range(nl):
                top_part.append(_add(top_part[-1], stacks[i]))
            bot_part = [None] * (nl + 1)
            bot_part[nl] = surf
            for i in range(nl - 1, -1, -1):
                bot_part[i] = _add(stacks[i], bot_part[i + 1])
            f_up = np.zeros(nl + 1); f_dn = np.zeros(nl + 1); f_dir = np.zeros(nl + 1)
            tau_cum = np.concatenate([[0.0], np.cumsum(taus)])
            for i in range(nl + 1):
                Ra, Ta, Ras, Tas = top_part[i]
                Rb = bot_part[i][0]
                idn = np.linalg.solve(eye - Ras @ Rb, Ta @ inc)
                iup = Rb @ idn
                lev = nl - i
                f_dn[lev] = 2.0 * np.pi * np.sum(w * mu * idn)
                f_up[lev] = 2.0 * np.pi * np.sum(w * mu * iup)
                f_dir[lev] = mu0 * np.exp(-tau_cum[i] / mu0)
                if i == 0:
                    up_toa = iup
            res = {'mu0': mu0, 'mu': mu, 'w': w, 'f_up': f_up, 'f_down': f_dn, 'f_down_direct': f_dir,
                   'rad_nadir_toa': up_toa[nstream - 1]}
        else:
            tot = surf
            for i in range(nl - 1, -1, -1):
                tot = _add(stacks[i], tot)
            up_toa = tot[0] @ inc
        rad_m[m] = up_toa[iview]
        if progress:
            progress(m)
    # Fourier synthesis + the m >= nmode tail of first-order scattering
    dphi = np.deg2rad([v[1] for v in views])
    rad_f = np.array([np.sum(rad_m[:, k] * np.cos(np.arange(nmode) * dphi[k])) for k in range(len(views))])
    tail = np.zeros(len(views))
    for k, (vza, dp) in enumerate(views):
        mv = vmu[k]
        cth = -mu0 * mv + np.sqrt(max(0.0, 1 - mu0 * mu0)) * np.sqrt(max(0.0, 1 - mv * mv)) * np.cos(np.deg2rad(dp))
        p_exact, p_trunc = [], []
        for lay in layers:
            pe = pt = 0.0
            for (e, o, pfun) in lay['comps']:
                if e * o <= 0:
                    continue
                pe += e * o * float(pfun(np.array([np.clip(cth, -1.0, 1.0)]))[0])
                pm = phase_fourier(pfun, np.array([mv]), np.array([-mu0]), nmode)[:, 0, 0]
                fac = np.where(np.arange(nmode) == 0, 1.0, 2.0)
                pt += e * o * float(np.sum(fac * pm * np.cos(np.arange(nmode) * np.deg2rad(dp))))
            p_exact.append(pe); p_trunc.append(pt)
        tail[k] = single_scatter_toa(layers, mu0, mv, p_exact) - single_scatter_toa(layers, mu0, mv, p_trunc)
    res['rad_views_ms_only'] = rad_f
    res['ss_tail'] = tail
    res['rad_views'] = rad_f + tail
    return res


# ---------------------------------------------------------------------------------------------------------------------
# direct beam as a SOURCE (any solar zenith angle), diffuse field on a double-Gauss grid
# ---------------------------------------------------------------------------------------------------------------------
def gauss_nodes(n):
    """n-point Gauss-Legendre rule on (0, 1).  Returns (mu, w), mu increasing."""
    x, w = L.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _layer_rts(tau, omega, ppp, ppm, sp0, sm0, mu, w, mu0, dt_max):
    """R, T and the beam source vectors (S+ leaving the top, S- leaving the bottom, for unit beam flux at the layer top)
    of a homogeneous layer by doubling.  sp0 / sm0: (2 - delta_m0) p^m(+-mu_i, -mu0) / (4 pi)."""
    n = mu.size
    eye = np.eye(n)
    if tau <= 0.0:
        return np.zeros((n, n)), eye.copy(), np.zeros(n), np.zeros(n)
    ndbl = max(0, int(np.ceil(np.log2(tau / dt_max))))
    dt = tau / (2.0 ** ndbl)
    minv = 1.0 / mu[:, None]
    R = 0.5 * omega * dt * minv * ppm * w[None, :]
    T = eye - dt * np.diag(1.0 / mu) + 0.5 * omega * dt * minv * ppp * w[None, :]
    Sp = omega * dt * sp0 / mu
    Sm = omega * dt * sm0 / mu
    eh = np.exp(-dt / mu0)
    for _ in range(ndbl):
        A = np.linalg.solve(eye - R @ R, eye)
        D = A @ (Sm + R @ Sp * eh)
        U = A @ (R @ Sm + Sp * eh)
        Sp, Sm = Sp + T @ U, Sm * eh + T @ D
        TA = T @ A
        R, T = R + TA @ R @ T, TA @ T
        eh = eh * eh
    return R, T, Sp, Sm


def _add_s(top, bot, e_top):
    """(R, T, R*, T*, S+, S-) of `top` over `bot`; e_top = beam transmission of `top`."""
    Ra, Ta, Ras, Tas, Spa, Sma = top
    Rb, Tb, Rbs, Tbs, Spb, Smb = bot
    n = Ra.shape[0]
    eye = np.eye(n)
    A = np.linalg.solve(eye - Ras @ Rb, eye)
    B = np.linalg.solve(eye - Rb @ Ras, eye)
    D = A @ (Sma + Ras @ Spb * e_top)
    U = B @ (Rb @ Sma + Spb * e_top)
    return (Ra + Tas @ B @ Rb @ Ta, Tb @ A @ Ta, Rbs + Tb @ A @ Ras @ Tbs, Tas @ B @ Tbs, Spa + Tas @ U, Smb * e_top + Tb @ D)


def solve_beam(layers, albedo, sza_deg, nstream=64, views=(), nmode=32, progress=None):
    """
    Plane-parallel atmosphere under a collimated beam at ANY solar zenith angle: the direct beam is carried
    analytically and enters the diffuse problem as a source (first-order scattering with the exact phase function).

    layers, albedo : as in `solve` (top to bottom; Lambertian surface)
    views          : sequence of (vza_deg, dphi_deg) as in `solve_views`
    returns f_up / f_down / f_down_direct at the levels from the SURFACE up (per unit flux density normal to the beam),
            'rad_views' at TOA for every view, 'ss_tail', 'mu0'.
    """
    mu_q, w_q = gauss_nodes(nstream)
    mu0 = float(np.cos(np.deg2rad(sza_deg)))
    vmu = [float(np.cos(np.deg2rad(v[0]))) for v in views]
    extra = sorted(set(vmu))
    mu = np.concatenate([mu_q, extra])
    w = np.concatenate([w_q, np.zeros(len(extra))])
    n = mu.size
    eye = np.eye(n)
    iview = [nstream + extra.index(m) for m in vmu]
    nl = len(layers)
    fac = np.where(np.arange(nmode) == 0, 1.0, 2.0)

    pf = {}

    def mats(pfun):
        key = id(pfun)
        if key not in pf:
            ppp = phase_fourier(pfun, mu, mu, nmode)
            ppm = phase_fourier(pfun, mu, -mu, nmode)
            s = 0.5 * ((ppp[0] + ppm[0]) * w[:, None]).sum(axis=0)
            bup = phase_fourier(pfun, mu, np.array([-mu0]), nmode)[:, :, 0]        # out upward, beam travelling down
            bdn = phase_fourier(pfun, -mu, np.array([-mu0]), nmode)[:, :, 0]       # out downward
            s0 = 0.5 * np.sum((bup[0] + bdn[0]) * w)
            pf[key] = (ppp / s[None, None, :], ppm / s[None, None, :], bup / s0, bdn / s0)
        return pf[key]

    taus = [(sum(c[0] for c in lay['comps']) + lay.get('absorb', 0.0)) * lay['dz'] for lay in layers]
    tau_cum = np.concatenate([[0.0], np.cumsum(taus)])
    dt_max = min(1.0e-6, 1.0e-3 * mu.min())
    rad_m = np.zeros((nmode, len(views)))
    res = None
    for m in range(nmode):
        stacks = []
        for lay, tau in zip(layers, taus):
            sca = sum(c[0] * c[1] for c in lay['comps'])
            kext = tau / lay['dz'] if lay['dz'] > 0 else 0.0
            if sca > 0:
                ppp = sum(c[0] * c[1] * mats(c[2])[0][m] for c in lay['comps']) / sca
                ppm = sum(c[0] * c[1] * mats(c[2])[1][m] for c in lay['comps']) / sca
                sp0 = fac[m] * sum(c[0] * c[1] * mats(c[2])[2][m] for c in lay['comps']) / sca / (4.0 * np.pi)
                sm0 = fac[m] * sum(c[0] * c[1] * mats(c[2])[3][m] for c in lay['comps']) / sca / (4.0 * np.pi)
                om = sca / kext
            else:
                ppp = ppm = np.zeros((n, n)); sp0 = sm0 = np.zeros(n); om = 0.0
            R, T, Sp, Sm = _layer_rts(tau, om, ppp, ppm, sp0, sm0, mu, w, mu0, dt_max)
            stacks.append((R, T, R, T, Sp, Sm))
        if m == 0:
            Rs = 2.0 * albedo * np.outer(np.ones(n), mu * w)
            surf = (Rs, np.zeros((n, n)), Rs, np.zeros((n, n)), np.full(n, albedo * mu0 / np.pi), np.zeros(n))
        else:
            zz = np.zeros((n, n))
            surf = (zz, zz, zz, zz, np.zeros(n), np.zeros(n))
        e_lay = [np.exp(-t / mu0) for t in taus]
        if m == 0:
            ident = (np.zeros((n, n)), eye.copy(), np.zeros((n, n)), eye.copy(), np.zeros(n), np.zeros(n))
            top_part, e_top = [ident], [1.0]
            for i in range(nl):
                top_part.append(_add_s(top_part[-1], stacks[i], e_top[-1]))
                e_top.append(e_top[-1] * e_lay[i])
            bot_part = [None] * (nl + 1)
            bot_part[nl] = surf
            for i in range(nl - 1, -1, -1):
                bot_part[i] = _add_s(stacks[i], bot_part[i + 1], e_lay[i])
            f_up = np.zeros(nl + 1); f_dn = np.zeros(nl + 1); f_dir = np.zeros(nl + 1)
            for i in range(nl + 1):
                Ras, Sma = top_part[i][2], top_part[i][5]
                Rb, Spb = bot_part[i][0], bot_part[i][4]
                D = np.linalg.solve(eye - Ras @ Rb, Sma + Ras @ Spb * e_top[i])
                U = np.linalg.solve(eye - Rb @ Ras, Rb @ Sma + Spb * e_top[i])
                lev = nl - i
                f_dir[lev] = mu0 * np.exp(-tau_cum[i] / mu0)
                f_dn[lev] = 2.0 * np.pi * np.sum(w * mu * D) + f_dir[lev]
                f_up[lev] = 2.0 * np.pi * np.sum(w * mu * U)
            up_toa = bot_part[0][4]
            res = {'mu0': mu0, 'mu': mu, 'w': w, 'f_up': f_up, 'f_down': f_dn, 'f_down_direct': f_dir}
        else:
            tot = surf
            for i in range(nl - 1, -1, -1):
                tot = _add_s(stacks[i], tot, e_lay[i])
            up_toa = tot[4]
        rad_m[m] = up_toa[iview]
        if progress:
            progress(m)
    dphi = np.deg2rad([v[1] for v in views])
    rad_f = np.array([np.sum(rad_m[:, k] * np.cos(np.arange(nmode) * dphi[k])) for k in range(len(views))])
    tail = np.zeros(len(views))
    for k, (vza, dp) in enumerate(views):
        mv = vmu[k]
        cth = -mu0 * mv + np.sqrt(max(0.0, 1 - mu0 * mu0)) * np.sqrt(max(0.0, 1 - mv * mv)) * np.cos(np.deg2rad(dp))
        p_exact, p_trunc = [], []
        for lay in layers:
            pe = pt = 0.0
            for (e, o, pfun) in lay['comps']:
                if e * o <= 0:
                    continue
                pe += e * o * float(pfun(np.array([np.clip(cth, -1.0, 1.0)]))[0])
                pm = mats(pfun)[2][:, iview[k]]
                pt += e * o * float(np.sum(fac * pm * np.cos(np.arange(nmode) * np.deg2rad(dp))))
            p_exact.append(pe); p_trunc.append(pt)
        tail[k] = single_scatter_toa(layers, mu0, mv, p_exact) - single_scatter_toa(layers, mu0, mv, p_trunc)
    res['rad_views_ms_only'] = rad_f
    res['ss_tail'] = tail
    res['rad_views'] = rad_f + tail
    return res
