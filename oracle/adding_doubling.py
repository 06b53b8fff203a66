"""
adding_doubling.py -- TEST INFRASTRUCTURE.  Deterministic (non Monte Carlo) plane-parallel solver used to
pin oracle_mc.cpp and the CUDA path, because the reference ships no golden vectors for the transport
itself ("validate against plane-parallel analytic/benchmark cases instead").

Scalar, azimuthally averaged (m = 0) adding-doubling (van de Hulst 1963; Hansen & Travis 1974) on a
Gauss-Radau quadrature of (0, 1] that contains mu = 1, so that nadir radiance (azimuth independent)
and all fluxes are obtained exactly from the m = 0 Fourier term.  The solar direction must be one of the
quadrature nodes (`radau_nodes(n)`; tests choose SZA = arccos(node)).

Layer model identical to the transport contract (include/b200rt.h): each layer holds several scattering
components (ext_k, omega_k, apf_k) plus a pure absorption coefficient; Lambertian surface below.
The role the reference gives to such a check: examples/00_er3t_bmk.py:470-579 (MCARaTS vs libRadtran).
"""

import numpy as np
from numpy.polynomial import legendre as L

__all__ = ['radau_nodes', 'solve', 'hg', 'rayleigh']


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


def _layer_rt(tau, omega, ppp, ppm, mu, w):
    """R, T of a homogeneous layer by doubling from an infinitesimal layer."""
    n = mu.size
    eye = np.eye(n)
    if tau <= 0.0:
        return np.zeros((n, n)), eye.copy()
    ndbl = max(0, int(np.ceil(np.log2(tau / 1.0e-6))))
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
