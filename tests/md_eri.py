"""Independent check of two-electron integrals over contracted Cartesian Gaussian shells: McMurchie-Davidson scheme
(Hermite expansion coefficients E_t^{ij}, Hermite Coulomb integrals R_tuv from the Boys function) in 40-digit mpmath.
Nothing is shared with the Rys-quadrature path of the oracle or of the CUDA kernels (SURVEY 8c asks for such a cross-check
for d / f quartets).  Normalisation follows the reference: contraction coefficients for primitives normalised on the
x^l component (basis_tools.F90:612-623), every other Cartesian component scaled to unit norm as normalize_ints does
(int2.F90:1187-1207): sqrt((2l-1)!! / ((2a-1)!! (2b-1)!! (2c-1)!!)).  Component order: constants.F90:66-83."""
import itertools

import mpmath as mp

mp.mp.dps = 40

# Cartesian component order of the reference (bf_names, constants.F90:66-83)
CART = {
    0: [(0, 0, 0)],
    1: [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
    2: [(2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1)],
    3: [(3, 0, 0), (0, 3, 0), (0, 0, 3), (2, 1, 0), (2, 0, 1), (1, 2, 0), (0, 2, 1), (1, 0, 2), (0, 1, 2), (1, 1, 1)],
}


def _dfact(n):  # (2n-1)!!
    r = 1
    for k in range(1, n + 1):
        r *= 2 * k - 1
    return r


def comp_norm(l, c):
    a, b, cc = c
    return mp.sqrt(mp.mpf(_dfact(l)) / (_dfact(a) * _dfact(b) * _dfact(cc)))


def hermite_E(i, j, a, b, xa, xb):
    """E_t^{ij}, t = 0..i+j, for a 1-D pair of Gaussians (exponents a, b at xa, xb)."""
    p = a + b
    xp = (a * xa + b * xb) / p
    xpa, xpb = xp - xa, xp - xb
    E = {(0, 0): [mp.e ** (-(a * b / p) * (xa - xb) ** 2)]}

    def get(ii, jj, t):
        v = E[(ii, jj)]
        return v[t] if 0 <= t < len(v) else mp.mpf(0)

    for ii in range(i):
        E[(ii + 1, 0)] = [get(ii, 0, t - 1) / (2 * p) + xpa * get(ii, 0, t) + (t + 1) * get(ii, 0, t + 1) for t in range(ii + 2)]
    for jj in range(j):
        for ii in range(i + 1):
            E[(ii, jj + 1)] = [get(ii, jj, t - 1) / (2 * p) + xpb * get(ii, jj, t) + (t + 1) * get(ii, jj, t + 1)
                               for t in range(ii + jj + 2)]
    return E


def boys(n, x):
    if x < mp.mpf("1e-30"):
        return mp.mpf(1) / (2 * n + 1)
    return mp.gammainc(n + mp.mpf(1) / 2, 0, x) / (2 * x ** (n + mp.mpf(1) / 2))


def hermite_R(L, alpha, X, Y, Z):
    """R_{tuv} = R^0_{tuv}, t+u+v <= L"""
    T = alpha * (X * X + Y * Y + Z * Z)
    R = {}
    for n in range(L + 1):
        R[(n, 0, 0, 0)] = (-2 * alpha) ** n * boys(n, T)
    for tot in range(1, L + 1):
        for t, u, v in itertools.product(range(tot + 1), repeat=3):
            if t + u + v != tot:
                continue
            for n in range(L - tot + 1):
                if t > 0:
                    val = X * R[(n + 1, t - 1, u, v)] + ((t - 1) * R[(n + 1, t - 2, u, v)] if t > 1 else 0)
                elif u > 0:
                    val = Y * R[(n + 1, t, u - 1, v)] + ((u - 1) * R[(n + 1, t, u - 2, v)] if u > 1 else 0)
                else:
                    val = Z * R[(n + 1, t, u, v - 1)] + ((v - 1) * R[(n + 1, t, u, v - 2)] if v > 1 else 0)
                R[(n, t, u, v)] = val
    return R


def shell_quartet(bs, i, j, k, l, mu=None):
    """(ij|kl) block [ni][nj][nk][nl] of contracted, unit-normalised CARTESIAN shells of BasisSet `bs` (as floats).
    mu: the operator is erf(mu r12)/r12 (the range-separated CAM pass, int_rys.F90:179-181, 225-227) instead of 1/r12:
    with 1/r = 2/sqrt(pi) int_0^inf exp(-u^2 r^2) du cut at u = mu, the Boys argument alpha = pq/(p+q) becomes
    alpha' = alpha mu^2 / (alpha + mu^2) and the integral picks up the factor sqrt(alpha'/alpha)."""
    sh = (i, j, k, l)
    L = [int(bs.am[s]) for s in sh]
    cen = [[mp.mpf(float(x)) for x in bs.centers[s]] for s in sh]
    prim = [[(mp.mpf(float(bs.ex[bs.g_offset[s] + m])), mp.mpf(float(bs.cc[bs.g_offset[s] + m]))) for m in range(bs.ncontr[s])] for s in sh]
    comps = [CART[x] for x in L]
    out = [[[[mp.mpf(0) for _ in comps[3]] for _ in comps[2]] for _ in comps[1]] for _ in comps[0]]
    Ltot = sum(L)
    for (a, ca), (b, cb), (c, cc), (d, cd) in itertools.product(*prim):
        p, q = a + b, c + d
        P = [(a * cen[0][x] + b * cen[1][x]) / p for x in range(3)]
        Q = [(c * cen[2][x] + d * cen[3][x]) / q for x in range(3)]
        alpha = p * q / (p + q)
        att = mp.mpf(1)
        if mu is not None:
            m2 = mp.mpf(mu) ** 2
            alpha_eff = alpha * m2 / (alpha + m2)
            att = mp.sqrt(alpha_eff / alpha)
            alpha = alpha_eff
        R = hermite_R(Ltot, alpha, P[0] - Q[0], P[1] - Q[1], P[2] - Q[2])
        Eab = [hermite_E(L[0], L[1], a, b, cen[0][x], cen[1][x]) for x in range(3)]
        Ecd = [hermite_E(L[2], L[3], c, d, cen[2][x], cen[3][x]) for x in range(3)]
        pref = 2 * mp.pi ** (mp.mpf(5) / 2) / (p * q * mp.sqrt(p + q)) * ca * cb * cc * cd * att
        for ia, A in enumerate(comps[0]):
            for ib, B in enumerate(comps[1]):
                ex, ey, ez = Eab[0][(A[0], B[0])], Eab[1][(A[1], B[1])], Eab[2][(A[2], B[2])]
                for ic, Cc in enumerate(comps[2]):
                    for idd, D in enumerate(comps[3]):
                        fx, fy, fz = Ecd[0][(Cc[0], D[0])], Ecd[1][(Cc[1], D[1])], Ecd[2][(Cc[2], D[2])]
                        s = mp.mpf(0)
                        for t, et in enumerate(ex):
                            for u, eu in enumerate(ey):
                                for v, ev in enumerate(ez):
                                    e1 = et * eu * ev
                                    if e1 == 0:
                                        continue
                                    for tt, ft in enumerate(fx):
                                        for uu, fu in enumerate(fy):
                                            for vv, fv in enumerate(fz):
                                                s += e1 * (-1) ** (tt + uu + vv) * ft * fu * fv * R[(0, t + tt, u + uu, v + vv)]
                        out[ia][ib][ic][idd] += pref * s
    res = [[[[float(out[ia][ib][ic][idd] * comp_norm(L[0], A) * comp_norm(L[1], B) * comp_norm(L[2], Cc) * comp_norm(L[3], D))
              for idd, D in enumerate(comps[3])] for ic, Cc in enumerate(comps[2])] for ib, B in enumerate(comps[1])]
           for ia, A in enumerate(comps[0])]
    return res
