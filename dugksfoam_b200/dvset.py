"""Discrete-velocity quadrature sets: Python 3 restatement of the reference's
pre-processing script src/scripts/setDV.py (Python 2).

dvGH: half-range Gauss-Hermite abscissae/weights (setDV.py:83-112)
dvNC: compound Newton-Cotes (Boole) rule          (setDV.py:114-138)

The shipped demo/cavity/constant/{Xis,weights} are dvGH(sqrt(2 R T0), 28)
(doc/demo.tex:28); tests/test_oracle.py::test_golden_quadrature pins this implementation
against them.

The three-term recurrence of setDV.py is evaluated in double precision there and loses
digits quickly (at 14 steps, the shipped 28-point set, its coefficients are already 3e-3
away from the exact ones; from 18 steps on a coefficient turns negative and numpy returns
NaN abscissae).  dvGH therefore follows setDV.py bit for bit where the reference's own
script works, REFUSES sets the script cannot produce, and offers `stable=True`: the same
recurrence in 120-digit decimal arithmetic (valid for any N; not what setDV.py would
print, so never the default).
"""
from __future__ import annotations

import numpy as np


def _gh_recurrence_exact(N: int):
    """a_i, b_i of setDV.py:86-97 in 120-digit decimal arithmetic."""
    from decimal import Decimal as D, getcontext
    ctx = getcontext().copy()
    getcontext().prec = 120
    try:
        pi = D("3.14159265358979323846264338327950288419716939937510582097494459230781640628620899862803482534"
               "2117067982148086513282306647093844609550582231725359408128")
        sp = pi.sqrt()
        a = [D(0)] * N
        b = [D(0)] * N
        a[0] = 1 / sp
        a[1] = 2 / sp / (pi - 2)
        b[1] = a[0] / (a[0] + a[1]) / 2
        for i in range(2, N):
            b[i] = (i - 1) + D(1) / 2 - b[i - 1] - a[i - 1] ** 2
            a[i] = (D(i * i) / 4 / b[i] - b[i - 1] - D(1) / 2) / a[i - 1] - a[i - 1]
        return np.array([float(x) for x in a]), np.array([float(x) for x in b])
    finally:
        getcontext().prec = ctx.prec


def dvGH(C: float, N2: int, stable: bool = False):
    """Half-range Gauss-Hermite set with 2*(N2//2) points scaled by C = sqrt(2RT).

    The recurrence for the half-range Hermite polynomials and the Golub-Welsch
    eigen-decomposition follow setDV.py:86-102; the weights are multiplied by
    exp(xi^2)*C so that sum_k w_k f(xi_k) approximates the plain integral of f
    (setDV.py:110)."""
    if N2 % 2 != 0 or N2 < 7:
        raise ValueError("Number of discrete velocities should be even, and at least 8")
    N = N2 // 2
    if stable:
        a, b = _gh_recurrence_exact(N)
    else:
        a = np.zeros(N)
        b = np.zeros(N)
        a[0] = 1.0 / np.sqrt(np.pi)
        a[1] = 2.0 / np.sqrt(np.pi) / (np.pi - 2.0)
        b[1] = a[0] / (a[0] + a[1]) / 2.0
        for i in range(2, N):
            b[i] = (i - 1) + 1.0 / 2.0 - b[i - 1] - a[i - 1] ** 2
            a[i] = (i ** 2 / 4.0 / b[i] - b[i - 1] - 1.0 / 2) / a[i - 1] - a[i - 1]
        if not (np.all(np.isfinite(a)) and np.all(b[1:] > 0)):
            raise ValueError(f"dvGH: the double-precision recurrence of setDV.py breaks down at N = {N2} "
                             "(negative off-diagonal coefficient, NaN abscissae); the reference script cannot "
                             "produce this set.  Use stable=True (extended-precision recurrence) or dvNC.")
    J = np.diag(a) + np.diag(np.sqrt(b[1:N]), 1) + np.diag(np.sqrt(b[1:N]), -1)
    v, V = np.linalg.eig(J)
    w = V[0, :] * V[0, :] * np.sqrt(np.pi) / 2.0
    order = np.argsort(v)
    v = v[order]
    w = w[order]
    Xis = np.hstack((-np.flipud(v), v))
    weights = np.hstack((np.flipud(w), w))
    weights = weights * np.exp(Xis ** 2) * C
    Xis = Xis * C
    if not (np.all(np.isfinite(Xis)) and np.all(np.isfinite(weights))):
        raise ValueError(f"dvGH: non-finite quadrature for N = {N2}")
    return Xis, weights


def dvNC(xiMax: float, N: int):
    """Compound Newton-Cotes (Boole) rule on [-xiMax, xiMax], N = 4Z+1 points
    (setDV.py:114-138; the 4Z+1 check is at :153)."""
    if N % 4 != 1:
        raise ValueError("The number of points should be 4*Z+1, Z = 1,2,3,...")
    nXi = N
    xiMin = -xiMax
    dv = (xiMax - xiMin) / (nXi - 1)
    nBy4 = (nXi - 1) // 4
    Xis = np.zeros(nXi)
    weights = np.zeros(nXi)
    for i in range(nXi):
        Xis[i] = xiMin + dv * i
    for i in range(nBy4):
        weights[4 * i + 0] = 14.0 / 90 * 4
        weights[4 * i + 1] = 32.0 / 90 * 4
        weights[4 * i + 2] = 12.0 / 90 * 4
        weights[4 * i + 3] = 32.0 / 90 * 4
    weights[0] = 7.0 / 90 * 4
    weights[nXi - 1] = 7.0 / 90 * 4
    for i in range(nXi):
        weights[i] = dv * weights[i]
    return Xis, weights
