"""oracle/fit_erfc.py — derives the erfc approximation used by csrc/mdk_pair.cu:erfcx_poly.

TEST INFRASTRUCTURE / provenance script.  Fits erfc(x) exp(x^2) = t P(t), t = 1/(1 + p x), by
relative least squares on Chebyshev nodes of x in [0, 4.2] and reports the error in exact arithmetic
and evaluated with float32 Horner steps.  Run: python oracle/fit_erfc.py
"""
import numpy as np
from scipy.special import erfcx

P, DEG, XMAX = 0.4, 8, 4.2


def fit(p=P, deg=DEG, xmax=XMAX, nodes=4000):
    tmin = 1 / (1 + p * xmax)
    k = np.arange(nodes)
    t = 0.5 * (1 + tmin) + 0.5 * (1 - tmin) * np.cos(np.pi * (k + 0.5) / nodes)
    x = (1 / t - 1) / p
    y = erfcx(x) / t
    A = np.vander(t, deg + 1, increasing=True) / y[:, None]
    coef, *_ = np.linalg.lstsq(A, np.ones_like(y), rcond=None)
    return coef


def errors(coef, p=P, xmax=XMAX):
    x = np.linspace(0, xmax, 200001)
    t64 = 1 / (1 + p * x)
    exact = np.polyval(coef[::-1], t64) * t64
    t32 = (1 / (1 + np.float32(p) * x.astype(np.float32))).astype(np.float32)
    acc = np.full_like(t32, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        acc = (acc * t32 + np.float32(c)).astype(np.float32)
    return np.abs(exact / erfcx(x) - 1).max(), np.abs((acc * t32).astype(np.float64) / erfcx(x) - 1).max()


if __name__ == '__main__':
    c = fit()
    for i, v in enumerate(c):
        print('c%d = %.10ef' % (i, v))
    print('max relative error: exact arithmetic %.2e, float32 Horner %.2e' % errors(c))
