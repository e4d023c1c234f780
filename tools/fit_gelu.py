"""Fit of the sigmoid-form GELU used by the bf16 GEMM epilogues (csrc/common.cuh: gelu_sig_pair).
Phi(u) = sigmoid(2 u P(u^2)); P = polynomial in s = u^2 on s <= 36, iteratively re-weighted least squares towards
the minimax error of Phi.  Prints the coefficients and the maximum errors of Phi, GELU and GELU' over [-8, 8]."""
import numpy as np
import numpy.polynomial.polynomial as P
from scipy.special import erfc


def Phi(u):
    return 0.5 * erfc(-u / np.sqrt(2))


def main(deg=4, U=6.0):
    u = np.linspace(1e-4, U, 20001)
    f = 0.5 * (np.log(Phi(u)) - np.log(Phi(-u))) / u
    x = 2 * u * u / (U * U) - 1
    V = np.polynomial.chebyshev.chebvander(x, deg)
    w0 = 2 * Phi(u) * Phi(-u) * u
    w = w0.copy()
    c = np.linalg.lstsq(V * w[:, None], f * w, rcond=None)[0]
    for _ in range(60):
        err = np.abs((V @ c - f) * w)
        w = w * (1 + 50 * err / err.max())
        c = np.linalg.lstsq(V * w[:, None], f * w, rcond=None)[0]
        w = w / w.max() * w0.max()
    xs = P.Polynomial([-1.0, 2 / (U * U)])
    coef = sum(ck * xs ** k for k, ck in enumerate(np.polynomial.chebyshev.cheb2poly(c))).coef
    uu = np.linspace(-8, 8, 400001)
    ss = np.minimum(uu * uu, U * U)
    y = uu * P.polyval(ss, coef)
    cdf = 1 / (1 + np.exp(-2 * y))
    yp = P.polyval(ss, coef) + 2 * ss * P.polyval(ss, P.polyder(coef)) * (uu * uu < U * U)
    g = cdf + uu * cdf * (1 - cdf) * 2 * yp
    gt = Phi(uu) + uu * np.exp(-uu * uu / 2) / np.sqrt(2 * np.pi)
    print("coefficients", ["%.10e" % v for v in coef])
    print("max |Phi err| %.2e  max |GELU err| %.2e  max |GELU' err| %.2e" %
          (np.abs(cdf - Phi(uu)).max(), np.abs(uu * cdf - uu * Phi(uu)).max(), np.abs(g - gt).max()))


if __name__ == "__main__":
    main()
