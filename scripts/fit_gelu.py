"""Fit of the GELU used by the linear1 epilogue (lam_slide_b200/csrc/ptx.cuh: gelu_fast).

gelu_erf(x) = x * Phi(x) = x * sigmoid(2 g(x)) with g(x) = atanh(erf(x / sqrt 2)); g is odd and is fitted by
x * (c0 + c1 x^2 + c2 x^4) minimising the max abs error of the GELU.  Prints the coefficients (and the pre-multiplied
-2 log2(e) * c used with ex2) and the fp32-emulated error."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erfc

x = np.linspace(-10, 10, 200001)
gelu = x * 0.5 * erfc(-x / np.sqrt(2))


def approx(c, x):
    t = np.minimum(x * x, 70.0)
    p = (c[2] * t + c[1]) * t + c[0]
    with np.errstate(over="ignore"):
        return x / (1 + np.exp(-2 * x * p))


def err(c):
    return np.max(np.abs(approx(c, x) - gelu))


if __name__ == "__main__":
    c = [np.sqrt(2 / np.pi), np.sqrt(2 / np.pi) * 0.044715, 0.0]
    for _ in range(8):
        c = minimize(err, c, method="Nelder-Mead", options=dict(xatol=1e-13, fatol=1e-15, maxiter=40000, maxfev=80000)).x
    print("c =", [float(v) for v in c], "max abs err =", err(c))
    print("-2 log2(e) c =", [float(v) for v in -2 * np.log2(np.e) * np.asarray(c)])
    # fp32 emulation of the device code
    xf = x.astype(np.float32)
    t = np.minimum(xf * xf, np.float32(70))
    k = (-2 * np.log2(np.e) * np.asarray(c)).astype(np.float32)
    p = (k[2] * t + k[1]) * t + k[0]
    with np.errstate(over="ignore"):
        y = xf / (np.float32(1) + np.exp2(xf * p))
    print("fp32 max abs err =", float(np.max(np.abs(y.astype(np.float64) - gelu))))
