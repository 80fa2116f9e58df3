#!/usr/bin/env python
"""The coefficients and error bounds of gelu_erf in leaf_b200/csrc/gemm_sm100.cuh: nn.GELU (erf form) as x * sigmoid(q(x)),
q(x) = x (c0 + c1 x^2 + c2 x^4), least-squares fit with a relative-error weighting (floor 2e-3), x^2 clamped at 36. CPU only."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf


def ref(x):
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def model(c, x):
    x2 = np.minimum(x * x, 36.0)
    return x / (1 + np.exp(-x * (c[0] + x2 * (c[1] + x2 * c[2]))))


xs = np.linspace(-7, 7, 28001)
w = np.minimum(1.0 / np.maximum(np.abs(ref(xs)), 2e-3), 50)
c = least_squares(lambda c: (model(c, xs) - ref(xs)) * w, [1.5949, 0.0741, -0.000717], xtol=1e-15, ftol=1e-15, gtol=1e-15).x
print("q(x)/x coefficients:", [float(v) for v in c])
print("folded with -log2(e):", [float(-1.4426950408889634 * v) for v in c])
x = np.linspace(-12, 12, 960001)
e = model(c, x) - ref(x)
print("max |error| %.3g at x = %.3f" % (np.abs(e).max(), x[np.abs(e).argmax()]))
for lo in (1e-2, 1e-3, 1e-4):
    m = np.abs(ref(x)) > lo
    print("max relative error where |y| > %g: %.3g   (bf16 rounding: 3.9e-3)" % (lo, np.abs(e[m] / ref(x)[m]).max()))
xf = x.astype(np.float32)
cf = (-1.4426950408889634 * c).astype(np.float32)
x2 = np.minimum(xf * xf, np.float32(36))
yf = xf / (np.float32(1) + np.exp2(xf * (cf[0] + x2 * (cf[1] + x2 * cf[2]))))
print("fp32 evaluation, folded constants: max |error| %.3g" % np.abs(yf - ref(x)).max())
