"""Fit of the sigmoid-form GELU used by the 16-bit-output GEMM epilogue (csrc/act.cuh gelu_sig2):
GELU(x) ~ x * sigmoid(x * P(min(x^2, c))) with P of degree 3 in u = x^2; iteratively re-weighted least squares towards
the minimax fit of the OUTPUT error.  Prints the coefficients (also pre-multiplied by -log2 e) and the max |error|."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

XC = 5.5


def gel(c, x):
    u = np.minimum(x * x, XC * XC)
    g = np.zeros_like(x)
    for ck in c[::-1]:
        g = g * u + ck
    return x / (1 + np.exp(-np.clip(g * x, -87, 87)))


x = np.linspace(-XC, XC, 20001)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
c = np.array([1.5957, 0.0713, 0.0, 0.0])
for p in (2, 4, 8, 16):
    for _ in range(30):
        e = np.abs(gel(c, x) - ref)
        w = 1 + (e / e.max()) ** p * 50
        c = least_squares(lambda cc: (gel(cc, x) - ref) * w * 1e3, c, method="lm", max_nfev=2000).x
xx = np.linspace(-40, 40, 800001)
rr = 0.5 * xx * (1 + erf(xx / np.sqrt(2)))
print("coefficients (c1, c3, c5, c7):", c)
print("times -log2(e):", -c * np.log2(np.e))
print("max |error| on [-40, 40]:", np.abs(gel(c, xx) - rr).max())
# fp32 evaluation of the device formula
f = np.float32
k = (-c * np.log2(np.e)).astype(np.float32)
xs = xx.astype(np.float32)
u = np.minimum(xs * xs, f(XC * XC))
g = ((k[3] * u + k[2]) * u + k[1]) * u + k[0]
out = xs * (f(1) / (f(1) + np.exp2(g * xs)))
print("fp32 evaluation, max |error|:", np.abs(out - rr).max())
