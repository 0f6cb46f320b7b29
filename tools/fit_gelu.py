"""Coefficients of the one-MUFU exact-erf GELU used by the GEGLU epilogue (csrc/common.cuh geglu2):
    gelu(g) = relu(g) - 0.5 |g| E(|g|),   E(u) = erfc(u / sqrt 2) ~ 2^-(d1 u + ... + d5 u^5)
Minimax-style fit (iteratively re-weighted least squares on the GELU's absolute error), then an fp32 emulation."""
from math import erf, sqrt

import numpy as np
from scipy.optimize import least_squares

u = np.linspace(0, 9, 9001)
target = np.array([1 - erf(v / sqrt(2)) for v in u])


def E(d, u):
    p = np.zeros_like(u)
    for k, dk in enumerate(d):
        p += dk * u ** (k + 1)
    return np.exp2(-p)


d = np.array([1.151, 0.4593, 0.05256, -0.0074, 0.00052])
w = np.ones_like(u)
for _ in range(80):
    d = least_squares(lambda d: w * (0.5 * u * (E(d, u) - target)), d, method="lm").x
    err = np.abs(0.5 * u * (E(d, u) - target))
    w = w * (1 + err / err.max())
    w /= w.mean()
print("coefficients:", ", ".join(f"{x:.10e}f" for x in d.astype(np.float32)), " max |gelu error| =", err.max())
d32 = d.astype(np.float32)
g = np.linspace(-12, 12, 240001).astype(np.float32)
uu = np.abs(g)
q = d32[4] * uu + d32[3]
for k in (2, 1, 0):
    q = q * uu + d32[k]
out = np.maximum(g, 0) - np.float32(0.5) * uu * np.exp2(-(q * uu))
ref = np.array([0.5 * v * (1 + erf(v / sqrt(2))) for v in g.astype(np.float64)])
print("fp32 emulation: max abs error", np.abs(out - ref).max())
