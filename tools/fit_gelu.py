"""Minimax fit behind csrc/gemm_umma.cu:gelu_sigmoid -- Phi(x) ~= sigmoid(x (a + b x^2 + c x^4)) against the exact-erf GELU.
Prints the coefficients (and the -log2(e)-scaled ones used by the kernel) and the worst-case errors."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

x = np.linspace(-9, 9, 400001)
ge = 0.5 * x * (1 + erf(x / np.sqrt(2)))
dge = 0.5 * (1 + erf(x / np.sqrt(2))) + x * np.exp(-x * x / 2) / np.sqrt(2 * np.pi)


def sig(c, x):
    xc = np.clip(x, -8, 8)
    return 1 / (1 + np.exp(-xc * (c[0] + xc * xc * (c[1] + xc * xc * c[2]))))


r = minimize(lambda c: np.abs(x * sig(c, x) - ge).max(), [1.6, 0.0694, 0.0], method="Nelder-Mead",
             options={"xatol": 1e-10, "fatol": 1e-13, "maxiter": 40000})
c = r.x
s = sig(c, x)
xc = np.clip(x, -8, 8)
d = s + xc * s * (1 - s) * (c[0] + 3 * c[1] * xc * xc + 5 * c[2] * xc ** 4)
print("a, b, c =", c, " max |gelu err| =", r.fun, " max |gelu' err| =", np.abs(d - dge)[np.abs(x) <= 8].max())
print("-log2e*(a,b,c) =", -np.log2(np.e) * c, " (a, 3b, 5c) =", c[0], 3 * c[1], 5 * c[2])
