"""Fit of the branch-free GELU of csrc/gemm_tc.cu (gelu_erf):  GELU(x) = max(x, 0) - a * Q(a),  a = min(|x|, 6),
Q(a) = erfc(a / sqrt 2) / 2 = 2^-(a p(a) + 1),  p = polynomial fit of -log2(erfc(a / sqrt 2)) / a on [0, 6], Lawson-reweighted for the
absolute error of a * Q(a).  Prints the float32 Horner coefficients (constant term first) and the error of the float32 evaluation
against the fp64 erf form.   python tools/fit_gelu.py [degree]"""
import sys
import numpy as np
from scipy.special import erf, erfc

deg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
A = 6.0
a = np.linspace(1e-6, A, 200001)
two_q = erfc(a / np.sqrt(2))
g = -np.log2(two_q) / a
t = 2 * a / A - 1
V = np.polynomial.chebyshev.chebvander(t, deg)
w = two_q * a * a + 1e-9
for _ in range(30):
    coef, *_ = np.linalg.lstsq(V * w[:, None], g * w, rcond=None)
    h = 0.5 * np.exp2(-a * (V @ coef))
    err = np.abs(a * h - a * 0.5 * two_q)
    w = w * (1 + err / err.max())
c = np.polynomial.polynomial.Polynomial(np.polynomial.chebyshev.cheb2poly(coef))(np.polynomial.polynomial.Polynomial([-1, 2 / A])).coef
af = a.astype(np.float32)
acc = np.full_like(af, np.float32(c[-1]))
for k in range(len(c) - 2, -1, -1):
    acc = acc * af + np.float32(c[k])
hf = np.exp2((-af * acc - np.float32(1.0)).astype(np.float32).astype(np.float64))
pos, neg = a - a * hf, -a * hf
ref_pos, ref_neg = a * 0.5 * (1 + erf(a / np.sqrt(2))), -a * 0.5 * erfc(a / np.sqrt(2))
print("degree", deg, "coefficients", [float(np.float32(v)) for v in c])
print("max |err|", max(np.abs(pos - ref_pos).max(), np.abs(neg - ref_neg).max()), "max rel err (x > 0.01)", (np.abs(pos - ref_pos) / ref_pos)[a > 0.01].max())
