"""Minimax (Lawson-weighted least squares) fit of GELU(x) ~= x * sigmoid(2 x (c1 + c3 x^2 + c5 x^4)) to the erf GELU;
the 3-term coefficients are the ones in csrc/common.cuh (gelu2).  python tools/fit_gelu.py"""
import numpy as np
from scipy.special import erf
from scipy.optimize import minimize, least_squares
x = np.linspace(-8, 8, 64001)
Phi = 0.5*(1+erf(x/np.sqrt(2)))
gelu = x*Phi
def model(c, x):
    s = x*x
    p = c[0]
    for k in c[1:]:
        pass
    # horner in s
    acc = c[-1]
    for k in c[-2::-1]:
        acc = acc*s + k
    y = x*acc
    return x/(1+np.exp(-2*y))
for nterm in (2,3,4):
    c0 = [np.sqrt(2/np.pi), np.sqrt(2/np.pi)*0.044715] + [0.0]*(nterm-2)
    # minimax via iterated weighted LSQ (Lawson)
    w = np.ones_like(x)
    c = np.array(c0)
    for it in range(60):
        r = least_squares(lambda c: np.sqrt(w)*(model(c,x)-gelu), c, xtol=1e-15, ftol=1e-15)
        c = r.x
        e = np.abs(model(c,x)-gelu)
        w = w*(e/e.max()+1e-3); w/=w.mean()
    e = np.abs(model(c,x)-gelu)
    print(nterm, c.tolist(), "max abs err gelu", e.max(), "at", x[e.argmax()])
    # relative to bf16 half ulp
    g = np.abs(gelu)+1e-30
print("std tanh", np.abs(model([np.sqrt(2/np.pi), np.sqrt(2/np.pi)*0.044715],x)-gelu).max())
