import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
np.set_printoptions(precision=17)
def fit(f, lo, hi, deg, n=4000):
    # near-minimax: interpolate at chebyshev nodes in variable s
    k = np.arange(n)
    t = np.cos(np.pi*(k+0.5)/n)
    s = lo + (hi-lo)*(t+1)/2
    y = f(s)
    c = C.chebfit(t, y, deg)
    # convert to power series in s
    p_t = C.cheb2poly(c)
    # t = (2s - (hi+lo))/(hi-lo)
    a = 2/(hi-lo); b = -(hi+lo)/(hi-lo)
    ps = np.zeros(1)
    tt = np.array([b, a])
    acc = np.array([1.0])
    out = np.zeros(deg+1)
    for i,ci in enumerate(p_t):
        out[:len(acc)] += ci*acc
        acc = P.polymul(acc, tt)
    return out
# atan(a) = a * Q(s), s=a^2 in [0,1]
def g(s):
    a = np.sqrt(np.maximum(s,1e-300))
    return np.where(s<1e-12, 1 - s/3, np.arctan(a)/a)
for deg in (7,8,9):
    co = fit(g, 0.0, 1.0, deg)
    a = np.linspace(0,1,2000001)
    s = a*a
    q = np.polyval(co[::-1], s)
    err = np.abs(a*q - np.arctan(a))
    print('atan deg',deg,'max abs err',err.max())
    print(list(co))
