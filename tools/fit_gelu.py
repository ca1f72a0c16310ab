"""Offline fit of the cheap GELU used by the tcgen05 rollout kernel (csrc/rollout_tc.cu).

erf(z) = 1 - exp(-q(z)),  q(z) = -ln(erfc(z)) ~ z * P(z)  on z in [0, ZMAX]; beyond ZMAX erf == 1 in fp32.
Weighted least squares + a few Remez-style reweighting sweeps so that the ABSOLUTE error of erf (hence of
GELU(x) = 0.5 x (1 + sign(x) erf(|x|/sqrt2))) is minimised.  Prints coefficients for q(z)*log2(e) in Horner order
and the max abs error of GELU evaluated in float32 arithmetic against the float64 exact-erf GELU.
"""
import numpy as np
from scipy.special import erfc, erf

ZMAX = 4.4
LOG2E = 1.4426950408889634


def fit(deg, sweeps=60):
    z = np.linspace(1e-6, ZMAX, 40001)
    q = -np.log(erfc(z))
    target = q / z                       # P(z)
    w = erfc(z) * z                      # d erf = erfc * dq = erfc * z * dP
    V = np.vander(z, deg + 1, increasing=True)
    extra = np.ones_like(z)
    for _ in range(sweeps):
        W = w * extra
        coef, *_ = np.linalg.lstsq(V * W[:, None], target * W, rcond=None)
        err = np.abs((V @ coef - target) * w)
        extra *= (1 + 4 * err / err.max()) ** 0.5
        extra /= extra.mean()
    return coef, err.max()


def gelu_f32(x, coef2):
    x = x.astype(np.float32)
    z = np.minimum(np.abs(x) * np.float32(0.7071067811865476), np.float32(ZMAX))
    p = np.float32(coef2[-1])
    for c in coef2[-2::-1]:
        p = np.float32(p * z + np.float32(c))
    t = np.float32(-(p * z))                      # -q(z) * log2(e)
    e = np.exp2(t.astype(np.float32)).astype(np.float32)
    er = np.float32(1.0) - e
    er = np.copysign(er, x)
    hx = np.float32(0.5) * x
    return (hx * er + hx).astype(np.float32)


if __name__ == "__main__":
    x = np.linspace(-9, 9, 2_000_001)
    exact = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    for deg in (4, 5, 6, 7, 8):
        coef, e_erf = fit(deg)
        coef2 = coef * LOG2E
        g = gelu_f32(x, coef2)
        err = np.abs(g.astype(np.float64) - exact)
        print(f"deg {deg}: max abs erf-fit err {e_erf:.2e}; GELU f32 max abs err {err.max():.2e} at x={x[err.argmax()]:.3f}")
        print("   coef (q*log2e, increasing powers of z):", ", ".join(f"{c:.9e}f" for c in coef2))


def packed_form(deg=5):
    """Coefficients of the form used by gelu_fast2 in csrc/rollout_tc.cu:
    zn = -min(|x|, L); t = zn * Pt(zn) - 1; GELU = max(x, 0) + zn * exp2(t)   (0.5 * erfc folded into the -1)."""
    coef, _ = fit(deg)
    a = coef * LOG2E
    s = 0.7071067811865476
    pt = [a[k] * s ** (k + 1) * (-1) ** k for k in range(deg + 1)]
    L = ZMAX / s
    x = np.linspace(-9, 9, 2_000_001).astype(np.float32)
    exact = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    zn = np.maximum(-np.abs(x), np.float32(-L)).astype(np.float32)
    p = np.float32(pt[-1])
    for c in pt[-2::-1]:
        p = (p * zn + np.float32(c)).astype(np.float32)
    t = (p * zn + np.float32(-1.0)).astype(np.float32)
    e = np.exp2(t.astype(np.float64)).astype(np.float32)
    g = (zn * e + np.maximum(x, np.float32(0))).astype(np.float32)
    err = np.abs(g.astype(np.float64) - exact)
    print(f"packed form deg {deg}: L = {L:.7f}; GELU f32 max abs err {err.max():.2e} at x={x[err.argmax()]:.3f}")
    print("   Pt (increasing powers of zn):", ", ".join(f"{c:.9e}f" for c in pt))


if __name__ == "__main__":
    packed_form(5)
