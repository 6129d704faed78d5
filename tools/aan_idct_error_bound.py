"""Worst-case error of the FP32 AAN inverse DCT used by k_inv_transform (jpezy_b200/csrc/dec_transform.cuh).

Inputs: dequantised coefficients F[v][u] (integers, |F| <= Fmax) pre-multiplied by aan[v]*aan[u]/8.
Output: sum/4 in the reference's scaling (src/decoder/jpezy_decoder.hpp:652-670), before the +128 level shift.
First-order bound as in tools/aan_error_bound.py; the bound scales linearly with Fmax, so it is reported per
unit of sum|F * aan*aan/8| ("input L1 norm"), which the kernel accumulates per block to form its guard band.
"""
import numpy as np

AAN = np.array([1.0] + [np.cos(k * np.pi / 16) * np.sqrt(2.0) for k in range(1, 8)])
C = dict(s2=1.414213562373095049, c1=1.847759065022573512, c2=1.082392200292393968, c3=2.613125929752753055)


def idct1d(d, mul, add):
    t0, t1, t2, t3 = d[0], d[2], d[4], d[6]
    t10, t11 = add(t0, t2, 1), add(t0, t2, -1)
    t13 = add(t1, t3, 1)
    t12 = add(mul(add(t1, t3, -1), C["s2"]), t13, -1)
    e0, e3 = add(t10, t13, 1), add(t10, t13, -1)
    e1, e2 = add(t11, t12, 1), add(t11, t12, -1)
    t4, t5, t6, t7 = d[1], d[3], d[5], d[7]
    z13, z10 = add(t6, t5, 1), add(t6, t5, -1)
    z11, z12 = add(t4, t7, 1), add(t4, t7, -1)
    o7 = add(z11, z13, 1)
    o11 = mul(add(z11, z13, -1), C["s2"])
    z5 = mul(add(z10, z12, 1), C["c1"])
    o10 = add(mul(z12, C["c2"]), z5, -1)          # 1.0824*z12 - z5
    o12 = add(z5, mul(z10, C["c3"]), -1)          # z5 - 2.6131*z10
    o6 = add(o12, o7, -1)
    o5 = add(o11, o6, -1)
    o4 = add(o10, o5, 1)
    return [add(e0, o7, 1), add(e1, o6, 1), add(e2, o5, 1), add(e3, o4, -1), add(e3, o4, 1), add(e2, o5, -1), add(e1, o6, -1), add(e0, o7, -1)]


def run(a, dt):
    mul = lambda v, k: (v * dt(k)).astype(dt)
    add = lambda p, q, s: (p + q if s > 0 else p - q).astype(dt)
    cols = [idct1d([a[:, v, u] for v in range(8)], mul, add) for u in range(8)]       # over v, for every u
    out = np.zeros(a.shape, dtype=dt)
    for y in range(8):
        row = idct1d([cols[u][y] for u in range(8)], mul, add)
        for x in range(8):
            out[:, y, x] = row[x]
    return out


def main():
    rng = np.random.default_rng(2)
    n = 100000
    F = np.zeros((n, 8, 8))
    # sparse, JPEG-like coefficient blocks with large magnitudes
    for k in range(n):
        nz = rng.integers(1, 20)
        idx = rng.integers(0, 64, nz)
        F[k].flat[idx] = rng.integers(-1000, 1001, nz)
    F[: n // 10] = rng.integers(-1016, 1017, size=(n // 10, 8, 8))
    pre = F * np.outer(AAN, AAN) / 8
    o64 = run(pre, np.float64)
    o32 = run(pre.astype(np.float32), np.float32).astype(np.float64)
    u = np.arange(8)
    Cm = np.cos((2 * u[None, :] + 1) * u[:, None] * np.pi / 16)
    cu = np.where(u == 0, 2 ** -0.5, 1.0)
    ref = np.einsum("vy,nvu,ux->nyx", Cm * cu[:, None], F, Cm * cu[:, None]) / 4
    print("flowgraph vs definition:", np.abs(o64 - ref).max())
    assert np.abs(o64 - ref).max() < 1e-8
    l1 = np.abs(pre).sum(axis=(1, 2))
    err = np.abs(o32 - o64).max(axis=(1, 2))
    print("observed max |f32 - f64| = %.3e ; max err / input-L1 = %.3e (units of 2^-24: %.2f)" % (err.max(), (err / l1).max(), (err / l1).max() * 2 ** 24))


if __name__ == "__main__":
    main()


# ---- first-order worst-case bound, per unit of input L1 norm -----------------------------------------
class Lin:
    def __init__(self, c, e=0.0, exact=False):
        self.c, self.e = c, e

    def mx(self):                      # max |value| over inputs with sum |pre_k| <= 1
        return np.abs(self.c).max()


U = 2.0 ** -24


def lmul(x, k):
    r = Lin(x.c * k, x.e * abs(k))
    r.e += 2 * U * r.mx()
    return r


def ladd(a, b, s):
    r = Lin(a.c + s * b.c, a.e + b.e)
    r.e += U * r.mx()
    return r


def bound_per_l1():
    x = [[Lin(np.eye(64)[v * 8 + u]) for u in range(8)] for v in range(8)]
    cols = [idct1d([x[v][u] for v in range(8)], lmul, ladd) for u in range(8)]
    worst = 0.0
    for y in range(8):
        row = idct1d([cols[u][y] for u in range(8)], lmul, ladd)
        worst = max(worst, max(r.e for r in row))
    return worst


if __name__ == "__main__":
    k = bound_per_l1()
    print("worst-case |error| <= %.3e * L1(pre)  = %.1f * 2^-24 * L1" % (k, k * 2 ** 24))


def bound_per_coefficient():
    """E[v][u]: worst-case output error (units of 2^-24) per unit of dequantised coefficient F[v][u]"""
    E = np.zeros((8, 8))
    for v0 in range(8):
        for u0 in range(8):
            scale = AAN[v0] * AAN[u0] / 8
            x = [[Lin(np.zeros(1) + (scale if (v, u) == (v0, u0) else 0.0), 2 * U * scale if (v, u) == (v0, u0) else 0.0) for u in range(8)] for v in range(8)]
            # (the input itself carries the rounding of F * (q*aan*aan/8): 2 ulp relative)
            cols = [idct1d([x[v][u] for v in range(8)], lmul, ladd) for u in range(8)]
            worst = 0.0
            for y in range(8):
                row = idct1d([cols[u][y] for u in range(8)], lmul, ladd)
                worst = max(worst, max(r.e for r in row))
            E[v0, u0] = worst * 2 ** 24
    return E


if __name__ == "__main__":
    np.set_printoptions(precision=1, linewidth=150, suppress=True)
    E = bound_per_coefficient()
    print("per-coefficient error sensitivity E (2^-24 per unit dequantised coefficient):\n", E)
    print("C initialiser:")
    for v in range(8):
        print("    " + ", ".join("%.1ff" % (np.ceil(E[v, u] * 10) / 10) for u in range(8)) + ",")
