"""Worst-case error of the FP32 AAN forward DCT used by k_fwd_transform (jpezy_b200/csrc/enc_transform.cuh).

Every intermediate of the flowgraph is a linear form of the 64 inputs x in [0,255].  An operation whose
result is not an exactly representable value contributes a rounding error <= 2^-24 * max|result| (plus the
same again for a multiplication by a rounded constant); errors propagate linearly with absolute gains.
The script evaluates that first-order bound for every output (i, j) in v units (the reference's
sum*cu*cv/4 scaling) and also measures the observed error on random / extreme blocks against float64.
DESIGN.md quotes its output; the guard band kGuardF32 must exceed the bound.
"""
import numpy as np

A = dict(c4=0.707106781186547524, c6=0.382683432365089772, c2m6=0.541196100146196985, c2p6=1.306562964876376528)
AAN = np.array([1.0] + [np.cos(k * np.pi / 16) * np.sqrt(2.0) for k in range(1, 8)])   # libjpeg's aanscalefactor


def dct1d(d, mul, add):
    """d: list of 8 values; mul(x, const) and add(a, b, sign) implement the arithmetic (so the same flowgraph
    runs on floats, on linear forms and on error bounds)"""
    t0, t7 = add(d[0], d[7], 1), add(d[0], d[7], -1)
    t1, t6 = add(d[1], d[6], 1), add(d[1], d[6], -1)
    t2, t5 = add(d[2], d[5], 1), add(d[2], d[5], -1)
    t3, t4 = add(d[3], d[4], 1), add(d[3], d[4], -1)
    t10, t13 = add(t0, t3, 1), add(t0, t3, -1)
    t11, t12 = add(t1, t2, 1), add(t1, t2, -1)
    o0, o4 = add(t10, t11, 1), add(t10, t11, -1)
    z1 = mul(add(t12, t13, 1), A["c4"])
    o2, o6 = add(t13, z1, 1), add(t13, z1, -1)
    u10, u11, u12 = add(t4, t5, 1), add(t5, t6, 1), add(t6, t7, 1)
    z5 = mul(add(u10, u12, -1), A["c6"])
    z2 = add(mul(u10, A["c2m6"]), z5, 1)
    z4 = add(mul(u12, A["c2p6"]), z5, 1)
    z3 = mul(u11, A["c4"])
    z11, z13 = add(t7, z3, 1), add(t7, z3, -1)
    o5, o3 = add(z13, z2, 1), add(z13, z2, -1)
    o1, o7 = add(z11, z4, 1), add(z11, z4, -1)
    return [o0, o1, o2, o3, o4, o5, o6, o7]


class Lin:
    """linear form in the 64 inputs + first-order absolute error bound"""
    def __init__(self, coef, err=0.0, exact_int=True):
        self.c, self.e, self.exact = coef, err, exact_int

    def maxabs(self):
        return 255.0 * max(self.c[self.c > 0].sum(), -self.c[self.c < 0].sum())


U = 2.0 ** -24


def lmul(x, k):
    r = Lin(x.c * k, x.e * abs(k), False)
    r.e += 2 * U * r.maxabs()          # rounding of the product + rounding of the constant
    return r


def ladd(a, b, s):
    r = Lin(a.c + s * b.c, a.e + b.e, a.exact and b.exact)
    if not r.exact:
        r.e += U * r.maxabs()
    return r


def bound():
    x = [[Lin(np.eye(64)[y * 8 + xx]) for xx in range(8)] for y in range(8)]
    rows = [dct1d(x[y], lmul, ladd) for y in range(8)]
    out = [[None] * 8 for _ in range(8)]
    for j in range(8):
        col = dct1d([rows[y][j] for y in range(8)], lmul, ladd)
        for i in range(8):
            out[i][j] = col[i]
    B = np.zeros((8, 8))
    for i in range(8):
        for j in range(8):
            B[i, j] = out[i][j].e / (8 * AAN[i] * AAN[j])       # v units
    return B


def observed(nblocks=200000, seed=1):
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, size=(nblocks, 8, 8)).astype(np.float64)
    blocks[: nblocks // 4] = rng.choice([0.0, 255.0], size=(nblocks // 4, 8, 8))     # extreme content
    f32 = blocks.astype(np.float32)

    def run(a, dt):
        mul = lambda v, k: (v * dt(k)).astype(dt)
        add = lambda p, q, s: (p + q if s > 0 else p - q).astype(dt)
        rows = [dct1d([a[:, y, xx] for xx in range(8)], mul, add) for y in range(8)]
        out = np.zeros(a.shape, dtype=dt)
        for j in range(8):
            col = dct1d([rows[y][j] for y in range(8)], mul, add)
            for i in range(8):
                out[:, i, j] = col[i]
        return out
    o32, o64 = run(f32, np.float32).astype(np.float64), run(blocks, np.float64)
    sc = 8 * np.outer(AAN, AAN)
    # cross-check the flowgraph against the textbook definition
    u = np.arange(8)
    C = np.cos((2 * u[None, :] + 1) * u[:, None] * np.pi / 16)
    cu = np.where(u == 0, 2 ** -0.5, 1.0)
    ref = np.einsum("iy,nyx,jx->nij", C, blocks, C) * cu[:, None] * cu[None, :] / 4
    assert np.abs(o64 / sc - ref).max() < 1e-6, "AAN flowgraph does not compute the reference's DCT"
    return np.abs(o32 - o64).max(axis=0) / sc


if __name__ == "__main__":
    np.set_printoptions(precision=2, linewidth=150)
    B = bound()
    O = observed()
    print("first-order worst-case bound, v units (max %.3e):\n" % B.max(), B)
    print("observed max |f32 - f64| over 200k random/extreme blocks, v units (max %.3e):\n" % O.max(), O)
