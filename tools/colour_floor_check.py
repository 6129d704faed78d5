"""Exhaustive check of the integer/FP32 colour path of k_inv_transform (jpezy_b200/csrc/dec_transform.cuh) against the
reference's FP64 expressions (src/decoder/jpezy_decoder.hpp:567-578, revise_value :672-676).

For chroma offsets a = Cb - 128, b = Cr - 128 in [-256, 256] (outside: the kernel's `wild` flag takes the exact path):
  R: y + floor(fl32(b) * fl32(1.4020))              B: y + floor(fl32(a) * fl32(1.7718))
  G: y + floor(-(3441 a + 7139 b) / 10000) in exact integer arithmetic, unless the quotient is exact and (a, b) != (0, 0)
then clamp to [0, 255].  Must equal revise_value(double expression) for every y in [-300, 600]."""
import numpy as np


def revise(v):
    return np.where(v < 0.0, 0, np.where(v > 255.0, 255, np.trunc(v))).astype(np.int64)


def main():
    a = np.arange(-256, 257, dtype=np.int64)
    y = np.arange(-300, 601, dtype=np.int64)
    fr = np.floor(a.astype(np.float32) * np.float32(1.4020)).astype(np.int64)
    fb = np.floor(a.astype(np.float32) * np.float32(1.7718)).astype(np.int64)
    Y, A = np.meshgrid(y, a, indexing="ij")
    ref_r = revise(Y.astype(np.float64) + A.astype(np.float64) * 1.4020)
    ref_b = revise(Y.astype(np.float64) + A.astype(np.float64) * 1.7718)
    assert (np.clip(Y + fr[None, :], 0, 255) == ref_r).all(), "R"
    assert (np.clip(Y + fb[None, :], 0, 255) == ref_b).all(), "B"
    bad = checked = skipped = 0
    for yy in range(-300, 601, 7):          # G: all (a, b) pairs for a stride of y (the offset does not depend on y)
        AA, BB = np.meshgrid(a, a, indexing="ij")
        n = -(3441 * AA + 7139 * BB)
        q = n // 10000                       # floor
        exact = (n % 10000 == 0) & ((AA != 0) | (BB != 0))
        ref = revise((np.float64(yy) - AA.astype(np.float64) * 0.3441) - BB.astype(np.float64) * 0.7139)
        got = np.clip(yy + q, 0, 255)
        bad += int(((got != ref) & ~exact).sum())
        checked += int((~exact).sum())
        skipped += int(exact.sum())
    assert bad == 0, "G: %d mismatches" % bad
    # the magic-number division used on the device: floor(n / 10000) = ((n + 10240000) * 109951163) >> 40 - 1024
    n = np.arange(-(3441 + 7139) * 256, (3441 + 7139) * 256 + 1, dtype=np.int64)
    assert ((((n + 10240000) * 109951163) >> 40) - 1024 == n // 10000).all()
    print("R, B: %d (y, offset) pairs identical; G: %d (y, a, b) triples identical, %d exact-quotient triples left to the FP64 path"
          % (Y.size, checked, skipped))


if __name__ == "__main__":
    main()
