"""How far does a decoder that starts from a guessed state run before it is synchronised with the true parse?  (test-side tool)

The chain of dependent re-decodes in k_sync_decode is as long as the largest such distance.  This simulates it on the
CPU for an S-photo frame: the true parse of the scan (from the oracle's file) gives the state (block in MCU, zig-zag index) at every
symbol boundary; decoders started at every K-th subsequence boundary with state (b = h, z = 0), h = 0..5, are stepped until they hit
a true state.  Prints the distribution of the distance for the single guess h = 0 (what the kernel does) and for the best of six.
usage: python tools/sync_distance_sim.py [W H [every]]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DC_L = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
DC_C = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
AC_L = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d], [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa])
AC_C = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77], [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa])


def lut16(bits, vals, ac):
    """[16-bit window] -> (bits consumed, zig-zag advance)  (advance 64 = end of block); 0 bits = no code"""
    cons = np.zeros(65536, dtype=np.uint8)
    adv = np.zeros(65536, dtype=np.uint8)
    code, k = 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            sym = vals[k]
            k += 1
            size, run = sym & 15, sym >> 4
            dz = 1 if not ac else (64 if sym == 0 else run + 1)
            lo, hi = code << (16 - ln), (code + 1) << (16 - ln)
            cons[lo:hi] = ln + size
            adv[lo:hi] = dz
            code += 1
        code <<= 1
    return cons, adv


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
    every = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    sub = 128
    import oracle as orc
    from jpezy_b200 import synth
    orc.build()
    o = orc.Oracle("canonical")
    r, g, b = synth.image(0, W, H, 0)
    scan = o.encode(r, g, b, W, H, scan_only=True)
    raw = np.frombuffer(scan, dtype=np.uint8)
    keep = np.ones(raw.size, dtype=bool)
    keep[1:] = ~((raw[1:] == 0) & (raw[:-1] == 0xff))
    u = raw[keep]
    nbits = u.size * 8
    bits = np.unpackbits(np.concatenate([u, np.zeros(8, dtype=np.uint8)]))
    # 16-bit window at every bit position
    win = np.zeros(nbits, dtype=np.uint32)
    for i in range(16):
        win = (win << 1) | bits[i:i + nbits]
    T = [lut16(*DC_L, False), lut16(*DC_C, False), lut16(*AC_L, True), lut16(*AC_C, True)]
    cons = [t[0][win].astype(np.int32) for t in T]
    adv = [t[1][win].astype(np.int32) for t in T]
    cons = [c.tolist() for c in cons]
    adv = [a.tolist() for a in adv]
    nb, ny = 6, 4

    def step(pos, bb, z):
        t = (0 if z == 0 else 2) + (1 if bb >= ny else 0)
        c = cons[t][pos]
        if c == 0:
            return pos + 1, bb, z
        z2 = z + adv[t][pos]
        if z2 >= 64:
            return pos + c, (bb + 1) % nb, 0
        return pos + c, bb, z2

    # true parse
    true = {}
    pos, bb, z = 0, 0, 0
    nmcu = ((W + 15) // 16) * ((H + 15) // 16)
    blocks = 0
    while blocks < nmcu * 6 and pos < nbits:
        true[pos] = (bb, z)
        p2, b2, z2 = step(pos, bb, z)
        if z2 == 0 and (z != 0 or b2 != bb):
            blocks += 1
        pos, bb, z = p2, b2, z2
    end = pos
    print("scan %d bytes, %d symbols, %.2f bits/symbol, %d subsequences of %d bits" % (u.size, len(true), end / len(true), end // sub, sub))
    dist = []
    for s in range(sub, end - 40 * sub, sub * every):
        row = []
        for h in range(6):
            pos, bb, z = s, h, 0
            lim = s + 64 * sub
            while pos < lim and true.get(pos) != (bb, z):
                pos, bb, z = step(pos, bb, z)
            row.append((pos - s + sub - 1) // sub)      # subsequences touched before the states agree
        dist.append(row)
    dist = np.array(dist)
    for hs in ((0,), (0, 3), (0, 4), (0, 2, 4), (0, 1, 2, 3), (0, 2, 4, 5), (0, 1, 2, 3, 4, 5)):
        d = dist[:, list(hs)].min(axis=1)
        print("best of b in %-20s" % (hs,), "mean %.2f  p99 %d  p99.9 %d  max %d   histogram" % (d.mean(), np.percentile(d, 99), np.percentile(d, 99.9), d.max()),
              np.bincount(d)[:24].tolist())
    nsync = (dist <= 4).sum(axis=1)
    print("hypotheses synchronised within 4 subsequences: histogram over starts", np.bincount(nsync, minlength=7).tolist())


if __name__ == "__main__":
    main()
