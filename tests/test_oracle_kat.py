"""CPU tests that pin the oracle (oracle/jpezy_oracle.cpp).

The reference ships no tests or golden vectors.  What can be pinned without running it:
its literal tables (parsed from the reference headers into tests/golden/ref_tables.json by
tests/golden/make_golden.py), the 644-byte header layout, hand-derived per-block bit strings, and
cross-decoder sanity.  Outputs of the reference's own code compiled here are covered by
tests/test_oracle_vs_ref.py.
"""
import hashlib
import io
import json
import os

import numpy as np
import pytest

import jpezy_b200 as J

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = json.load(open(os.path.join(G, "ref_tables.json")))
VEC = json.load(open(os.path.join(G, "oracle_vectors.json")))


def unstuff(data):
    return data.replace(b"\xff\x00", b"\xff")


def bits_of(data):
    """bit string of an entropy-coded segment with the FF00 stuffing removed"""
    return "".join("{:08b}".format(x) for x in unstuff(data))


def test_tables_match_reference_literals(oracle):
    zz, qy, qc, cos, ds = oracle.tables()
    assert zz.tolist() == REF["ZZ"]
    assert qy.tolist() == REF["YQuantumTb"] and qc.tolist() == REF["CQuantumTb"]
    assert sum(REF["YQuantumTb"]) == 3688 and sum(REF["CQuantumTb"]) == 5505
    assert sorted(REF["ZZ"]) == list(range(64))
    assert ds.hex() == "0x1.6a09e667f3bccp-1"          # decision O2
    assert cos[0] == 1.0 and abs(cos[4 * 8 + 0] - 2 ** -0.5) < 1e-15


@pytest.mark.parametrize("chroma,prefix", [(0, "Y"), (1, "C")])
def test_encoder_lut_derivation_matches_reference_luts(oracle, chroma, prefix):
    # canonical codes derived from the DHT payloads reproduce every literal (size, code) entry of
    # src/encoder/huffman_table.hpp:26-195 under idx = run*10 + size + (run==15), EOB = 0, ZRL = 151
    dcs, dcc, acs, acc = oracle.enc_lut(chroma)
    assert dcs.tolist() == REF[prefix + "DcSizeT"] and dcc.tolist() == REF[prefix + "DcCodeT"]
    assert acs.tolist() == REF[prefix + "AcSizeT"] and acc.tolist() == REF[prefix + "AcCodeT"]


def test_header_is_644_bytes_and_carries_the_reference_dht_arrays(oracle):
    h = oracle.header(512, 512)
    assert len(h) == 644
    assert h[:2] == b"\xff\xd8" and h[2:4] == b"\xff\xe0" and h[6:11] == b"JFIF\0" and h[11:13] == b"\x01\x02"
    assert h[13] == 1 and h[14:18] == b"\x00\x60\x00\x60" and h[18:20] == b"\0\0"
    assert h[20:24] == b"\xff\xfe\x00\x13" and h[24:41] == b"Encoded by jpezy\0"
    assert oracle.header(8, 8, gray=True)[24:41] == b"Encoded by JPEZY\0"
    p = 41
    for t, q in ((0, "YQuantumTb"), (1, "CQuantumTb")):
        assert h[p:p + 5] == bytes([0xff, 0xdb, 0x00, 0x43, t])
        assert list(h[p + 5:p + 69]) == [REF[q][REF["ZZ"][i]] for i in range(64)]
        p += 69
    for name in ("YDcDht", "CDcDht", "YAcDht", "CAcDht"):
        n = len(REF[name])
        assert list(h[p:p + n]) == REF[name], name
        p += n
    assert h[p:p + 19] == bytes([0xff, 0xc0, 0, 17, 8, 2, 0, 2, 0, 3, 0, 0x22, 0, 1, 0x11, 1, 2, 0x11, 1])
    p += 19
    assert h[p:] == bytes([0xff, 0xda, 0, 12, 3, 0, 0, 1, 0x11, 2, 0x11, 0, 63, 0]) and p + 14 == 644
    # marker order of README.md:89-112: APP0, COM, DQT, DQT, DHT x4, SOF0, SOS
    w = oracle.header(3840, 2160)
    assert w[p - 19 + 5:p - 19 + 9] == bytes([2160 >> 8, 2160 & 255, 3840 >> 8, 3840 & 255])


def test_block_known_answers(oracle):
    def scan(blocks):     # blocks: list of (block_index_in_mcu, zz_index, value)
        c = np.zeros((1, 6, 64), dtype=np.int16)
        for k, n, v in blocks:
            c[0, k, n] = v
        return oracle.scan_from_coefs(c)
    # all-zero MCU: Y blocks = DC cat 0 "00" + EOB "1010" (x4), chroma = "00" + "00" (x2)  -> 32 bits
    z = scan([])
    assert bits_of(z) == "001010" * 4 + "0000" * 2
    # DC +1 in Y0: cat 1 = "010" + "1"; the next Y block sees diff -1 = "010" + "0"
    s = bits_of(scan([(0, 0, 1)]))
    assert s.startswith("010" + "1" + "1010" + "010" + "0" + "1010" + "00" + "1010")
    # AC: run 0 size 1 (+1) in Y0 at zz 1: code 00 + "1" ; then EOB
    s = bits_of(scan([(0, 1, 1)]))
    assert s.startswith("00" + "00" + "1" + "1010")
    # coefficient 63 non-zero: 62 zeros = 3 ZRL (11111111001) + run 14 size 1 (1111111111101011 0 for -1), no EOB
    s = bits_of(scan([(0, 63, -1)]))
    assert s.startswith("00" + "11111111001" * 3 + "1111111111101011" + "0" + "001010")
    # chroma: 16 zeros before zz 17 -> ZRL 1111111010, then run 0 / size 1 = "01" + "1", EOB "00"; pad with 1s
    s = bits_of(scan([(4, 17, 1)]))
    assert s == "001010" * 4 + "00" + "1111111010" + "01" + "1" + "00" + "0000" + "111"
    # the byte 0xFF produced by code bits is followed by a stuffed 0x00 (bofstream behaviour, decision O4)
    raw = scan([(0, 63, -1)])
    assert b"\xff\x00" in raw and len(unstuff(raw)) == len(raw) - raw.count(b"\xff\x00")


def test_pad_policy_only_changes_the_last_byte(oracle):
    r, g, b = J.synth.image(0, 40, 24)
    c = oracle.coefs(r, g, b, 40, 24)
    one, zero = oracle.scan_from_coefs(c, pad_ones=True), oracle.scan_from_coefs(c, pad_ones=False)
    assert one[:-2] == zero[:-2] and len(one) - len(zero) in (0, 1)


@pytest.mark.parametrize("name", sorted(VEC))
def test_golden_vectors(oracle, name):
    v = VEC[name]
    r, g, b = J.synth.image(v["family"], v["W"], v["H"])
    assert hashlib.sha256(r.tobytes() + g.tobytes() + b.tobytes()).hexdigest() == v["input_sha256"]
    f = oracle.encode(r, g, b, v["W"], v["H"], gray=v["gray"])
    assert len(f) == v["file_bytes"] and hashlib.sha256(f).hexdigest() == v["file_sha256"]
    c = oracle.coefs(r, g, b, v["W"], v["H"], gray=v["gray"])
    assert hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest() == v["coefs_sha256"]
    _, _, R, Gp, B = oracle.decode(f, gray=v["gray"])
    assert hashlib.sha256(R.tobytes() + Gp.tobytes() + B.tobytes()).hexdigest() == v["decoded_sha256"]


@pytest.mark.parametrize("W,H,family", [(200, 120, 0), (64, 64, 1), (17, 33, 2), (1, 1, 0), (16, 16, 0)])
def test_stream_consistency(oracle, W, H, family):
    r, g, b = J.synth.image(family, W, H)
    f = oracle.encode(r, g, b, W, H)
    c = oracle.coefs(r, g, b, W, H)
    assert f[:644] == oracle.header(W, H) and f[-2:] == b"\xff\xd9"
    assert f[644:-2] == oracle.scan_from_coefs(c) == oracle.encode(r, g, b, W, H, scan_only=True)
    assert (oracle.decode_coefs(f) == c).all()          # Huffman decode inverts Huffman encode
    Wd, Hd, R, Gp, B = oracle.decode(f)
    assert (Wd, Hd) == (W, H) and R.size == ((W + 15) // 16 * 16) * ((H + 15) // 16 * 16)


def test_gray_mode(oracle):
    W, H = 48, 32
    r, g, b = J.synth.image(0, W, H)
    c = oracle.coefs(r, g, b, W, H, gray=True)
    assert (c[:, 4:, :] == 0).all() and (c[:, :4] == oracle.coefs(r, g, b, W, H)[:, :4]).all()
    f = oracle.encode(r, g, b, W, H, gray=True)
    _, _, R, Gp, B = oracle.decode(f, gray=True)
    assert (R == Gp).all() and (R == B).all()


def test_flat_block_dc_is_off_by_one_ulp(oracle):
    # (1/sqrt 2)^2 < 0.5 in double, so a flat block of value p has DCT DC = 8p - ulp -> int() = 8p - 1
    W = H = 16
    for p, want in [(200, (8 * 72 - 1) // 16), (130, (8 * 2 - 1) // 16), (64, -((8 * 64 - 1) // 16))]:
        v = np.full((H, W), p, dtype=np.uint8)
        c, raw = oracle.coefs(v, v, v, W, H, want_raw=True)
        y = p - 128 if p != 200 else 72
        assert c[0, 0, 0] == want, (p, c[0, 0, 0], raw[0, 0, 0])


def test_cross_decoder_sanity(oracle):
    # sanity, not parity: libjpeg-turbo (Pillow) and OpenCV open the oracle's files and agree roughly
    from PIL import Image
    W, H = 200, 120
    r, g, b = J.synth.image(0, W, H)
    f = oracle.encode(r, g, b, W, H)
    im = Image.open(io.BytesIO(f))
    im.load()
    assert im.size == (W, H) and im.mode == "RGB"
    pil = np.asarray(im.convert("RGB")).astype(int)
    _, _, R, Gp, B = oracle.decode(f)
    mine = np.stack([x[: W * H].reshape(H, W) for x in (R, Gp, B)], -1).astype(int)
    assert np.abs(mine - pil).mean() < 4.0
    assert np.abs(mine[..., 0] - r).mean() < 8.0
    try:
        import cv2
        bgr = cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
        assert bgr is not None and bgr.shape == (H, W, 3)
    except ImportError:
        pass


def test_shipped_flags_variant_differences_are_reported(oracle, oracle_shipped):
    # decision O3: canonical = strict IEEE; the reference's Release flags allow FMA contraction.  Count, don't hide.
    W, H = 256, 64
    r, g, b = J.synth.image(2, W, H)
    a, s = oracle.coefs(r, g, b, W, H), oracle_shipped.coefs(r, g, b, W, H)
    n = int((a != s).sum())
    print("coefficients differing between strict and as-shipped flags on the adversarial image: %d of %d" % (n, a.size))
    assert np.abs(a.astype(int) - s).max() <= 1
