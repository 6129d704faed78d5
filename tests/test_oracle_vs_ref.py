"""The oracle restatement against the reference's OWN code.

oracle/_ref holds /root/reference/src compiled unmodified against oracle/shim (stand-ins for the absent
SrookCppLibraries / Boost headers, see oracle/shim/README.md).  It cannot travel as source, so:
  * tests/golden/ref_vectors.json (made by tests/golden/make_golden.py in the build container) pins its outputs; the
    first group of tests checks the oracle against those committed vectors on any box;
  * where oracle/_ref exists (build container; its binaries also travel to the GPU box) the second group compares
    the oracle with the reference live on more inputs, including the unmodified jpezy_encode / jpezy_decode CLIs.
What this does not pin: behaviour inside the absent third-party library (decisions O1..O7).
"""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import jpezy_b200 as J
import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = json.load(open(os.path.join(G, "ref_vectors.json")))
CASES = sorted(k for k in VEC if not k.startswith("_"))
needs_ref = pytest.mark.skipif(orc.ref_dir() is None, reason="oracle/_ref not built (needs /root/reference)")


def sha(*arrs):
    return hashlib.sha256(b"".join(a.tobytes() if hasattr(a, "tobytes") else a for a in arrs)).hexdigest()


# ---- committed vectors ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_oracle_encoder_reproduces_reference_files(oracle, name):
    v = VEC[name]
    r, g, b = J.synth.image(v["family"], v["W"], v["H"])
    assert sha(r, g, b) == v["input_sha256"], "synthetic generator changed: regenerate tests/golden"
    f = oracle.encode(r, g, b, v["W"], v["H"], gray=v["gray"])
    assert len(f) == v["file_bytes"] == v["wrote_size"]          # decision O7: wrote_size() == file length
    assert sha(f) == v["file_sha256"]
    if "file_hex" in v:
        assert f.hex() == v["file_hex"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("gray", [False, True])
def test_oracle_decoder_reproduces_reference_planes(oracle, name, gray):
    v = VEC[name]
    r, g, b = J.synth.image(v["family"], v["W"], v["H"])
    f = oracle.encode(r, g, b, v["W"], v["H"], gray=v["gray"])
    W, H, R, G_, B = oracle.decode(f, gray=gray)
    assert (W, H) == (v["W"], v["H"]) and R.size == v["plane_bytes"]
    assert sha(R, G_, B) == v["decoded_gray_sha256" if gray else "decoded_sha256"]


def test_oracle_constants_equal_the_stand_in_constants(oracle):
    # O1/O2: run-time std::cos / std::sqrt (oracle) against GCC's correctly rounded constant folding (shim)
    _, _, _, cos, ds = oracle.tables()
    c = VEC["_constants"]
    assert [float(x).hex() for x in cos] == c["cos_table_hex"]
    assert float(ds).hex() == c["dis_sqrt_hex"] == "0x1.6a09e667f3bccp-1"


# ---- live comparison ------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    return orc.Reference()


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_live_random_sizes(oracle, ref, seed):
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(1, 150)), int(rng.integers(1, 110))
    fam = int(rng.integers(0, 3))
    gray = bool(rng.integers(0, 2))
    r, g, b = J.synth.image(fam, W, H, frame=seed)
    if seed % 2:        # arbitrary content, not just the synthetic families
        r, g, b = (rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(3))
    f_ref, wrote = ref.encode(r, g, b, W, H, gray=gray)
    f = oracle.encode(r, g, b, W, H, gray=gray)
    assert f == f_ref and wrote == len(f)
    for gm in (False, True):
        a = ref.decode(f, gray=gm)
        o = oracle.decode(f, gray=gm)
        assert a[:2] == o[:2] and all((x == y).all() for x, y in zip(a[2:], o[2:]))


@needs_ref
def test_live_reference_rejects_what_the_oracle_rejects(oracle, ref):
    r, g, b = J.synth.image(1, 64, 48)
    f = oracle.encode(r, g, b, 64, 48)
    cut = f[: 644 + (len(f) - 646) // 3]
    assert ref.decode(cut) is None
    with pytest.raises(RuntimeError):
        oracle.decode(cut)


@needs_ref
def test_live_cli_round_trip_matches_oracle(oracle, ref, tmp_path):
    """the reference's unmodified main.cpp files: P3 in -> JPEG -> P3 out"""
    W, H = 37, 21
    r, g, b = J.synth.image(0, W, H)
    ppm = tmp_path / "in.ppm"
    rgb = np.stack([r, g, b], -1).reshape(-1, 3)
    ppm.write_text("P3\n# comment line\n%d %d\n255\n" % (W, H) + "".join("%d %d %d\n" % tuple(p) for p in rgb))
    for flags, gray in (([], False), (["--gray"], True)):
        jpg, out = tmp_path / "o.jpg", tmp_path / "o.ppm"
        p = subprocess.run([ref.encode_exe, str(ppm), str(jpg)] + flags, capture_output=True, text=True)
        assert p.returncode == 0 and ("Output size: %d" % jpg.stat().st_size) in p.stdout
        f = jpg.read_bytes()
        assert f == oracle.encode(r, g, b, W, H, gray=gray)
        p = subprocess.run([ref.decode_exe, str(jpg), str(out)] + flags, capture_output=True, text=True)
        assert p.returncode == 0 and "Decoded image: Netpbm image data, size = %d x %d" % (W, H) in p.stdout
        lines = out.read_text().split("\n")
        assert lines[:4] == ["P3", "# Decoded by jpezy", "%d %d" % (W, H), "255"]
        vals = np.array(" ".join(lines[4:]).split(), dtype=np.int64).reshape(-1, 3)
        _, _, R, G_, B = oracle.decode(f, gray=gray)
        assert (vals[:, 0] == R[: W * H]).all() and (vals[:, 1] == G_[: W * H]).all() and (vals[:, 2] == B[: W * H]).all()
