"""The drop-in CLIs (jpezy_b200/bin/jpezy_encode, jpezy_decode: host C++ mirror of the reference's classes over the C ABI)
against the reference's own CLIs (oracle/_ref, built from /root/reference/src/*/main.cpp unmodified) and the oracle.

CPU part: argv handling, usage/exit codes, the P3 reader's grammar (echo mode needs no device) and the loud failure
without a CUDA device.  GPU part (-m gpu): files byte-identical to the reference's, decoded PPM identical, same console
lines modulo the timing figures.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import jpezy_b200 as J
import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "jpezy_b200", "bin")
ENC, DEC = os.path.join(BIN, "jpezy_encode"), os.path.join(BIN, "jpezy_decode")
needs_ref = pytest.mark.skipif(orc.ref_dir() is None, reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "jpezy_b200", "cli"), "-s"])


def run(*args):
    return subprocess.run(list(args), capture_output=True, text=True)


def write_ppm(path, r, g, b, header="P3\n%d %d\n255\n", tail="\n"):
    H, W = r.shape
    rgb = np.stack([r, g, b], -1).reshape(-1, 3)
    body = "\n".join("%d %d %d" % tuple(p) for p in rgb)
    path.write_text(header % (W, H) + body + tail)


def strip_times(s):
    s = re.sub(r"Processing time: [0-9.e+-]+\(sec\)", "Processing time: T(sec)", s)
    return re.sub(r"Total processing time: [0-9.e+-]+", "Total processing time: T", s)


# ---- CPU ------------------------------------------------------------------------------------------
def test_usage_and_exit_codes():
    p = run(ENC)
    assert p.returncode == 1 and p.stderr.startswith("Usage: jpezy_encode <input.ppm>")
    p = run(ENC, "a.ppm", "out.bmp")
    assert p.returncode == 1 and "Usage: jpezy_encode" in p.stderr
    p = run(DEC, "a.jpg")
    assert p.returncode == 1 and p.stderr.startswith("Usage: jpezy_decode <input.(jpg | jpeg)>")
    p = run(DEC, "a.png", "b.ppm")
    assert p.returncode == 1
    p = run(ENC, "/nonexistent/in.ppm", "/tmp/x.jpg")
    assert p.returncode == 1 and "The file is not found or the formatting error" in p.stderr


PPM_CASES = {
    "plain": "P3\n2 2\n255\n1 2 3\n4 5 6\n7 8 9\n10 11 12\n",
    "comments_everywhere": "# c0\nP3\n# c1\n2 2\n# c2\n255\n1 2 3 4 5 6\n# dropped 99 99 99\n7 8 9 10 11 12\n",
    "data_line_with_hash_is_dropped": "P3\n2 1\n255\n1 2 3 # tail\n4 5 6\n7 8 9\n",
    "no_final_newline_drops_last_line": "P3\n2 2\n255\n1 2 3\n4 5 6\n7 8 9\n10 11 12",
    "trailing_blank_and_crlf": "P3\n2 2\n255\n1 2 3 \n4 5 6\r\n7 8 9 10 11 12\n",
    "values_wrap_mod_256": "P3\n1 1\n65535\n256 511 1000\n",
    "stoi_oddities": "P3\n2 2\n 255\n+5 12x -1\n\n7\t8 9\n010 0x10 99999999999\n1 2 3\n",
    "interior_empty_token_throws": "P3\n1 1\n255\n1  2 3\n",
    "not_p3": "P6\n1 1\n255\n1 2 3\n",
    "size_line_three_tokens": "P3\n2 2 \n255\n1 2 3\n",
}


@needs_ref
@pytest.mark.parametrize("name", sorted(PPM_CASES))
def test_p3_reader_grammar_matches_reference_cli(tmp_path, name):
    """echo mode (<output.ppm>) runs the P3 reader and the P3 echo writer only: no device needed"""
    ref = orc.Reference()
    src = tmp_path / "in.ppm"
    src.write_text(PPM_CASES[name])
    a = run(ref.encode_exe, str(src), str(tmp_path / "ref.ppm"))
    b = run(ENC, str(src), str(tmp_path / "mine.ppm"))
    assert a.returncode == b.returncode
    assert strip_times(a.stdout) == strip_times(b.stdout) and a.stderr == b.stderr
    if a.returncode == 0:
        assert (tmp_path / "ref.ppm").read_text() == (tmp_path / "mine.ppm").read_text()


def _big_ppm(seed, bad_at=None, tail_newline=True):
    """~2.5 MB of pixel lines (above the reader's 1 MiB threshold for splitting the body over the host's cores) with the
    grammar's quirks sprinkled in: comment lines, data lines with '#', trailing blanks, tabs, CRLF, several pixels per line"""
    rng = np.random.default_rng(seed)
    W, H = 640, 330
    vals = rng.integers(0, 256, size=W * H * 3)
    out = ["P3\n# made for the test\n%d %d\n255\n" % (W, H)]
    i, n, line_no = 0, len(vals), 0
    while i < n:
        k = int(rng.integers(1, 13)) * 3 if line_no % 7 else 3
        toks = [str(v) for v in vals[i:i + k]]
        i += k
        sep = "\t" if line_no % 11 == 3 else " "
        line = sep.join(toks)
        if line_no % 13 == 5:
            line += " "                      # one trailing blank is forgiven
        if line_no % 17 == 9:
            line += "\r"                     # CRLF: the '\r' separates a trailing empty token
        out.append(line + "\n")
        if line_no % 19 == 4:
            out.append("# comment between the data lines\n")
        if line_no % 23 == 6:
            out.append("1 2 3 # a data line with a hash is dropped whole\n")
        if bad_at is not None and line_no == bad_at:
            out.append("4  5 6\n")            # interior empty token: std::stoi throws
        line_no += 1
    text = "".join(out)
    return text if tail_newline else text + "7 8 9"


@needs_ref
@pytest.mark.parametrize("case", ["clean", "no_final_newline", "bad_token_early", "bad_token_late"])
def test_p3_reader_on_a_file_split_over_host_threads(tmp_path, case):
    """the multi-threaded reader / echo writer must behave like the reference's sequential ones, exceptions included"""
    ref = orc.Reference()
    text = {"clean": lambda: _big_ppm(1), "no_final_newline": lambda: _big_ppm(2, tail_newline=False),
            "bad_token_early": lambda: _big_ppm(3, bad_at=40), "bad_token_late": lambda: _big_ppm(4, bad_at=20000)}[case]()
    assert len(text) > (1 << 20)
    src = tmp_path / "in.ppm"
    src.write_text(text)
    a = run(ref.encode_exe, str(src), str(tmp_path / "ref.ppm"))
    b = run(ENC, str(src), str(tmp_path / "mine.ppm"))
    assert a.returncode == b.returncode
    assert strip_times(a.stdout) == strip_times(b.stdout) and a.stderr == b.stderr
    if a.returncode == 0:
        assert (tmp_path / "ref.ppm").read_bytes() == (tmp_path / "mine.ppm").read_bytes()


def test_decode_io_writer_host_only(tmp_path):
    """the decoder CLI's P3 writer, single- and multi-threaded, against a plain iostream rendering (no device needed)"""
    exe = tmp_path / "decode_io_check"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           "-o", str(exe), os.path.join(ROOT, "tests", "host", "decode_io_check.cpp")])
    p = run(str(exe))
    assert p.returncode == 0, p.stdout
    assert p.stdout.count(" ok ") == 4


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    src = tmp_path / "in.ppm"
    src.write_text(PPM_CASES["plain"])
    p = run(ENC, str(src), str(tmp_path / "o.jpg"))
    assert p.returncode == 1 and "no CUDA device" in p.stderr


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("W,H,family,gray", [(64, 48, 0, False), (37, 21, 1, False), (200, 120, 2, False), (129, 65, 0, True)])
def test_cli_round_trip_matches_reference(tmp_path, oracle, W, H, family, gray):
    r, g, b = J.synth.image(family, W, H)
    src = tmp_path / "in.ppm"
    write_ppm(src, r, g, b, header="P3\n# made by test\n%d %d\n255\n")
    flags = ["--gray"] if gray else []
    p = run(ENC, str(src), str(tmp_path / "m.jpg"), *flags)
    assert p.returncode == 0, p.stderr
    f = (tmp_path / "m.jpg").read_bytes()
    assert f == oracle.encode(r, g, b, W, H, gray=gray)
    assert ("Output size: %d %s" % (len(f), "srook::byte" if gray else "byte")) in p.stdout
    q = run(DEC, str(tmp_path / "m.jpg"), str(tmp_path / "m.ppm"), *flags)
    assert q.returncode == 0, q.stderr
    assert ("Decoded image: Netpbm image data, size = %d x %d, pixmap, ASCII text" % (W, H)) in q.stdout
    lines = (tmp_path / "m.ppm").read_text().split("\n")
    assert lines[:4] == ["P3", "# Decoded by jpezy", "%d %d" % (W, H), "255"]
    vals = np.array(" ".join(lines[4:]).split(), dtype=np.int64).reshape(-1, 3)
    _, _, R, G_, B = oracle.decode(f, gray=gray)
    assert (vals[:, 0] == R[: W * H]).all() and (vals[:, 1] == G_[: W * H]).all() and (vals[:, 2] == B[: W * H]).all()
    if orc.ref_dir() is not None:           # same console lines and files as the reference's own CLIs
        ref = orc.Reference()
        pr = run(ref.encode_exe, str(src), str(tmp_path / "r.jpg"), *flags)
        assert strip_times(pr.stdout) == strip_times(p.stdout) and (tmp_path / "r.jpg").read_bytes() == f
        for extra in ([], ["-v"]):
            qr = run(ref.decode_exe, str(tmp_path / "r.jpg"), str(tmp_path / "r.ppm"), *flags, *extra)
            qm = run(DEC, str(tmp_path / "m.jpg"), str(tmp_path / "m2.ppm"), *flags, *extra)
            assert strip_times(qr.stdout) == strip_times(qm.stdout)
            assert (tmp_path / "r.ppm").read_text() == (tmp_path / "m2.ppm").read_text()


@pytest.mark.gpu
def test_cli_decode_failure_paths(tmp_path, oracle):
    r, g, b = J.synth.image(1, 64, 48)
    f = oracle.encode(r, g, b, 64, 48)
    bad = tmp_path / "cut.jpg"
    bad.write_bytes(f[: 644 + (len(f) - 646) // 3])
    p = run(DEC, str(bad), str(tmp_path / "o.ppm"))
    assert p.returncode == 1 and "decode failed" in p.stderr
    junk = tmp_path / "junk.jpg"
    junk.write_bytes(b"\xff\xd8\xff\xd9")          # EOI before SOS: analyze_header throws, decode() returns empty
    p = run(DEC, str(junk), str(tmp_path / "o.ppm"))
    assert p.returncode == 1 and "decode failed" in p.stderr
