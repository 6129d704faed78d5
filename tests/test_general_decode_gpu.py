"""-m gpu: baseline JPEGs that jpezy's own encoder never writes but its decoder accepts: other sampling
factors, one component, quality-scaled quantisation tables, component ids 1..3 -- produced here with Pillow, decoded through
the drop-in CLI (host marker parser of include/jpezy/jpezy_decoder.hpp + jpezyb200_decode) and compared with the oracle
restatement of the reference decoder (which equals the reference's own decoder on these files, tests/test_oracle_vs_ref.py
and the live check below).  Bar: identical samples."""
import io
import os
import subprocess

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEC = os.path.join(ROOT, "jpezy_b200", "bin", "jpezy_decode")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "jpezy_b200", "cli"), "-s"])


def picture(W, H, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([(x * 3 + y) % 256, (x + 2 * y) % 256, (x * y // 7) % 256], -1).astype(np.int64)
    img = img // 2 + rng.integers(0, 128, (H, W, 3))
    img[: H // 3, : W // 2] = [200, 30, 90]          # a flat patch: DC-only blocks, values on rounding boundaries
    return img.clip(0, 255).astype(np.uint8)


def jpeg_bytes(img, mode, **kw):
    buf = io.BytesIO()
    im = PIL.fromarray(img if mode == "RGB" else img[..., 0], mode)
    im.save(buf, "JPEG", **kw)
    return buf.getvalue()


CASES = [("444", "RGB", dict(quality=75, subsampling=0)), ("422", "RGB", dict(quality=75, subsampling=1)),
         ("420", "RGB", dict(quality=75, subsampling=2)), ("420_q95", "RGB", dict(quality=95, subsampling=2)),
         ("444_q30", "RGB", dict(quality=30, subsampling=0)), ("gray", "L", dict(quality=80)),
         ("gray_q100", "L", dict(quality=100)), ("422_optimized_tables", "RGB", dict(quality=60, subsampling=1, optimize=True)),
         # Cb and Cr with DIFFERENT quantisation tables (Tq 1 and 2) on the 2x2 / 1x1 / 1x1 layout of the production inverse kernel
         ("420_three_qtables", "RGB", dict(subsampling=2, qtables=[[8 + (i % 8) + 2 * (i // 8) for i in range(64)],
                                                                  [11 + 3 * (i % 8) + (i // 8) for i in range(64)],
                                                                  [29 + (i % 8) + 5 * (i // 8) for i in range(64)]]))]


@pytest.mark.parametrize("name,mode,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("W,H", [(70, 45), (16, 16), (8, 8), (129, 257), (640, 360)])
@pytest.mark.parametrize("gray_out", [False, True])
def test_general_baseline_decode_matches_reference_decoder(tmp_path, oracle, name, mode, kw, W, H, gray_out):
    if gray_out and (W, H) != (70, 45):
        pytest.skip("--gray output checked on one size")
    f = jpeg_bytes(picture(W, H, 7), mode, **kw)
    jpg, out = tmp_path / "in.jpg", tmp_path / "out.ppm"
    jpg.write_bytes(f)
    Wd, Hd, R0, G0, B0 = oracle.decode(f, gray=gray_out)
    p = subprocess.run([DEC, str(jpg), str(out)] + (["--gray"] if gray_out else []), capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = out.read_text().split("\n")
    assert lines[2] == "%d %d" % (W, H)
    vals = np.array(" ".join(lines[4:]).split(), dtype=np.int64).reshape(-1, 3)
    d = [np.abs(vals[:, k] - a[: W * H]) for k, a in enumerate((R0, G0, B0))]
    assert max(int(x.max()) for x in d) <= 1, "more than 1 away from the reference decoder"
    assert sum(int((x != 0).sum()) for x in d) == 0
    if orc.ref_dir() is not None and (W, H) == (70, 45):        # the reference's own CLI writes the same PPM
        ref = orc.Reference()
        q = subprocess.run([ref.decode_exe, str(jpg), str(tmp_path / "ref.ppm")] + (["--gray"] if gray_out else []), capture_output=True, text=True)
        assert q.returncode == 0 and (tmp_path / "ref.ppm").read_text() == out.read_text()


@pytest.mark.parametrize("name,mode,kw", [c for c in CASES if c[0] in ("444", "422", "420_q95", "gray")], ids=["444", "422", "420_q95", "gray"])
def test_general_baseline_decode_one_chain_per_guess(tmp_path, oracle, name, mode, kw):
    """the multi-guess first synchronisation launch (JPEZY_B200_DEC_HYP=1 = JPEZYB200_OPT_SYNC_GUESSES of every new context) on
    MCUs of 3, 4 and 6 blocks (one block: it falls back to the single guess)"""
    W, H = 640, 360
    f = jpeg_bytes(picture(W, H, 11), mode, **kw)
    jpg, out = tmp_path / "in.jpg", tmp_path / "out.ppm"
    jpg.write_bytes(f)
    Wd, Hd, R0, G0, B0 = oracle.decode(f)
    p = subprocess.run([DEC, str(jpg), str(out)], capture_output=True, text=True, env=dict(os.environ, JPEZY_B200_DEC_HYP="1"))
    assert p.returncode == 0, p.stderr
    vals = np.array(" ".join(out.read_text().split("\n")[4:]).split(), dtype=np.int64).reshape(-1, 3)
    assert (vals[:, 0] == R0[: W * H]).all() and (vals[:, 1] == G0[: W * H]).all() and (vals[:, 2] == B0[: W * H]).all()


def test_decode_again_when_the_enqueued_synchronisation_launches_were_not_enough(tmp_path, oracle):
    """a high-quality file needs more self-synchronisation launches than a context that enqueues a single one provides: the device
    reports JPEZYB200_EAGAIN and jpezyb200_decode decodes again with the host-polled loop.  (The writing pass of the first attempt
    runs from wrong states and trips over impossible runs; that must not turn "again" into "corrupt".)"""
    W, H = 640, 360
    f = jpeg_bytes(picture(W, H, 7), "L", quality=100)
    jpg, out = tmp_path / "in.jpg", tmp_path / "out.ppm"
    jpg.write_bytes(f)
    Wd, Hd, R0, G0, B0 = oracle.decode(f)
    env = dict(os.environ, JPEZY_B200_SYNC_ROUNDS="1")
    p = subprocess.run([DEC, str(jpg), str(out)], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    vals = np.array(" ".join(out.read_text().split("\n")[4:]).split(), dtype=np.int64).reshape(-1, 3)
    assert (vals[:, 0] == R0[: W * H]).all() and (vals[:, 1] == G0[: W * H]).all() and (vals[:, 2] == B0[: W * H]).all()


DRI_CASES = [("420_dri2", "RGB", dict(quality=75, subsampling=2, restart_marker_blocks=2)),
             ("444_dri_row", "RGB", dict(quality=80, subsampling=0, restart_marker_rows=1)),
             ("422_dri7", "RGB", dict(quality=50, subsampling=1, restart_marker_blocks=7)),
             ("gray_dri5", "L", dict(quality=90, restart_marker_blocks=5)),
             ("420_dri_larger_than_image", "RGB", dict(quality=75, subsampling=2, restart_marker_blocks=60000))]


@pytest.mark.parametrize("name,mode,kw", DRI_CASES, ids=[c[0] for c in DRI_CASES])
@pytest.mark.parametrize("W,H", [(70, 45), (16, 16), (333, 200)])
def test_restart_intervals(tmp_path, oracle, name, mode, kw, W, H):
    """DRI files (src/decoder/jpezy_decoder.hpp:152-163,400-404): segments located while un-stuffing, one thread per segment"""
    f = jpeg_bytes(picture(W, H, 11), mode, **kw)
    assert (b"\xff\xdd" in f) == ("larger" not in name or True)
    jpg, out = tmp_path / "in.jpg", tmp_path / "out.ppm"
    jpg.write_bytes(f)
    _, _, R0, G0, B0 = oracle.decode(f)
    p = subprocess.run([DEC, str(jpg), str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    vals = np.array(" ".join(out.read_text().split("\n")[4:]).split(), dtype=np.int64).reshape(-1, 3)
    assert all((vals[:, k] == a[: W * H]).all() for k, a in enumerate((R0, G0, B0)))


def test_truncated_restart_stream_fails_cleanly(tmp_path):
    f = jpeg_bytes(picture(64, 48, 3), "RGB", quality=75, subsampling=2, restart_marker_blocks=2)
    jpg = tmp_path / "dri.jpg"
    jpg.write_bytes(f[: len(f) * 2 // 3])
    p = subprocess.run([DEC, str(jpg), str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert p.returncode == 1 and "decode failed" in p.stderr
