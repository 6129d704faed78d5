"""-m gpu: the pipelined host-buffer batch entry points (jpezyb200_encode_batch / jpezyb200_decode_batch) must give exactly
what the per-image calls give, for one group and for many (the three-stream pipeline with double buffers)."""
import numpy as np
import pytest
import torch

import jpezy_b200 as J
from jpezy_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group_bytes", [96 << 20, 1, 3 * 200 * 120 * 2])        # one group / one image per group / two per group
@pytest.mark.parametrize("gray", [False, True])
def test_batch_equals_per_image(ctx, oracle, group_bytes, gray):
    W, H, N = 200, 120, 7
    imgs = [J.synth.image(k % 3, W, H, frame=k) for k in range(N)]
    r, g, b = (np.ascontiguousarray(np.stack([im[c] for im in imgs])) for c in range(3))
    slot = max(W * H * 3, 10240)
    scans = np.zeros((N, slot), dtype=np.uint8)
    nbytes = np.zeros(N, dtype=np.uint64)
    try:
        ctx.set_option(capi.OPT_BATCH_GROUP_BYTES, group_bytes)
        ctx.encode_batch(r, g, b, W, H, N, gray, scans, slot, nbytes)
        frame = J.default_frame(W, H)
        pl = capi.plane_bytes(frame)
        R, G, B = (np.full((N, pl), 0x55, dtype=np.uint8) for _ in range(3))
        st = np.full(N, -1, dtype=np.int32)
        ctx.decode_batch(scans, slot, nbytes, N, frame, gray, R, G, B, pl, st)
    finally:
        ctx.set_option(capi.OPT_BATCH_GROUP_BYTES, 96 << 20)
    assert (st == 0).all()
    for k, im in enumerate(imgs):
        want, _ = ctx.encode(im[0], im[1], im[2], W, H, gray=gray)
        assert scans[k, : int(nbytes[k])].tobytes() == want
        if k < 3:
            assert want == oracle.encode(im[0], im[1], im[2], W, H, gray=gray, scan_only=True)
        r0, g0, b0 = ctx.decode(want, frame, gray=gray)
        assert (R[k] == r0).all() and (G[k] == g0).all() and (B[k] == b0).all()


def test_batch_with_pinned_buffers_and_small_slots(ctx):
    W, H, N = 320, 192, 12
    npx = W * H
    r, g, b = (torch.empty((N, H, W), dtype=torch.uint8).pin_memory() for _ in range(3))
    for k in range(N):
        im = J.synth.image(1, W, H, frame=k)         # noise: ~2.5 bit/px
        for t, a in zip((r, g, b), im):
            t[k].copy_(torch.from_numpy(a))
    slot = 4096                                       # far too small: every image must report UINT64_MAX
    scans = torch.zeros((N, slot), dtype=torch.uint8).pin_memory()
    nbytes = np.zeros(N, dtype=np.uint64)
    with pytest.raises(J.JpezyError) as e:
        ctx.encode_batch(r, g, b, W, H, N, False, scans, slot, nbytes)
    assert e.value.code == capi.ECAPACITY and (nbytes == np.iinfo(np.uint64).max).all()
    slot = npx
    scans = torch.zeros((N, slot), dtype=torch.uint8).pin_memory()
    ctx.set_option(capi.OPT_BATCH_GROUP_BYTES, 3 * npx * 5)
    try:
        ctx.encode_batch(r, g, b, W, H, N, False, scans, slot, nbytes)
    finally:
        ctx.set_option(capi.OPT_BATCH_GROUP_BYTES, 96 << 20)
    for k in (0, 5, 11):
        want, _ = ctx.encode(r[k].numpy(), g[k].numpy(), b[k].numpy(), W, H)
        assert scans[k, : int(nbytes[k])].numpy().tobytes() == want


def test_device_chain_with_read_sizes(ctx, oracle):
    """encode_batch_dev -> read_sizes (the one host round trip) -> decode_batch_dev: the fused device-resident round trip of
    bench.py; segments and decoded planes must equal the per-image host calls."""
    W, H, N = 176, 144, 5
    imgs = [J.synth.image(k % 3, W, H, frame=10 + k) for k in range(N)]
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    try:
        d_in = torch.from_numpy(np.stack([np.stack([im[c] for im in imgs]) for c in range(3)])).cuda()       # [3][N][H][W]
        slot = W * H * 3
        d_scan = torch.zeros((N, slot), dtype=torch.uint8, device="cuda")
        d_nb = torch.zeros(N, dtype=torch.int64, device="cuda")
        frame = J.default_frame(W, H)
        pl = capi.plane_bytes(frame)
        d_out = torch.full((3, N, pl), 0x55, dtype=torch.uint8, device="cuda")
        d_st = torch.full((N,), -1, dtype=torch.int32, device="cuda")
        ctx.encode_batch_dev(d_in[0], d_in[1], d_in[2], W, H, N, False, d_scan, slot, d_nb, None, stream=st.cuda_stream)
        nb = np.zeros(N, dtype=np.uint64)
        ctx.read_sizes(d_nb, N, nb, stream=st.cuda_stream)
        assert (nb == d_nb.cpu().numpy().astype(np.uint64)).all()
        ctx.decode_batch_dev(d_scan, slot, nb, N, frame, False, d_out[0], d_out[1], d_out[2], pl, d_st, stream=st.cuda_stream)
        torch.cuda.synchronize()
        assert (d_st.cpu().numpy() == 0).all()
        scans, out = d_scan.cpu().numpy(), d_out.cpu().numpy()
        for k, im in enumerate(imgs):
            want, _ = ctx.encode(im[0], im[1], im[2], W, H)
            assert scans[k, : int(nb[k])].tobytes() == want
            if k == 0:
                assert want == oracle.encode(im[0], im[1], im[2], W, H, scan_only=True)
            r0, g0, b0 = ctx.decode(want, frame)
            assert (out[0, k] == r0).all() and (out[1, k] == g0).all() and (out[2, k] == b0).all()
    finally:
        torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.parametrize("nimg", [1, 11])
def test_device_resident_lengths_equal_host_lengths(ctx, nimg):
    """jpezyb200_decode_batch_dev2 (segment lengths read from device memory: the encoder's own d_scan_bytes output) must give
    exactly what jpezyb200_decode_batch_dev gives with the lengths in a host array; a segment longer than the launch was
    sized for is refused with JPEZYB200_ECAPACITY, the others of the batch still decode"""
    W, H = 208, 112
    frame = J.default_frame(W, H)
    pl = capi.plane_bytes(frame)
    slot = W * H * 3
    d_in = torch.empty((3, nimg, H, W), dtype=torch.uint8, device="cuda")
    for k in range(nimg):
        im = J.synth.image(k % 2, W, H, frame=k)        # photo / noise alternate: very different segment lengths
        for c in range(3):
            d_in[c, k].copy_(torch.from_numpy(im[c]))
    d_scan = torch.zeros((nimg, slot), dtype=torch.uint8, device="cuda")
    d_nb = torch.zeros(nimg, dtype=torch.int64, device="cuda")
    ctx.encode_batch_dev(d_in[0], d_in[1], d_in[2], W, H, nimg, False, d_scan, slot, d_nb, None)
    torch.cuda.synchronize()
    h_nb = d_nb.cpu().numpy().astype(np.uint64)
    out_h = torch.zeros((3, nimg, pl), dtype=torch.uint8, device="cuda")
    out_d = torch.full((3, nimg, pl), 0x33, dtype=torch.uint8, device="cuda")
    st_h = torch.full((nimg,), -1, dtype=torch.int32, device="cuda")
    st_d = torch.full((nimg,), -1, dtype=torch.int32, device="cuda")
    ctx.decode_batch_dev(d_scan, slot, h_nb, nimg, frame, False, out_h[0], out_h[1], out_h[2], pl, st_h)
    ctx.decode_batch_dev2(d_scan, slot, d_nb, int(h_nb.max()), nimg, frame, False, out_d[0], out_d[1], out_d[2], pl, st_d)
    torch.cuda.synchronize()
    assert (st_h == 0).all() and (st_d == 0).all()
    assert torch.equal(out_h, out_d)
    if nimg > 1:
        # sized for the shortest segment + 1 byte: only the images that fit are decoded
        cap = int(h_nb.min()) + 1
        st_c = torch.full((nimg,), -1, dtype=torch.int32, device="cuda")
        ctx.decode_batch_dev2(d_scan, slot, d_nb, cap, nimg, frame, False, out_d[0], out_d[1], out_d[2], pl, st_c)
        torch.cuda.synchronize()
        st = st_c.cpu().numpy()
        fits = h_nb <= cap
        assert fits.any() and (~fits).any()
        assert (st[fits] == 0).all() and (st[~fits] == capi.ECAPACITY).all()
        for k in np.nonzero(fits)[0]:
            assert torch.equal(out_h[:, k], out_d[:, k])
