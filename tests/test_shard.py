"""One image on several GPUs, sharded by MCU rows (jpezy_b200/shard.py, jpezyb200_shard_encode_a..d).

CPU: the host-side rules (row partition, byte ownership, shared-byte completion, stuffing after alignment) against the
oracle's whole-image stream, single process and over a world_size-2 gloo group.
GPU (-m gpu): N emulated ranks on one device produce the byte-identical stream of the single-GPU encoder.
"""
import os
import sys

import numpy as np
import pytest

import jpezy_b200 as J
from jpezy_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unstuffed_bits(scan, nbits):
    raw = scan.replace(b"\xff\x00", b"\xff")
    return "".join("{:08b}".format(x) for x in raw)[:nbits]


def local_bitstrings(oracle, r, g, b, W, H, nranks, gray=False):
    """each rank's un-stuffed local bit string, cut out of the oracle's whole-image stream at the MCU-row boundaries
    (the bit position of a boundary = bit count of the stream of the MCUs before it: predictors chain from the start)"""
    import ctypes as C
    coefs = oracle.coefs(r, g, b, W, H, gray=gray)
    HU = (W + 15) // 16

    def nbits_of(nmcu):
        if nmcu == 0:
            return 0
        c = np.ascontiguousarray(coefs[:nmcu])
        out = np.zeros(c.size * 4 + 1024, dtype=np.uint8)
        n, nb = C.c_size_t(0), C.c_uint64(0)
        assert oracle.lib.orc_scan_from_coefs(c.ctypes.data_as(C.POINTER(C.c_int16)), nmcu, 1, out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              out.size, C.byref(n), C.byref(nb)) == 0
        return int(nb.value)
    whole = oracle.encode(r, g, b, W, H, gray=gray, scan_only=True)
    parts = shard.partition_mcu_rows((H + 15) // 16, nranks)
    cuts = [nbits_of(row0 * HU) for row0, _ in parts] + [nbits_of(coefs.shape[0])]
    bits = unstuffed_bits(whole, cuts[-1])
    return [bits[cuts[k]: cuts[k + 1]] for k in range(nranks)], whole


def test_partition_and_pixel_rows():
    assert shard.partition_mcu_rows(68, 8) == [(0, 9), (9, 9), (18, 9), (27, 9), (36, 8), (44, 8), (52, 8), (60, 8)]
    assert shard.partition_mcu_rows(2048, 8)[-1] == (1792, 256)
    assert shard.pixel_rows(1080, 60, 8) == (960, 120)          # the last shard of a 1080-row image: rows 960..1079
    with pytest.raises(ValueError):
        shard.partition_mcu_rows(3, 4)
    lay = shard.stream_layout([13, 8, 24])
    assert [g["first_own"] for g in lay] == [0, 2, 3] and [g["nown"] for g in lay] == [2, 1, 3] and [g["d"] for g in lay] == [0, 3, 3]


@pytest.mark.parametrize("W,H,family,nranks,gray", [(64, 48, 0, 2, False), (200, 120, 1, 3, False), (136, 72, 2, 4, False),
                                                   (48, 130, 1, 8, False), (333, 77, 0, 2, True), (64, 64, 1, 4, False)])
def test_host_model_reproduces_whole_image_stream(oracle, W, H, family, nranks, gray):
    r, g, b = J.synth.image(family, W, H)
    local, whole = local_bitstrings(oracle, r, g, b, W, H, nranks, gray)
    assert shard.stitch_bitstrings(local) == whole
    whole0 = oracle.encode(r, g, b, W, H, gray=gray, pad_ones=False, scan_only=True)
    assert shard.stitch_bitstrings(local, pad_ones=False) == whole0


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle as orc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grp = shard.DistGroup(dist, "cpu")
        o = orc.Oracle()
        W, H = 120, 100
        r, g, b = J.synth.image(1, W, H)
        local, whole = local_bitstrings(o, r, g, b, W, H, world)
        mine = local[rank]                      # what this rank's device phase b would have produced
        info = torch.tensor([len(mine), int((mine + "0" * 8)[:8], 2)], dtype=torch.int64)
        all_info = torch.zeros((world, 2), dtype=torch.int64)
        grp.all_gather(all_info, info)          # all-gather #2
        lay = shard.stream_layout([int(x) for x in all_info[:, 0]])[rank]
        head = "1" * 8 if rank + 1 == world else "{:08b}".format(int(all_info[rank + 1, 1]))
        virt = mine + head
        piece = bytearray()
        for i in range(lay["nown"]):
            byte = int(virt[8 * i + lay["d"]: 8 * i + lay["d"] + 8].ljust(8, "0"), 2)
            piece.append(byte)
            if byte == 0xFF:
                piece.append(0)
        nb = torch.tensor([len(piece)], dtype=torch.int64)
        all_nb = torch.zeros(world, dtype=torch.int64)
        grp.all_gather(all_nb, nb)              # all-gather #3
        base = int(all_nb[:rank].sum())
        pieces = [None] * world
        dist.all_gather_object(pieces, (base, bytes(piece)))
        if rank == 0:
            out = bytearray(int(all_nb.sum()))
            for bs, p in pieces:
                out[bs: bs + len(p)] = p
            q.put(bytes(out) == whole and grp.broadcast_object("x") == "x")
        else:
            grp.broadcast_object(None)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_group_stitches_the_reference_stream():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(k, 2, port, q)) for k in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("W,H,family,nranks,gray", [(64, 48, 0, 2, False), (200, 120, 1, 3, False), (136, 72, 2, 4, False),
                                                   (48, 130, 1, 8, False), (333, 77, 0, 2, True), (640, 360, 1, 8, False),
                                                   (1920, 1080, 0, 8, False), (512, 512, 1, 1, False)])
def test_sharded_encode_is_byte_identical_to_single_gpu(ctx, oracle, W, H, family, nranks, gray):
    import torch
    r, g, b = J.synth.image(family, W, H)
    want, _ = ctx.encode(r, g, b, W, H, gray=gray)
    ctxs = [J.Context(0) for _ in range(nranks)]
    try:
        planes = tuple(torch.from_numpy(x).cuda() for x in (r, g, b))
        got, bits = shard.encode_sharded_local(ctxs, planes, W, H, gray=gray)
    finally:
        for c in ctxs:
            c.close()
    assert got == want
    if W * H <= 640 * 360:
        assert got == oracle.encode(r, g, b, W, H, gray=gray, scan_only=True)
    assert len(bits) == nranks and all(t >= 24 for t in bits)


@pytest.mark.gpu
def test_sharded_encode_pad_zero_and_overflow(ctx, oracle):
    import torch
    from jpezy_b200 import capi
    W, H = 96, 80
    r, g, b = J.synth.image(1, W, H)
    ctxs = [J.Context(0) for _ in range(3)]
    try:
        for c in ctxs:
            c.set_option(capi.OPT_PAD_ONES, 0)
        planes = tuple(torch.from_numpy(x).cuda() for x in (r, g, b))
        got, _ = shard.encode_sharded_local(ctxs, planes, W, H)
        assert got == oracle.encode(r, g, b, W, H, pad_ones=False, scan_only=True)
        with pytest.raises(RuntimeError):
            shard.encode_sharded_local(ctxs, planes, W, H, dst_cap=100)
        # ONE rank runs out of its own scratch: every rank must see it (no silently short segment on rank 0)
        ctxs[1].set_option(capi.OPT_SHARD_SCRATCH_BYTES, 64)
        flags = shard.encode_sharded_local(ctxs, planes, W, H, return_overflow_flags=True)
        assert flags == [1, 1, 1]
        ctxs[1].set_option(capi.OPT_SHARD_SCRATCH_BYTES, 0)
        got, _ = shard.encode_sharded_local(ctxs, planes, W, H)
        assert got == oracle.encode(r, g, b, W, H, pad_ones=False, scan_only=True)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,family,nranks,gray", [(64, 48, 0, 2, False), (200, 120, 1, 3, False), (333, 77, 0, 2, True),
                                                   (640, 360, 1, 8, False), (40, 25, 2, 2, False), (1024, 1024, 0, 8, True)])
def test_sharded_decode_matches_single_gpu(ctx, oracle, W, H, family, nranks, gray):
    """transform stage sharded by MCU rows (entropy stage replicated): same planes as the whole-image decode, tail cleared"""
    r, g, b = J.synth.image(family, W, H)
    scan, _ = ctx.encode(r, g, b, W, H, gray=gray)
    frame = J.default_frame(W, H)
    want = ctx.decode(scan, frame, gray=gray)
    ctxs = [J.Context(0) for _ in range(nranks)]
    try:
        got = shard.decode_sharded_local(ctxs, scan, frame, gray=gray)
    finally:
        for c in ctxs:
            c.close()
    assert all((a == b_).all() for a, b_ in zip(got, want))
