"""One image, several GPUs: the encoder sharded by MCU rows (BASELINE.json config 5).

Host-side orchestration of the four device phases declared in include/jpezy_b200.h (jpezyb200_shard_encode_a..d).
The reference is single threaded; what is preserved is its output: ONE restart-less entropy-coded segment
(src/encoder/jpezy_encoder.hpp:58-67), byte-identical to the single-GPU stream.  Between the phases three tiny
all-gathers carry the two image-wide prefix dependencies across the shards:

    #1  last quantised DCs (3 x int32)        the running predictors pre_DC[3] (src/encoder/jpezy_encoder.hpp:180-181)
    #2  {local bit count, first 8 bits}       the bit cursor of the single segment + the bits that complete a shared byte
    #3  owned + stuffed byte count            FF 00 stuffing depends on the global byte alignment, so it comes last

The collectives are supplied by a `group` object:
    DistGroup   one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo in the CPU tests)
    LocalGroup  all ranks inside one process on one device, in lockstep (tests; validating the N-rank byte stream on one GPU)
The stitched stream is written by every rank directly into one destination buffer; with DistGroup that buffer lives on
rank 0 and is mapped into the other processes through CUDA IPC, so phase d's stores travel over NVLink.
"""
import numpy as np


def partition_mcu_rows(mcu_rows_total, nranks):
    """Contiguous ranges of MCU rows, as equal as possible (the first `rem` ranks get one more): [(row0, nrows)] * nranks.
    Every rank must own at least one MCU row: a byte of the stream may then be shared by at most two ranks."""
    if nranks < 1 or mcu_rows_total < nranks:
        raise ValueError("need at least one MCU row per rank (%d rows, %d ranks)" % (mcu_rows_total, nranks))
    q, rem = divmod(mcu_rows_total, nranks)
    out, r0 = [], 0
    for k in range(nranks):
        n = q + (1 if k < rem else 0)
        out.append((r0, n))
        r0 += n
    return out


def pixel_rows(H, row0, nrows):
    """image rows a shard must hold: [16*row0, min(H, 16*(row0+nrows))) -- rows past H are edge replicated by the kernel"""
    y0 = 16 * row0
    return y0, min(H, 16 * (row0 + nrows)) - y0


def stream_layout(bit_counts):
    """Host restatement of the ownership rule of phase c (k_shard_geom), used by the CPU tests and for reporting:
    rank k owns the globally aligned bytes [ceil(B_k/8), ceil(B_{k+1}/8)), B_k = sum of the previous ranks' bit counts.
    -> list of dicts {bit_base, first_own, nown, d} (d = local bit offset of owned byte 0)"""
    out, base = [], 0
    for t in bit_counts:
        first = (base + 7) // 8
        out.append({"bit_base": base, "first_own": first, "nown": (base + t + 7) // 8 - first, "d": first * 8 - base})
        base += t
    return out


def stitch_bitstrings(local_bits, pad_ones=True):
    """Reference model of phases c+d on the host: `local_bits` = each rank's un-stuffed local bit string ('0'/'1'), already
    coded with the right DC predictors.  Every rank extracts its owned aligned bytes (completing a shared byte with the
    next rank's head, the last one with pad bits), stuffs them, and the pieces are concatenated.  -> bytes"""
    lay = stream_layout([len(b) for b in local_bits])
    out = bytearray()
    for k, (bits, g) in enumerate(zip(local_bits, lay)):
        head = ("1" if pad_ones else "0") * 8 if k + 1 == len(local_bits) else (local_bits[k + 1] + "0" * 8)[:8]
        virt = bits + head
        for i in range(g["nown"]):
            byte = int(virt[8 * i + g["d"]: 8 * i + g["d"] + 8].ljust(8, "0"), 2)
            out.append(byte)
            if byte == 0xFF:
                out.append(0)
    return bytes(out)


class DistGroup:
    """torch.distributed process group: one rank per process (NCCL on the GPU box, gloo in the CPU tests)"""

    def __init__(self, dist, device):
        self.dist, self.device = dist, device
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def all_gather(self, out, inp):
        self.dist.all_gather_into_tensor(out.view(-1), inp.view(-1).contiguous())

    def broadcast_object(self, obj, src=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]

    def barrier(self):
        self.dist.barrier()


class ShardedEncoder:
    """Rank-local driver: `ctx` is this process's jpezy_b200.Context, `group` a DistGroup."""

    def __init__(self, ctx, group, dst_cap):
        import torch
        self.torch, self.ctx, self.group, self.dst_cap = torch, ctx, group, int(dst_cap)
        dev = group.device
        n = group.size
        self.last_dc = torch.zeros(3, dtype=torch.int32, device=dev)
        self.all_dc = torch.zeros((n, 3), dtype=torch.int32, device=dev)
        self.zero_dc = torch.zeros(3, dtype=torch.int32, device=dev)
        self.info = torch.zeros(2, dtype=torch.int64, device=dev)
        self.all_info = torch.zeros((n, 2), dtype=torch.int64, device=dev)
        self.nbytes = torch.zeros(1, dtype=torch.int64, device=dev)
        self.all_bytes = torch.zeros(n, dtype=torch.int64, device=dev)
        self.total = torch.zeros(1, dtype=torch.int64, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.phase_events = None
        # the stitched stream lives on rank 0; the other ranks map it (CUDA IPC) and store into it over NVLink
        self._owned = self._mapped = None
        if group.size == 1:
            self.dst_tensor = torch.zeros(self.dst_cap, dtype=torch.uint8, device=dev)
            self.dst = self.dst_tensor.data_ptr()
        elif group.rank == 0:
            self._owned, handle = ctx.ipc_alloc(self.dst_cap)
            self.dst = self._owned
            group.broadcast_object(handle, src=0)
        else:
            handle = group.broadcast_object(None, src=0)
            self._mapped = ctx.ipc_open(handle)
            self.dst = self._mapped

    def encode(self, d_r, d_g, d_b, W, H, row0, nrows, y_origin, gray=False, stream=None):
        """all ranks call this together; planes hold image rows y_origin.. of this rank's shard.  `stream` must be the raw handle
        of torch's CURRENT stream and not the legacy default stream (0 / None selects the context's private stream, which
        the collectives of torch.distributed are not ordered with)"""
        if not stream:
            raise ValueError("ShardedEncoder.encode needs torch's current non-default stream (torch.cuda.Stream)")
        c, g = self.ctx, self.group
        ev = self.phase_events          # None, or 8 CUDA events bracketing the seven phases (bench.py: which phase limits the scaling)
        mark = (lambda k: ev[k].record()) if ev is not None else (lambda k: None)
        mark(0)
        c.shard_encode_a(d_r, d_g, d_b, W, H, row0, nrows, y_origin, gray, self.last_dc, stream=stream)
        mark(1)
        g.all_gather(self.all_dc, self.last_dc)
        mark(2)
        dc_init = self.all_dc[g.rank - 1] if g.rank > 0 else self.zero_dc
        c.shard_encode_b(dc_init, self.info, stream=stream)
        mark(3)
        g.all_gather(self.all_info, self.info)
        mark(4)
        c.shard_encode_c(self.all_info, g.rank, g.size, self.nbytes, stream=stream)
        mark(5)
        g.all_gather(self.all_bytes, self.nbytes)
        mark(6)
        c.shard_encode_d(self.all_bytes, self.dst, self.dst_cap, self.total, self.overflow, stream=stream)
        mark(7)

    PHASES = ("a_transform_lastdc", "allgather1_dc", "b_bits_scatter", "allgather2_bits", "c_geometry_ffcount", "allgather3_bytes", "d_stuff_push")

    def phase_ms(self):
        """after a synchronise: duration of the seven phases of the last encode() made with phase_events set"""
        ev = self.phase_events
        return {name: ev[k].elapsed_time(ev[k + 1]) for k, name in enumerate(self.PHASES)}

    def result(self):
        """rank 0, after a barrier: (bytes of the stitched segment, per-rank bit counts).  Raises on overflow."""
        self.torch.cuda.synchronize()
        self.group.barrier()
        self.torch.cuda.synchronize()
        if int(self.overflow.item()):
            raise RuntimeError("stitched stream does not fit dst_cap=%d (or a rank's scratch)" % self.dst_cap)
        n = int(self.total.item())
        bits = [int(x) for x in self.all_info[:, 0].cpu().tolist()]
        if self.group.rank != 0:
            return None, bits
        if self.group.size == 1:
            return self.dst_tensor[:n].cpu().numpy().tobytes(), bits
        return self._wrap(self.dst, n).cpu().numpy().tobytes(), bits

    def _wrap(self, ptr, n):
        """view n bytes at a raw device pointer as a torch uint8 tensor (no copy)"""
        torch = self.torch

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(_Arr(), device=self.group.device)

    def close(self):
        if self._mapped:
            self.ctx.ipc_close(self._mapped)
            self._mapped = None
        self.group.barrier()
        if self._owned:
            self.ctx.ipc_free(self._owned)
            self._owned = None


def encode_sharded_local(ctxs, planes, W, H, gray=False, dst_cap=None, device="cuda", return_overflow_flags=False):
    """All ranks inside one process on one device, in lockstep (LocalGroup): validates the N-rank byte stream where only one
    GPU is available.  ctxs: one Context per emulated rank; planes: (r, g, b) uint8 tensors [H, W] on `device`.
    -> (segment bytes, per-rank bit counts)"""
    import torch
    n = len(ctxs)
    parts = partition_mcu_rows((H + 15) // 16, n)
    dst_cap = int(dst_cap or max(W * H * 3, 10240))
    dst = torch.zeros(dst_cap, dtype=torch.uint8, device=device)
    last = torch.zeros((n, 3), dtype=torch.int32, device=device)
    zero = torch.zeros(3, dtype=torch.int32, device=device)
    info = torch.zeros((n, 2), dtype=torch.int64, device=device)
    nb = torch.zeros(n, dtype=torch.int64, device=device)
    total = torch.zeros(1, dtype=torch.int64, device=device)
    ovf = torch.zeros(n, dtype=torch.int32, device=device)
    # one explicit stream for all emulated ranks: a NULL stream would select every context's own private stream, and the
    # phases of different ranks must be ordered with each other
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    st = stream.cuda_stream
    shards = []
    for (row0, nrows) in parts:       # every rank holds only its own pixel rows
        y0, ny = pixel_rows(H, row0, nrows)
        shards.append(tuple(p[y0: y0 + ny].contiguous() for p in planes) + (y0,))
    for k, (row0, nrows) in enumerate(parts):
        r, g, b, y0 = shards[k]
        ctxs[k].shard_encode_a(r, g, b, W, H, row0, nrows, y0, gray, last[k], stream=st)
    for k in range(n):
        ctxs[k].shard_encode_b(last[k - 1] if k else zero, info[k], stream=st)
    for k in range(n):
        ctxs[k].shard_encode_c(info, k, n, nb[k: k + 1], stream=st)
    for k in range(n):
        ctxs[k].shard_encode_d(nb, dst, dst_cap, total, ovf[k: k + 1], stream=st)
    torch.cuda.synchronize()
    if return_overflow_flags:        # (tests: the flag of every emulated rank)
        return [int(x) for x in ovf.cpu().tolist()]
    if int(ovf.sum().item()):
        raise RuntimeError("stitched stream does not fit dst_cap=%d (or a rank's scratch)" % dst_cap)
    return dst[: int(total.item())].cpu().numpy().tobytes(), [int(x) for x in info[:, 0].cpu().tolist()]


class ShardedDecoder:
    """One image decoded by all ranks of `group` (jpezyb200_shard_decode_dev).  The entropy-coded segment is broadcast
    from rank 0 (the one real exchange: S bytes over NVLink) and decoded whole by every rank (replicas, see
    include/jpezy_b200.h); the transform stage is sharded by MCU rows and every rank stores its pixel rows straight into
    rank 0's planes, mapped through CUDA IPC."""

    def __init__(self, ctx, group, frame, scan_cap):
        import torch
        from . import capi
        self.torch, self.ctx, self.group, self.frame = torch, ctx, group, frame
        self.plane_len = capi.plane_bytes(frame)
        dev = group.device
        self.scan = torch.zeros(int(scan_cap), dtype=torch.uint8, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        hs, vs = max(frame.hs[:frame.ncomp]), max(frame.vs[:frame.ncomp])
        vu = -(-((frame.height + 7) // 8) // vs)
        self.row0, self.nrows = partition_mcu_rows(vu, group.size)[group.rank]
        self._owned = self._mapped = None
        nbytes = 3 * self.plane_len
        if group.size == 1:
            self.planes_tensor = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self.base = self.planes_tensor.data_ptr()
        elif group.rank == 0:
            self._owned, handle = ctx.ipc_alloc(nbytes)
            self.base = self._owned
            group.broadcast_object(handle, src=0)
        else:
            handle = group.broadcast_object(None, src=0)
            self._mapped = ctx.ipc_open(handle)
            self.base = self._mapped

    def decode(self, scan_bytes, gray=False, stream=None):
        """rank 0 holds the segment in self.scan[:scan_bytes]; all ranks call this together"""
        if not stream:
            raise ValueError("ShardedDecoder.decode needs torch's current non-default stream (torch.cuda.Stream)")
        g = self.group
        if g.size > 1:
            g.dist.broadcast(self.scan[:scan_bytes], src=0)
        pl = self.plane_len
        self._last = (scan_bytes, gray, stream)
        self.ctx.shard_decode_dev(self.scan, scan_bytes, self.frame, gray, self.row0, self.nrows, self.base, self.base + pl,
                                  self.base + 2 * pl, pl, self.status, stream=stream)

    def result(self):
        """rank 0, after a barrier: the three planes as numpy arrays (reference layout: stride W, padded length)"""
        from . import capi
        self.torch.cuda.synchronize()
        if int(self.status.item()) == capi.EAGAIN:
            # the enqueued synchronisation launches were not enough for this stream: this rank decodes again with the
            # host-polled loop (always terminates); the segment is already here, the other ranks are not involved
            scan_bytes, gray, stream = self._last
            pl = self.plane_len
            self.ctx.set_option(capi.OPT_SYNC_ROUNDS, 0)
            try:
                self.ctx.shard_decode_dev(self.scan, scan_bytes, self.frame, gray, self.row0, self.nrows, self.base, self.base + pl,
                                          self.base + 2 * pl, pl, self.status, stream=stream)
                self.torch.cuda.synchronize()
            finally:
                self.ctx.set_option(capi.OPT_SYNC_ROUNDS, 3)
        self.group.barrier()
        self.torch.cuda.synchronize()
        if int(self.status.item()):
            raise RuntimeError("entropy-coded segment could not be decoded (status %d)" % int(self.status.item()))
        if self.group.rank != 0:
            return None
        n = 3 * self.plane_len
        if self.group.size == 1:
            flat = self.planes_tensor.cpu().numpy()
        else:
            class _Arr:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (int(self.base), False), "version": 2}
            flat = self.torch.as_tensor(_Arr(), device=self.group.device).cpu().numpy()
        pl = self.plane_len
        return flat[:pl].copy(), flat[pl: 2 * pl].copy(), flat[2 * pl:].copy()

    def close(self):
        if self._mapped:
            self.ctx.ipc_close(self._mapped)
            self._mapped = None
        self.group.barrier()
        if self._owned:
            self.ctx.ipc_free(self._owned)
            self._owned = None


def decode_sharded_local(ctxs, scan, frame, gray=False, device="cuda"):
    """N ranks emulated on one device: every context decodes the segment and writes its MCU rows into the same planes."""
    import torch
    from . import capi
    n = len(ctxs)
    pl = capi.plane_bytes(frame)
    hs, vs = max(frame.hs[:frame.ncomp]), max(frame.vs[:frame.ncomp])
    vu = -(-((frame.height + 7) // 8) // vs)
    planes = torch.full((3, pl), 0xAA, dtype=torch.uint8, device=device)       # poisoned: every byte must be written or cleared
    d_scan = torch.from_numpy(np.frombuffer(scan, dtype=np.uint8).copy()).to(device)
    status = torch.zeros(n, dtype=torch.int32, device=device)
    stream = torch.cuda.Stream()          # (see encode_sharded_local)
    torch.cuda.synchronize()
    st = stream.cuda_stream
    for k, (row0, nrows) in enumerate(partition_mcu_rows(vu, n)):
        ctxs[k].shard_decode_dev(d_scan, len(scan), frame, gray, row0, nrows, planes[0], planes[1], planes[2], pl, status[k: k + 1], stream=st)
    torch.cuda.synchronize()
    if int(status.abs().sum().item()):
        raise RuntimeError("decode failed")
    h = planes.cpu().numpy()
    return h[0], h[1], h[2]
