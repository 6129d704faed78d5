#!/usr/bin/env python
"""Live per-kernel durations and gaps of the device round trip (CUPTI activity records through torch.profiler).

ncu serialises and cold-caches every launch; this runs the same step un-instrumented apart from CUPTI's activity
buffer and prints, per kernel name, the mean duration and the mean idle gap in front of it, in launch order of one
step.  Usage: python tools/kernel_timeline.py [--workload c2|c3] [--family 0|1] [--steps 20]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--family", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    from torch.profiler import ProfilerActivity, profile
    import jpezy_b200 as J
    from bench import WORKLOADS

    wl = dict(WORKLOADS[args.workload])
    W, H, B, gray = wl["W"], wl["H"], args.batch or wl["batch"], wl["gray"]
    ctx = J.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    frame = J.default_frame(W, H)
    plane_len = J.plane_bytes(frame)
    npx = W * H
    slot = max(npx, 65536)
    ring = 4
    d_in = torch.empty((ring, 3, B, H, W), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((ring, 3, B, plane_len), dtype=torch.uint8, device="cuda")
    d_scan = torch.zeros((ring, B, slot), dtype=torch.uint8, device="cuda")
    d_nbytes = torch.zeros((ring, B), dtype=torch.int64, device="cuda")
    d_status = torch.zeros((ring, B), dtype=torch.int32, device="cuda")
    for k in range(ring):
        ctx.synth_dev(d_in[k, 0], d_in[k, 1], d_in[k, 2], W, H, nimg=B, first_frame=k * B, family=args.family, stream=sp)
    torch.cuda.synchronize()

    def step(k):
        ctx.encode_batch_dev(d_in[k, 0], d_in[k, 1], d_in[k, 2], W, H, B, gray, d_scan[k], slot, d_nbytes[k], None, stream=sp)
        nb = d_nbytes[k].cpu().numpy().astype(np.uint64)
        ctx.decode_batch_dev(d_scan[k], slot, nb, B, frame, gray, d_out[k, 0], d_out[k, 1], d_out[k, 2], plane_len, d_status[k], stream=sp)

    for i in range(8):
        step(i % ring)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(i % ring)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at every k_fwd_transform
    per = collections.OrderedDict()
    prev_end = None
    order = []
    idx_in_step = 0
    for e in evs:
        name = e.name.split("(")[0].replace("jz::", "").replace("void ", "")
        if name.startswith("k_fwd_transform"):
            idx_in_step = 0
            prev_end = None
        key = (idx_in_step, name)
        dur = e.time_range.end - e.time_range.start
        gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
        per.setdefault(key, []).append((dur, gap))
        prev_end = e.time_range.end
        idx_in_step += 1
    tot = 0.0
    print(f"{'#':>3} {'kernel':28} {'n':>4} {'dur us':>9} {'gap us':>9}")
    for (i, name), v in sorted(per.items()):
        d = sum(x[0] for x in v) / len(v)
        g = sum(x[1] for x in v) / len(v)
        tot += d + g
        print(f"{i:3d} {name:28} {len(v):4d} {d:9.2f} {g:9.2f}")
    print(f"sum of mean (dur + gap) per step: {tot:.1f} us")
    print("sync iterations (launch 0 / 1):", ctx.stat(J.capi.STAT_SYNC_ITERS0), ctx.stat(J.capi.STAT_SYNC_ITERS1))


if __name__ == "__main__":
    main()
