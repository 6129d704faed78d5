"""where does the end-to-end step go?  host-side timings of the encode / decode calls, alone and in the 2-lane pipeline; PCIe rate with
kernels running next to the copies."""
import os, sys, time, threading, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import jpezy_b200 as J
from jpezy_b200 import capi
W, H = 3840, 2160
npx = W * H
ctx = J.Context(0); L = ctx.lib
frame = J.default_frame(W, H); pl = J.plane_bytes(frame)
r, g, b = J.synth.image(0, W, H, 0)
hin = [torch.from_numpy(x.reshape(-1).copy()).pin_memory() for x in (r, g, b)]
hscan = torch.empty(npx * 3, dtype=torch.uint8).pin_memory()
hout = [torch.empty(pl, dtype=torch.uint8).pin_memory() for _ in range(3)]
nb, nbits = C.c_size_t(0), C.c_uint64(0)
def enc(c=ctx, hs=hscan, n=nb):
    c._chk(L.jpezyb200_encode(c.h, hin[0].data_ptr(), hin[1].data_ptr(), hin[2].data_ptr(), W, H, 0, hs.data_ptr(), hs.numel(), C.byref(n), C.byref(nbits)))
def dec(c=ctx, hs=hscan, n=nb, ho=hout):
    c._chk(L.jpezyb200_decode(c.h, hs.data_ptr(), n.value, C.byref(frame), 0, ho[0].data_ptr(), ho[1].data_ptr(), ho[2].data_ptr(), pl))
for _ in range(3): enc(); dec()
t0 = time.perf_counter()
for _ in range(20): enc()
t1 = time.perf_counter()
for _ in range(20): dec()
t2 = time.perf_counter()
print("encode call alone %.3f ms, decode call alone %.3f ms (copies alone: %.3f ms each at 46 GB/s)" % ((t1 - t0) / 20 * 1e3, (t2 - t1) / 20 * 1e3, 3 * npx / 46e9 * 1e3))
# copies with kernels running next to them
n = 3 * npx
h_up = torch.empty(n, dtype=torch.uint8).pin_memory(); h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda"); d_dn = torch.zeros(n, dtype=torch.uint8, device="cuda")
s_up, s_dn, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
d_in = torch.empty((3, 1, H, W), dtype=torch.uint8, device="cuda")
ctx.synth_dev(d_in[0], d_in[1], d_in[2], W, H, nimg=1, first_frame=0, family=0, stream=s_k.cuda_stream)
nm = capi.num_mcus(W, H)
d_coefs = torch.empty((nm, 6, 64), dtype=torch.int16, device="cuda"); d_scan = torch.zeros(npx, dtype=torch.uint8, device="cuda")
d_nb = torch.zeros(1, dtype=torch.int64, device="cuda"); d_st = torch.zeros(1, dtype=torch.int32, device="cuda")
d_out = torch.zeros((3, pl), dtype=torch.uint8, device="cuda")
def kernels():
    ctx.encode_batch_dev(d_in[0], d_in[1], d_in[2], W, H, 1, False, d_scan, npx, d_nb, None, stream=s_k.cuda_stream)
    ctx.decode_batch_dev2(d_scan, npx, d_nb, 400000, 1, frame, False, d_out[0], d_out[1], d_out[2], pl, d_st, stream=s_k.cuda_stream)
for with_k in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s_up): d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_dn): h_dn.copy_(d_dn, non_blocking=True)
        if with_k:
            for _ in range(3): kernels()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("copies both ways%s: %.3f ms per pair = %.1f GB/s each way" % (" with 3 round trips of kernels next to each pair" if with_k else "", dt * 1e3, n / dt / 1e9))
# 3 separate plane copies vs one
for parts in (1, 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        for q in range(parts):
            a, z = q * n // parts, (q + 1) * n // parts
            with torch.cuda.stream(s_up): d_up[a:z].copy_(h_up[a:z], non_blocking=True)
            with torch.cuda.stream(s_dn): h_dn[a:z].copy_(d_dn[a:z], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("copies in %d parts: %.3f ms" % (parts, dt * 1e3))
# L independent workers, each: encode then decode, own contexts and buffers
for nl in (1, 2, 3, 4):
    ws = []
    for k in range(nl):
        c = J.Context(0)
        hs = torch.empty(npx * 3, dtype=torch.uint8).pin_memory()
        ho = [torch.empty(pl, dtype=torch.uint8).pin_memory() for _ in range(3)]
        ws.append((c, hs, C.c_size_t(0), ho))
    def work(w, n):
        torch.cuda.set_device(0)
        c, hs, nbw, ho = w
        for _ in range(n):
            enc(c, hs, nbw); dec(c, hs, nbw, ho)
    for w in ws: work(w, 2)
    ths = [threading.Thread(target=work, args=(w, 12)) for w in ws]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print("%d workers (encode then decode each): %.3f ms per step = %.1f GPix/s" % (nl, dt / (12 * nl) * 1e3, npx * 12 * nl / dt / 1e9))
    for w in ws: w[0].close()
