"""Times jpezyb200_transform_inv_dev on a 3840x2160 frame of every sampling layout the device decoder accepts (synthetic coefficients:
DC + two small AC coefficients in a third of the blocks).  usage: python tools/inv_general_bench.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import jpezy_b200 as J
from jpezy_b200 import capi
ctx = J.Context(0)
W, H = 3840, 2160
st = torch.cuda.Stream(); sp = st.cuda_stream
torch.cuda.set_stream(st)
for name, hs, vs, ncomp in (("420", (2,1,1), (2,1,1), 3), ("444", (1,1,1), (1,1,1), 3), ("422", (2,1,1), (1,1,1), 3), ("gray1", (1,), (1,), 1)):
    f = J.default_frame(W, H)
    f.ncomp = ncomp
    for i in range(ncomp):
        f.hs[i] = hs[i]; f.vs[i] = vs[i]
    if ncomp == 1:
        pass
    hmax, vmax = max(hs), max(vs)
    hu, vu = -(-W // (8*hmax)), -(-H // (8*vmax))
    nb = sum(h*v for h, v in zip(hs, vs))
    pl = capi.plane_bytes(f)
    rng = np.random.default_rng(1)
    c = np.zeros((hu*vu*nb, 64), dtype=np.int16)
    c[:, 0] = rng.integers(-60, 60, size=c.shape[0])
    m = rng.random(c.shape[0]) < 0.3
    c[m, 1] = rng.integers(-3, 4, size=int(m.sum())); c[m, 2] = rng.integers(-2, 3, size=int(m.sum()))
    d = torch.from_numpy(c).cuda()
    out = torch.zeros((3, pl), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.transform_inv_dev(d, f, 1, False, out[0], out[1], out[2], pl, stream=sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(20):
        ctx.transform_inv_dev(d, f, 1, False, out[0], out[1], out[2], pl, stream=sp)
    e1.record(st); torch.cuda.synchronize()
    print(name, "nb", nb, "us per launch %.1f" % (e0.elapsed_time(e1) / 20 * 1e3))
