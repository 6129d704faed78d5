"""A/B run of the transform kernels on 64 HD frames: JPEZYB200_OPT_TRANSFORM = 2 selects the one-thread-per-block forward
kernel (the inverse kernel has a single production variant; its time is printed as a repeatability check).
usage: python tools/ab_transform.py [family]"""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import jpezy_b200 as J
from jpezy_b200 import capi
W,H,B=1920,1080,64
ctx=J.Context(0)
stream=torch.cuda.Stream(); torch.cuda.set_stream(stream); sp=stream.cuda_stream
frame=J.default_frame(W,H); pl=J.plane_bytes(frame); nm=capi.num_mcus(W,H)
ring=3
d_in=torch.empty((ring,3,B,H,W),dtype=torch.uint8,device="cuda")
d_out=torch.zeros((ring,3,B,pl),dtype=torch.uint8,device="cuda")
d_coefs=torch.empty((ring,B,nm,6,64),dtype=torch.int16,device="cuda")
for k in range(ring):
    ctx.synth_dev(d_in[k,0],d_in[k,1],d_in[k,2],W,H,nimg=B,first_frame=k*B,family=int(sys.argv[1]) if len(sys.argv)>1 else 0,stream=sp)
    ctx.transform_fwd_dev(d_in[k,0],d_in[k,1],d_in[k,2],W,H,B,False,d_coefs[k],stream=sp)
torch.cuda.synchronize()
def timed(fn,it=20):
    for i in range(3): fn(i%ring)
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(it): fn(i%ring)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b)/it
ref=None
for var in (2,0,2,0):
    ctx.set_option(capi.OPT_TRANSFORM,var)
    ti=timed(lambda k: ctx.transform_inv_dev(d_coefs[k],frame,B,False,d_out[k,0],d_out[k,1],d_out[k,2],pl,stream=sp))
    tf=timed(lambda k: ctx.transform_fwd_dev(d_in[k,0],d_in[k,1],d_in[k,2],W,H,B,False,d_coefs[k],stream=sp))
    chk=int(d_out[0].to(torch.int64).sum().item())
    print("variant",var,"inv ms %.4f fwd ms %.4f checksum %d"%(ti,tf,chk))
