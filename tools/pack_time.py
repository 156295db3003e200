#!/usr/bin/env python
"""DEVELOPMENT: time of npde_pack / npde_unpack at 64 x 1025^2 (python tools/pack_time.py on the GPU box)."""
import sys, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo')
npde = importlib.import_module("neural-pde-solver_b200")
x = torch.rand(64,1025,1025,device='cuda'); it = npde.ConvIterator("clamp",3).cuda()
p = it.pack(x); out = torch.empty_like(x)
for name,fn in (("pack",lambda: it.pack(x)),("unpack",lambda: it.unpack(p,out=out))):
    for _ in range(3): fn()
    ts=[]
    for _ in range(10):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(name, "%.1f us"%(min(ts)*1e3))
assert torch.equal(it.unpack(it.pack(x)), x)
