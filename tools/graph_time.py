import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import raymarchdenoisercuda_b200 as rmd
from raymarchdenoisercuda_b200.synth import synth_frame
for name,(W,H,seed) in {"1080p":(1920,1080,0x5EED0001),"4k":(3840,2160,0x5EED0002)}.items():
    dev=[[torch.from_numpy(x.view(np.int32) if x.dtype==np.uint32 else x).cuda() for x in synth_frame(W,H,seed,f)] for f in range(2)]
    out=[torch.empty((H,W,4),dtype=torch.float32,device='cuda') for _ in range(2)]
    p=rmd.FilterParams(type=rmd.FilterType.WAVELET,depth=5,radius=2)
    ctx=rmd.SvgfContext(W,H)
    for i in range(6): ctx.frame(*dev[i&1],out[i&1],p)
    torch.cuda.synchronize()
    s=torch.cuda.current_stream()
    def t(fn,n=20):
        fn(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n): fn()
        e1.record(s); torch.cuda.synchronize()
        return e0.elapsed_time(e1)/n/2*1e3
    eager=t(lambda:(ctx.frame(*dev[0],out[0],p),ctx.frame(*dev[1],out[1],p)))
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ctx.frame(*dev[0],out[0],p); ctx.frame(*dev[1],out[1],p)
    graph=t(lambda:g.replay())
    print(name,'eager us/frame',round(eager,1),'graph us/frame',round(graph,1))
