import os, sys, time, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from codec_eval_b200 import _lib
from codec_eval_b200.metrics import GpuMetrics, MetricConfig
sys.argv=['bench.py']
import bench
urefs,dists,ref_index=bench.make_pairs('cfg2',0)
w,h=768,512; n=dists.shape[0]; img=w*h*3
h_ref=torch.from_numpy(urefs).pin_memory(); h_dist=torch.from_numpy(dists).pin_memory()
ctx=GpuMetrics(0); L=ctx._L
pairs=(_lib.CePair*n)()
for i in range(n): pairs[i]=_lib.CePair(h_ref.data_ptr()+int(ref_index[i])*img, h_dist.data_ptr()+i*img, img,img,w,h,int(ref_index[i]),0)
out=(_lib.CeResult*n)(); cfg=MetricConfig.all()._c()
for ch in sys.argv[1:] or ['1','2','3','4','6','8']:
    pass
for ch in ['1','2','3','4','6','8']:
    os.environ['CE_HOST_CHUNKS']=ch
    for _ in range(3): L.ce_evaluate_batch(ctx._h,pairs,n,C.byref(cfg),80.0,out)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): L.ce_evaluate_batch(ctx._h,pairs,n,C.byref(cfg),80.0,out)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print('chunks',ch,'ms',round(dt*1e3,2),'MPix-pairs/s',round(n*w*h/1e6/dt,1))
