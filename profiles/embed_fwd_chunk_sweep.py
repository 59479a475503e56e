import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import acquisition_focus_b200 as afb
from oracle import cases
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.mean(ts)
for (c,S) in ((16,128),(32,64),(64,32)):
    case=cases.embed_case(S,c,6,2,seed=300+S)
    x=case['x'].cuda(); aff=torch.stack(case['affines'],0).cuda()
    nbytes=2*6*c*S**3*4
    for chunk in (0,):
        os.environ['AFB_EMBED_FWD_CHUNK']=str(chunk)
        t=timeit(lambda: afb.embed_slices(x,aff,6))
        print(f"c={c} S={S} chunk={chunk}: {t:.3f} ms {nbytes/t/1e6:.0f} GB/s")
    t=timeit(lambda: torch.zeros(2,6*c,S,S,S,device='cuda'))
    print(f"   torch.zeros same size: {t:.3f} ms {nbytes/t/1e6:.0f} GB/s")
