import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from exoplanet_b200 import gp, synthetic, _lib
lib=_lib.require_cuda()
dev="cuda"
d=lambda x: torch.as_tensor(np.ascontiguousarray(x),dtype=torch.float64,device=dev)
pr=synthetic.gp_problem(1000000)
t,y,dg=d(pr["t"]),d(pr["y"]),d(pr["diag"])
flush=torch.empty(192*1024*1024//8,dtype=torch.float64,device=dev)
def tm(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot=0
    for _ in range(n):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot+=a.elapsed_time(b)
    return tot/n
for B in (1,16):
    yb=y[None,:].repeat(B,1).contiguous()
    ws=torch.empty(lib.xo_celerite_sho_workspace(1000000,B)//8+1,dtype=torch.float64,device=dev)
    wg=torch.empty(lib.xo_celerite_sho_grad_workspace(1000000,B)//8+1,dtype=torch.float64,device=dev)
    r={}
    a=tm(lambda: r.__setitem__("l",gp.sho_loglike(t,yb,dg,pr["S0"],pr["w0"],pr["Q"],workspace=ws)))
    b=tm(lambda: r.__setitem__("g",gp.sho_grad(t,yb,dg,pr["S0"],pr["w0"],pr["Q"],workspace=wg)))
    print(os.environ.get("XO_B200_LIB","default")[-16:], "B=%d loglike %.4f ms grad %.4f ms ll=%.10e gh=%s"%(B,a,b,r["l"][0].item(), r["g"][1][0].cpu().numpy()))
