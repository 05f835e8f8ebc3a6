import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),"oracle")); sys.path.insert(0, os.path.join(os.getcwd(),"tests"))
import numpy as np, torch, ctypes as Ct
import xo_oracle as O
from exoplanet_b200 import fused, synthetic, elements, _lib
DEV="cuda"
dev=lambda x: torch.as_tensor(np.ascontiguousarray(x),dtype=torch.float64,device=DEV)
for P,texp in ((1,None),(3,0.02)):
    N,C=30000,6
    p=synthetic.transit_params(C,P,seed=13)
    t=synthetic.cadence_times(N); y=synthetic.noise(N)
    data=fused.TransitData(t,y,np.full(N,5e-4),device=DEV)
    rs=1.0+0.1*np.arange(C)/C; ms=1.0-0.05*np.arange(C)/C
    par=fused.pack_params(dev(p["period"]),dev(p["t0"]),dev(p["ecc"]),dev(p["omega"]),dev(p["b"]),dev(p["ror"]),dev(p["m_planet"]),dev(p["K"]))
    star=fused.pack_star(dev(rs),dev(ms),dev(p["u1"]),dev(p["u2"])); mean=dev(p["mean"])
    model=fused.TransitModel(data,texp=texp)
    logp,gpar,gstar,gmean=model.logp_grad(par,star,mean)
    el_o=O.elements_from_params(p["period"],p["t0"],p["ecc"],p["omega"],p["b"],p["ror"],r_star=rs[:,None],m_star=ms[:,None],m_planet=p["m_planet"],texp=texp)
    lo=O.transit_logp_grad(t,y,np.full(N,1/5e-4**2),el_o,O.get_cl(p["u1"],p["u2"]),p["mean"],*O.stencil(texp),lognorm=data.lognorm)
    print("P",P,"logp rel err vs oracle",np.max(np.abs(logp.cpu().numpy()-lo[0])/np.abs(lo[0])))
    elem=torch.empty((C,P,16),dtype=torch.float64,device=DEV); clk=torch.empty((C,3),dtype=torch.float64,device=DEV)
    rc=_lib.load().xo_elements_from_params(Ct.c_void_p(par.data_ptr()),Ct.c_void_p(star.data_ptr()),C,P,0.0 if texp is None else texp,1,Ct.c_void_p(elem.data_ptr()),Ct.c_void_p(clk.data_ptr()),None)
    torch.cuda.synchronize()
    e=elem.cpu().numpy()
    rel=np.abs(e[...,:11]-el_o[...,:11])/np.maximum(np.abs(el_o[...,:11]),1e-300)
    print("  elements max rel err per column", rel.max(axis=(0,1)))
    print("  window abs err", np.abs(e[...,11:13]-el_o[...,11:13]).max())
