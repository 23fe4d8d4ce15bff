"""DEV TOOL (not imported by the product, not an oracle): CPU prototype of the price coordination in
pyhybridcontrol_b200/csrc/coupling.cu -- projected subgradient ascent on the dual of the centralised micro-grid problem,
agents solved with HiGHS.  `python tools/coupling_proto.py` prints bounds and the certified gap."""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import mld as omld, condense as oc, assemble as oa, solve as osv
from pyhybridcontrol_b200.examples.residential_mg_with_pv_and_dewhs import synthetic as syn

def agent_prob(p, Nt, T, w, lam, price):
    mats = syn.dewh_scalars(p, const_heat=True)
    m = dict(A=[[mats[0]]], B1=[[mats[1]]], B4=[[mats[2]]], b5=[[mats[3]]], E=[[1.0],[-1.0]], F1=[[0.0],[0.0]], Psi=[[-1.0,0.0],[0.0,-1.0]], f5=[[p["T_h_max"]],[-p["T_h_min"]]])
    full,d,vt = omld.complete({k:np.array(v,float) for k,v in m.items()}, nu_l=1)
    tot = (price*p["P_h_Nom"]).sum()
    return oa.build_problem(oc.condense(full,d,Nt), d, vt, Nt, np.array([T]), w, atoms=dict(q_u=lam*p["P_h_Nom"], q_mu=[10*tot, tot]))

def run(N_h, N_p, iters=100, seed=0, theta=1.0):
    Nt=N_p+1
    params=[syn.dewh_agent_params(seed*100+b) for b in range(N_h)]
    T0=np.array([syn.dewh_initial_state(seed*100+b) for b in range(N_h)])
    dem=np.stack([syn.dhw_demand_profile(Nt, seed=seed*100+b) for b in range(N_h)])
    price=syn.price_profile(Nt, seed=seed)
    P=np.array([p["P_h_Nom"] for p in params])
    k=np.arange(Nt)
    pv = -0.6*P.sum()*np.clip(np.sin((k-2)/Nt*2*np.pi),0,None)      # surplus hump
    res = 0.1*P.sum()*np.ones(Nt)
    r = pv+res
    lo=np.zeros(Nt); hi=np.full(Nt,P.sum())
    lam=price.copy()
    UB=np.inf; LB=-np.inf; best=None
    hist=[]
    for it in range(iters):
        objs=[];U=[]
        for b in range(N_h):
            st,obj,v=osv.solve_milp(agent_prob(params[b],Nt,T0[b],dem[b],lam,price))
            objs.append(obj); U.append(np.round(v[::3]))
        U=np.array(U); agg=P@U
        sobj=sum(objs)
        pen=sobj-lam@agg
        primal=(price*np.maximum(0,agg+r)).sum()+pen
        kink=np.clip(-r,lo,hi)
        cand=np.stack([lo,hi,kink]); vals=price*np.maximum(0,cand+r)-lam*cand
        j=vals.argmin(0); a=cand[j,k]; dualagg=vals.min(0).sum()
        dual=sobj+dualagg
        LB=max(LB,dual)
        if primal<UB: UB=primal; best=U.copy()
        g=agg-a
        hist.append((dual,primal))
        if g@g<1e-9: break
        alpha=theta*(UB-dual)/(g@g)
        lam=np.clip(lam+alpha*g,0,price)
    return LB,UB,hist,dict(params=params,T0=T0,dem=dem,price=price,P=P,r=r,Nt=Nt,best=best)

if __name__=="__main__":
    for N_h in (4,16):
        t=time.time()
        LB,UB,hist,ctx=run(N_h,12,iters=60)
        print(N_h,"LB",LB,"UB",UB,"gap",(UB-LB)/abs(UB), "decentral primal", hist[0][1], "iters",len(hist), time.time()-t)
