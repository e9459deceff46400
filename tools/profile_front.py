"""GPU debug: clock64() phase breakdown of the warp factor kernel on the largest top-level fronts of the OPF-10k case."""
import sys, os, ctypes as C
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT+"/oracle", ROOT+"/tests"): sys.path.insert(0,p)
import numpy as np, torch
import madnlp_jl_b200 as pkg
from madnlp_jl_b200 import kkt as K
from mf_emulator import Symbolic
W=pkg.workloads; capi=pkg.capi; lib=capi.lib
case=sys.argv[1] if len(sys.argv)>1 else "case10000_goc"
model,st=W.acopf_case(case); it=W.ipm_iterates(model,st,1,seed=0)[0]
class CB: pass
cb=CB(); cb.nvar,cb.ncon=st.nvar,st.ncon; cb.jac_I,cb.jac_J,cb.hess_I,cb.hess_J=st.jac_I,st.jac_J,st.hess_I,st.hess_J; cb.ind_ineq,cb.ind_lb,cb.ind_ub=st.ind_ineq,st.ind_lb,st.ind_ub
kg=K.SparseCondensedKKTSystem(cb); kg.initialize()
dev=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
kg.get_jacobian().copy_(dev(it.jac)); kg.get_hessian().copy_(dev(it.hess)); kg.reg.copy_(dev(it.reg)); kg.du_diag.copy_(dev(it.du_diag))
kg.l_diag.copy_(dev(it.l_diag)); kg.u_diag.copy_(dev(it.u_diag)); kg.l_lower.copy_(dev(it.l_lower)); kg.u_lower.copy_(dev(it.u_lower))
kg.compress_jacobian(); kg.compress_hessian(); kg.set_aug_diagonal_(); kg.build_kkt(); kg.linear_solver.factorize(); print(kg.linear_solver.inertia())
S=Symbolic(kg.n,kg.aug_com.colptr,kg.aug_com.rowval)
ws=S.sn_first[1:]-S.sn_first[:-1]; fs=(S.rows_ptr[1:]-S.rows_ptr[:-1]).astype(int)
ch=S.children()
order=np.argsort(-(ws*fs))[:6]
names=["desc","zero","A","children","pivots","update-block+hand-off","panel-store"]
for sn in list(order)+[int(np.argmax(S.sn_level))]:
    reps=4; st_=np.zeros(8*reps,dtype=np.int64)
    capi.check(lib.b2_debug_profile_front(kg.linear_solver._h,int(sn),reps,st_.ctypes.data))
    st_=st_.reshape(reps,8); d=np.diff(st_,axis=1)
    print("sn",sn,"w",ws[sn],"f",fs[sn],"nchild",len(ch[sn]),"level",S.sn_level[sn])
    for r in (0,reps-1): print("   rep",r," ".join("%s=%d"%(n,v) for n,v in zip(names,d[r])),"total",st_[r,7]-st_[r,0],"cycles")
