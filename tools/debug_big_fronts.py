"""GPU debug: compare the device factor front by front with the numpy replay (tests/mf_emulator.py)."""
import sys, os, ctypes as C
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT+"/oracle", ROOT+"/tests"): sys.path.insert(0,p)
import numpy as np, torch
import madnlp_oracle as o, madnlp_jl_b200 as pkg
from mf_emulator import Symbolic
from madnlp_jl_b200.linear_solvers import B200SparseSolver, DeviceCSC
W=pkg.workloads; capi=pkg.capi; lib=capi.lib
nx=int(sys.argv[1]) if len(sys.argv)>1 else 14
N,n_tot,m,I,J,V=W.augmented_grid_kkt(nx,nx,nx)
cp,rv,mp=o.coo_to_csc(I,J,N,N); nz=np.zeros(len(rv)); o.transfer(nz,V,mp)
csc=DeviceCSC(N,N,cp,rv,torch.from_numpy(nz).cuda())
Kf=o.tril_to_full(cp,rv,nz,N); b=np.random.default_rng(0).standard_normal(N)
for smax in (160,64,8):
    opts=dict(kkt_n_primal=n_tot,small_front_max=smax,use_cuda_graph=0)
    M=B200SparseSolver(csc,B200SparseSolver.default_options(**opts)); M.factorize(); print("smax",smax,"inertia",M.inertia(),M.stats()["n_big_fronts"])
    S=Symbolic(N,cp,rv,**opts); S.factorize(nz)
    L=np.zeros(S.lval_size); d=np.zeros(N); capi.check(lib.b2_debug_get_factor(M._h,L.ctypes.data,d.ctypes.data))
    print("  d err",np.abs(d-S.d).max()/np.abs(S.d).max())
    worst=[]
    for s in range(S.ns):
        w=S.sn_first[s+1]-S.sn_first[s]; f=int(S.rows_ptr[s+1]-S.rows_ptr[s])
        Pg=L[S.lp_off[s]:S.lp_off[s]+f*w].reshape(w,f).T; Pe=S.L[S.lp_off[s]:S.lp_off[s]+f*w].reshape(w,f).T
        e=np.abs(np.tril(Pg,-1)-np.tril(Pe,-1)).max() if f>1 else 0.0
        ed=np.abs(d[S.sn_first[s]:S.sn_first[s+1]]-S.d[S.sn_first[s]:S.sn_first[s+1]]).max()
        worst.append((max(e,ed),s,w,f,int(S.sn_level[s])))
    worst.sort(reverse=True)
    print("  worst fronts (err, sn, w, f, level):",[(float("%.2e"%a),b_,c,dd,e_) for a,b_,c,dd,e_ in worst[:6]])
    # first level at which an error appears
    bad=[x for x in worst if x[0]>1e-9]
    if bad: print("  lowest bad level:",min(x[4] for x in bad),"fronts:",sorted([(x[4],x[1],x[2],x[3],float('%.1e'%x[0])) for x in bad])[:8])
    x=M.solve_linear_system(torch.from_numpy(b).cuda()).cpu().numpy()
    print("  gpu solve residual",np.abs(Kf@x-b).max()/(abs(Kf).max()*np.abs(x).max()+np.abs(b).max()))
    # emulator solve with GPU factor
    S.L, S.d = L, d
    xe=S.solve(b); print("  emulator-solve(with gpu factor) residual",np.abs(Kf@xe-b).max()/(abs(Kf).max()*np.abs(xe).max()+np.abs(b).max()))
