"""Five-point solver: the reference's own source (oracle/_ref/libgcr_ref_solvers.so) against the oracle's restatement, with the
reference's root finder plugged into the restatement and with the restatement's own -- the statistics quoted in DESIGN.md
section 5 (profiles/r2_reference_source_pinning.txt).  CPU only: python tools/five_point_root_finder_stats.py"""
import os
import sys, ctypes as C, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ('graph-cut-ransac_b200','oracle','tests'): sys.path.insert(0,os.path.join(ROOT,d))
import oracle as O
from pygcransac import synth
import test_reference_solver_sources as T
src=C.CDLL(O.build_ref_solvers()); orc=O.Oracle()
def resid(E):
    E=E/np.linalg.norm(E); return np.abs(2*E@E.T@E-np.trace(E@E.T)*E).max()
rng = np.random.default_rng(2)
px, Fm, Em, K, inl = synth.epipolar_pair(700, 0.4, seed=5)
pts = np.ascontiguousarray(orc.normalize_correspondences(px, K, K)); idx=np.flatnonzero(inl)
N=3000
conv_pairs=[]; unconv=0; ref_models=0; missing_in_orc=0; extra_orc=0; extra_orc_bad=0; cnt_same=0; worst_conv=0; w2=[]
for t in range(N):
    s = (rng.choice(idx, 5, replace=False) if t % 3 else rng.choice(700, 5, replace=False)).astype(np.uint32)
    ref = T._call(src.gcr_refsrc_solve_e5_nister, pts, s, 10, 3, 3)
    out=np.zeros(90); k=src.gcr_refsrc_oracle_e5_with_reference_sturm(T._p(pts),T._p(s),0,T._p(out)); rs=out[:9*k].reshape(k,3,3).copy()
    own = orc.solve_e5(pts, s)
    cnt_same += len(ref)==len(rs)
    if len(ref)==len(rs):
        for a,b in zip(ref,rs):
            if resid(a)<1e-9 and resid(b)<1e-9: w2.append(np.abs(a-b).max())
    for a in ref:
        ref_models+=1
        if resid(a)>=1e-9: unconv+=1; continue
        d=min([np.abs(a-b).max() for b in own] or [9])
        conv_pairs.append(d)
    matched=set()
    for b in own:
        d=min([np.abs(a-b).max() for a in ref] or [9])
        if d>1e-5:
            extra_orc+=1
            if resid(b)>1e-9: extra_orc_bad+=1
print("samples",N,"same count (restatement+ref sturm vs source)",cnt_same)
print("converged pairs restatement+ref-sturm vs source: n",len(w2),"quantiles",np.quantile(w2,[.5,.9,.99,.999,1]))
print("reference models",ref_models,"of which not converged (resid>=1e-9)",unconv)
cp=np.array(conv_pairs); print("converged reference models matched by the oracle: quantiles",np.quantile(cp,[.5,.9,.99,.999,1]), "unmatched(>1e-5)",(cp>1e-5).sum())
print("oracle models without a reference counterpart",extra_orc,"of which resid>1e-9",extra_orc_bad)
