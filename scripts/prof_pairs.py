"""Per-phase latency of the full pipelines on one host thread (gcr_run_stats.phase_seconds): F plain / F default / H / E."""
import sys, time
sys.path.insert(0,'graph-cut-ransac_b200')
import numpy as np
import pygcransac
from pygcransac import synth
def run(name, fn, pts, **kw):
    for rep in range(2):
        t0=time.perf_counter(); st=[]
        for i,p in enumerate(pts):
            fn(p, seed=i+1, **kw); st.append(pygcransac.last_statistics)
        dt=(time.perf_counter()-t0)/len(pts)
    ph=np.mean([s["phase_seconds"] for s in st],axis=0)*1e3
    print(name, "%.2f ms/pair"%(1e3*dt), "phases ms [setup batch loop-rest LO IRLS final]", np.round(ph,2), "gc %.2f"%(1e3*np.mean([s["graph_cut_seconds"] for s in st])), "exact", np.mean([s["exact_rescores"] for s in st]), "gc#", np.mean([s["graph_cut_number"] for s in st]))
F=lambda p,**kw: pygcransac.findFundamentalMatrix(p,1024,1024,1024,1024,[],threshold=0.75,spatial_coherence_weight=0.0,neighborhood=0,neighborhood_size=4.0,use_sprt=False,sampler=0,**kw)
H=lambda p,**kw: pygcransac.findHomography(p,1024,1024,1024,1024,[],threshold=3.0,spatial_coherence_weight=0.975,neighborhood=0,neighborhood_size=4.0,use_sprt=False,sampler=0,**kw)
E=lambda p,**kw: pygcransac.findEssentialMatrix(p,synth.K_MAT,synth.K_MAT,1024,1024,1024,1024,[],threshold=0.75,spatial_coherence_weight=0.1,neighborhood=0,neighborhood_size=4.0,use_sprt=False,sampler=0,max_iters=100000,**kw)
import os
if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
    g = np.load(os.path.join("tests", "golden", "cfg1_matches.npz"))
    pts = np.ascontiguousarray(g["pts"], dtype=np.float64); sz = [int(v) for v in g["size"]]
    for sampler in (2, 1, 0):
        C = lambda p, **kw: pygcransac.findHomography(p, *sz, [], threshold=3.0, max_iters=5000, min_iters=50, sampler=sampler, **kw)
        run("cfg1 sampler %d" % sampler, C, [pts] * 32)
        print("   iterations", pygcransac.last_statistics["iteration_number"], "batches", pygcransac.last_statistics["gpu_batches"], "lo_samples", pygcransac.last_statistics["lo_samples"])
    sys.exit(0)
run("F plain", F, [synth.epipolar_pair(4096,0.5,s+1)[0] for s in range(16)], use_degensac=False, bundle_adjustment=False)
run("F full ", F, [synth.epipolar_pair(4096,0.5,s+1)[0] for s in range(16)])
run("H      ", H, [synth.homography_pair(2048,0.3,s+1)[0] for s in range(16)])
run("E      ", E, [synth.epipolar_pair(2048,0.85,s+1)[0] for s in range(8)])
