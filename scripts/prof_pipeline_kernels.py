"""Kernel-time split of ONE full-pipeline call (run under `ncu --metrics gpu__time_duration.sum --csv`): which kernels a
findEssentialMatrix / findFundamentalMatrix / findHomography pair spends its GPU time in.  usage: prof_pipeline_kernels.py e|f|h"""
import sys
sys.path.insert(0, "graph-cut-ransac_b200")
import pygcransac
from pygcransac import synth
which = sys.argv[1] if len(sys.argv) > 1 else "e"
if which == "e":
    p = synth.epipolar_pair(2048, 0.85, 1)[0]
    pygcransac.findEssentialMatrix(p, synth.K_MAT, synth.K_MAT, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=0.1,
                                   neighborhood=0, neighborhood_size=4.0, use_sprt=False, sampler=0, max_iters=100000, seed=1)
elif which == "f":
    p = synth.epipolar_pair(4096, 0.5, 1)[0]
    pygcransac.findFundamentalMatrix(p, 1024, 1024, 1024, 1024, [], threshold=0.75, neighborhood=0, neighborhood_size=4.0, sampler=0, seed=1)
else:
    p = synth.homography_pair(2048, 0.3, 1)[0]
    pygcransac.findHomography(p, 1024, 1024, 1024, 1024, [], threshold=3.0, sampler=0, seed=1)
print(pygcransac.last_statistics)
