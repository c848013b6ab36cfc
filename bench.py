#!/usr/bin/env python
"""bench.py -- hypothesis x point evaluations / s (and image pairs / s) of the GC-RANSAC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-extras]

Default (what the driver runs): the headline line is BASELINE.json configs[1] -- fundamental matrix, 7-point + MSAC,
synthetic N = 4096 correspondences, 50 % outliers, threshold 0.75 px, single image pair -- and the SAME JSON line carries an
`extra` block with the other BASELINE configurations measured in the same process: h4k (the homography twin), e2k
(configs[2]), h2k_batch (configs[3]: 2048 pairs per GPU, pairs/s resident and end to end), f1m_sharded (configs[4]:
hypotheses sharded, NCCL all-gather of the FP64 winner records; asserts N-rank result == 1-rank result bit for bit) and the
full pipelines through pygcransac.find* (pairs/s) with their reference CPU arms.

One "step" = `rounds_per_step` consecutive fixed-K passes of the hot path (gcr_round: K1 draw the minimal samples -> K2
solve + validity tests -> K3 score every model against all points -> K4 FP64 winner), each over FRESH samples of the
counter-based stream, so that K steps make a timed region of >= 1 s (clocks are sampled during it).  `value` counts only
REAL models (valid model count x N), not empty root slots.  At N > 1 every rank runs the per-GPU workload on its own
pairs (pairs shard with no collective: weak scaling); f1m_sharded is strong scaling.

The JSON line follows the driver contract (+ roofline, cpu_baseline, e2e, clocks, gpu_launches, extra).
--impl reference times the reference's CPU implementation of the same pass (oracle/_ref: the reference's scoring /
driver headers compiled verbatim + restated estimators; falls back to the oracle port) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
# One hardware work queue per stream: the pipelines run one context (= stream) per host thread, and with the default of 8
# connections sixteen streams share queues, so a latency-bound kernel of one pair holds up another pair's launches
# (f4k_pairs 2.2 k -> 2.8 k pairs/s, e2k_pairs 330 -> 370).  Must be in the environment before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

# H-MSAC 21, Sampson-MSAC 35, Sampson-MAGSAC++ 43 flop per evaluation (counted from the reference lines)
FLOP_PER_EVAL = {("h", "msac"): 21.0, ("f", "msac"): 35.0, ("e", "msac"): 35.0, ("e", "magsac"): 43.0,
                 ("f", "magsac"): 43.0, ("h", "magsac"): 29.0}

WORKLOADS = {
    # name: problem, n points, outlier ratio, threshold, samples per round, pairs per GPU, rounds per step
    "f4k": dict(problem="f", n=4096, outliers=0.5, thr=0.75, samples=262144, pairs=1, rounds=128,
                desc="F 7-pt + MSAC, N=4096, 50% outliers, thr 0.75 px, single pair (BASELINE configs[1])"),
    # the north star's other scoring target ("0.5 of FP32 peak on 4k-point homography and fundamental scoring")
    "h4k": dict(problem="h", n=4096, outliers=0.5, thr=3.0, samples=262144, pairs=1, rounds=256,
                desc="H 4-pt + MSAC, N=4096, 50% outliers, thr 3.0 px, single pair (the homography twin of configs[1])"),
    "h2k_batch": dict(problem="h", n=2048, outliers=0.3, thr=3.0, samples=1024, pairs=2048, rounds=16,
                      desc="batched H 4-pt + MSAC, 2048 pairs/GPU x N=2048, 30% outliers, thr 3.0 px, K=1024 + argmax + final "
                           "least-squares refit per pair (configs[3])"),
    # configs[2]: E, 5-point + MAGSAC++ on K-normalised correspondences (threshold / focal length), 85 % outliers
    "e2k": dict(problem="e", scoring="magsac", n=2048, outliers=0.85, thr=0.75 / 1000.0, samples=65536, pairs=1, rounds=64,
                desc="E 5-pt (Nister) + MAGSAC++, N=2048, 85% outliers, known K, thr 0.75 px / f, single pair (configs[2])"),
    # full pipelines (adaptive stop, LO, graph-cut, IRLS, final fit) through pygcransac.find*: image pairs / s
    "f4k_pairs": dict(problem="f", n=4096, outliers=0.5, thr=0.75, pairs=128, pipeline=True,
                      desc="findFundamentalMatrix full default pipeline (7-pt + MSAC, LO + graph-cut with BA refits, DEGENSAC, "
                           "post-run BA, conf 0.99), N=4096, 50% "
                           "outliers, thr 0.75 px, distinct synthetic pairs (configs[1] pipeline)"),
    "h2k_pairs": dict(problem="h", n=2048, outliers=0.3, thr=3.0, pairs=128, pipeline=True,
                      desc="findHomography full pipeline (4-pt + MSAC, LO + graph-cut, conf 0.99), N=2048, 30% outliers, "
                           "thr 3.0 px, distinct synthetic pairs"),
    # BASELINE configs[0]: the reference's own test call on its bundled image pair (tests/test_homography.py:75-85:
    # threshold 3.0, P-NAPSAC sampler, max_iters 5000), every call a fresh seed; the SIFT matches are the committed fixture
    "cfg1_pairs": dict(problem="h", n=762, outliers=None, thr=3.0, pairs=128, pipeline=True, max_iters=5000, sampler=2,
                       fixture="cfg1_matches.npz",
                       desc="findHomography on the bundled tests/test_data pair (762 SIFT matches), threshold 3.0 px, the "
                            "reference's test call (P-NAPSAC sampler, max_iters 5000, conf 0.99), one call per seed (configs[0])"),
    "e2k_pairs": dict(problem="e", n=2048, outliers=0.85, thr=0.75, pairs=64, pipeline=True, max_iters=100000, scoring="magsac",
                      desc="findEssentialMatrix full pipeline with MAGSAC++ scoring (GCRANSAC<E estimator, grid, "
                           "MAGSACScoringFunction>: 5-pt, Stewenius refits in LO, conf 0.99, max_iters 100000), N=2048, 85% "
                           "outliers, known K, thr 0.75 px, distinct synthetic pairs (configs[2] pipeline)"),
    # configs[4]: ONE problem, its hypotheses sharded over the ranks (strong scaling), NCCL all-gather of the winner records
    "f1m_sharded": dict(problem="f", n=1048576, outliers=0.5, thr=0.75, samples=1048576, pairs=1, shard="hyps", rounds=1,
                        desc="single F problem, N=1048576 x K=1048576 7-pt samples sharded over the GPUs, MSAC, "
                             "NCCL all-gather of the (FP64 score, index, model) winner records (configs[4])"),
}
# what the default run adds to the headline line's `extra` block: (workload, steps cap)
EXTRA_ROUNDS = [("h4k", 8), ("e2k", 8), ("h2k_batch", 8), ("f1m_sharded", 3)]
EXTRA_PIPELINES = ["cfg1_pairs", "f4k_pairs", "h2k_pairs", "e2k_pairs"]


def fixture_sizes(wl):
    """(h1, w1, h2, w2) of a fixture workload's image pair, else the synthetic 1024 x 1024."""
    if wl.get("fixture"):
        g = np.load(os.path.join(ROOT, "tests", "golden", wl["fixture"]))
        return tuple(int(v) for v in g["size"])
    return (1024, 1024, 1024, 1024)


def make_points(wl, rank):
    from pygcransac import synth
    if wl.get("fixture"):
        g = np.load(os.path.join(ROOT, "tests", "golden", wl["fixture"]))
        one = np.ascontiguousarray(g["pts"], dtype=np.float64)
        return np.ascontiguousarray(np.stack([one] * wl["pairs"]))
    pts = []
    distinct = min(wl["pairs"], 128)         # a larger batch repeats 128 distinct synthetic pairs with a small per-copy shift
    for p in range(distinct):
        seed = (0 if wl.get("shard") == "hyps" else 1000 * rank) + p + 1   # a sharded problem is the SAME on every rank
        if wl["problem"] == "f":
            pts.append(synth.epipolar_pair(wl["n"], wl["outliers"], seed)[0])
        elif wl["problem"] == "e" and wl.get("pipeline"):
            pts.append(synth.epipolar_pair(wl["n"], wl["outliers"], seed)[0])       # pixels: the API normalises by K itself
        elif wl["problem"] == "e":
            # normalizeCorrespondences (py/gcransac_python.cpp:2519-2600): x' = K^-1 (x, y, 1)
            px = synth.epipolar_pair(wl["n"], wl["outliers"], seed)[0]
            c, f = synth.SIZE / 2, synth.FOCAL
            pts.append((px - c) / f)
        else:
            pts.append(synth.homography_pair(wl["n"], wl["outliers"], seed)[0])
    out = np.empty((wl["pairs"],) + pts[0].shape)
    for p in range(wl["pairs"]):
        out[p] = pts[p % distinct] + (p // distinct) * 1e-3
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, sm_load, mx, power, reasons = [], [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
                if len(r) > 9 and float(r[9]) >= 50.0:
                    sm_load.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        use = sm_load if sm_load else sm
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(sm_load),
                "power_w_max": max(power) if power else None}


def k3_dram_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K3 launch, from the committed `ncu --set full` capture of this
    workload (the newest profiles/r*_ncu_k3_*.txt; scripts/gpu_pass.sh regenerates it).  None if there is no capture."""
    name = {"f4k": "f4k", "h4k": "h4k", "h2k_batch": "h2k", "e2k": "e2k"}.get(workload)
    if name is None:
        return None
    try:
        total = 0.0
        import glob
        for line in open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_k3_%s.txt" % name)))[-1]):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[2]]
        return total
    except (OSError, ValueError, KeyError, IndexError):
        return None


# ---------------------------------------------------------------------------------------------- CPU arms ---------
_CPU_IMPL = {}


def cpu_samples_per_core(wl):
    """Bounded CPU sample: ~1 s of single-core work per step (65536 7-point / 8192 4-point samples at N = 4096, scaled by N)."""
    base = 65536 if wl["problem"] == "f" else 8192                    # 5-point: ~10 models of FP64 Sturm solving per sample
    return max(8, int(base * 4096 / wl["n"]))


def cpu_round_worker(args):
    """One CPU slice: `count` samples of the same workload through the reference scoring loop (or the port)."""
    kind, problem, pts, seed, first, count, thr, scoring = args
    if kind not in _CPU_IMPL:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        _CPU_IMPL[kind] = (O, O.Ref() if kind == "reference" else O.Oracle())
    O, impl = _CPU_IMPL[kind]
    score, idx, scored, _ = impl.round({"f": O.F, "h": O.H, "e": O.E}[problem], O.MAGSAC if scoring == "magsac" else O.MSAC,
                                       pts, seed, first, count, thr)
    return scored, score, idx


def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build_oracle()
    return "reference" if O.build_ref() else "port"


class CpuArm:
    """The reference's single-threaded CPU path, one slice of the sample stream per host core (persistent pool)."""

    def __init__(self, wl, pts0):
        import multiprocessing as mp
        self.kind = cpu_kind()
        self.cores = os.cpu_count() or 1
        self.wl, self.pts0 = wl, pts0
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.run(self.cores * 8)  # load the library in every worker

    def run(self, samples_total, first=0):
        """Returns (evals, wall seconds, samples actually run)."""
        per = max(1, samples_total // (self.cores * 4))
        jobs = [(self.kind, self.wl["problem"], self.pts0, 12345, first + c * per, per, self.wl["thr"],
                 self.wl.get("scoring", "msac"))
                for c in range(self.cores * 4)]
        t0 = time.perf_counter()
        out = self.pool.map(cpu_round_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out) * self.wl["n"], wall, per * len(jobs)

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline_rounds(wl, pts0, seconds_scale=10):
    """cpu_baseline object of a round workload: the reference scoring loop on a bounded sample, all host cores."""
    arm = CpuArm(wl, pts0)
    c_evals, c_wall, used = arm.run(cpu_samples_per_core(wl) * arm.cores * seconds_scale)
    arm.close()
    return {"value": c_evals / c_wall, "unit": "evals/s", "cores": arm.cores, "kind": arm.kind,
            "sample": "%d samples of pair 0 (solve + %s over all %d points), sliced over the host cores, %.1f s wall" % (
                used, wl.get("scoring", "msac").upper(), wl["n"], c_wall)}


def reference_rounds(args, wl):
    """--impl reference for a round workload: the CPU arm on this box's host cores."""
    pts0 = make_points(dict(wl, pairs=1), 0)[0]
    arm = CpuArm(wl, pts0)
    per_step = cpu_samples_per_core(wl) * arm.cores                 # ~1 s of all-core CPU work per step
    for s in range(args.warmup):
        arm.run(per_step // 8)
    t_total, evals_total, used = 0.0, 0, 0
    for s in range(args.steps):
        evals, wall, used = arm.run(per_step, first=s * per_step)
        t_total += wall
        evals_total += evals
    arm.close()
    value = evals_total / t_total
    return {
        "impl": "reference", "metric": "hypothesis x point evals/s", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "samples_per_step": used},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": "%d samples/step of the same pair (minimal solve + scoring over all %d points), sliced "
                                   "over the host cores" % (used, wl["n"])},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def _pairs_cpu_worker(args):
    """One image pair through the reference's verbatim GCRANSAC<>::run on one host core."""
    problem, pts, thr, weight, seed, max_iters, sampler, sizes, magsac = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    if "ref" not in _CPU_IMPL:
        _CPU_IMPL["ref"] = O.Ref()
    eps = np.finfo(np.float64).eps if problem == "h" else 0.0
    h1, w1, h2, w2 = sizes
    cs = np.array([(w1 + eps) / 4.0, (h1 + eps) / 4.0, (w2 + eps) / 4.0, (h2 + eps) / 4.0])
    t0 = time.perf_counter()
    if sampler == 2:                                                             # configs[0]: the reference's own P-NAPSAC class
        out = _CPU_IMPL["ref"].run(O.H, pts, cs, 4, thr, conf=0.99, spatial_weight=weight, max_iters=max_iters, min_iters=50,
                                   lo_number=50, seed=seed, sampler=2, image_sizes=(w1, h1, w2, h2))
        return out["inliers"], out["stats"]["iterations"], time.perf_counter() - t0
    if problem == "e":
        from pygcransac import synth
        npts = (pts - synth.SIZE / 2) / synth.FOCAL                              # K^-1 normalisation (utils.h:65-110)
        out = _CPU_IMPL["ref"].run(O.E, npts, cs, 4, thr / synth.FOCAL, conf=0.99, spatial_weight=weight, max_iters=max_iters,
                                   min_iters=50, lo_number=50, seed=seed, grid_pts=pts, scoring_kind=2 if magsac else 0, f_ba=4)
        return out["inliers"], out["stats"]["iterations"], time.perf_counter() - t0
    # F: the reference's default pipeline on both arms -- BA non-minimal solver, DEGENSAC, post-run bundle adjustment
    out = _CPU_IMPL["ref"].run(O.H if problem == "h" else O.F, pts, cs, 4, thr, conf=0.99, spatial_weight=weight,
                               max_iters=max_iters, min_iters=50, lo_number=50, seed=seed, use_degensac=problem == "f",
                               f_ba=3 if problem == "f" else 0)
    return out["inliers"], out["stats"]["iterations"], time.perf_counter() - t0


PIPELINE_WEIGHT = {"h": 0.975, "f": 0.0, "e": 0.1}          # the Python defaults (py/bindings.cpp:1147-1178)


def pairs_line_base(args, wl, world):
    return {"metric": "image pairs/s", "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "data": "real (the reference's bundled image pair, committed SIFT matches)" if wl.get("fixture") else "synthetic",
            "config": {"workload": wl["desc"], "pairs_per_step_per_gpu": wl["pairs"],
                       "spatial_coherence_weight": PIPELINE_WEIGHT[wl["problem"]]}}


def reference_pairs(args, wl, steps=None, warmup=None):
    """The reference's verbatim GCRANSAC<>::run (oracle/_ref) on the workload's pairs, one pair per host core."""
    import multiprocessing as mp
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    weight, max_iters, n_pairs = PIPELINE_WEIGHT[wl["problem"]], wl.get("max_iters", 10000), wl["pairs"]
    pts = make_points(wl, 0)
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores)
    jobs = lambda s: [(wl["problem"], pts[p], wl["thr"], weight, 100 * s + p + 1, max_iters, wl.get("sampler", 0),
                       fixture_sizes(wl), wl.get("scoring") == "magsac") for p in range(n_pairs)]
    for s in range(warmup):
        pool.map(_pairs_cpu_worker, jobs(s)[:cores], chunksize=1)
    t0 = time.perf_counter()
    outs = []
    for s in range(steps):
        outs += pool.map(_pairs_cpu_worker, jobs(warmup + s), chunksize=1)
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    value = n_pairs * steps / wall
    line = pairs_line_base(args, wl, args.gpus)
    line.update({"impl": "reference", "value": value, "ms_per_step": 1e3 * wall / steps, "dtype": "f64", "steps": steps,
                 "warmup": warmup,
                 "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference",
                                  "sample": "%d pairs/step, one pair per host core; mean %.1f ms per pair on one core, "
                                            "mean %d iterations" % (n_pairs, 1e3 * np.mean([o[2] for o in outs]),
                                                                    np.mean([o[1] for o in outs]))},
                 "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return line


# ---------------------------------------------------------------------------------------------- GPU session ------
class Session:
    """torch / torch.distributed plumbing, set up once per process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
        self._peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        if self.world == 1:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()

    def fp32_peak(self, ctx):
        """Measured FP32-FMA roofline of this GPU (gcr_ubench_fp32, pure FFMA2 chains for >= 300 ms right before the timed
        regions) + the ceiling of K3's own instruction mix; the derived figure beside it."""
        if self._peak is None:
            info = ctx.device_info()
            sustained, burst, ms = ctx.ubench_fp32(0, 300.0)
            mix_sustained, _, _ = ctx.ubench_fp32(1, 100.0)
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    mhz = float(json.load(f)["sm_max_mhz"])
            except (OSError, KeyError, ValueError):
                mhz = info["sm_clock_khz"] / 1e3
            self._peak = {"sustained": sustained, "burst": burst, "k3_mix_sustained": mix_sustained,
                          "derived": info["sm_count"] * 128 * 2 * mhz * 1e6 / 1e12,
                          "source": "measured in this process: gcr_ubench_fp32 (dependent FFMA2 chains, 4 CTAs x 256 threads per "
                                    "SM) back to back for %.0f ms -- the sustained figure; best single launch %.2f; derived %d SMs x "
                                    "128 x 2 x %.0f MHz beside it (MEASURED_PEAKS.json has no FP32 figure)" % (
                                        ms, burst, info["sm_count"], mhz)}
        return self._peak

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def gpu_rounds(sess, args, name, wl, steps, warmup, with_cpu, headline):
    """A fixed-K round workload on this rank's GPU: resident leg, e2e leg, K3 roofline.  Returns the JSON line (rank 0) or None."""
    torch = sess.torch
    from pygcransac import _cabi
    rank, world, local_rank = sess.rank, sess.world, sess.local_rank
    prob = {"f": _cabi.F, "h": _cabi.H, "e": _cabi.E}[wl["problem"]]
    scoring = _cabi.MAGSAC if wl.get("scoring") == "magsac" else _cabi.MSAC
    flags = _cabi.SCORE_EXACT_COUNTS if args.exact_counts else 0
    rounds = args.rounds if args.rounds > 0 else wl["rounds"]
    ctx = _cabi.Context(local_rank)
    # a real (non-default) stream: handle 0 would mean "restore the context's own stream" to gcr_ctx_set_stream, and
    # the CUDA events below must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.variant >= 0:
        ctx.set_score_variant(args.variant)
    if args.k3_cap > 0:
        ctx.set_tuning(_cabi.TUNE_K3_CTAS_PER_SM, args.k3_cap)
    peak = sess.fp32_peak(ctx)

    pts_host = torch.from_numpy(make_points(wl, rank)).pin_memory()          # [pairs, n, 4] f64, pinned
    pts_dev = pts_host.cuda(non_blocking=True)
    n_pairs, n = wl["pairs"], wl["n"]
    keys = torch.zeros(n_pairs, dtype=torch.int64, device="cuda")
    best = torch.zeros(n_pairs, 9, dtype=torch.float64, device="cuda")
    keys_host = torch.zeros(n_pairs, dtype=torch.int64).pin_memory()
    best_host = torch.zeros(n_pairs, 9, dtype=torch.float64).pin_memory()
    flush = sess.flush
    torch.cuda.synchronize()

    hyps_sharded = wl.get("shard") == "hyps"
    if hyps_sharded:
        from pygcransac import sharded
        winners = torch.zeros(n_pairs, 16, dtype=torch.int64, device="cuda")      # gcr_winner records (128 B each)
        winners_host = torch.zeros(n_pairs, 16, dtype=torch.int64).pin_memory()
        sharded.init_comm(ctx, rank, world)

    # configs[3] = "fixed K hypotheses per pair + argmax + one final least-squares refit" (SURVEY.md 8d): the refit tail
    # for every pair of the batch, device-resident, two launches (inliers of the winners, one fit per pair)
    refit_tail = n_pairs > 1 and not wl.get("shard")
    if refit_tail:
        t_idx = torch.empty(n_pairs, n, dtype=torch.int32, device="cuda")
        t_cnt = torch.empty(n_pairs, dtype=torch.int32, device="cuda")
        t_sc = torch.empty(n_pairs, dtype=torch.float64, device="cuda")
        t_refit = torch.empty(n_pairs, 9, dtype=torch.float64, device="cuda")
        t_ok = torch.empty(n_pairs, dtype=torch.uint8, device="cuda")

    def tail(c, lo, hi):
        c.final_refit_pairs(prob, best[lo:hi].data_ptr(), wl["thr"], t_idx[lo:hi].data_ptr(), t_cnt[lo:hi].data_ptr(),
                            t_sc[lo:hi].data_ptr(), t_refit[lo:hi].data_ptr(), t_ok[lo:hi].data_ptr())

    def one_round(i, fl=None):
        fl = flags if fl is None else fl
        if refit_tail:
            ev = ctx.round_device(prob, scoring, _cabi.UNIFORM, i * wl["samples"], wl["samples"], 12345, wl["thr"], fl, 0,
                                  keys.data_ptr(), best.data_ptr())
            tail(ctx, 0, n_pairs)
            return ev
        if hyps_sharded:
            return ctx.round_sharded_device(prob, scoring, _cabi.UNIFORM, i * wl["samples"], wl["samples"], 12345,
                                            wl["thr"], fl, 65536, winners.data_ptr())
        return ctx.round_device(prob, scoring, _cabi.UNIFORM, i * wl["samples"], wl["samples"], 12345, wl["thr"],
                                fl, 0, keys.data_ptr(), best.data_ptr())

    def step_resident(i, fl=None):
        return sum(one_round(i * rounds + r, fl) for r in range(rounds))

    # e2e for a batch of pairs: the pairs are independent, so the upload of one chunk of pairs overlaps the scoring of
    # the previous ones -- E2E_SIDE contexts (streams) take the chunks in turn (the user-level pattern for host-resident
    # batches); the FP32 upload entry point halves the bytes (the reference's own tests hand over float32 arrays)
    E2E_CHUNKS, E2E_SIDE = args.e2e_chunks, args.e2e_streams
    overlap = n_pairs >= 8 * E2E_CHUNKS and not hyps_sharded
    use_f32 = overlap and args.upload == "f32"
    if overlap:
        side = []
        for _ in range(E2E_SIDE):
            c2 = _cabi.Context(local_rank)
            s2 = torch.cuda.Stream()
            c2.set_stream(s2.cuda_stream)
            if args.variant >= 0:
                c2.set_score_variant(args.variant)
            if args.k3_cap > 0:
                c2.set_tuning(_cabi.TUNE_K3_CTAS_PER_SM, args.k3_cap)
            side.append((c2, s2))
        chunk = n_pairs // E2E_CHUNKS
        pts_host32 = pts_host.to(torch.float32).pin_memory() if use_f32 else None
    h2d_bytes_per_round = int(pts_host.numel() * (4 if use_f32 else 8))

    # ONE pair, many rounds: consecutive rounds alternate between a few contexts (streams, model slots, workspaces), so the
    # solver kernel of round i + 1 (FP64 pipe, latency-bound) runs beside the scoring kernel of round i (FP32 FMA pipe) and the
    # upload of the next round's points hides behind both.  The resident leg stays on ONE stream so that K3's launch times
    # (the roofline) are undisturbed; its multi-stream figure is reported beside it (config.multi_stream_resident).
    two_lanes = (not overlap) and (not hyps_sharded) and n_pairs == 1 and args.lanes >= 2
    lanes = []
    if two_lanes:
        pairs_cs = [(ctx, stream)]
        for _ in range(args.lanes - 1):
            cB = _cabi.Context(local_rank)
            sB = torch.cuda.Stream()
            cB.set_stream(sB.cuda_stream)
            if args.variant >= 0:
                cB.set_score_variant(args.variant)
            if args.k3_cap > 0:
                cB.set_tuning(_cabi.TUNE_K3_CTAS_PER_SM, args.k3_cap)
            pairs_cs.append((cB, sB))
        for c_, s_ in pairs_cs:
            lanes.append({"ctx": c_, "stream": s_, "keys": torch.zeros(1, dtype=torch.int64, device="cuda"),
                          "best": torch.zeros(1, 9, dtype=torch.float64, device="cuda"),
                          "keys_host": torch.zeros(1, dtype=torch.int64).pin_memory(),
                          "best_host": torch.zeros(1, 9, dtype=torch.float64).pin_memory()})

    def lane_round(i, lane, with_copies):
        c_ = lane["ctx"]
        with torch.cuda.stream(lane["stream"]):
            if with_copies:
                upload(c_, 0, 1)
            ev = c_.round_device(prob, scoring, _cabi.UNIFORM, i * wl["samples"], wl["samples"], 12345, wl["thr"], flags, 0,
                                 lane["keys"].data_ptr(), lane["best"].data_ptr())
            if with_copies:
                lane["keys_host"].copy_(lane["keys"], non_blocking=True)
                lane["best_host"].copy_(lane["best"], non_blocking=True)
        return ev

    def step_lanes(i, with_copies):
        start = torch.cuda.Event()
        start.record(stream)
        for lane in lanes[1:]:
            lane["stream"].wait_event(start)
        ev = sum(lane_round(i * rounds + r, lanes[r % len(lanes)], with_copies) for r in range(rounds))
        for lane in lanes[1:]:
            done = torch.cuda.Event()
            done.record(lane["stream"])
            stream.wait_event(done)
        return ev

    def upload(c, lo, hi):
        if use_f32:
            c.check(c.lib.gcr_points_upload_f32(c.h, pts_host32[lo:hi].data_ptr(), n, 4, hi - lo))
        else:
            c.check(c.lib.gcr_points_upload(c.h, pts_host[lo:hi].data_ptr(), n, 4, hi - lo))

    def round_e2e(i, with_copies=True):
        """one round through the C ABI with HOST buffers: H2D of the points, the round, D2H of the result"""
        if not overlap:
            upload(ctx, 0, n_pairs)
            ev = one_round(i)
            if hyps_sharded:
                winners_host.copy_(winners, non_blocking=True)
            else:
                keys_host.copy_(keys, non_blocking=True)
                best_host.copy_(best, non_blocking=True)
            return ev
        ev = 0
        for c in range(E2E_CHUNKS):
            c2, s2 = side[c % E2E_SIDE]
            lo, hi = c * chunk, n_pairs if c == E2E_CHUNKS - 1 else (c + 1) * chunk
            with torch.cuda.stream(s2):
                if with_copies:
                    upload(c2, lo, hi)
                ev += c2.round_device(prob, scoring, _cabi.UNIFORM, i * wl["samples"], wl["samples"], 12345 + c, wl["thr"], flags, 0,
                                      keys[lo:hi].data_ptr(), best[lo:hi].data_ptr())
                tail(c2, lo, hi)
                if with_copies:
                    keys_host[lo:hi].copy_(keys[lo:hi], non_blocking=True)
                    best_host[lo:hi].copy_(t_refit[lo:hi], non_blocking=True)
        return ev

    def step_e2e(i, with_copies=True):
        if two_lanes:
            return step_lanes(i, with_copies)
        if not overlap:
            return sum(round_e2e(i * rounds + r) for r in range(rounds))
        # the side streams start after the step's start event and the step's end event waits for all of them; inside the
        # step every stream works through its chunks of every round back to back (a chunk's buffers are only ever touched
        # by its own stream, in order), so the upload of the next round's chunk overlaps the scoring of this round's
        start = torch.cuda.Event()
        start.record(stream)
        for c2, s2 in side:
            s2.wait_event(start)
        ev = sum(round_e2e(i * rounds + r, with_copies) for r in range(rounds))
        for c2, s2 in side:
            done = torch.cuda.Event()
            done.record(s2)
            stream.wait_event(done)
        return ev

    def timed(step_fn, first, nsteps):
        """nsteps steps, each bracketed by its own CUDA event pair on the launching stream; the L2 flush (a 256 MiB device
        memset) runs BETWEEN steps, outside the event pairs.  Returns (sum of step times in ms, slot evaluations)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        slot = 0
        for i, (a, b) in enumerate(evs):
            flush.zero_()
            a.record()
            slot += step_fn(first + i)
            b.record()
        stream.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs), slot

    # ---- resident leg: inputs in HBM before the timed region ----
    ctx.points_set_device(pts_dev.data_ptr(), n, 4, n_pairs)
    for i in range(warmup):
        flush.zero_()
        step_resident(i)
    sess.barrier()
    ctx.counters(reset=True)
    ctx.profile_enable(True)
    ctx.profile_read()
    launches0 = ctx.launch_count()
    clocks = ClockSampler(local_rank).start()
    sess.barrier()
    ms, slot_evals = timed(step_resident, warmup, steps)
    sess.barrier()
    launches = ctx.launch_count() - launches0
    score_ms, score_launches, _ = ctx.profile_read()
    rescored, max_err = ctx.select_stats()
    counters = ctx.counters(reset=True)
    valid_models, dropped = counters[0], counters[1]
    if dropped:
        raise SystemExit("bench.py: %d FP64 re-check evaluations were dropped (work list full)" % dropped)
    evals = valid_models * n

    # ---- K3 with FP64-exact per-model inlier counts (threshold-band re-check): the same kernel family, one short leg ----
    frac_counts = None
    if headline and not args.exact_counts and not hyps_sharded:
        ctx.profile_read()
        timed(lambda i: step_resident(i, _cabi.SCORE_EXACT_COUNTS), warmup + steps, max(1, min(steps, 3)))
        c_ms, _, _ = ctx.profile_read()
        c_valid = ctx.counters(reset=True)[0]
        frac_counts = (c_valid * n, c_ms)
    ctx.profile_enable(False)

    # ---- the resident pass once more with the rounds alternating between two streams (what the e2e leg does) ----
    ms_two = None
    if two_lanes:
        # with several streams in flight the scoring kernel keeps 3 of its 5 CTA slots per SM (gcr_set_tuning): alone it loses
        # ~1 % that way, but the other streams' solver / select / upload kernels then find room beside it (e2e -3 %)
        if args.k3_cap == 0:
            for lane in lanes:
                lane["ctx"].set_tuning(_cabi.TUNE_K3_CTAS_PER_SM, args.lanes_k3_cap)
        for lane in lanes[1:]:
            lane["ctx"].points_set_device(pts_dev.data_ptr(), n, 4, 1)
        for i in range(2):
            step_lanes(i, False)
        for lane in lanes:
            lane["ctx"].counters(reset=True)
        n2 = max(1, min(steps, 5))
        ms_two, _ = timed(lambda i: step_lanes(i, False), warmup + steps, n2)
        ev_two = sum(lane["ctx"].counters(reset=True)[0] for lane in lanes) * n
        ms_two = (ms_two / (n2 * rounds), ev_two / (ms_two * 1e-3))

    if overlap and E2E_CHUNKS == E2E_SIDE:
        # a batch of pairs: the same, with the chunks of pairs the e2e leg uses (every side context keeps its chunk resident)
        for c in range(E2E_CHUNKS):
            lo, hi = c * chunk, n_pairs if c == E2E_CHUNKS - 1 else (c + 1) * chunk
            with torch.cuda.stream(side[c][1]):
                side[c][0].points_set_device(pts_dev[lo:hi].data_ptr(), n, 4, hi - lo)
        torch.cuda.synchronize()
        for i in range(2):
            step_e2e(i, False)
        for c2, _ in side:
            c2.counters(reset=True)
        n2 = max(1, min(steps, 5))
        ms_two, _ = timed(lambda i: step_e2e(i, False), warmup + steps, n2)
        ev_two = sum(c2.counters(reset=True)[0] for c2, _ in side) * n
        ms_two = (ms_two / (n2 * rounds), ev_two / (ms_two * 1e-3))

    # ---- hardware proof of SURVEY 8(e) row 2: the N-rank winner record == the 1-rank record, bit for bit ----
    sharded_equal, rec = None, None
    if hyps_sharded:
        one_round(0)
        got = winners.clone()
        ctx.comm_destroy()                                              # without a communicator: the chunked 1-GPU pass
        one_round(0)
        stream.synchronize()
        sharded_equal = bool(torch.equal(got, winners))
        rec = got.cpu().numpy().reshape(-1).view(_cabi.WINNER_DTYPE)[0]
        if not sharded_equal:
            raise SystemExit("bench.py: rank %d: the %d-rank winner record differs from the single-GPU record" % (rank, world))
        sharded.init_comm(ctx, rank, world)

    # ---- e2e leg: host buffers, H2D + D2H inside the timed region ----
    for i in range(warmup):
        flush.zero_()
        step_e2e(i)
    sess.barrier()
    ctx.counters(reset=True)
    if overlap:
        for c2, _ in side:
            c2.counters(reset=True)
    if two_lanes:
        for lane in lanes[1:]:
            lane["ctx"].counters(reset=True)
    ms_e2e, _ = timed(step_e2e, warmup, steps)
    sess.barrier()
    clock_info = clocks.stop()
    evals_e2e = ctx.counters(reset=True)[0] * n
    if overlap:
        evals_e2e = sum(c2.counters(reset=True)[0] for c2, _ in side) * n
    if two_lanes:
        evals_e2e += sum(lane["ctx"].counters(reset=True)[0] for lane in lanes[1:]) * n

    # ---- max over ranks, sum of work ----
    ms, ms_e2e = sess.reduce([ms, ms_e2e], "max")
    evals_all, evals_e2e, launches, slot_evals = sess.reduce([evals, evals_e2e, launches, slot_evals], "sum")

    line = None
    if rank == 0:
        flop = FLOP_PER_EVAL[(wl["problem"], wl.get("scoring", "msac"))]
        # dominant kernel = K3 (k_score_main): algorithmic flops of the REAL models this rank scored / its own event time
        achieved = evals * flop / (score_ms * 1e-3) / 1e12 if score_ms > 0 else None
        total_rounds = steps * rounds
        line = {
            "metric": "hypothesis x point evals/s", "value": evals_all / (ms * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "strong" if hyps_sharded else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "name": name, "samples_per_round_per_pair": wl["samples"], "pairs_per_gpu": n_pairs,
                       "rounds_per_step": rounds, "ms_per_round": ms / total_rounds, "timed_region_s": ms * 1e-3,
                       "pairs_per_s": (world * n_pairs * total_rounds / (ms * 1e-3)) if n_pairs > 1 else None,
                       "pairs_per_s_e2e": (world * n_pairs * total_rounds / (ms_e2e * 1e-3)) if n_pairs > 1 else None,
                       "step": "%d x gcr_round (K1-K4) over fresh samples" % rounds +
                               (" + inliers of every pair's winner + one least-squares refit per pair" if refit_tail else "") +
                               (" (gcr_round_sharded: chunks of 65536 samples, FP64 winner records, ncclAllGather)" if hyps_sharded else ""),
                       "models_scored_per_round": evals_all / n / total_rounds, "model_slots_per_round": slot_evals / n / total_rounds,
                       "k3_exact_counts": bool(args.exact_counts),
                       "winner": "every slot within the FP32 margin of the top re-scored in FP64 (K4): %d re-scores in the timed "
                                 "region, max FP32 score error seen %.1e" % (rescored, max_err),
                       "k3_variant": args.variant,
                       "l2_flush": "256 MiB device memset between steps (outside the per-step event pairs); inside a step the "
                                   "points (<= 128 KB per pair) stay L2-resident by design"},
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak["sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peak["sustained"]) if achieved else None,
                         "frac_of_derived_peak": (achieved / peak["derived"]) if achieved else None,
                         "peak_burst": peak["burst"], "peak_derived": peak["derived"], "k3_mix_ceiling": peak["k3_mix_sustained"],
                         "traffic": k3_dram_traffic(name),
                         "traffic_note": "DRAM bytes of one K3 launch from the committed ncu --set full capture (profiles/): the "
                                         "points and models are L2-resident, the FMA pipe is the bound",
                         "kernel": "k_score_main (K3), without per-model inlier counts (the winner's exact count comes from K4)",
                         "flop_per_eval": flop, "k3_ms_per_launch": score_ms / max(1, score_launches),
                         "k3_share_of_step": score_ms / ms if world == 1 else None, "peak_source": peak["source"],
                         "note": "FP32-FMA-pipe bound (projective divide / Sampson denominator; not a tensor-core "
                                 "contraction, HBM traffic ~0): peak is the FP32 roofline, not hbm/tensor"},
            "clocks": clock_info,
            "e2e": {"value": evals_e2e / (ms_e2e * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": h2d_bytes_per_round * rounds,
                    "d2h_bytes_per_step": int(n_pairs * (128 if hyps_sharded else 8 + 72)) * rounds,
                    "ms_per_round": ms_e2e / total_rounds,
                    "h2d_gbs": h2d_bytes_per_round * total_rounds / (ms_e2e * 1e-3) / 1e9,
                    "overlap": ("%d chunks of pairs on %d contexts / streams: uploads overlap scoring; %s upload" % (
                        E2E_CHUNKS, E2E_SIDE, args.upload)) if overlap else
                        ("rounds alternate between %d contexts / streams:" % len(lanes) + " the solver kernel and the upload of round i + 1 run "
                         "beside the scoring kernel of round i (which keeps %d of its 5 CTA slots per SM meanwhile)" % (
                             args.k3_cap or args.lanes_k3_cap)) if two_lanes else "none"},
            "gpu_launches": int(launches),
        }
        if ms_two and n_pairs > 1:
            line["config"]["pairs_per_s_multi_stream"] = world * n_pairs / (ms_two[0] * 1e-3)
        if ms_two:
            line["config"]["multi_stream_resident"] = {"streams": len(lanes) if two_lanes else E2E_SIDE, "ms_per_round": ms_two[0],
                                                      "evals_per_s": ms_two[1] * world,  # (rank 0's rate x ranks: every rank runs the same pass)
                                                      "note": "the resident pass with consecutive rounds alternating between the streams (K2 of "
                                                              "round i + 1 beside K3 of round i), as the e2e leg runs; `value` and the roofline "
                                                              "are the one-stream pass, where K3's launch time is undisturbed"}
        if frac_counts:
            a_c = frac_counts[0] * flop / (frac_counts[1] * 1e-3) / 1e12 if frac_counts[1] > 0 else None
            line["roofline"]["frac_with_exact_counts"] = a_c / peak["sustained"] if a_c else None
        if hyps_sharded:
            line["config"]["sharded_equals_single_gpu"] = sharded_equal
            line["config"]["winner_record"] = {"score": float(rec["score"]), "index": int(rec["index"]), "count": int(rec["count"])}
        if with_cpu:
            line["cpu_baseline"] = cpu_baseline_rounds(wl, pts_host[0].numpy().copy(), 10 if headline else 3)
    if overlap:
        for c2, _ in side:
            c2.close()
    if two_lanes:
        for lane in lanes[1:]:
            lane["ctx"].close()
    ctx.close()
    return line


def gpu_pairs(sess, args, name, wl, steps, warmup):
    """Full-pipeline workloads: image pairs / s through pygcransac.find* (batched GPU driver: adaptive stop, LO with the
    reference's BK min-cut on the host, IRLS, final fit), a few host threads per GPU each with its own context / stream
    taking pairs from a queue; pairs shard over the ranks with no collective.  The API takes host arrays, so every call
    includes its H2D / D2H: value == e2e."""
    import pygcransac
    from pygcransac import synth
    rank, world, local_rank = sess.rank, sess.world, sess.local_rank
    weight, max_iters, n_pairs = PIPELINE_WEIGHT[wl["problem"]], wl.get("max_iters", 10000), wl["pairs"]
    pts = make_points(wl, rank)
    if wl["problem"] == "e":
        find = lambda c, *a, **kw: pygcransac.findEssentialMatrix(c, synth.K_MAT, synth.K_MAT, *a, **kw)
    else:
        find = pygcransac.findHomography if wl["problem"] == "h" else pygcransac.findFundamentalMatrix

    from concurrent.futures import ThreadPoolExecutor
    workers = args.workers if args.workers > 0 else max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    pygcransac._load_driver()
    contexts_before = set(pygcransac._contexts)
    pool = ThreadPoolExecutor(workers)
    extra = {}
    sizes = fixture_sizes(wl)

    slow = []

    def one_pair(job):
        s, p = job
        t_pair = time.perf_counter()
        M, mask = find(pts[p], *sizes, [], threshold=wl["thr"], conf=0.99,
                       spatial_coherence_weight=weight, max_iters=max_iters, min_iters=50, use_sprt=False, sampler=wl.get("sampler", 0),
                       neighborhood=0, neighborhood_size=4.0, lo_number=50, seed=100 * s + p + 1, device=local_rank,
                       scoring=1 if wl.get("scoring") == "magsac" else 0)
        st = pygcransac._tls.last_statistics
        for k in ("exact_rescores", "gpu_batches", "graph_cut_number", "lo_samples"):
            extra[k] = extra.get(k, 0) + st[k]                          # (approximate under threads: diagnostics only)
        extra["graph_cut_seconds"] = extra.get("graph_cut_seconds", 0.0) + st["graph_cut_seconds"]
        if os.environ.get("GCR_BENCH_SLOW_PAIRS") and s >= warmup:
            slow.append((round(1e3 * (time.perf_counter() - t_pair), 1), s, p, st["iteration_number"], st["degensac_runs"], st["gpu_batches"],
                         st["graph_cut_number"], [round(1e3 * v, 1) for v in st["phase_seconds"]]))
        return int(mask.sum()), st["iteration_number"]

    batch_api = args.pairs_api == "batch"
    if batch_api and world > 1:   # the ranks of one node share its host cores
        os.environ.setdefault("GCR_BATCH_THREADS", str(max(2, (os.cpu_count() or 2) // world - 2)))
    if wl["problem"] == "e":
        find_batch = lambda c, *a, **kw: pygcransac.findEssentialMatrixBatch(c, synth.K_MAT, synth.K_MAT, *a, **kw)
    else:
        find_batch = pygcransac.findHomographyBatch if wl["problem"] == "h" else pygcransac.findFundamentalMatrixBatch

    def one_pass(s):
        """batch: ONE call for the step's pairs -- they sit on the GPU together and advance in lock-step, every kind of
        device operation one launch for all of them (host/gcr_batch_engine.h).  threads: `workers` host threads, each with
        its own gcr_ctx / stream, take the pairs of the step from a queue (pygcransac.find*, one pair per call)."""
        if batch_api:
            seeds = np.array([100 * s + p + 1 for p in range(n_pairs)], dtype=np.uint64)
            M, masks, counts = find_batch(pts, *sizes, threshold=wl["thr"], conf=0.99, spatial_coherence_weight=weight,
                                          max_iters=max_iters, min_iters=50, use_sprt=False, sampler=wl.get("sampler", 0), neighborhood=0,
                                          neighborhood_size=4.0, lo_number=50, seeds=seeds, device=local_rank,
                                          scoring=1 if wl.get("scoring") == "magsac" else 0, seats=args.seats, engines=args.engines)
            stats = pygcransac.last_batch_statistics
            for k in ("exact_rescores", "gpu_batches", "graph_cut_number", "lo_samples"):
                extra[k] = extra.get(k, 0) + sum(getattr(stats[p], k) for p in range(n_pairs))
            extra["graph_cut_seconds"] = extra.get("graph_cut_seconds", 0.0) + sum(stats[p].graph_cut_seconds for p in range(n_pairs))
            return int(counts.sum()), sum(stats[p].iteration_number for p in range(n_pairs))
        out = list(pool.map(one_pair, [(s, p) for p in range(n_pairs)]))
        return sum(o[0] for o in out), sum(o[1] for o in out)

    for s in range(warmup):
        one_pass(s)
    launch_total = lambda: (sum(c.launch_count() for (d, _), c in list(pygcransac._contexts.items()) if d == local_rank) +
                            pygcransac.batch_launch_count())
    launches0 = launch_total()
    clocks = ClockSampler(local_rank).start()
    sess.barrier()
    # no cyclic garbage collection inside the timed passes: a full collection holds the GIL for ~100 ms once the process has
    # run a few workloads, and sixteen threads that each need the GIL between two 0.3 ms calls all stand still for it
    import gc
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    inl = its = 0
    step_ms = []
    for s in range(steps):
        ts = time.perf_counter()
        a, b = one_pass(warmup + s)
        step_ms.append(round(1e3 * (time.perf_counter() - ts), 2))
        inl, its = inl + a, its + b
    sess.barrier()
    wall = time.perf_counter() - t0
    gc.enable()
    clock_info = clocks.stop()
    launches = launch_total() - launches0
    pool.shutdown()
    for key in set(pygcransac._contexts) - contexts_before:      # the contexts of this pool's threads (one per thread)
        pygcransac._contexts.pop(key).close()
    if slow:
        for row in sorted(slow, reverse=True)[:8]:
            print("slow pair: ms %s step %d pair %d iterations %d degensac_runs %d batches %d graph cuts %d phases %s" % row, file=sys.stderr)
    wall, = sess.reduce([wall], "max")
    total_pairs, launches = sess.reduce([n_pairs * steps, launches], "sum")
    if rank != 0:
        return None
    value = total_pairs / wall
    line = pairs_line_base(args, wl, world)
    line.update({"value": value, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * wall / steps,
                 "dtype": "f32 ranking / f64 decisions", "clocks": clock_info, "gpu_launches": int(launches),
                 "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": int(pts.nbytes),
                         "d2h_bytes_per_step": int(n_pairs * (72 + wl["n"]))},
                 "roofline": None})
    line["config"].update({"name": name, "per_pair_mean": {k: v / (n_pairs * (steps + warmup)) for k, v in extra.items()}})
    line["config"].update({"pairs_api": "find*Batch: one call per step, lock-step multi-pair driver" if batch_api else
                           "find*: one pair per call from a pool of host threads"})
    line["config"].update({"step_ms": step_ms, "host_threads_per_gpu": workers, "mean_inliers": inl / (n_pairs * steps), "mean_iterations": its / (n_pairs * steps),
                           "timing": "host wall clock around synchronous API calls (each returns host results), "
                                     "max over ranks"})
    return line


def compact(line):
    """What an `extra` entry keeps of a workload's full line."""
    out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"],
           "steps": line["steps"], "scaling": line["scaling"], "e2e": line["e2e"]["value"], "workload": line["config"]["workload"],
           "gpu_launches": line.get("gpu_launches")}
    cfg = line["config"]
    for k in ("rounds_per_step", "ms_per_round", "timed_region_s", "pairs_per_s", "pairs_per_s_e2e", "pairs_per_s_multi_stream", "multi_stream_resident", "step_ms", "pairs_api", "sharded_equals_single_gpu",
              "winner_record", "mean_inliers", "mean_iterations", "host_threads_per_gpu", "models_scored_per_round"):
        if cfg.get(k) is not None:
            out[k] = cfg[k]
    if line["e2e"].get("h2d_gbs") is not None:
        out["e2e_h2d_gbs"] = line["e2e"]["h2d_gbs"]
    if line.get("roofline"):
        r = line["roofline"]
        out["roofline"] = {k: r[k] for k in ("achieved", "peak", "frac", "frac_of_derived_peak", "k3_ms_per_launch", "k3_share_of_step",
                                             "flop_per_eval")}
    if line.get("clocks"):
        out["clocks"] = {k: line["clocks"].get(k) for k in ("sm_mhz", "reasons", "samples_under_load")}
    if line.get("cpu_baseline"):
        out["cpu_baseline"] = line["cpu_baseline"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS),
                    help="one workload only; default: the f4k headline + every other configuration in `extra`")
    ap.add_argument("--no-extras", action="store_true", help="default run without the `extra` block")
    ap.add_argument("--samples", type=int, default=0, help="minimal samples per round per pair (default: workload's)")
    ap.add_argument("--rounds", type=int, default=0, help="rounds per step (default: workload's, sized for a >= 1 s timed region)")
    ap.add_argument("--variant", type=int, default=-1, help="force a K3 kernel variant (see gcr_set_score_variant)")
    ap.add_argument("--upload", default="f32", choices=["f32", "f64"], help="batched e2e leg: host array type handed to the C ABI")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="batched e2e leg: chunks of pairs per round")
    ap.add_argument("--e2e-streams", type=int, default=4, help="batched e2e leg: contexts / streams taking the chunks in turn")
    ap.add_argument("--exact-counts", action="store_true",
                    help="K3 also produces FP64-exact per-model inlier counts (threshold-band re-check); not an output of the "
                         "round -- the winner's FP64 score / exact count come from K4 either way")
    ap.add_argument("--k3-cap", type=int, default=0, help="experiment: resident K3 CTAs per SM (0 = all that fit)")
    ap.add_argument("--lanes-k3-cap", type=int, default=3,
                    help="K3 CTAs per SM while the rounds alternate between several streams (multi-stream resident pass, e2e leg)")
    ap.add_argument("--lanes", type=int, default=4, choices=[1, 2, 3, 4],
                    help="single-pair workloads, e2e leg: consecutive rounds alternate between this many contexts / streams")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--pairs-api", default="threads", choices=["batch", "threads"],
                    help="pipeline workloads: pygcransac.find*Batch (one call per step) or find* from a pool of host threads")
    ap.add_argument("--pairs", type=int, default=0, help="pipeline workloads: image pairs per step per GPU (default: workload's)")
    ap.add_argument("--seats", type=int, default=0, help="batch API: pairs resident on the GPU at a time (0 = library default)")
    ap.add_argument("--engines", type=int, default=0, help="batch API: contexts / dispatcher threads (0 = library default)")
    ap.add_argument("--workers", type=int, default=0, help="pipeline workloads: host threads (one gcr_ctx each) per GPU")
    args = ap.parse_args()
    name = args.workload or "f4k"
    wl = dict(WORKLOADS[name])
    if args.samples:
        wl["samples"] = args.samples
    if args.pairs and wl.get("pipeline"):
        wl["pairs"] = args.pairs
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        line = reference_pairs(args, wl) if wl.get("pipeline") else reference_rounds(args, wl)
        print(json.dumps(line))
        return

    sess = Session()
    with_cpu = not args.no_cpu and sess.world == 1            # the CPU legs run on rank 0 at N = 1 only (contract)
    # Every CPU arm runs AFTER all the GPU work of the process: the arms fork one worker per host core, and a process that
    # has forked pays for it afterwards (copy-on-write faults on every page it touches again) -- the multi-threaded
    # pipelines showed passes of 130 ms among passes of 42 ms when a reference arm had run before them.
    cpu_jobs = []
    if wl.get("pipeline"):
        line = gpu_pairs(sess, args, name, wl, args.steps, args.warmup)
    else:
        line = gpu_rounds(sess, args, name, wl, args.steps, args.warmup, with_cpu=False, headline=True)
        if with_cpu and rank == 0:
            cpu_jobs.append(("rounds", line, wl, 10, None))
    if not args.workload and not args.no_extras:
        extra = {}
        for ename, cap in EXTRA_ROUNDS:
            t0 = time.perf_counter()
            ewl = dict(WORKLOADS[ename])
            l2 = gpu_rounds(sess, args, ename, ewl, max(1, min(args.steps, cap)), min(args.warmup, 2), with_cpu=False, headline=False)
            if rank == 0:
                extra[ename] = compact(l2)
                extra[ename]["bench_wall_s"] = time.perf_counter() - t0
                if with_cpu:
                    cpu_jobs.append(("rounds", extra[ename], ewl, 3, None))
        for ename in EXTRA_PIPELINES:
            t0 = time.perf_counter()
            ewl = dict(WORKLOADS[ename])
            l2 = gpu_pairs(sess, args, ename, ewl, 6, 4)
            if rank == 0:
                extra[ename] = compact(l2)
                extra[ename]["bench_wall_s"] = time.perf_counter() - t0
                if with_cpu:
                    cpu_jobs.append(("pairs", extra[ename], ewl, 0, l2["value"]))
            sess.barrier()
        if rank == 0:
            line["extra"] = extra
            # the other half of BASELINE's metric, gathered where a reader of the headline line looks first
            line["config"]["pairs_per_s"] = {k: (v.get("pairs_per_s") if v["unit"] == "evals/s" else v["value"])
                                             for k, v in extra.items() if v.get("pairs_per_s") or v["unit"] == "pairs/s"}
            line["config"]["pairs_per_s_e2e"] = {k: (v.get("pairs_per_s_e2e") if v["unit"] == "evals/s" else v["e2e"])
                                                 for k, v in extra.items() if v.get("pairs_per_s_e2e") or v["unit"] == "pairs/s"}
    for kind, target, jwl, scale, gpu_value in cpu_jobs:
        if kind == "rounds":
            target["cpu_baseline"] = cpu_baseline_rounds(jwl, make_points(dict(jwl, pairs=1), 0)[0].copy(), scale)
        else:
            ref = reference_pairs(args, jwl, steps=1, warmup=1)   # (one warm-up pass of one pair per core: the workers start warm)
            target["cpu_baseline"] = ref["cpu_baseline"]
            target["speedup_vs_cpu_baseline"] = gpu_value / ref["value"]
    if rank == 0:
        print(json.dumps(line))
    sess.close()


if __name__ == "__main__":
    main()
