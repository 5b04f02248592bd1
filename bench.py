#!/usr/bin/env python
"""bench.py -- aligned pairs/second through the overlap-maximisation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload blj256|lj38]
                    [--impl ours|reference] [--pairs P]

One "step" = one pass of the hot path (coords -> best grid index / displacement or rotation)
over one batch of P synthetic pairs per GPU.  Output: ONE JSON line on rank 0.

  value        whole-job pairs/s with the coordinates already resident in HBM (device API,
               timed with CUDA events on the stream the kernels run on, max over ranks)
  e2e          same metric through the host-buffer C-ABI call (fo_*_align_pairs): the H2D copy of
               every step's coordinates and the D2H copy of its results are inside the timed region
  roofline     dominant kernel: executed FP64 flop / CUDA-event duration vs the measured DFMA peak
  cpu_baseline the CPU oracle (C restatement of the reference) on a bounded sample, all host cores

--impl reference times the reference's CPU algorithm (the oracle port; the reference itself is
Fortran/f2py + numpy and cannot be built/run on the GPU box) on the host cores for the same
workload.  Under torchrun only rank 0 runs it.
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
sys.path.insert(0, ROOT)

METRIC = "aligned pairs/sec"
BOX_BLJ = 5.975206329


# ----------------------------------------------------------------------------- workloads

class Blj256:
    """BASELINE.json configs[1] shape (examples/BLJ256: N=256, 204 A + 52 B, box 5.975206329,
    default sigma, n=9, F=40) batched as configs[4] (independent synthetic pairs, SURVEY 8d C5:
    partner = base + uniform random translation + N(0,0.05^2) jitter, wrapped, permuted within
    species, seed 256)."""
    name = "blj256"
    natoms = 256
    default_pairs = 16384

    def __init__(self, nwave=9):
        g = np.load(os.path.join(ROOT, "tests", "golden", "periodic_blj256.npz"))
        self.base = g["pos1"]
        self.box = np.ones(3) * BOX_BLJ
        self.perm = [np.arange(204), np.arange(204, 256)]
        self.n = nwave
        self.F = 40
        if nwave != 9:  # fine k-grid (configs[4]): F = next fast length >= 2 (2 n + 1)
            import fastoverlap_b200 as fob
            self.F = int(fob.load_library().fo_next_fast_len(2 * (2 * nwave + 1)))
        self.sigma = float((np.prod(self.box) / 256) ** (1. / 3) / 3)

    def describe(self, pairs):
        return {"workload": "BLJ256 PeriodicAlign, batched independent pairs (configs[1] system, "
                            "configs[4] batching)", "natoms": 256, "species": [204, 52],
                "nwave": self.n, "nfspace": self.F, "pairs_per_step_per_gpu": pairs,
                "seed": 256, "l2_policy": "inputs + intermediates per step exceed L2 "
                "(coords %.0f MB, structure-factor bank %.0f MB per chunk of 3256 pairs through host buffers, up "
                "to three times that device-resident)" % (
                    2 * pairs * 256 * 24 / 1e6, 3256 * 2 * 2 * 3610 * 16 / 1e6)}

    def make(self, pairs, rank):
        rng = np.random.default_rng(256 + 7919 * rank)
        shift = rng.uniform(0, 1, size=(pairs, 1, 3)) * self.box
        posB = self.base[None] + shift + rng.normal(scale=0.05, size=(pairs, 256, 3))
        posB -= np.round(posB / self.box) * self.box
        # permute within species
        for i in range(pairs):
            order = np.concatenate([rng.permutation(204), 204 + rng.permutation(52)])
            posB[i] = posB[i, order]
        posA = np.broadcast_to(self.base, posB.shape).copy()
        return posA, posB, shift[:, 0, :]

    # -- ours
    def setup(self, ctx):
        import fastoverlap_b200 as fob
        self.al = fob.PeriodicAlign(256, self.box, self.perm, n=self.n, ctx=ctx)
        self.params = self.al._params()

    def run_dev(self, ctx, dA, dB, P, out):
        ctx.per_align_pairs_dev(self.params, dA.data_ptr(), dB.data_ptr(), P, out[0].data_ptr(),
                                out[1].data_ptr(), out[2].data_ptr())

    def out_tensors(self, torch, P):
        return (torch.empty((P, 3), dtype=torch.int64, device="cuda"),
                torch.empty(P, dtype=torch.float64, device="cuda"),
                torch.empty((P, 3), dtype=torch.float64, device="cuda"))

    def run_host(self, ctx, A, B):
        return ctx.per_align_pairs(self.params, A, B)

    def d2h_bytes(self, P):
        return P * (24 + 8 + 24 + 4)

    def run_aligned(self, ctx, A, B, nthreads):
        """Full alignment: GPU hot path + native host refinement (LAP <-> mean displacement)."""
        return self.al.align_batch(A, B, nthreads=nthreads)[0]

    def check(self, res, shift):
        """Positive control: the known translation is recovered to within a grid cell."""
        fr = res[2]
        d = fr * self.box / self.F - shift
        d -= np.round(d / self.box) * self.box
        return bool(float(np.abs(d).max()) < self.box[0] / self.F)

    # FP64 work of the dominant kernel (structure factors, per_sf2_kernel), per pair: the 8-real-sum
    # formulation needs 2 structures x 256 atoms x 100 (i,j) x [4 DMUL + 10 l x 8 FMA]  (DESIGN.md
    # "K_sf").  The tensor-core kernel executes 24/20 of the FMAs (column padding 20 -> 24); the
    # padding is NOT counted here.
    dominant = "per_sf"
    dominant_pipe = "fp64_tensor"

    def dominant_flops_per_pair(self):
        return 2 * 256 * 100 * (4 * 1 + 80 * 2)

    # algorithmic (un-symmetrised): 2 x N x K x 8 flop
    def algorithmic_flops_per_pair(self):
        return 2 * 256 * 6859 * 8

    # -- oracle
    def run_oracle(self, oracle, A, B, nthreads=0):
        return oracle.per_align_pairs(A, B, self.box, self.n, self.F, self.sigma, self.perm,
                                      nthreads=nthreads)


class Lj38:
    """BASELINE.json configs[0] shape (examples/LJ38: N=38, sigma=0.3, Jmax=15, direct
    SphericalAlign coefficients, both orientations) batched over synthetic perturbed minima
    (SURVEY 8d C3 recipe: minimum + N(0,0.05^2), recentred, random rotation + permutation,
    seed 20171013)."""
    name = "lj38"
    natoms = 38
    default_pairs = 16384

    def __init__(self):
        g = np.load(os.path.join(ROOT, "tests", "golden", "spherical_lj38.npz"))
        self.minima = [g["pos1"] - g["pos1"].mean(0), g["pos2"] - g["pos2"].mean(0)]
        self.Jmax = 15
        self.sigma = 0.3

    def describe(self, pairs):
        return {"workload": "LJ38 SphericalAlign (direct coefficients, normal + inverted "
                            "orientation), batched independent pairs", "natoms": 38,
                "Jmax": self.Jmax, "sigma": self.sigma, "pairs_per_step_per_gpu": pairs,
                "seed": 20171013, "l2_policy": "L2 flushed by the >L2 coefficient scratch "
                "written every step"}

    @staticmethod
    def _rot(rng):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        a, b, c, d = q
        return np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                         [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                         [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])

    def make(self, pairs, rank):
        rng = np.random.default_rng(20171013 + 7919 * rank)
        A = np.empty((pairs, 38, 3))
        B = np.empty((pairs, 38, 3))
        for i in range(pairs):
            a = self.minima[i % 2] + rng.normal(scale=0.05, size=(38, 3))
            b = self.minima[i % 2] + rng.normal(scale=0.05, size=(38, 3))
            b = b.dot(self._rot(rng).T)[rng.permutation(38)]
            A[i] = a - a.mean(0)
            B[i] = b - b.mean(0)
        return A, B, None

    def setup(self, ctx):
        import fastoverlap_b200 as fob
        self.sa = fob.SphericalAlign(self.sigma, self.Jmax, ctx=ctx)

    def run_dev(self, ctx, dA, dB, P, out):
        ctx.sph_align_pairs_dev(dA.data_ptr(), dB.data_ptr(), P, 38, self.Jmax, self.sigma, True,
                                out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())

    def out_tensors(self, torch, P):
        return (torch.empty((P, 2, 3), dtype=torch.int64, device="cuda"),
                torch.empty((P, 2), dtype=torch.float64, device="cuda"),
                torch.empty((P, 2, 3), dtype=torch.float64, device="cuda"))

    def run_host(self, ctx, A, B):
        return ctx.sph_align_pairs(A, B, self.Jmax, self.sigma, invert=True)

    def d2h_bytes(self, P):
        return P * 2 * (24 + 8 + 24) + P * 4

    def run_aligned(self, ctx, A, B, nthreads):
        """Full alignment: GPU hot path + native host refinement (LAP + Kearsley, both orientations)."""
        return self.sa.align_batch(A, B, nthreads=nthreads)[0]

    def check(self, res, extra):
        return bool(np.all(np.isfinite(res[1])))

    def extra_measurements(self, ctx, torch, timed, dA, dB, A, B, P, steps):
        """numpy orientation rule (sphericalAlignment.py:178-187): hot path + continuous refinement of
        both orientations on the device (fo_sph_align_pairs_refined_dev), then the full alignment with
        a single host LAP + Kearsley per pair."""
        import fastoverlap_b200 as fob
        bi = torch.empty((P, 2, 3), dtype=torch.int64, device="cuda")
        bv = torch.empty((P, 2), dtype=torch.float64, device="cuda")
        fr = torch.empty((P, 2, 3), dtype=torch.float64, device="cuda")
        eu = torch.empty((P, 2, 3), dtype=torch.float64, device="cuda")
        ov = torch.empty((P, 2), dtype=torch.float64, device="cuda")
        step = lambda: ctx.sph_align_pairs_refined_dev(dA.data_ptr(), dB.data_ptr(), P, 38, self.Jmax, self.sigma,
                                                       True, bi.data_ptr(), bv.data_ptr(), fr.data_ptr(),
                                                       eu.data_ptr(), ov.data_ptr())
        for _ in range(3):
            step()
        n = max(3, steps // 4)
        ctx.profile_begin()
        ms, _, _ = timed(step, n)
        prof = ctx.profile_end()
        sa = fob.SphericalAlign(self.sigma, self.Jmax, ctx=ctx, orientation="overlap")
        ns = min(P, 4096)
        nthr = os.cpu_count() or 1
        sa.align_batch(A[:ns], B[:ns], nthreads=nthr)
        dts = []
        for _ in range(3):
            t0 = time.perf_counter()
            d = sa.align_batch(A[:ns], B[:ns], nthreads=nthr)[0]
            dts.append(time.perf_counter() - t0)
        ref_ms, ref_n = prof.get("sph_refine", (0.0, 0))
        return {"numpy_orientation_rule": {
            "value": P * n / (ms * 1e-3), "unit": "pairs/s", "steps": n,
            "what": "device-resident hot path + continuous rotation refinement (damped Newton, "
                    "fo_refine.cu) of both orientations",
            "refine_ms_per_step": ref_ms / n,
            "aligned_with_host_refine": ns / float(np.median(dts)), "aligned_pairs": ns,
            "median_distance": float(np.median(d))}}

    dominant = "sph_isoft"
    dominant_pipe = "fp64_tensor"

    def dominant_flops_per_pair(self):
        # iSOFT, both orientations (DESIGN.md "K_isoft"): executed real FMA count x 2
        from fastoverlap_b200.spherical import isoft_executed_flops
        return isoft_executed_flops(self.Jmax, True)

    def algorithmic_flops_per_pair(self):
        L = self.Jmax
        nnz = (L + 1) * (2 * L + 1) * (2 * L + 3) // 3
        B2 = 2 * (L + 1)
        return 2 * (4 * B2 * nnz + 2 * B2 * B2 * 5 * B2 * np.log2(B2))

    def run_oracle(self, oracle, A, B, nthreads=0):
        return oracle.sph_align_pairs(A, B, self.Jmax, self.sigma, invert=True, nthreads=nthreads)


WORKLOADS = {"blj256": Blj256, "lj38": Lj38}


# ----------------------------------------------------------------------------- helpers

class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"]):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": reasons}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def cpu_baseline(wl, sample_pairs, nthreads=None):
    """The oracle (CPU restatement of the reference algorithm) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    if nthreads is None:
        nthreads = os.cpu_count() or 1  # explicit: torchrun exports OMP_NUM_THREADS=1
    A, B, _ = wl.make(sample_pairs, 1000)
    oracle.lib()
    wl.run_oracle(oracle, A[:2], B[:2], nthreads)  # warm-up (page in, omp pool)
    t = time.perf_counter()
    res = wl.run_oracle(oracle, A, B, nthreads)
    dt = time.perf_counter() - t
    return sample_pairs / dt, int(res[-1]), dt


# ----------------------------------------------------------------------------- reference arm

def run_reference(args, wl):
    rank, local, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step: sized from a short probe so the whole run stays within minutes
    v0, used, _ = cpu_baseline(wl, max(8, 2 * cores))
    budget_s = 150.0  # whole run (warm-up + K steps) within a few minutes
    per_step = int(max(cores, min(4096, v0 * budget_s / (args.steps + 1))))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    A, B, _ = wl.make(per_step, 1000)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        wl.run_oracle(oracle, A, B, cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        wl.run_oracle(oracle, A, B, cores)
    dt = time.perf_counter() - t
    value = per_step * args.steps / dt
    cfg = wl.describe(per_step)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": used, "kind": "port",
                            "sample": "%d pairs per step x %d steps, OpenMP over pairs, C oracle "
                                      "(reference is Fortran/f2py: no Fortran compiler in the "
                                      "image)" % (per_step, args.steps)},
           "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- our arm

def run_ours(args, wl):
    import torch
    import fastoverlap_b200 as fob
    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its banner there)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(local)
    ctx = fob.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    wl.setup(ctx)
    P = args.pairs or wl.default_pairs
    A, B, extra = wl.make(P, rank)
    hA = torch.from_numpy(A).pin_memory()
    hB = torch.from_numpy(B).pin_memory()
    dA, dB = hA.cuda(), hB.cuda()
    out = wl.out_tensors(torch, P)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        sampler = None
        barrier()
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local)
            sampler.start()
            time.sleep(0.25)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        barrier()
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), clocks

    dev_step = lambda: wl.run_dev(ctx, dA, dB, P, out)
    host_res = [None]

    def host_step():
        host_res[0] = wl.run_host(ctx, hA.numpy(), hB.numpy())

    # ---- device-resident throughput (value) + per-kernel event timing over the same region
    for _ in range(max(3, args.warmup)):
        dev_step()
    l0 = ctx.launch_count()
    ctx.profile_begin()
    ms_dev, _, clocks = timed(dev_step, args.steps, sample_clocks=True)
    prof = ctx.profile_end()
    launches = ctx.launch_count() - l0
    value = world * P * args.steps / (ms_dev * 1e-3)

    # ---- end to end through the host-buffer C ABI (e2e): wall clock == device work + copies
    for _ in range(max(3, args.warmup)):
        host_step()
    e2e_steps = max(3, args.steps // 4)
    _, wall_ms, _ = timed(host_step, e2e_steps)
    e2e = world * P * e2e_steps / (wall_ms * 1e-3)
    ok = wl.check(host_res[0], extra)
    # device and host paths must agree bit for bit
    same = bool(np.array_equal(out[0].cpu().numpy(), host_res[0][0]))

    # ---- full alignment incl. the host refinement pool, on a sample (rank 0, N = 1 only)
    aligned = None
    if world == 1:
        ns = min(P, 4096)
        nthr = os.cpu_count() or 1
        try:  # a secondary figure: its failure must not take the bench line with it
            wl.run_aligned(ctx, A[:ns], B[:ns], nthr)  # warm-up at the timed size (scratch, thread pool)
            dts = []
            for _ in range(3):
                t0 = time.perf_counter()
                dists = wl.run_aligned(ctx, A[:ns], B[:ns], nthr)
                dts.append(time.perf_counter() - t0)
            dt = float(np.median(dts))
            aligned = {"value": ns / dt, "unit": "pairs/s", "pairs": ns, "host_threads": nthr,
                       "repeats": 3, "median_distance": float(np.median(dists)),
                       "what": "GPU hot path + native host refinement (Jonker-Volgenant LAP, "
                               "mean displacement / Kearsley) to the final distance; chunks of the batch "
                               "pipelined, host pool one chunk behind the GPU"}
        except Exception as e:  # noqa: BLE001
            aligned = {"error": "%s: %s" % (type(e).__name__, e)}

    extra_meas = {}
    if world == 1 and hasattr(wl, "extra_measurements"):
        extra_meas = wl.extra_measurements(ctx, torch, timed, dA, dB, A, B, P, args.steps)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak_vec = ctx.measure_fp64_peak()
    peak_tensor = ctx.measure_fp64_tensor_peak()
    peak = peak_tensor if getattr(wl, "dominant_pipe", "") == "fp64_tensor" else peak_vec
    dom_ms, dom_n = prof.get(wl.dominant, (0.0, 0))
    total_prof = sum(v[0] for v in prof.values())
    roof = None
    if dom_n:
        pairs_timed = P * args.steps
        achieved = wl.dominant_flops_per_pair() * pairs_timed / (dom_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "pipe": "fp64 tensor pipe (DMMA.8x8x4); the path is FP64 throughout, so the "
                "peak is the measured FP64 tensor throughput, not the bf16 figure of MEASURED_PEAKS.json",
                "kernel": wl.dominant, "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "measured in this run: FP64 tensor pipe (mma.sync.m8n8k4.f64 "
                               "microbenchmark, fo_measure_fp64_tensor_peak); MEASURED_PEAKS.json has no "
                               "FP64 figure",
                "peak_fp64_vector_tflops": peak_vec, "peak_fp64_tensor_tflops": peak_tensor,
                "flops_counted": "useful FP64 of the symmetry-reduced algorithm (FMA=2, MUL=1); "
                                 "tile padding executed by the kernel is not counted",
                "algorithmic_unsymmetrised_tflops": wl.algorithmic_flops_per_pair() * pairs_timed /
                (ms_dev * 1e-3) / 1e12,
                "kernel_ms_per_launch": dom_ms / dom_n, "kernel_share_of_step": dom_ms / total_prof,
                "kernel_shares": {k: v[0] / total_prof for k, v in prof.items()},
                "traffic": None,
                "hbm_frac_of_measured": None}
        try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tr = json.load(f)[wl.dominant]
            # the capture holds the bytes of one launch of tr["units_per_launch"] units; the launches of this run
            # may be larger (chunk size), so scale by the units an average launch of the timed region processed
            units_timed = pairs_timed * (2 if tr["unit"] == "structure" else 1)
            units_per_launch = units_timed / dom_n
            roof["traffic"] = tr["bytes_per_launch"] * units_per_launch / tr["units_per_launch"]
            roof["traffic_source"] = ("profiles/r01_traffic.json (ncu dram__bytes_read+write: %.0f bytes per %s, "
                                      "captured on a launch of %d; x %.0f %ss per launch here)" % (
                                          tr["bytes_per_launch"] / tr["units_per_launch"], tr["unit"],
                                          tr["units_per_launch"], units_per_launch, tr["unit"]))
        except Exception:
            pass
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                roof["hbm_peak_gbs"] = json.load(f).get("hbm_gbs")
            if roof.get("traffic") and roof.get("hbm_peak_gbs"):
                roof["hbm_frac_of_measured"] = (roof["traffic"] / (roof["kernel_ms_per_launch"] * 1e-3) / 1e9 /
                                                roof["hbm_peak_gbs"])
        except Exception:
            pass
    base = None
    if world == 1 and not args.no_cpu_baseline:
        v, used, dt = cpu_baseline(wl, args.cpu_sample)
        base = {"value": v, "unit": "pairs/s", "cores": used, "kind": "port",
                "sample": "%d pairs of the same workload, C oracle (CPU restatement of the "
                          "reference algorithm), OpenMP over pairs, %.1f s" % (args.cpu_sample, dt)}
    cfg = wl.describe(P)
    cfg["parallelism"] = "pairs sharded over %d GPU(s), no collective" % world
    res = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
           "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * P * wl.natoms * 24),
                   "d2h_bytes_per_step": int(wl.d2h_bytes(P)), "steps": e2e_steps,
                   "api": "fo_%s_align_pairs (host buffers)" % ("per" if wl.name == "blj256" else "sph")},
           "aligned_with_host_refine": aligned, "roofline": roof, "cpu_baseline": base,
           "checks": {"positive_control": bool(ok), "device_vs_host_identical": same}}
    res.update(extra_meas)
    print(json.dumps(res), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="blj256", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per step per GPU")
    ap.add_argument("--cpu-sample", type=int, default=1536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nwave", type=int, default=9, help="blj256 only: k-grid half width n (default 9, F = 40)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload](args.nwave) if args.workload == "blj256" else WORKLOADS[args.workload]()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
