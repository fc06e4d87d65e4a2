#!/usr/bin/env python
"""bench.py -- particle interactions/s of the FP64 U/J sweep (UJ_direct) on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a
synthetic random vortex particle cloud (jittered lattice in a 1x1x7 box, recipe of the
reference's scripts/benchmark_fmm2.jl:11-34), N = 2^20 particles by default, strong
scaling over 1/2/4/8 GPUs: targets are block-sharded over ranks and every step starts
with an all-gather (NCCL) of the 8 x N source buffer.

One step = one U/J sweep of the whole field: N^2 ordered (source, target) interactions
(SURVEY 8d).  `value` = N^2 / device time with particle state resident in HBM; `e2e`
is the same sweep through the public host API with host buffers (H2D + D2H inside the
timed region).  The SFS sweep and the other kernel families are measured after the
timed region and reported in extra keys.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--particles N] [--kernel NAME]
  python bench.py --impl reference ...   # the CPU restatement of the reference on host cores
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle interactions/sec (FP64 UJ_direct)"
UNIT = "interactions/s"
# algorithmic flops per U/J interaction: singular / gaussian / gaussianerf / winckelmans
F_UJ = {"singular": 68, "gaussian": 75, "gaussianerf": 78, "winckelmans": 82}
F_SFS = {"singular": 44, "gaussian": 48, "gaussianerf": 48, "winckelmans": 53}
WORKLOAD = "C4 synthetic random vortex particle cloud (jittered lattice 1x1x7, scripts/benchmark_fmm2.jl recipe)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--particles", dest="n", type=int, default=1 << 20)
    ap.add_argument("--kernel", default="winckelmans", choices=sorted(F_UJ))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip SFS / other-kernel / CPU extras")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------- clocks
class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# -------------------------------------------------------------- reference arm
def cpu_sample(n, kernel, seconds, threads=None, repeats=1):
    """The CPU restatement of the reference's UJ_direct pair loop (oracle port: Julia is not
    available on this box) on ALL host cores (omp_get_num_procs -- not OMP_NUM_THREADS, which
    torchrun sets to 1): all n sources x a target slice sized to about `seconds` of work, cut
    into >= 4 contiguous target blocks per thread.  ONE sampling policy for both CPU legs
    (`cpu_baseline` and `--impl reference`).  Returns (interactions/s, description, cores)."""
    from vpm_import import load
    from oracle import oracle
    vpm = load()
    threads = threads or oracle.num_procs()
    pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel])
    sb = vpm.source_system_to_buffer(pf)

    def run(nt):
        tb = np.zeros((16, nt), order="F")
        tb[0:3] = pf.get_X()[:, :nt]
        t = time.perf_counter()
        oracle.direct_buffers(tb, 0, nt, sb, 0, n, kernel, True, True, threads)
        return time.perf_counter() - t

    g = 8 * threads  # every thread gets whole blocks
    probe = min(n, 64 * threads)
    run(min(n, g))  # spin the thread pool up
    dt = run(probe)
    rate = probe * n / dt
    nt = int(min(n, max(probe, seconds * rate / n // g * g)))
    best = run(nt)
    if best < 0.6 * seconds and nt < n:  # the probe under-estimated the rate: size once more
        nt = int(min(n, max(nt, seconds * (nt * n / best) / n // g * g)))
        best = run(nt)
    for _ in range(repeats - 1):
        best = min(best, run(nt))
    return nt * n / best, f"all {n} sources x first {nt} targets of the same cloud ({nt * n:.3g} interactions, {best:.1f} s, {threads} threads)", threads


def reference_arm(args):
    """--impl reference: the CPU restatement on all host cores, rank 0 only, same workload,
    metric and sampling policy as the b200 arm's cpu_baseline leg."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    threads = oracle.num_procs()
    times = []
    per_step_seconds = max(4.0, min(15.0, 90.0 / max(1, args.steps + args.warmup)))
    desc = ""
    for i in range(args.warmup + args.steps):
        v, desc, _ = cpu_sample(args.n, args.kernel, per_step_seconds, threads)
        if i >= args.warmup:
            times.append(v)
    value = float(np.mean(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": args.n * args.n / value * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.n, args.kernel, args.gpus),
        "note": "ms_per_step is the full N^2 sweep extrapolated from the bounded sample; the CPU arm uses "
                "the host's cores whatever --gpus says",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": desc + "; reference-equivalent C restatement (oracle/), Julia not available"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_config(n, kernel, world):
    """the `config` object, identical in both arms (same keys, same workload)"""
    return {"workload": WORKLOAD, "n_particles": n, "kernel": kernel,
            "parallelism": f"targets block-sharded over {world} GPU(s), all-gather of the 8xN source buffer per sweep",
            "l2": "flushed between timed steps (256 MiB write); per-step CUDA events, max over ranks",
            "interaction": "one ordered (source,target) pair visit of the U+J loop; N^2 per step"}


def rvpm_step_c2(vpm, h):
    """BASELINE.json configs[1]: two leapfrogging rings (test/runtests_leapfrog.jl:42-46 geometry,
    nc=6 -> 33 800 particles), RK3 + reformulated VPM (f=0, g=1/5) + SFS + Pedrizzetti relaxation.
    The O(N) stage updates run on the host (numpy restatement in tests/physics.py -- in production
    they stay in the reference's Julia code); every UJ / SFS evaluation is a host-API call to the
    GPU, so this is the end-to-end time of one time step as the reference would see it."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import physics
    R = 0.7906
    pf = vpm.fields.ring_field(Nphi=100, nc=6, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906, kernel=vpm.gaussianerf)
    pf.particles[36, :pf.np] = 1.0  # constant SFS coefficient
    P = pf.particles
    h.check(h.lib.vpm_pin_host(h.ptr, P.ctypes.data, P.nbytes))
    gpu_ms = [0.0]

    def UJ(pfield, **kw):
        vpm.UJ_direct(pfield, **kw)
        gpu_ms[0] += h.timing()["total_ms"]

    dt = 1e-3
    zeta0 = 1.0 / (2 * np.pi) ** 1.5
    physics.rk3_step(pf, dt, UJ, f=0.0, g=0.2, relax=True, sfs=True, zeta0=zeta0)
    gpu_ms[0] = 0.0
    k = 3
    t = time.perf_counter()
    for _ in range(k):
        physics.rk3_step(pf, dt, UJ, f=0.0, g=0.2, relax=True, sfs=True, zeta0=zeta0)
    wall = (time.perf_counter() - t) / k
    h.check(h.lib.vpm_unpin_host(h.ptr, P.ctypes.data))
    # the same step with the field resident on the device (vpm_field_step)
    rf = vpm.ResidentField(pf, handle=h)
    kw = dict(integration="rungekutta3", f=0.0, g=0.2, sfs=True, Cs=1.0, relaxation="pedrizzetti", relax=True)
    rf.nextstep(dt, **kw)
    t = time.perf_counter()
    for _ in range(k):
        rf.nextstep(dt, **kw)
    dev_wall = (time.perf_counter() - t) / k
    # ... and with the DynamicSFS pseudo-3-level procedure (configs[1] as the reference runs it:
    # 4 x (U/J + SFS) + 1 x U/J per step), also resident
    kwd = dict(integration="rungekutta3", f=0.0, g=0.2, sfs="dynamic", clip_backscatter=True, alpha=0.999,
               sfs_rlxf=0.005, minC=0.0, maxC=1.0, force_positive=True, relaxation="correctedpedrizzetti", relax=True)
    rf.nextstep(dt, **kwd)
    t = time.perf_counter()
    for _ in range(k):
        rf.nextstep(dt, **kwd)
    dyn_wall = (time.perf_counter() - t) / k
    rf.download()
    return {"n_particles": pf.np, "kernel": "gaussianerf", "ms_per_step": wall * 1e3, "gpu_ms_per_step": gpu_ms[0] / k,
            "device_resident_ms_per_step": dev_wall * 1e3,
            "device_resident_dynamic_sfs_ms_per_step": dyn_wall * 1e3,
            "what": "3 x (U/J + SFS) + 1 x U/J (relaxation) through UJ_direct(pfield) with host buffers; "
                    "O(N) RK3/rVPM/relaxation updates on the host in numpy; constant SFS coefficient "
                    "(the dynamic procedure adds one more U/J + SFS call per step)"}


def nearfield_extra(vpm, h, n):
    """FMM near-field hook (BASELINE.json configs[4] shape at this N): uniform-octree leaves,
    theta = 0.4 near-field list, one vpm_p2p_leafpairs call from host buffers."""
    out = {}
    pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans)
    h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
    for ncrit in (128, 1024):
        # f-3: tree + theta-MAC list built on the device, near field over the resident lists
        vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, fetch=False)
        t = time.perf_counter()
        vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, fetch=False)
        t_tree = time.perf_counter() - t
        vpm.UJ_nearfield(pf, reset=True)
        t = time.perf_counter()
        vpm.UJ_nearfield(pf, reset=True)
        t_near = time.perf_counter() - t
        f3 = {"device_tree_ms": t_tree * 1e3, "nearfield_call_ms": t_near * 1e3,
              "e2e_interactions_per_s": h.timing()["uj_pairs"] / (t_tree + t_near),
              "what": "vpm_leaflists_build + vpm_uj_nearfield from the pinned host matrix (tree, list, near field on the device)"}
        # Hook 3: the same lists handed in by the caller, as FastMultipole would
        ll = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4)
        order = ll["sort_index"]
        sb = np.asfortranarray(vpm.source_system_to_buffer(pf)[:, order])
        tb = np.zeros((16, n), order="F")
        tb[0:3] = pf.get_X()[:, order]
        leaves = (ll["leaf_begin"], ll["leaf_end"])
        sizes = ll["leaf_end"] - ll["leaf_begin"]
        dl = ll["direct_list"]
        pairs = int((sizes[dl[:, 0]].astype(np.int64) * sizes[dl[:, 1]]).sum())
        vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.winckelmans)
        t = time.perf_counter()
        vpm.nearfield_device(tb, leaves, sb, leaves, dl, vpm.winckelmans)
        dt = time.perf_counter() - t
        tm = h.timing()
        out[f"ncrit_{ncrit}"] = {"leaves": int(len(sizes)), "mean_leaf": float(sizes.mean()), "list_pairs": int(len(dl)),
                                 "interactions": pairs, "kernel_interactions_per_s": pairs / (tm["uj_ms"] * 1e-3),
                                 "e2e_interactions_per_s": pairs / dt, "device_lists": f3}
    h.check(h.lib.vpm_unpin_host(h.ptr, pf.particles.ctypes.data))
    return out


def near_fraction(pf, cutoff_s, samples=400000, seed=1):
    """fraction of ordered pairs with r / sigma_source < cutoff_s (sampled)"""
    n = pf.np
    rng = np.random.Generator(np.random.PCG64(seed))
    i, j = rng.integers(0, n, samples), rng.integers(0, n, samples)
    X, sig = pf.get_X()[:, :n], pf.particles[6, :n]
    s = np.linalg.norm(X[:, i] - X[:, j], axis=0) / sig[j]
    return float(np.mean(s < cutoff_s))


FAR_CUTOFF = {"gaussianerf": 9.0, "gaussian": 3.45}


def family_rates(vpm, h, pf, peak_tflops, kernels, reps=2):
    """pair-kernel rate of each family on the field pf, far-field shortcut on and off.  roofline_frac for the
    shortcut-OFF run credits the family's flops to every pair (every pair took the regularised evaluation);
    for the shortcut-ON run the flops are credited per branch: the sampled fraction of pairs inside the
    regularised range gets the family's count, the rest the singular kernel's 68 (a LOWER bound of the work
    done: the branch is per warp, so some far pairs of mixed warps also took the regularised path)."""
    import torch
    from flowvpm_jl_b200 import sharding
    n = pf.np
    src8 = torch.from_numpy(np.ascontiguousarray(vpm.source_system_to_buffer(pf).T)).cuda()
    out = {}
    for name in kernels:
        f = sharding.ShardedField(h, src8, n, 0, 1, vpm.KERNELS[name].id)
        row = {}
        for flags, tag in ((0, "shortcut_on"), (vpm._cabi.FLAG_NO_FARFIELD_SHORTCUT, "shortcut_off")):
            if flags and name not in FAR_CUTOFF:
                continue
            f.uj(flags)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                f.uj(flags)
                torch.cuda.synchronize()
                best = min(best, h.timing()["uj_ms"])
            rate = n * n / (best * 1e-3)
            if name in FAR_CUTOFF and not flags:
                fn = near_fraction(pf, FAR_CUTOFF[name])
                flop = fn * F_UJ[name] + (1 - fn) * F_UJ["singular"]
                row["pairs_inside_cutoff"] = fn
            else:
                flop = F_UJ[name]
            row[tag] = {"interactions_per_s": rate, "kernel_ms": best, "flop_credited_per_pair": flop,
                        "roofline_frac": rate * flop / 1e12 / peak_tflops}
        out[name] = row
    del src8
    return out


def dense_fields(vpm, h, peak_tflops):
    """fields where a large share of the pairs is INSIDE the regularised range (the configurations that
    actually use gaussianerf: CoreSpreading requires it, src/FLOWVPM.jl:265-267), next to the sparse C4 cloud
    of the headline: BASELINE config 2 (two rings, 33 800 particles) and a compact blob (65 536 particles in
    a unit cube, sigma = 3 lattice spacings)."""
    R = 0.7906
    c2 = vpm.fields.ring_field(Nphi=100, nc=6, R=R, Rcross=0.1 * R, rings=2, dZ=0.7906)
    rng = np.random.Generator(np.random.PCG64(5))
    n = 65536
    blob = vpm.ParticleField(n)
    blob.particles[0:3, :n] = rng.random((3, n))
    blob.particles[3:6, :n] = rng.standard_normal((3, n)) / n
    blob.particles[6, :n] = 3.0 / 40 * (1 + 0.1 * (rng.random(n) - 0.5))
    blob.np = n
    out = {}
    for name, pf in (("c2_two_rings_33800", c2), ("blob_65536", blob)):
        out[name] = {"n_particles": pf.np, "pairs_with_s_below_9": near_fraction(pf, 9.0),
                     "by_kernel": family_rates(vpm, h, pf, peak_tflops, ["gaussianerf", "gaussian", "winckelmans"])}
    return out


def single_process_handle(vpm, n, kernel, ngpu):
    """the deployment north_star describes: ONE process (the Julia caller) drives all GPUs through one handle.
    UJ_direct(pfield; sfs=true) on the page-locked 46 x N host matrix: uploads rows X, Gamma, sigma, U/J + SFS
    sweeps sharded over the devices (NCCL broadcast of the sources, all-gather of J), rows 10:27 and 40:42
    written back -- everything inside the timed call."""
    h = vpm.Handle(device_ids=list(range(ngpu)))
    try:
        pf = vpm.fields.cloud_field(n, kernel=vpm.KERNELS[kernel])
        P = pf.particles
        h.check(h.lib.vpm_pin_host(h.ptr, P.ctypes.data, P.nbytes))
        out = {"n_gpus": ngpu, "n_particles": n, "kernel": kernel}
        for sfs, tag in ((False, "uj"), (True, "uj_sfs")):
            vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs, handle=h)
            t = time.perf_counter()
            vpm.UJ_direct(pf, sfs=sfs, reset=True, reset_sfs=sfs, handle=h)
            dt = time.perf_counter() - t
            out[tag] = {"s_per_call": dt, "interactions_per_s": (2 if sfs else 1) * n * n / dt,
                        "h2d_bytes": n * 7 * 8, "d2h_bytes": n * (18 + (3 if sfs else 0)) * 8}
        h.check(h.lib.vpm_unpin_host(h.ptr, P.ctypes.data))
        out["what"] = "vpm.Handle(device_ids=range(N)) in rank 0's process; UJ_direct(pfield) on the pinned host matrix"
        return out
    finally:
        h.close()


def c5_nearfield(vpm, ngpu, logn=24, ncrits=(128, 512)):
    """BASELINE config 5: 2^24 particles, leaf lists + near field on `ngpu` GPUs from one process"""
    h = vpm.Handle(device_ids=list(range(ngpu)))
    try:
        n = 1 << logn
        pf = vpm.fields.cloud_field(n, kernel=vpm.winckelmans)
        h.check(h.lib.vpm_pin_host(h.ptr, pf.particles.ctypes.data, pf.particles.nbytes))
        out = {"n_particles": n, "n_gpus": ngpu}
        for ncrit in ncrits:
            vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h, fetch=False)
            t = time.perf_counter()
            info = vpm.leaf_lists(pf, ncrit=ncrit, theta=0.4, handle=h, fetch=False)
            t_tree = time.perf_counter() - t
            vpm.UJ_nearfield(pf, reset=True, handle=h)
            t = time.perf_counter()
            vpm.UJ_nearfield(pf, reset=True, handle=h)
            dt = time.perf_counter() - t
            tm = h.timing()
            out[f"ncrit_{ncrit}"] = {"leaves": info["n_leaves"], "list_pairs": info["n_pairs"], "interactions": tm["uj_pairs"],
                                     "device_tree_s": t_tree, "nearfield_call_s": dt, "pair_kernel_ms_dev0": tm["uj_ms"],
                                     "e2e_interactions_per_s": tm["uj_pairs"] / dt}
        h.check(h.lib.vpm_unpin_host(h.ptr, pf.particles.ctypes.data))
        return out
    finally:
        h.close()


def ncu_traffic(n, kernel, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pair kernel from the committed ncu
    --set full capture (profiles/uj_pairs_traffic.json), if it was taken at this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "uj_pairs_traffic.json")) as fh:
            t = json.load(fh)
        if t["n_particles"] == n and t["kernel"] == kernel and world == 1:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    return None


# ------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from vpm_import import load

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvpm_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    vpm = load()
    from flowvpm_jl_b200 import sharding
    h = vpm.Handle(device_ids=[local_rank])
    vpm.set_handle(h)
    lib = h.lib
    n, kernel = args.n, vpm.KERNELS[args.kernel]
    dev = torch.device("cuda", local_rank)

    # ---- synthetic field (same seed on every rank), shard resident in HBM
    pf = vpm.fields.cloud_field(n, kernel=kernel)
    src8 = vpm.source_system_to_buffer(pf)  # 8 x n, the reference's source-buffer layout
    t0, t1 = sharding.shard_bounds(n, world, rank)
    host_local = torch.from_numpy(np.ascontiguousarray(src8[:, t0:t1].T)).pin_memory()
    field = sharding.ShardedField(h, host_local.to(dev), n, rank, world, kernel.id)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(fn, k, w, key="uj_ms"):
        """W warm-ups, then K steps each bracketed by CUDA events on the launching stream,
        L2 flushed between steps; returns (list of ms, list of pair-kernel ms)."""
        for _ in range(w):
            fn()
        sync_all()
        ms, kms = [], []
        for _ in range(k):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            a.record()
            fn()
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
            kms.append(h.timing()[key])
        sync_all()
        return ms, kms

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- roofline denominator: FP64 FMA pipe, measured on this GPU right now
    dfma = np.zeros(1)
    dms = np.zeros(1)
    import ctypes as C
    h.check(lib.vpm_measure_dfma_peak(h.ptr, dfma.ctypes.data_as(C.POINTER(C.c_double)),
                                      dms.ctypes.data_as(C.POINTER(C.c_double))))
    dfma_per_s = float(dfma[0])

    # ---- timed region: K U/J sweeps, state resident
    clocks = ClockSampler(local_rank)
    clocks.start()
    field.launches = 0
    ms, kms = timed_steps(lambda: field.uj(0), args.steps, max(3, args.warmup))
    clk = clocks.stop()
    total_ms = reduce_max(float(np.sum(ms)))
    ms_per_step = total_ms / args.steps
    value = n * n / (ms_per_step * 1e-3)
    launches = 3 * args.steps  # prep_uj_records + uj_pairs_kernel + uj_finish_kernel per step
    kernel_ms = float(np.mean(kms))
    my_pairs = (min(n, t1) - t0) * (world * field.c)
    achieved_tflops = my_pairs / (kernel_ms * 1e-3) * F_UJ[args.kernel] / 1e12
    peak_tflops = 2 * dfma_per_s / 1e12

    # ---- e2e: the same sweep through the host API with host buffers
    if world == 1:
        P = pf.particles
        h.check(lib.vpm_pin_host(h.ptr, P.ctypes.data, P.nbytes))

        def e2e_step():
            vpm.UJ_direct(pf, reset=True)
        h2d, d2h = n * 7 * 8, n * 18 * 8
    else:
        host_out = torch.empty((field.c, 12), dtype=torch.float64).pin_memory()

        def e2e_step():
            field.src8_local[: t1 - t0].copy_(host_local, non_blocking=True)
            field.uj(0)
            host_out.copy_(field.out12, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = (t1 - t0) * 8 * 8, field.c * 12 * 8
    e2e_k = max(1, min(args.steps, 2))
    e2e_step()
    sync_all()
    tw = time.perf_counter()
    for _ in range(e2e_k):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = reduce_max((time.perf_counter() - tw) / e2e_k)
    e2e_value = n * n / e2e_s

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(n, args.kernel, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_k, "api": "UJ_direct(pfield) -> vpm_uj_direct (pinned host matrix)" if world == 1
                else "sharding.ShardedField.uj with pinned H2D/D2H of the local shard"},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "fp64_fma", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / peak_tflops, "traffic": ncu_traffic(n, args.kernel, world),
                     "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: "
                                     "80 B/source record + 112 B/target per source split)",
                     "kernel": f"uj_pairs_kernel<{args.kernel}>", "kernel_ms": kernel_ms,
                     "flop_per_interaction": F_UJ[args.kernel],
                     "peak_source": "2 x DFMA/s measured live by vpm_measure_dfma_peak on this GPU "
                                    "(MEASURED_PEAKS.json has no FP64 figure; theoretical 37.2 TFLOP/s at 1965 MHz)",
                     "dfma_per_s": dfma_per_s,
                     "note": "the path is FP64-FMA-pipe bound, not HBM or tensor bound (SURVEY 8d)"},
    }

    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        # SFS sweep over the J just computed
        sms, skms = timed_steps(lambda: field.sfs(vpm._cabi.FLAG_TRANSPOSED), 1, 1, key="sfs_ms")
        extras["sfs"] = {"interactions_per_s": n * n / (np.mean(sms) * 1e-3), "ms": float(np.mean(sms)),
                         "kernel": args.kernel,
                         "roofline_frac": n * n / (np.mean(skms) * 1e-3) * F_SFS[args.kernel] / 1e12 / peak_tflops}
        extras["by_kernel"] = family_rates(vpm, h, pf, peak_tflops, [k for k in sorted(F_UJ) if k != args.kernel])
        extras["dense_fields"] = dense_fields(vpm, h, peak_tflops)
        # the constant-bank variant of the headline kernel (VPM_OPT_UJ_CONST: source records through __constant__
        # memory, 768 per launch -- fewer three-register FP64 instructions; not the default)
        h.set_option(vpm._cabi.OPT_UJ_CONST, 1)
        try:
            cm, ckm = timed_steps(lambda: field.uj(0), 1, 1)
        finally:
            h.set_option(vpm._cabi.OPT_UJ_CONST, 0)
        extras["constant_bank_variant"] = {
            "kernel": args.kernel, "interactions_per_s": n * n / (np.mean(cm) * 1e-3), "ms": float(np.mean(cm)),
            "roofline_frac": n * n / (np.mean(cm) * 1e-3) * F_UJ[args.kernel] / 1e12 / peak_tflops,
            "launches_per_sweep": h.timing()["kernel_launches"],
            "what": "vpm_set_option(VPM_OPT_UJ_CONST, 1); whole sweep (all its launches) on the device clock"}
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import smalln_bench
        extras["small_n_latency"] = {"what": "one UJ_direct(pfield; sfs=true, reset=true, reset_sfs=true) call, wall clock "
                                             "per call (us), host matrix pageable / page-locked, next to the CPU port on all cores",
                                     "gaussianerf": smalln_bench.measure(h, "gaussianerf")}
        # optional FP32-arithmetic sweep (VPM_FLAG_FP32, north star's 1e-5 mode) against the FP32 FMA pipe
        import ctypes as C
        fv, fms = C.c_double(), C.c_double()
        h.check(lib.vpm_measure_ffma_peak(h.ptr, 0, C.byref(fv), C.byref(fms)))
        m, km = timed_steps(lambda: field.uj(vpm._cabi.FLAG_FP32), 2, 1)
        extras["fp32_mode"] = {"interactions_per_s": n * n / (np.mean(m) * 1e-3), "kernel": args.kernel,
                               "ffma_per_s_measured": fv.value,
                               "fp32_lane_ops_per_interaction": 46,
                               "frac_of_ffma_issue_peak": n * n / (np.mean(km) * 1e-3) * 46 / fv.value,
                               "what": "uj_pairs_kernel_f32: hi/lo split positions, packed f32x2 pair loop (23 FFMA2/FMUL2/"
                                       "FADD2 per pair), FP64 flush per 128-source tile; parity bar 1e-5 (tests/test_fp32_gpu.py)"}
        uj_ms = ms_per_step
        extras["rvpm_step_ms_estimate"] = {"value": 5 * uj_ms + 4 * extras["sfs"]["ms"],
                                           "what": "5 U/J + 4 SFS sweeps (RK3 + DynamicSFS + relaxation, SURVEY 3.1), O(N) host work excluded"}
        extras["rvpm_step_c2"] = rvpm_step_c2(vpm, h)
        # the metric's second half at the bench size: one whole rVPM time step on the device
        # (RK3 + reformulated VPM + DynamicSFS + relaxation = 5 U/J + 4 SFS sweeps, vpm_field_step)
        rf = vpm.ResidentField(pf, handle=h)
        tw = time.perf_counter()
        rf.nextstep(1e-4, integration="rungekutta3", f=0.0, g=0.2, sfs="dynamic", clip_backscatter=True,
                    force_positive=True, alpha=0.999, relaxation="correctedpedrizzetti", relax=True)
        extras["rvpm_step"] = {"n_particles": n, "kernel": args.kernel, "s_per_step": time.perf_counter() - tw,
                               "what": "one device-resident step: rungekutta3 + rVPM (f=0, g=1/5) + DynamicSFS "
                                       "(pseudo3level_positive, clipping_backscatter) + correctedpedrizzetti "
                                       "relaxation = 5 U/J + 4 SFS sweeps (9 N^2 pair visits); 8 GPUs: "
                                       "profiles/r1_rvpm_step_1M_8gpu.json"}
        extras["fmm_nearfield"] = nearfield_extra(vpm, h, n)
        line["extras"] = extras
        v, desc, cores = cpu_sample(n, args.kernel, args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": desc + "; reference-equivalent C restatement (oracle/), Julia not available"}
    elif rank == 0:
        line["cpu_baseline"] = None

    if world > 1:
        # single-process legs: rank 0 alone drives all GPUs through ONE handle while the other ranks wait
        # on the HOST (a key in torch.distributed's store): an NCCL barrier would leave a spinning kernel on
        # every other GPU, time-slicing against rank 0's kernels there
        sync_all()
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            extras = {}
            try:
                extras["single_process_handle"] = single_process_handle(vpm, n, args.kernel, world)
                if world >= 8 and not args.no_extras:
                    extras["c5_nearfield_2p24"] = c5_nearfield(vpm, world)
            except Exception as exc:  # noqa: BLE001 -- an extra must not take the headline line down
                extras["single_process_error"] = repr(exc)
            line["extras"] = extras
            store.set("single_process_legs_done", "1")
        else:
            store.wait(["single_process_legs_done"], datetime.timedelta(minutes=30))

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
