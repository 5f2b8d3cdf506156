#!/usr/bin/env python
"""bench.py -- fp64 residual+Jacobian assembly throughput of the thermal hex-Q1 path.

Workload (BASELINE.json configs[1]): 3-D steady thermal, hex Q1, 128^3 inline brick, residual +
Jacobian, source 12 pi^2 sin sin sin, all-boundary strong Dirichlet.  One step = one
assembleJacRes(compute_jacobian, compute_residual) over every element of the rank's mesh, written
into the CSR values / residual vector (overwrite mode, i.e. the caller's zeroing is not needed).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                          CPU restatement of the reference path
                                                                (oracle/, all host cores as processes)
N > 1 is launched by torch.distributed.run; the mesh is N bricks stacked along z (weak scaling),
every rank assembles its slab and the shared-row contributions are summed with the NCCL halo sum.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 residual+Jacobian elements/sec"
UNIT = "elements/s"


WORKLOADS = {"thermal": "thermal hex-Q1", "le": "linear elasticity hex-Q1 (3 dofs/node)", "leq2": "linear elasticity hex-Q2 (81 dofs/element, 27 Gauss points)", "ns": "Navier-Stokes hex-Q1 ux/pr/uy/uz, SUPG+PSPG, reference uz rows",
             "maxwell": "Maxwell hex HCURL E (12 edge dofs) + HDIV B (6 face dofs), one DIRK-1,2 stage"}
_WORKLOAD = "thermal"


def workload_name(n, world, nz=None):
    nz = n if nz is None else nz
    return "%s %dx%dx%d inline brick, %s, residual+Jacobian%s" % (WORKLOADS[_WORKLOAD], n, n, nz * world, "transient stage" if _WORKLOAD == "maxwell" else "steady", "" if world == 1 else " (%d z-slabs of %dx%dx%d)" % (world, n, n, nz))


def measure_traffic(args, kernel_regex, n_kernels):
    """DRAM bytes (read + write) of one launch of the dominant kernel(s), measured by an ncu pass over a child run of this same
    command (1 GPU, one timed step): dram__bytes_read.sum + dram__bytes_write.sum.  Returns None when ncu is unavailable."""
    import csv
    import io
    import shutil
    if shutil.which("ncu") is None:
        return None
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:" + kernel_regex, "-s", str(3 * n_kernels),
           "-c", str(n_kernels), "--csv", sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", _WORKLOAD, "--n", str(args.n), "--steps", "1", "--warmup", "3"]
    for kv in args.opt:
        cmd += ["--opt", kv]
    for kv in args.fn:
        cmd += ["--fn", kv]
    if args.perturb:
        cmd += ["--perturb", str(args.perturb)]
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env).stdout
        rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 5]
        hdr = next(r for r in rows if "Metric Name" in r)
        im, iv, iu = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for r in rows:
            if r is not hdr and r[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        return total if total > 0 else None
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.proc = None
        self.path = "/tmp/bench_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            top = sorted(sm)[len(sm) // 2:]          # the loaded half of the samples
            out["sm_mhz"] = float(np.median(top))
            out["sm_max_mhz"] = mx
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------------------
# CPU legs (the ONLY places that touch oracle/)
# --------------------------------------------------------------------------------------------------
def _oracle_cfg(n, nz):
    mesh = {"dimension": 3, "element type": "hex", "xmin": 0.0, "xmax": 1.0, "ymin": 0.0, "ymax": 1.0, "zmin": 0.0, "zmax": float(nz) / n, "NX": n, "NY": n, "NZ": nz}
    if _WORKLOAD == "leq2":
        return {"Mesh": mesh, "Physics": {"modules": "linearelasticity", "Dirichlet conditions": {v: {"all boundaries": "0.0"} for v in ("dx", "dy", "dz")}},
                "Discretization": {"order": {"dx": 2, "dy": 2, "dz": 2}, "quadrature": 4},
                "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)*sin(pi*z)", "source dy": "sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)",
                              "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"},
                "Solver": {"solver": "steady-state", "workset size": 100}}
    if _WORKLOAD == "le":
        return {"Mesh": mesh, "Physics": {"modules": "linearelasticity", "Dirichlet conditions": {v: {"all boundaries": "0.0"} for v in ("dx", "dy", "dz")}},
                "Discretization": {"order": {"dx": 1, "dy": 1, "dz": 1}, "quadrature": 2},
                "Functions": {"lambda": "1.0", "mu": "1.0", "source dx": "sin(pi*x)*sin(pi*y)", "source dy": "sin(2*pi*x)*sin(2*pi*y)", "source dz": "sin(3*pi*x)*sin(3*pi*y)*sin(3*pi*z)"},
                "Solver": {"solver": "steady-state", "workset size": 100}}
    if _WORKLOAD == "maxwell":
        return {"Mesh": mesh, "Physics": {"modules": "maxwell", "Dirichlet conditions": {}},
                "Discretization": {"order": {"E": 1, "B": 1}, "quadrature": 2},
                "Functions": {"current x": "sin(2*pi*z)"}, "Solver": {"solver": "transient", "workset size": 100}}
    if _WORKLOAD == "ns":
        return {"Mesh": mesh, "Physics": {"modules": "navier stokes", "useSUPG": True, "usePSPG": True,
                                          "Dirichlet conditions": {v: {"all boundaries": "0.0"} for v in ("ux", "uy", "uz")}},
                "Discretization": {"order": {"ux": 1, "pr": 1, "uy": 1, "uz": 1}, "quadrature": 2},
                "Functions": {"source ux": "1.0", "viscosity": "1.0", "density": "1.0"}, "Solver": {"solver": "steady-state", "workset size": 100}}
    return {"Mesh": {"dimension": 3, "element type": "hex", "xmin": 0.0, "xmax": 1.0, "ymin": 0.0, "ymax": 1.0, "zmin": 0.0, "zmax": float(nz) / n,
                     "NX": n, "NY": n, "NZ": nz},
            "Physics": {"modules": "thermal", "Dirichlet conditions": {"T": {"all boundaries": "0.0"}}},
            "Discretization": {"order": {"T": 1}, "quadrature": 2},
            "Functions": {"thermal source": "12*(pi*pi)*sin(2*pi*x)*sin(2*pi*y)*sin(2*pi*z)"},
            "Solver": {"solver": "steady-state", "workset size": 100}}


_W = {}


def _worker_init(n, nz):
    from oracle import pyoracle
    op = pyoracle.OracleProblem(_oracle_cfg(n, nz))
    rng = np.random.default_rng(20261017)
    _W["op"] = op
    _W["u"] = rng.uniform(-1.0, 1.0, op.num_dofs)
    _W["res"] = np.zeros(op.num_dofs)
    _W["jac"] = np.zeros(op.nnz)
    _W["kw"] = {}
    if _WORKLOAD == "maxwell":   # the same DIRK-1,2 stage the GPU arm assembles
        op.set_time(True, time=0.3, dt=0.01, stage=0, A=[[0.5]], b=[1.0], c=[0.5], bdf=(1.0, -1.0))
        _W["kw"] = dict(sol_prev=[rng.uniform(-1.0, 1.0, op.num_dofs)], sol_stage=[_W["u"]])


def _worker_step(_):
    op = _W["op"]
    _W["res"][:] = 0.0
    _W["jac"][:] = 0.0
    t0 = time.perf_counter()
    op.assemble_jacres(_W["u"], res=_W["res"], jac=_W["jac"], **_W["kw"])
    return time.perf_counter() - t0, op.num_elems


def cpu_baseline_serial(n, budget_s=15.0):
    """Oracle (scalar C++ restatement, 1 thread == Kokkos::Serial) on a z-slab of the same mesh."""
    from oracle import pyoracle
    pyoracle.build()
    nz = max(1, min(n, 16 if _WORKLOAD == "thermal" else (1 if _WORKLOAD == "leq2" else 4)))
    if _WORKLOAD == "leq2":
        n = min(n, 24)      # the oracle's SFad<81> pass over a 24 x 24 x 1 slab already takes seconds
    _worker_init(n, nz)
    _worker_step(0)
    t_total, elems, reps = 0.0, 0, 0
    while t_total < budget_s and reps < 50:
        dt, ne = _worker_step(0)
        t_total += dt
        elems += ne
        reps += 1
    return {"value": elems / t_total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle assembleJacRes on a %dx%dx%d z-slab of the workload, %d passes, %.1f s" % (n, n, nz, reps, t_total)}


def run_reference(args):
    """`--impl reference`: the reference cannot be built here (Trilinos + MPI are absent), so this times
    the oracle port, one process per host core like `mpiexec -n <cores>` of the serial build, each on
    its own z-slab of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle
    pyoracle.build()
    n = args.n
    cores = max(1, min(os.cpu_count() or 1, 64))
    nz = 4 if _WORKLOAD == "thermal" else (1 if _WORKLOAD == "leq2" else 2)
    if _WORKLOAD == "leq2":
        n = min(n, 24)
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_worker_init, initargs=(n, nz)) for _ in range(cores)]
    try:
        def step():
            t0 = time.perf_counter()
            rs = [p.apply_async(_worker_step, (0,)) for p in pools]
            out = [r.get() for r in rs]
            return time.perf_counter() - t0, sum(o[1] for o in out)
        for _ in range(args.warmup):
            step()
        t_total, elems = 0.0, 0
        for _ in range(args.steps):
            dt, ne = step()
            t_total += dt
            elems += ne
    finally:
        for p in pools:
            p.terminate()
    value = elems / t_total
    full = float(args.n) ** 3
    sample = "%d processes x (%dx%dx%d z-slab of the workload) per step = %.4f of the workload's %d elements per step" % (cores, n, n, nz, cores * n * n * nz / full, int(full))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(n, 1), "sample": sample, "sampled_fraction_per_step": cores * n * n * nz / full},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def ring_name(plan, general):
    """What a sweep step stages per element."""
    if general:
        return "n/a (general path)"
    if plan.stat("class_ring"):
        return "class (%d values + residual)" % plan.stat("class_ring")
    if plan.stat("metric_ring"):
        return "metric (%d metric entries + load + state)" % plan.stat("metric_ring")
    return "full (upper triangle + residual)"


def bind_to_gpu_numa_node(local):
    """Pins this process (and so the pinned host buffers it allocates next) to the NUMA node the GPU hangs off: with one process per
    GPU and buffers placed wherever the launcher left the process, the device -> host copies of the end-to-end path cross the
    socket interconnect (per-GPU D2H fell from ~53 to ~11.5 GB/s at 8 GPUs).  Returns the node or None."""
    try:
        import torch
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local)
        h = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mrhyde_b200.problems import ElasticityQ2Brick, MaxwellBrick, SystemBrick, ThermalBrick

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the assembly path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local) if os.environ.get("MRHYDE_B200_NUMA_BIND", "1") == "1" else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    nz = n
    if args.scaling == "strong" and world > 1:
        if n % world:
            raise RuntimeError("--scaling strong needs n divisible by the number of GPUs")
        nz = n // world
    options = {"accumulate": "false"}
    if world > 1 and _WORKLOAD == "thermal" and os.environ.get("MRHYDE_B200_OVERLAP_HALO", "0") == "1":
        options["overlap halo"] = "true"   # opt-in until the overlapped exchange has been measured
    for kv in args.opt:
        k, v = kv.split("=", 1)
        options[k] = v
    functions = dict(kv.split("=", 1) for kv in args.fn)
    if _WORKLOAD == "leq2":
        prob = ElasticityQ2Brick(n, device=local, options=options, rank=rank, nranks=world, nz=nz)
    elif _WORKLOAD == "maxwell":
        prob = MaxwellBrick(n, device=local, options=options, functions=functions or None, rank=rank, nranks=world, nz=nz)
    elif _WORKLOAD == "thermal":
        prob = ThermalBrick(3, [n, n, nz], device=local, rank=rank, nranks=world, options=options, functions=functions or None, perturb=args.perturb)
    else:
        prob = SystemBrick({"le": "linearelasticity", "ns": "navier stokes"}[_WORKLOAD], 3, [n, n, nz], device=local, rank=rank, nranks=world, options=options)
    plan = prob.plan
    general = plan.stat("general") == 1
    if world > 1:
        uid = torch.from_numpy(plan.comm_unique_id()).to(dev) if rank == 0 else torch.zeros(128, dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        plan.comm_init(uid.cpu().numpy(), rank, world)
        plan.set_halo(prob.col_gids)
    d_u = torch.from_numpy(prob.state()).to(dev)
    d_res = torch.empty(prob.n_rows, dtype=torch.float64, device=dev)
    d_jac = torch.empty(prob.nnz, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    time_spec = None
    if _WORKLOAD == "maxwell":   # one DIRK-1,2 stage (regression/maxwell/PlaneWave's integrator) from a previous-step state
        from mrhyde_b200.capi import TimeSpec
        d_up = (0.5 * d_u).contiguous()
        time_spec = TimeSpec(time=0.3, deltat=0.01, stage=0, A=[[0.5]], b=[1.0], c=[0.5], bdf=(1.0, -1.0), sol_prev=[d_up], sol_stage=[d_u])

    def step():
        plan.assemble_jacres(d_u, d_res, d_jac, stream=stream, time=time_spec)
        if world > 1:
            plan.halo_sum(d_res, d_jac, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.traffic_child:   # profiled by the parent's ncu pass: warm-up launches, one more, done
        for _ in range(max(3, args.warmup) + args.steps):
            step()
        torch.cuda.synchronize()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    plan.kernel_time(reset=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[-1])
    ms_median = float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]))   # median of the repetitions
    barrier()
    kern_ms, kern_n = plan.kernel_time(reset=True)
    launches_per_step = plan.stat("kernel_launches_per_assemble") + (plan.stat("halo_launches_per_sum") if world > 1 else 0)
    t = torch.tensor([ms, ms_median], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_median = float(t[0].item()), float(t[1].item())
    total_elems = prob.n_elem * world
    value = total_elems * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call: pinned host sol in, res + J values out, every step
    h_u = torch.from_numpy(prob.state()).pin_memory()
    h_res = torch.empty(prob.n_rows, dtype=torch.float64).pin_memory()
    h_jac = torch.empty(prob.nnz, dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    h_time, h2d_vectors = None, 1
    if _WORKLOAD == "maxwell":
        from mrhyde_b200.capi import TimeSpec
        h_up = (0.5 * h_u).pin_memory()
        h_time = TimeSpec(time=0.3, deltat=0.01, stage=0, A=[[0.5]], b=[1.0], c=[0.5], bdf=(1.0, -1.0), sol_prev=[h_up.numpy()], sol_stage=[h_u.numpy()])
        h2d_vectors = 2   # the stage state and the previous step's state go to the device every step
    plan.assemble_jacres_host(h_u.numpy(), h_res.numpy(), h_jac.numpy(), time=h_time)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.assemble_jacres_host(h_u.numpy(), h_res.numpy(), h_jac.numpy(), time=h_time)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = total_elems * e2e_steps / e2e_s
    # the same load for ~1 s more so the clock sampler sees the kernel under load.  The iteration count comes from the
    # max-reduced step time, so every rank runs the same number of halo sums (a wall-clock loop can leave one rank an
    # iteration short and the send/recv pairs unmatched)
    n_load = max(1, min(20000, int(1.0 / max(1e-6, ms / args.steps * 1e-3))))
    for i in range(n_load):
        step()
        if i % 20 == 19:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    assert bool(torch.isfinite(d_res).all()) and bool(torch.isfinite(d_jac[:: 97]).all())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = prob.algorithmic_bytes()
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel(s), measured now by an ncu pass over a child run of this command (not read from a file)
        traffic = None
        if world == 1 and not args.no_traffic:
            traffic = measure_traffic(args, "gen_element_kernel|gen_pull_kernel" if general else "mrh_thermal_q1", 2 if general else 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms / args.steps, "ms_per_step_median": ms_median, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n, world, nz), "elements_per_gpu": prob.n_elem, "rows_per_gpu": prob.n_rows, "nnz_per_gpu": prob.nnz,
                           "l2": "inputs+outputs %.2f GB per GPU exceed the 126 MB L2 (no flush needed)" % (alg_bytes / 1e9),
                           "output_mode": "overwrite (accumulate=false)", "chains": plan.stat("n_chains"), "columns": plan.stat("n_columns"),
                           "segments": plan.stat("n_segments"), "ring_capacity": plan.stat("ring_capacity"), "threads_per_block": plan.stat("threads_per_block"),
                           "smem_bytes": plan.stat("smem_bytes"), "row_patterns": plan.stat("n_patterns"), "ring": ring_name(plan, general), "kernel_build": "nvrtc plan-specialised" if plan.stat("jit") else "ahead-of-time",
                           "elements_incl_halo": plan.stat("n_elem_with_halo"), "plan_options": {k: v for k, v in options.items() if k != "accumulate"}, "functions": functions, "perturb": args.perturb, "parallelism": ("z-slabs x%d + halo sum (%s)" % (world, "p2p peer stores over NVLink" if plan.stat("halo_p2p") else "NCCL send/recv")) if world > 1 else "1 GPU"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                             "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, measured in this run" if traffic else None,
                             "kernel": "gen_element_kernel + gen_pull_kernel (general path, whole assemble call)" if general else "mrh_thermal_q1_3d", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * prob.n_rows * world * h2d_vectors, "d2h_bytes_per_step": 8 * (prob.n_rows + prob.nnz) * world,
                        "steps": e2e_steps, "api": "mrhyde_b200_assemble_jacres_host (pinned host buffers)", "numa_node_of_rank0": numa_node},
                "gpu_launches": int(args.steps * launches_per_step), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_serial(n)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_STDOUT_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    # Libraries below us write to file descriptor 1 (NCCL prints its version banner there when a communicator is created):
    # keep the original stdout for the JSON line and point fd 1 at stderr for everything else.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=0, help="elements per brick edge (per GPU); default 128 (thermal), 64 (le), 96 (ns), 64 (leq2), 64 (maxwell): the BASELINE sizes")
    ap.add_argument("--workload", default="thermal", choices=sorted(WORKLOADS), help="thermal = BASELINE configs[1] (the headline); le / ns: other modules through the general path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="plan option key=value (tuning experiments), repeatable")
    ap.add_argument("--fn", action="append", default=[], help="Functions entry name=expression overriding the workload's deck (e.g. 'thermal diffusion=1.0+0.5*x*x'), repeatable")
    ap.add_argument("--perturb", type=float, default=0.0, help="thermal: move interior nodes by up to this fraction of h (general cells instead of boxes)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: n^3 elements PER GPU (z-slabs stacked); strong: n^3 elements in total, n/N layers per GPU")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu pass that measures the dominant kernel's DRAM traffic (roofline.traffic = null)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    global _WORKLOAD
    _WORKLOAD = args.workload
    if args.n <= 0:
        args.n = {"thermal": 128, "le": 64, "ns": 96, "leq2": 64, "maxwell": 64}[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
