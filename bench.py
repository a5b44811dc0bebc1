#!/usr/bin/env python
"""bench.py — converged game instances / second of the batched newton_solve! path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this framework (CUDA, sm_100a)
  python bench.py --impl reference [...]                       the reference algorithm's CPU restatement on host cores

Headline: a "step" is one newton_solve! of one batch of BASELINE config B = 1024 × 3-player DoubleIntegratorGame, N=40
per GPU (weak scaling, every rank draws its OWN instances: seed 1234 + rank; `--scaling strong` splits a fixed
8192-instance set N ways instead).  For N>1 every step ends with the path's single all-gather of the result slabs,
pushed over NVLink peer memory by the copy engines (agb_allgather; NCCL all_gather_into_tensor if IPC is unavailable).
`value`    device-timed (CUDA events on the launching stream, inputs resident in HBM, L2 flushed between steps).
`e2e`      the same metric through the public host-buffer API (x0 + initial iterate H2D from pinned memory, solve,
           trajectories/duals/multipliers/stats D2H) timed by the host clock.
`roofline` the solve is FP64-issue/latency bound, not HBM bound: achieved = executed FP64 flops of the solve kernel
           (ncu-counted DFMA/DADD/DMUL per Newton step × Newton steps of the launch) ÷ its launch duration, against the
           FP64 vector peak MEASURED on this GPU by agb_measure_fp64_peak; the SURVEY §8(d) KKT-band model and the real
           DRAM traffic are reported next to it.
`other_configs`  BASELINE configs C, D (MPC), E measured outside B's timed region (device-timed, e2e, converged
           fraction, bounded CPU-arm sample on the same inputs).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged game instances/sec (3-player DoubleIntegratorGame, N=40, batch 1024 per GPU)"
UNIT = "instances/s"


def workload(batch, seed):
    import algames_b200 as ab
    return ab.workloads.config_b(batch=batch, seed=seed)


def bytes_per_newton_step(p, n, m, N):
    b = n * p + m + n
    return 8 * (2 * 3 * b * b * (N - 1) + 4 * (N - 1) * b)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's plain-C restatement of the reference algorithm (oracle/algames_oracle.c) on host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_solve(cfg, x0, xf, Z0, L0, opts, threads=0, conlam=None, conmu=None):
    """Solve instances of any config with the C oracle.  Returns (out dict, seconds, threads used)."""
    import algames_b200 as ab
    from oracle import c_oracle
    model, N, dt, obj, con = cfg[:5]
    p, J, B = model.p, ab.problem._joint, x0.shape[0]
    desc = ab.problem._make_desc(model, N, dt, obj, con)
    tile = lambda v: np.tile(v, (B, 1))
    xfj = tile(J(obj.xf, p, 4)) if xf is None else xf
    t0 = time.perf_counter()
    out, used = c_oracle.newton_solve(desc, opts.to_c(), x0, xfj, tile(J(obj.Q, p, 4)), tile(J(obj.R, p, 2)), tile(J(obj.uf, p, 2)),
                                      Z0, L0, nthreads=threads, conlam=conlam, conmu=conmu)
    return out, time.perf_counter() - t0, used


def initial_iterate(opts, B, N, n, m, p, seed=None):
    rng = np.random.default_rng(opts.seed if seed is None else seed)
    return opts.amplitude_init * rng.random((B, N, n + m)), opts.amplitude_init * rng.random((B, p, N - 1, n))


def cpu_sample(n_inst, threads=0, seed=1234):
    """Solve `n_inst` config-B instances with the C oracle on `threads` host threads (0 = all cores).
    Returns (converged/s, seconds, converged, Newton steps, threads used)."""
    cfg = workload(n_inst, seed)
    model, N, dt, obj, con, opts, x0, _ = cfg
    Z0, L0 = initial_iterate(opts, n_inst, N, model.n, model.m, model.p)
    out, dt_s, used = cpu_solve(cfg, x0, None, Z0, L0, opts, threads)
    conv = int((out["status"] == 0).sum())
    return conv / dt_s, dt_s, conv, float(out["stats"][:, 6].sum()), used


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_inst = args.batch                       # the same 1024-instance batch the GPU arm solves per step
    for _ in range(args.warmup):
        cpu_sample(n_inst)
    t_tot, conv_tot, newton_tot, used = 0.0, 0, 0.0, 1
    for _ in range(args.steps):
        _, dt, conv, nn, used = cpu_sample(n_inst)
        t_tot += dt; conv_tot += conv; newton_tot += nn
    value = conv_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "B: batch=%d per GPU x 3-player DoubleIntegratorGame N=40 dt=0.1, collision cost + collision avoidance, jittered x0" % n_inst,
                   "options": "reference defaults", "seed": "1234 (rank 0's instance set of the GPU arm)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{args.steps} steps x {n_inst} config-B instances; oracle/algames_oracle.c: C restatement of Algames.jl "
                                   "newton_solve! (explicit KKT Jacobian + band LU with partial pivoting per Newton step), one pthread per core; "
                                   "the Julia reference itself cannot run in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "newton_steps_per_s": newton_tot / t_tot,
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi equivalent through NVML)
# ---------------------------------------------------------------------------------------------------------------
class Clocks:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.times, self.reasons, self.stop, self.max = [], [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop and self.nv:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.times.append(time.perf_counter())
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.t.join(timeout=1)

    def summary(self, t0=None, t1=None):
        """Median SM clock of the samples taken while the device executed the timed region [t0, t1] (host clock)."""
        inside = [c for c, t in zip(self.samples, self.times) if t0 is not None and t0 <= t <= t1]
        use = inside if inside else self.samples
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "samples_in_timed_region": len(inside)}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process plumbing: device, streams, L2 flush buffer, torch.distributed."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sp = self.stream.cuda_stream
        self.flush = torch.empty(20 << 20, dtype=torch.float64, device=self.dev)     # 160 MiB > 126 MB L2

    def pinned(self, shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype or self.torch.float64).pin_memory()

    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def allreduce(self, vals, op="sum"):
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN}[op])
        return [float(v) for v in t.tolist()]

    def gather_list(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [float(v)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]


def connect_peers(ctx, gb, batch):
    """agb_peer_init + IPC exchange over torch.distributed.  Returns 'peer' or (IPC refused) 'nccl'."""
    torch, dist = ctx.torch, ctx.dist
    ok = 1
    try:
        gb.peer_init(ctx.world, ctx.rank, [batch] * ctx.world)
        mine = gb.peer_export()
    except Exception as e:                                   # noqa: BLE001 — any CUDA IPC refusal means "fall back"
        print(f"[bench] rank {ctx.rank}: CUDA IPC export refused ({e}); falling back to NCCL all-gather", file=sys.stderr)
        ok, mine = 0, bytes(64)
    t = torch.tensor(list(mine), dtype=torch.uint8, device=ctx.dev)
    allh = [torch.zeros_like(t) for _ in range(ctx.world)]
    dist.all_gather(allh, t)
    if ok:
        try:
            gb.peer_connect([bytes(h.cpu().numpy().tobytes()) for h in allh])
        except Exception as e:                               # noqa: BLE001
            print(f"[bench] rank {ctx.rank}: CUDA IPC open refused ({e}); falling back to NCCL all-gather", file=sys.stderr)
            ok = 0
    return "peer" if ctx.allreduce([ok], "min")[0] > 0.5 else "nccl"


def run_headline(ctx, args):
    """Config B: `steps` solves of one batch per GPU; returns the JSON line's fields (rank 0 prints)."""
    import algames_b200 as ab
    from algames_b200 import distributed as D
    torch, dist, world, rank, dev = ctx.torch, ctx.dist, ctx.world, ctx.rank, ctx.dev
    strong = args.scaling == "strong"
    if strong:
        total = args.strong_total
        lo, hi = D.shard_bounds(total, world, rank)
        cfg = workload(total, 1234)
        model, N, dt, obj, con, opts, x0_all, _ = cfg
        x0, B = x0_all[lo:hi], hi - lo
        Z0a, L0a = initial_iterate(opts, total, N, model.n, model.m, model.p)
        Z0, L0 = Z0a[lo:hi], L0a[lo:hi]
        if ctx.allreduce([B], "min")[0] != ctx.allreduce([B], "max")[0]:
            raise SystemExit("--strong-total must be divisible by the number of GPUs")
    else:
        B = args.batch
        seed = 1234 + (0 if args.same_instances else rank)          # every rank draws its own instance set
        cfg = workload(B, seed)
        model, N, dt, obj, con, opts, x0, _ = cfg
        Z0, L0 = initial_iterate(opts, B, N, model.n, model.m, model.p, seed=opts.seed + (0 if args.same_instances else rank))
    n, m, p = model.n, model.m, model.p
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=ctx.local)
    h_x0 = ctx.pinned((B, n)); h_x0.numpy()[:] = x0
    h_Z0 = ctx.pinned((B, N, n + m)); h_Z0.numpy()[:] = Z0
    h_L0 = ctx.pinned((B, p, N - 1, n)); h_L0.numpy()[:] = L0
    gb.set_instance_params(x0=h_x0.numpy())
    gb.set_initial(h_Z0.numpy(), h_L0.numpy())
    stream, sp, flush = ctx.stream, ctx.sp, ctx.flush
    views = D.result_views(gb)
    slab = D.results_slab(gb)
    gather = "none"
    if world > 1 and not args.no_gather:
        gather = connect_peers(ctx, gb, B) if args.gather == "peer" else "nccl"
    comm = torch.cuda.Stream(device=dev) if gather == "nccl" else None
    stage = [torch.empty_like(slab) for _ in range(2)] if gather == "nccl" else None
    gathered = [torch.empty((world * slab.numel(),), dtype=torch.float64, device=dev) for _ in range(2)] if gather == "nccl" else None
    comm_done = [None, None]

    def step(k):
        gb.newton_solve_async(opts, sp)
        if gather == "peer":
            # the step's single collective: push this rank's result slab into every rank's gather buffer (copy engines,
            # ordered after the solve by an event; no SM is taken from the next step's solve)
            gb.allgather(sp)
        elif gather == "nccl":
            b = k & 1
            if comm_done[b] is not None:
                stream.wait_event(comm_done[b])               # staging buffer b is free again
            stage[b].copy_(slab, non_blocking=True)
            ready = torch.cuda.Event(); ready.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                D.all_gather_slabs(stage[b], out=gathered[b])
                comm_done[b] = torch.cuda.Event(); comm_done[b].record(comm)

    def drain():
        if gather == "peer":
            gb.allgather_wait()                               # this rank's pushes have landed (the barrier covers the peers')
        elif gather == "nccl":
            stream.wait_stream(comm)

    with torch.cuda.stream(stream):
        for k in range(max(args.warmup, 3)):
            flush.zero_()                       # (also warms the fill kernel: lazy module loading costs ~15 ms once)
            step(k)
        drain()
        ctx.sync_all()
        launches0 = gb.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        clk = Clocks(ctx.local)
        # the timed region is short (tens of ms): start sampling clocks / throttle reasons before it so that several
        # samples fall inside it; the sampler sleeps between NVML calls and does not touch the CUDA context
        with clk:
            time.sleep(0.02)
            t_wall0 = time.perf_counter()
            ev0.record(stream)
            for k in range(args.steps):
                flush.zero_()                       # L2 flush between steps (inside the timed region)
                kev[k][0].record(stream)
                step(k)
                kev[k][1].record(stream)            # brackets the solve kernel (+ the staging copy for the NCCL gather)
            if gather == "nccl":
                stream.wait_stream(comm)
            ev1.record(stream)
            torch.cuda.synchronize(dev)
            drain()                                 # peer pushes of this rank done; then every rank's (barrier)
            t_region_host = time.perf_counter() - t_wall0
        clocks = clk.summary(t_wall0, t_wall0 + t_region_host)
        ctx.sync_all()
        t_wall = time.perf_counter() - t_wall0
        launches = gb.launch_count() - launches0
    region_ms = max(ev0.elapsed_time(ev1), 1e3 * t_region_host if gather == "peer" else 0.0)   # peer pushes end after ev1
    kernel_list = [a.elapsed_time(b) for a, b in kev]
    kernel_ms = float(np.mean(kernel_list))
    gaps = {"first_ms": ev0.elapsed_time(kev[0][0]), "between_ms": float(np.mean([kev[k][1].elapsed_time(kev[k + 1][0]) for k in range(args.steps - 1)])) if args.steps > 1 else 0.0,
            "tail_ms": kev[-1][1].elapsed_time(ev1), "kernel_max_ms": float(np.max(kernel_list))}
    status = views["status"].clone()
    stats = views["stats"].clone()
    conv_local = float((status == 0).sum().item())
    newton_local = float(stats[:, 6].sum().item())
    total_ms = ctx.allreduce([region_ms], "max")[0]
    conv, newton = ctx.allreduce([conv_local, newton_local])
    per_rank_kernel = ctx.gather_list(kernel_ms)
    per_rank_region = ctx.gather_list(region_ms)
    if gather == "peer":                        # every rank's slab is in this rank's gather buffer
        torch.cuda.synchronize(dev)
        ptr, offs = gb.gathered_view()
        flat = D.device_tensor(ptr, (offs[-1] // 8,), "f8", ctx.local)
        for r in range(world):
            seg = flat[offs[r] // 8: offs[r + 1] // 8]
            if r == rank:
                assert torch.equal(seg, slab), "all-gather returned a different result slab"
            else:
                assert bool(torch.isfinite(seg[: B * N * (n + m)]).all()) and float(seg[: B * N * (n + m)].abs().sum()) > 0.0, "peer slab missing"
    elif gather == "nccl":
        mine = gathered[(args.steps - 1) & 1].view(world, -1)[rank]
        assert torch.equal(mine, slab), "all-gather returned a different result slab"
    value = conv * args.steps / (total_ms / 1e3)

    # ---- end-to-end through the host-buffer API (pinned host <-> device copies inside the timed region)
    nrow = gb.nrow
    h_out = {"Z": ctx.pinned((B, N, n + m)).numpy(), "L": ctx.pinned((B, p, N - 1, n)).numpy(),
             "stats": ctx.pinned((B, 10)).numpy(), "status": ctx.pinned((B,), torch.int32).numpy()}
    if nrow:
        h_out["conlam"], h_out["conmu"] = ctx.pinned((B, N - 1, nrow)).numpy(), ctx.pinned((B, N - 1, nrow)).numpy()

    def e2e_step():
        # one public API call: pinned host inputs -> solve -> pinned host results (chunked copy/solve pipeline inside)
        gb.solve_from_host(opts, h_x0.numpy(), h_Z0.numpy(), h_L0.numpy(), out=h_out)
        return int((h_out["status"] == 0).sum())

    for _ in range(2):
        e2e_step()
    ctx.sync_all()
    t0 = time.perf_counter()
    e2e_conv = 0
    for _ in range(args.steps):
        e2e_conv += e2e_step()
    ctx.sync_all()
    e2e_t = ctx.allreduce([time.perf_counter() - t0], "max")[0]
    e2e_c = ctx.allreduce([float(e2e_conv)])[0]
    h2d = 8 * (B * n + B * N * (n + m) + B * p * (N - 1) * n)
    d2h = 8 * (B * N * (n + m) + B * p * (N - 1) * n + 2 * B * (N - 1) * nrow + B * 10) + 4 * B

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        # FP64 vector peak measured on this GPU right now (register-resident DFMA chains)
        import ctypes as C
        tf, ms = C.c_double(), C.c_float()
        rc = gb.lib.agb_measure_fp64_peak(ctx.local, 1 << 16, C.byref(tf), C.byref(ms))
        fp64_peak = float(tf.value) if rc == 0 else None
        ncu = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                ncu = json.load(open(tpath))
            except Exception:
                ncu = {}
        newton_per_launch = newton / world
        flops_step = ncu.get("fp64_flops_per_newton_step")            # ncu: (dadd + dmul + 2·dfma thread instructions) / Newton steps of the captured launch
        achieved = flops_step * newton_per_launch / (kernel_ms / 1e3) / 1e12 if flops_step else None
        bps = bytes_per_newton_step(p, n, m, N)
        kkt_model_gbs = bps * newton_per_launch / (kernel_ms / 1e3) / 1e9
        traffic = ncu.get("dram_bytes_per_launch")
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("B: batch=%d per GPU x 3-player DoubleIntegratorGame N=40 dt=0.1, collision cost + collision avoidance, jittered x0" % B)
                       + (" (strong scaling: one %d-instance set, seed 1234, split over the GPUs)" % args.strong_total if strong else ""),
                       "seed": "1234 (same set on every rank)" if (strong or args.same_instances) else "1234 + rank (every rank draws its own instances)",
                       "l2": "flushed (160 MiB write) between steps, flush inside the timed region", "options": "reference defaults",
                       "collective": {"peer": "one all-gather of the result slabs per step, pushed by the copy engines over NVLink peer memory (agb_allgather, CUDA IPC), overlapped with the next step's solve",
                                      "nccl": "one ncclAllGather of the result slab per step on a side stream, overlapped with the next step's solve",
                                      "none": "none"}[gather]},
            "converged_fraction": conv / (B * world), "newton_steps_per_s": newton * args.steps / (total_ms / 1e3),
            "newton_steps_per_instance": newton / (B * world),
            "e2e": {"value": e2e_c / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "kernel_ms": kernel_ms, "region_gaps": gaps,
            "per_rank": {"kernel_ms": per_rank_kernel, "region_ms": per_rank_region,
                         "slowest_over_fastest_kernel": max(per_rank_kernel) / min(per_rank_kernel)},
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64-issue", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (achieved / fp64_peak) if (achieved and fp64_peak) else None, "traffic": traffic,
                         "peak_source": "agb_measure_fp64_peak on this GPU during this run (8 independent DFMA chains per thread, 2048 threads per SM)",
                         "kernel": "agb_newton_solve_kernel<3, DoubleIntegrator, small layout>",
                         "fp64_flops_per_newton_step": flops_step,
                         "flops_source": ncu.get("source", "profiles/traffic.json (ncu smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on of the solve kernel)"),
                         "fp64_pipe_busy_pct": ncu.get("fp64_pipe_busy_pct"), "issue_active_pct": ncu.get("issue_active_pct"),
                         "warp_inst_per_newton_step": ncu.get("warp_inst_per_newton_step"), "lanes_per_inst": ncu.get("lanes_per_inst"),
                         "waves": B / float(sms * ncu.get("ctas_per_sm", 4)),
                         "hbm": {"traffic_gbs": (traffic / (kernel_ms / 1e3) / 1e9) if traffic else None, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                                 "frac": (traffic / (kernel_ms / 1e3) / 1e9 / hbm_peak) if traffic else None},
                         "kkt_model_gbs": kkt_model_gbs,
                         "note": "the solve never materialises the KKT band: real DRAM traffic is the compulsory iterate in/out (hbm.frac), the kernel is bound by "
                                 "FP64 issue / dependent-chain latency. kkt_model_gbs = SURVEY 8(d) accounting (%d algorithmic B per Newton step of a band solver) "
                                 "x Newton steps per launch / launch time, kept for comparison with band implementations; it is NOT a bandwidth" % bps},
            "clocks": clocks,
        }
    gb.close()
    return line


def time_solves(ctx, gb, opts, reps):
    """Device time of `reps` back-to-back cold solves (L2 flushed before each), ms per solve."""
    torch = ctx.torch
    with torch.cuda.stream(ctx.stream):
        ctx.flush.zero_(); gb.newton_solve_async(opts, ctx.sp)        # warm-up
        torch.cuda.synchronize(ctx.dev)
        evs = []
        for _ in range(reps):
            ctx.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream); gb.newton_solve_async(opts, ctx.sp); b.record(ctx.stream)
            evs.append((a, b))
        torch.cuda.synchronize(ctx.dev)
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def run_cold_config(ctx, name, cfg, cpu_n, reps, gather_once=False):
    """One BASELINE shape solved cold on every rank's own shard: device-timed, e2e through the host API, converged
    fraction, and (rank 0, N=1 only) the C-port CPU arm on the first `cpu_n` instances of the same inputs."""
    import algames_b200 as ab
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    model, N, dt, obj, con, opts, x0, xf = cfg
    B, n, m, p = x0.shape[0], model.n, model.m, model.p
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=ctx.local)
    Z0, L0 = initial_iterate(opts, B, N, n, m, p, seed=opts.seed + rank)
    gb.set_instance_params(x0=x0, xf=xf)
    gb.set_initial(Z0, L0)
    mode = "none"
    if gather_once and world > 1:
        mode = connect_peers(ctx, gb, B)
    ms = time_solves(ctx, gb, opts, reps)
    gather_ms = 0.0
    if mode == "peer":                           # ONE all-gather at the end of the sweep, device-timed
        ctx.sync_all()
        t0 = time.perf_counter(); gb.allgather(ctx.sp); gb.allgather_wait(); ctx.sync_all()
        gather_ms = 1e3 * (time.perf_counter() - t0)
    from algames_b200 import distributed as D
    v = D.result_views(gb)
    conv_l, newton_l = float((v["status"] == 0).sum().item()), float(v["stats"][:, 6].sum().item())
    ms_max = ctx.allreduce([ms], "max")[0]
    gather_ms = ctx.allreduce([gather_ms], "max")[0]
    conv, newton = ctx.allreduce([conv_l, newton_l])
    # e2e: pinned host buffers through agb_solve_from_host
    h_x0, h_Z0, h_L0 = ctx.pinned((B, n)), ctx.pinned((B, N, n + m)), ctx.pinned((B, p, N - 1, n))
    h_x0.numpy()[:] = x0; h_Z0.numpy()[:] = Z0; h_L0.numpy()[:] = L0
    nrow = gb.nrow
    h_out = {"Z": ctx.pinned((B, N, n + m)).numpy(), "L": ctx.pinned((B, p, N - 1, n)).numpy(),
             "stats": ctx.pinned((B, 10)).numpy(), "status": ctx.pinned((B,), torch.int32).numpy()}
    if nrow:
        h_out["conlam"], h_out["conmu"] = ctx.pinned((B, N - 1, nrow)).numpy(), ctx.pinned((B, N - 1, nrow)).numpy()
    gb.solve_from_host(opts, h_x0.numpy(), h_Z0.numpy(), h_L0.numpy(), out=h_out)
    ctx.sync_all()
    t0 = time.perf_counter()
    gb.solve_from_host(opts, h_x0.numpy(), h_Z0.numpy(), h_L0.numpy(), out=h_out)
    e2e_conv_l = float((h_out["status"] == 0).sum())
    ctx.sync_all()
    e2e_t = ctx.allreduce([time.perf_counter() - t0], "max")[0]
    e2e_conv = ctx.allreduce([e2e_conv_l])[0]
    total_ms = ms_max + gather_ms
    rec = {"workload": name, "instances": int(B * world), "instances_per_gpu": int(B), "n_gpus": world,
           "value": conv / (total_ms / 1e3), "unit": "converged instances/s", "ms_per_solve": ms_max,
           "converged_fraction": conv / (B * world), "newton_steps_per_instance": newton / (B * world),
           "newton_steps_per_s": newton / (total_ms / 1e3), "layout": None,
           "e2e": {"value": e2e_conv / e2e_t, "unit": "converged instances/s",
                   "h2d_bytes": 8 * (B * n + B * N * (n + m) + B * p * (N - 1) * n),
                   "d2h_bytes": 8 * (B * N * (n + m) + B * p * (N - 1) * n + 2 * B * (N - 1) * nrow + B * 10) + 4 * B}}
    if gather_once and world > 1:
        rec["collective"] = {"kind": mode, "what": "ONE all-gather of every rank's result slab at the end of the sweep", "ms": gather_ms,
                             "bytes_per_rank": int(gb.device_view().results_bytes)}
    gb.close()
    if rank == 0 and world == 1 and cpu_n > 0:
        c = min(cpu_n, B)
        out, dt_s, used = cpu_solve(cfg, x0[:c], None if xf is None else xf[:c], Z0[:c], L0[:c], opts)
        cc = int((out["status"] == 0).sum())
        rec["cpu_baseline"] = {"value": cc / dt_s, "unit": "converged instances/s", "cores": used, "kind": "port",
                               "sample": f"first {c} instances of the same inputs in {dt_s:.1f} s (oracle/algames_oracle.c)",
                               "converged_fraction": cc / c,
                               "status_agreement_with_device": float((out["status"] == h_out["status"][:c]).mean())}
    return rec


def run_mpc_config(ctx, streams, resolves, cpu_streams, cpu_resolves):
    """BASELINE config D: `streams` MPC streams per GPU × `resolves` warm-started re-solves (shift = 1, multipliers carried),
    x0 ← x_2 + N(0, 1e-3) after every solve.  Device loop: solve + advance enqueued back to back, no host sync."""
    import algames_b200 as ab
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    total = streams * world
    cfg_all = ab.workloads.config_d(batch=total)
    model, N, dt, obj, con, opts, x0a, xfa = cfg_all
    lo, hi = rank * streams, (rank + 1) * streams
    x0, xf = x0a[lo:hi], xfa[lo:hi]
    n = model.n
    gb = ab.GameBatch(model, N, dt, obj, con, streams, device=ctx.local)
    first = ab.Options(**{**opts.to_dict(), "dual_reset": True})
    warm = ab.Options(**{**opts.to_dict(), "dual_reset": False, "shift": 1})
    dist_np = 1e-3 * np.random.default_rng(3456 + rank).standard_normal((resolves, streams, n))   # the same disturbances for both loops
    dist_dev = torch.from_numpy(dist_np).to(ctx.dev)
    from algames_b200 import distributed as D
    v = D.result_views(gb)

    def loop(collect):
        gb.set_instance_params(x0=x0, xf=xf)
        gb.random_initial(opts.amplitude_init, opts.seed + rank)
        conv = torch.zeros((), dtype=torch.float64, device=ctx.dev)
        newton = torch.zeros((), dtype=torch.float64, device=ctx.dev)
        torch.cuda.synchronize(ctx.dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hs = torch.cuda.ExternalStream(gb.stream(), device=ctx.dev)      # the whole loop on the handle's own stream
        with torch.cuda.stream(hs):
            a.record(hs)
            for t in range(resolves):
                gb.newton_solve_async(first if t == 0 else warm, hs.cuda_stream)
                if collect:
                    conv += (v["status"] == 0).sum(); newton += v["stats"][:, 6].sum()
                gb.mpc_advance_async(1, dist_dev[t].data_ptr())
            b.record(hs)
        torch.cuda.synchronize(ctx.dev)
        return a.elapsed_time(b), float(conv.item()), float(newton.item())

    # (1) the fused loop: agb_mpc_run_async — every stream's CTA runs all its re-solves inside ONE launch of the solve kernel
    stats_dev = torch.empty((resolves, streams, 10), dtype=torch.float64, device=ctx.dev)
    status_dev = torch.empty((resolves, streams), dtype=torch.int32, device=ctx.dev)
    xs_dev = torch.empty((resolves, streams, n), dtype=torch.float64, device=ctx.dev)

    def fused():
        gb.set_instance_params(x0=x0, xf=xf)
        gb.random_initial(opts.amplitude_init, opts.seed + rank)
        torch.cuda.synchronize(ctx.dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hs = torch.cuda.ExternalStream(gb.stream(), device=ctx.dev)
        with torch.cuda.stream(hs):
            a.record(hs)
            gb.mpc_run_async(first, resolves, 1, dist_dev.data_ptr(), stats_dev.data_ptr(), status_dev.data_ptr(), xs_dev.data_ptr(), hs.cuda_stream)
            b.record(hs)
        torch.cuda.synchronize(ctx.dev)
        return a.elapsed_time(b), float((status_dev == 0).sum().item()), float(stats_dev[:, :, 6].sum().item()), float(stats_dev[:, :, 6].max().item())

    fused()                                          # warm-up pass (same work)
    ms_f, conv_f, newton_f, newton_max = fused()
    ms_f_max = ctx.allreduce([ms_f], "max")[0]
    conv_fa, newton_fa = ctx.allreduce([conv_f, newton_f])
    newton_max = ctx.allreduce([newton_max], "max")[0]
    # (2) the step-wise loop (one launch per re-solve; a step lasts as long as its slowest stream), for comparison
    loop(False)
    ms, conv_l, newton_l = loop(True)
    ms_max = ctx.allreduce([ms], "max")[0]
    conv, newton = ctx.allreduce([conv_l, newton_l])
    # e2e: disturbances supplied from the host, every re-solve's stats, status and executed state copied back (agb_mpc_run)
    gb.set_instance_params(x0=x0, xf=xf)
    gb.random_initial(opts.amplitude_init, opts.seed + rank)
    ctx.sync_all()
    t0 = time.perf_counter()
    stats, status, xs = gb.mpc_run(first, resolves, 1, dist_np)
    ctx.sync_all()
    e2e_t = ctx.allreduce([time.perf_counter() - t0], "max")[0]
    e2e_conv = ctx.allreduce([float((status == 0).sum())])[0]
    same = bool(np.array_equal(status, status_dev.cpu().numpy()) and np.array_equal(stats, stats_dev.cpu().numpy()))
    # e2e of the step-wise host loop (trajectories, stats and status to the host after every re-solve)
    t0 = time.perf_counter()
    stats_s, status_s, xs_s = ab.mpc.mpc_run(gb, opts, x0, resolves, xf=xf, disturbances=dist_np)
    ctx.sync_all()
    e2e_step_t = ctx.allreduce([time.perf_counter() - t0], "max")[0]
    same_step = bool(np.array_equal(status_s, status) and np.array_equal(stats_s, stats))
    rec = {"workload": "D: MPC ramp merge, 3-player UnicycleGame N=40, shift=1, dual_reset=false, x0 <- x_2 + N(0,1e-3) after every solve",
           "streams": int(total), "streams_per_gpu": int(streams), "resolves": int(resolves), "n_gpus": world,
           "value": conv_fa / (ms_f_max / 1e3), "unit": "converged re-solves/s", "ms_per_resolve": ms_f_max / resolves,
           "converged_fraction": conv_fa / (total * resolves), "newton_steps_per_resolve": newton_fa / (total * resolves),
           "newton_steps_max_per_resolve": newton_max,
           "loop": "agb_mpc_run_async: ONE launch of the solve kernel, every stream's CTA runs its re-solves back to back in shared memory "
                   "(solve, x0 <- x_2 + disturbance, shift, re-solve) — no stream waits for another; timed with CUDA events",
           "stepwise": {"value": conv / (ms_max / 1e3), "ms_per_resolve": ms_max / resolves,
                        "loop": "agb_newton_solve_async + agb_mpc_advance_async per re-solve (5 launches each), no host sync: a step lasts as long as its slowest stream",
                        "e2e_value": e2e_conv / e2e_step_t, "identical_to_fused": same_step},
           "e2e": {"value": e2e_conv / e2e_t, "unit": "converged re-solves/s", "resolves": int(resolves), "identical_to_device_loop": same,
                   "what": "agb_mpc_run: disturbances from host memory, per-re-solve stats, status and executed states copied back"}}
    gb.close()
    if rank == 0 and world == 1 and cpu_streams > 0:
        # CPU arm on the same kind of loop: C oracle, cpu_streams streams × cpu_resolves re-solves
        c = min(cpu_streams, streams)
        cfg = (model, N, dt, obj, con, opts, x0[:c], xf[:c])
        rng = np.random.default_rng(3456)
        Z0, L0 = initial_iterate(opts, c, N, n, model.m, model.p)
        xcur, lam, mu, done, conv_c, t_cpu, used = x0[:c].copy(), None, None, 0, 0, 0.0, 1
        for t in range(cpu_resolves):
            out, dt_s, used = cpu_solve(cfg, xcur, xf[:c], Z0, L0, first if t == 0 else warm, conlam=lam, conmu=mu)
            t_cpu += dt_s; conv_c += int((out["status"] == 0).sum()); done += c
            xcur = out["Z"][:, 1, :n] + 1e-3 * rng.standard_normal((c, n))
            Z0 = np.concatenate([out["Z"][:, 1:], np.zeros_like(out["Z"][:, :1])], axis=1)
            L0 = np.concatenate([out["L"][:, :, 1:], np.zeros_like(out["L"][:, :, :1])], axis=2)
            lam, mu = out["conlam"], out["conmu"]
        rec["cpu_baseline"] = {"value": conv_c / t_cpu, "unit": "converged re-solves/s", "cores": used, "kind": "port",
                               "sample": f"{c} streams x {cpu_resolves} re-solves in {t_cpu:.1f} s (oracle/algames_oracle.c, same warm-start loop)",
                               "converged_fraction": conv_c / done}
    return rec


def run_other_configs(ctx, args):
    import algames_b200 as ab
    W = ab.workloads
    out = []
    world, rank = ctx.world, ctx.rank
    per = args.other_batch
    # E: Monte-Carlo highway sweep, 8192 distinct instances per GPU (65 536 on 8 GPUs), ONE all-gather at the end
    cfg = W.config_e(batch=per * world)
    sl = slice(rank * per, (rank + 1) * per)
    cfg_r = cfg[:6] + (cfg[6][sl], cfg[7][sl])
    r = run_cold_config(ctx, "E: %d x 3-player UnicycleGame N=60, highway (walls, collision avoidance, control bounds), seed 4567; rank r owns instances [r*%d, (r+1)*%d)" % (per * world, per, per),
                        cfg_r, args.cpu_sample_other, 2, gather_once=True)
    out.append(r)
    # D: MPC, 1024 streams per GPU × 200 re-solves (4096 streams on 4 GPUs)
    out.append(run_mpc_config(ctx, args.mpc_streams, args.mpc_resolves, 64, 20))
    # C: 4-player unicycle, single-GPU configuration of BASELINE.json
    if world == 1:
        cfg = W.config_c(batch=per)
        out.append(run_cold_config(ctx, "C: %d x 4-player UnicycleGame N=50, collision avoidance + control bounds, seed 2345" % per, cfg,
                                   min(args.cpu_sample_other, 256), 1))
        # Q: QuadrotorGame through the band solver; CPU arm = the NumPy oracle on ONE instance (the C port
        # has no quadrotor model), so its figure is a single-core one
        if args.quad_batch > 0:
            out.append(run_quadrotor_config(ctx, args.quad_batch, args.quad_cpu))
    return out


def run_quadrotor_config(ctx, B, cpu_n):
    import algames_b200 as ab
    torch = ctx.torch
    cfg = ab.workloads.config_q(batch=B, N=12, p=2)
    model, N, dt, obj, con, opts, x0, xf = cfg
    n, m, p = model.n, model.m, model.p
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=ctx.local)
    Z0, L0 = initial_iterate(opts, B, N, n, m, p)
    gb.set_instance_params(x0=x0)
    gb.set_initial(Z0, L0)
    ms = time_solves(ctx, gb, opts, 1)
    out = gb.solve_from_host(opts, x0, Z0, L0)          # warm-up of the chunked host pipeline
    t0 = time.perf_counter()
    out = gb.solve_from_host(opts, x0, Z0, L0)
    e2e_t = time.perf_counter() - t0
    conv = int((out["status"] == 0).sum())
    rec = {"workload": "Q: %d x 2-player QuadrotorGame N=12 (12 states, 4 rotor commands per player), spherical collision avoidance, rotor bounds, 3-D wall, cylinder; band solver" % B,
           "instances": B, "n_gpus": 1, "value": conv / (ms / 1e3), "unit": "converged instances/s", "ms_per_solve": ms,
           "converged_fraction": conv / B, "newton_steps_per_instance": float(out["stats"][:, 6].mean()),
           "e2e": {"value": conv / e2e_t, "unit": "converged instances/s"}}
    gb.close()
    if cpu_n > 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle.algames_oracle as O
        import parity
        t0 = time.perf_counter()
        cc = 0
        for b in range(cpu_n):
            op = parity.oracle_problem(model, N, dt, obj, con, opts, x0[b])
            O.newton_solve(op, Z0=Z0[b], L0=L0[b])
            cc += int(op.converged)
            agree = bool(op.converged == (out["status"][b] == 0)) and abs(float(np.abs(np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1) - out["Z"][b]).max())) < 1e-6
        dt_s = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": cc / dt_s, "unit": "converged instances/s", "cores": 1, "kind": "port",
                               "sample": f"{cpu_n} instance(s) of the same inputs in {dt_s:.1f} s (oracle/algames_oracle.py, NumPy, one core; the C port has no quadrotor model)",
                               "agrees_with_device_to_1e-6": agree}
    return rec


def run_gpu(args):
    ctx = Ctx()
    line = run_headline(ctx, args)
    others = None
    if not args.no_other_configs:
        try:
            others = run_other_configs(ctx, args)
        except Exception as e:                               # noqa: BLE001 — the headline line must still be printed
            import traceback
            traceback.print_exc()
            others = {"error": f"{type(e).__name__}: {e}"}
    if ctx.rank == 0:
        if others is not None:
            line["other_configs"] = others
        if ctx.world == 1 and not args.no_cpu_baseline:
            v, dt_s, c, nn, used = cpu_sample(args.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                                    "sample": f"{args.cpu_sample} config-B instances in {dt_s:.1f} s; oracle/algames_oracle.c (C restatement of "
                                              "Algames.jl newton_solve!: explicit KKT Jacobian + band LU per Newton step), one pthread per core"}
        print(json.dumps(line), file=_OUT, flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="instances per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: one --strong-total instance set split over the GPUs")
    ap.add_argument("--strong-total", type=int, default=8192)
    ap.add_argument("--same-instances", action="store_true", help="diagnostic: every rank solves the seed-1234 set (round-1 behaviour)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"], help="the per-step all-gather: copy-engine pushes over peer memory, or NCCL")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="instances of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-sample-other", type=int, default=1024, help="instances of the CPU samples of the other configs")
    ap.add_argument("--other-batch", type=int, default=8192, help="instances per GPU of configs C and E")
    ap.add_argument("--mpc-streams", type=int, default=1024, help="config D streams per GPU")
    ap.add_argument("--mpc-resolves", type=int, default=200)
    ap.add_argument("--quad-batch", type=int, default=512, help="config Q (QuadrotorGame, band solver) instances; 0 skips it")
    ap.add_argument("--quad-cpu", type=int, default=1, help="instances of config Q solved by the NumPy oracle for its CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic: skip the all-gather at N>1")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything any library writes to fd 1 (NCCL prints its version banner there at
    # NCCL_DEBUG=VERSION and WARN) goes to stderr instead, and the line itself is written to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
