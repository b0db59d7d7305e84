#!/usr/bin/env python3
"""Benchmark of the PuckSwarm per-tick hot path (BASELINE.json): robot-steps/sec, 16 robots / 200 pucks x N arenas.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one think interval of RunOffline.step (src/arena/RunOffline.java:72-95) over the whole batch:
Arena.step (sense + think + move, CacheCons on device) followed by stepInterval = 3 World.step(1/14, 3, 100)
ticks. One robot-step = one robot advanced through one physics tick. At N = 1 the workload is BASELINE config
[2]: 4096 arenas x 16 robots x 200 pucks, seeds = arena index, spawned by the reference's rejection sampling;
for N > 1 every rank owns 4096 arenas (weak scaling, arenas shard by index, no data-path collective; the
statistics all-reduce over NCCL runs once at the end of the timed region).

value : device-resident state, CUDA events on the engine's stream around K steps, max over ranks.
e2e   : the same K steps through the C-ABI with HOST buffers: every step uploads the bodies' poses/velocities
        from pinned host memory (ps_set_state), steps, and reads them back (ps_get_state) plus the step counters.
--impl reference: the CPU restatement of the reference (oracle/, "port": the Java reference cannot be built in this
        image) on all host cores, one arena per thread, on a bounded sample of the same workload.
"""
import argparse
import json
import os

# the engine runs arena groups on many streams: give CUDA enough hardware queues before the context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_ROBOTS, N_PUCKS, ARENAS_PER_GPU, STEP_INTERVAL = 16, 200, 4096, 3
STEADY_STEPS = 200      # think steps timed for also.steady_state (600 ticks) after >= 300 warm-up ticks
FAST_POS_ITERS = 10     # also.fast_mode: position-iteration cap (the reference passes 100)
PARITY_ARENAS = 8       # arenas of the timed batch replayed by the oracle after the clocks stop
METRIC = "robot-steps/sec, 16 robots/200 pucks x N arenas"
UNIT = "robot-steps/s"


def workload_config(n_arenas, parallelism):
    return {
        "workload": "BASELINE config[2]: %d arenas x 16 robots x 200 pucks (2 colours), CacheCons on device, camera + local map "
                    "every 3rd tick, World.step(1/14,3,100), enclosure scale 1, seeds = arena index" % n_arenas,
        "arenas": n_arenas, "robots_per_arena": N_ROBOTS, "pucks_per_arena": N_PUCKS, "step_interval": STEP_INTERVAL,
        "ticks_per_step": STEP_INTERVAL, "rng": "java.util.Random shared per arena (parity mode)",
        "parallelism": parallelism,
        "cache": "inputs larger than L2: per-arena state x arenas = ~1 GB per GPU vs 126 MB L2, no flush needed",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port_run(n_threads, warm_ticks, ticks, budget_s):
    """The oracle (CPU restatement) on n_threads arenas of the benchmark shape, one arena per thread, over the same simulated
    interval as the GPU arm (warm_ticks after the reset, then `ticks`), cut short when budget_s runs out. Returns robot-steps/s."""
    from puckswarm_b200 import abi
    import oracle_binding
    cfg = abi.config_defaults(n_arenas=n_threads, n_robots=N_ROBOTS, n_pucks=N_PUCKS, controller=abi.PS_CTRL_CACHECONS)
    o = oracle_binding.Oracle(cfg, threads=n_threads)
    o.reset(np.arange(n_threads))
    t0 = time.perf_counter()
    done_warm = 0
    while done_warm < warm_ticks and time.perf_counter() - t0 < budget_s:
        n = min(30, warm_ticks - done_warm)
        o.step(n)
        done_warm += n
    t0 = time.perf_counter()
    done = 0
    while done < ticks and (done == 0 or time.perf_counter() - t0 < budget_s):
        n = min(30, ticks - done)
        o.step(n)
        done += n
    dt = time.perf_counter() - t0
    return n_threads * N_ROBOTS * done / dt, done, dt, done_warm


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only). A step is one think
    interval (3 ticks) of one arena per host thread; the same W warm-up and K timed steps as the GPU arm, within a time budget."""
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build_oracle()
    cores = os.cpu_count() or 1
    value, ticks, dt, warm = cpu_port_run(cores, STEP_INTERVAL * args.warmup, STEP_INTERVAL * args.steps, 90.0)
    sample = "%d arenas (one per host thread) x 16 robots x 200 pucks, %d timed ticks after %d warm-up ticks (%.1f s)" % (cores, ticks, warm, dt)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * cores * N_ROBOTS * STEP_INTERVAL / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(ARENAS_PER_GPU * args.gpus, "host threads, one arena per core"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of JBox2D 2.1.2.2 + PuckSwarm sensing/controllers (oracle/), not the JVM reference: no JDK / JBox2D jar in this image",
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """Only the JSON line may reach stdout: libraries that print there (NCCL's version banner) are sent to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults follow the metric's definition: >= 300 warm-up ticks, then a few hundred think steps
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arenas", type=int, default=ARENAS_PER_GPU, help="arenas per GPU (default: BASELINE config[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (stepInterval = 1, external controller, configs[3] / [4])")
    ap.add_argument("--no-steady", action="store_true", help="skip also.steady_state (the >= 300 warm-up ticks regime) when W / K are shorter")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle replay of arenas of the timed batch")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="A/B runs only: skip the end-to-end leg (the line then carries e2e = null and is not a bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from puckswarm_b200 import abi, engine, dist as psd
    import __graft_entry__ as g

    if not os.path.exists(engine.LIB_PATH):
        g.build_engine()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    A = args.arenas
    lo = rank * A   # arenas shard by index: rank r owns [r * A, (r + 1) * A)
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_engine(n_arenas, first_seed, **kw):
        c = abi.config_defaults(n_arenas=n_arenas, n_robots=kw.pop("n_robots", N_ROBOTS), n_pucks=kw.pop("n_pucks", N_PUCKS),
                                controller=kw.pop("controller", abi.PS_CTRL_CACHECONS), rng_mode=abi.PS_RNG_JAVA_SHARED, **kw)
        e = engine.Engine(c, device=local_rank)
        e.reset(np.arange(first_seed, first_seed + n_arenas))
        return e

    def timed_steps(e, n_steps, ticks_per_step=STEP_INTERVAL, sample_clocks=False):
        """n_steps think intervals with device-resident state: CUDA events on the engine's stream, barrier + synchronize on both
        sides, max over ranks. Returns (ms, launches, per-kernel ms, per-kernel launch counts, clocks)."""
        stream = torch.cuda.ExternalStream(e.stream, device=local_rank)
        e.sync()
        e.profile(True)
        e.kernel_ms()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        l0 = e.kernel_launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(n_steps):
            e.step(ticks_per_step, sync=False)
        e.join()
        ev1.record(stream)
        e.sync()   # raises on a capacity flag in any arena: a degraded batch is not a measurement
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = e.kernel_launches - l0
        kms, kcnt = e.kernel_ms()
        e.profile(False)
        return ms, launches, kms, kcnt, clocks

    def rooflines(e, kms, kcnt, stats, clocks, regime):
        """roofline (World.step pipeline, HBM) and roofline_sense (camera stage, shared memory) of one timed region."""
        peak, peak_src = measured_peaks()
        c_hat = stats.n_touching_points / max(1, stats.n_arenas)   # live touching manifold points per arena
        bytes_per_arena_tick = 2.0 * (24.0 * (N_ROBOTS + N_PUCKS) + 24.0 * c_hat + 96.0 * N_ROBOTS)
        # one World.step launch sequence of one arena group = k_phys_pre + island kernels + k_phys_post, each timed by its own
        # event pair on the group's stream
        phys_ms = STEP_INTERVAL * sum(kms[k] / max(1, kcnt[k]) for k in ("phys_pre", "phys_islands", "phys_post"))
        arenas_per_launch = e.cfg.n_arenas / max(1, e.num_groups)   # every arena group launches its own kernel sequence
        alg_bytes_per_launch = bytes_per_arena_tick * arenas_per_launch * STEP_INTERVAL
        achieved = alg_bytes_per_launch / (phys_ms / 1000.0) / 1e9 if phys_ms > 0 else 0.0
        top = ("sense", "think", "physics", "other")
        total_k = sum(kms[k] for k in top)
        roofline = {
            "kernel": "World.step pipeline (k_phys_pre + island kernels + k_phys_post) x %d ticks" % STEP_INTERVAL,
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": None, "regime": regime,
            "algorithmic_bytes_per_arena_tick": bytes_per_arena_tick, "touching_points_per_arena": c_hat,
            "kernel_ms_per_launch": phys_ms, "arenas_per_launch": arenas_per_launch, "arena_groups": e.num_groups,
            "kernel_share_of_step": {k: (kms[k] / total_k if total_k > 0 else 0.0) for k in top},
            "world_step_stage_ms_per_tick": {k: kms[k] / max(1, kcnt[k]) for k in ("phys_pre", "phys_islands", "phys_post")},
            "kernel_ms_total": kms, "kernel_launch_counts": kcnt,
        }
        # DRAM bytes the World.step kernels move: cannot be counted outside a profiler, so it comes from the ncu capture of
        # the same regime (profiles/k_physics_traffic.json names the capture), scaled to this launch like `achieved`
        ncu_traffic = os.path.join(ROOT, "profiles", "k_physics_traffic.json")
        if os.path.exists(ncu_traffic):
            try:
                j = json.load(open(ncu_traffic))
                per_arena_tick = (j.get("by_regime") or {}).get(regime, {}).get("dram_bytes_per_arena_tick")
                if per_arena_tick is not None:
                    roofline["traffic"] = per_arena_tick * arenas_per_launch * STEP_INTERVAL
                    roofline["traffic_bytes_per_arena_tick"] = per_arena_tick
                    roofline["traffic_source"] = (j.get("by_regime") or {}).get(regime, {}).get("source")
            except Exception:
                pass
        # shared-memory roofline of the camera stage (algorithmic smem bytes per robot-sense)
        sense_dbg = e.sense_stats().reshape(-1, 8).astype(np.float64)
        p_act = 4556.0
        k_bar = float(sense_dbg[:, 4].mean()) / p_act            # candidate bodies examined per pixel (measured)
        smem_bytes_per_sense = p_act * 9.0 + p_act * k_bar * 12.0 + 3825.0 + 2294.0 * 3.0 + 19040.0
        sense_ms = kms["sense"] / max(1, kcnt["sense"])             # the camera + local-map stage of one arena group
        robots_per_launch = arenas_per_launch * N_ROBOTS
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9   # GB/s: 148 SMs x 128 B/clk x SM clock
        smem_achieved = smem_bytes_per_sense * robots_per_launch / (sense_ms / 1000.0) / 1e9 if sense_ms > 0 else 0.0
        roofline_sense = {
            "kernel": "camera + local-map stage (camera, occupancy, blobs, local map) per arena group", "bound": "smem",
            "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s", "frac": smem_achieved / smem_peak,
            "peak_source": "148 SMs x 128 B/clk x sampled SM clock", "candidates_per_pixel": k_bar,
            "algorithmic_bytes_per_robot_sense": smem_bytes_per_sense, "kernel_ms_per_launch": sense_ms, "robots_per_launch": robots_per_launch,
            "note": "candidate binning cuts K from the ~12 the survey assumed to the measured value, so the algorithmic byte count (and this fraction) "
                    "shrinks with it; the stage is issue/latency bound (profiles/)",
        }
        # issue-slot view of the same stage (the smem byte count shrank with K, so the smem fraction says little): executed warp
        # instructions and issue utilisation cannot be counted outside a profiler; they come from the ncu capture profiles/k_sense_issue.json names
        issue_json = os.path.join(ROOT, "profiles", "k_sense_issue.json")
        if os.path.exists(issue_json):
            try:
                roofline_sense["issue_slots"] = json.load(open(issue_json))
            except Exception:
                pass
        return roofline, roofline_sense

    def e2e_run(n_arenas, first_seed, warm_steps, timed_steps_n):
        """The same steps through the C-ABI with HOST buffers: per step ps_set_state(body) from pinned memory, ps_step(3),
        ps_get_state(body, arena) into pinned memory; host wall clock around the timed steps, max over ranks."""
        # PS_E2E_HANDLES > 1 splits the arenas over several handles served round-robin (their copies overlap the others' steps)
        n_handles = int(os.environ.get("PS_E2E_HANDLES", "1"))
        halves, h2d, d2h = [], 0, 0
        for hi in range(n_handles):
            a0, a1 = hi * n_arenas // n_handles, (hi + 1) * n_arenas // n_handles
            e2 = make_engine(a1 - a0, first_seed + a0)
            st = e2.new_state()
            pinned_body = torch.empty(st.body.shape, dtype=torch.float32).pin_memory()
            pinned_arena = torch.empty(st.arena.nbytes, dtype=torch.uint8).pin_memory()
            del st
            body, arena = pinned_body.numpy(), pinned_arena.numpy().view(abi.ARENA_DTYPE)
            vin, vout = abi.ps_state_view(), abi.ps_state_view()
            vin.body = body.ctypes.data_as(C.POINTER(C.c_float))
            vout.body = body.ctypes.data_as(C.POINTER(C.c_float))
            vout.arena = arena.ctypes.data_as(C.POINTER(abi.ps_arena_state))
            e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))
            h2d += body.nbytes
            d2h += body.nbytes + arena.nbytes
            halves.append((e2, vin, vout, pinned_body, pinned_arena))

        def e2e_round(first):
            for e2, vin, vout, _, _ in halves:
                if not first:
                    e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))   # D2H: poses + velocities + step counters (synchronous)
                e2._check(e2.L.ps_set_state(e2.h, C.byref(vin)))          # H2D: poses + velocities of every body
                e2._check(e2.L.ps_step(e2.h, STEP_INTERVAL))

        def e2e_drain():
            for e2, vin, vout, _, _ in halves:
                e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))

        e2e_round(True)
        for _ in range(warm_steps - 1):
            e2e_round(False)
        e2e_drain()
        barrier()
        t0 = time.perf_counter()
        e2e_round(True)
        for _ in range(timed_steps_n - 1):
            e2e_round(False)
        e2e_drain()                       # the last step's results are on the host when the clock stops
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
        for e2, *_ in halves:
            e2.sync()
            e2.close()
        return {"value": world * n_arenas * N_ROBOTS * STEP_INTERVAL * timed_steps_n / (ms / 1000.0), "unit": UNIT,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "call": "per step and per handle: ps_set_state(body) + ps_step(3) + ps_get_state(body, arena), pinned host buffers; "
                        "%d handle(s) served round-robin; host wall clock" % n_handles}

    import ctypes as C
    eng = make_engine(A, lo)

    # ---- device-resident throughput ("value"): W warm-up steps after the reset, then K timed steps ----
    for _ in range(args.warmup):
        eng.step(STEP_INTERVAL, sync=False)
    ms_max, launches, kms, kcnt, clocks = timed_steps(eng, args.steps, sample_clocks=True)
    local_stats = eng.compute_stats()
    total_stats = psd.allreduce_stats(local_stats, device=dev) if world > 1 else local_stats
    robot_steps = world * A * N_ROBOTS * STEP_INTERVAL * args.steps
    value = robot_steps / (ms_max / 1000.0)
    main_regime = "steady" if args.warmup * STEP_INTERVAL >= 300 else "early"
    roofline, roofline_sense = rooflines(eng, kms, kcnt, total_stats, clocks, main_regime)
    ticks_done = (args.warmup + args.steps) * STEP_INTERVAL

    # ---- the steady state the metric is defined on (>= 300 warm-up ticks, then >= 600 timed ticks), whatever
    #      --steps / --warmup were: the same engine keeps running until tick 300, then STEADY_STEPS think steps are timed ----
    also = {}
    if args.warmup * STEP_INTERVAL >= 300 and args.steps * STEP_INTERVAL >= 600:
        steady = {"value": value, "unit": UNIT, "ms_per_step": ms_max / args.steps, "warmup_ticks": args.warmup * STEP_INTERVAL,
                  "timed_ticks": args.steps * STEP_INTERVAL, "same_as": "main line", "roofline": roofline, "roofline_sense": roofline_sense}
    elif not args.no_steady:
        while ticks_done < 300:
            eng.step(STEP_INTERVAL, sync=False)
            ticks_done += STEP_INTERVAL
        warm_ticks = ticks_done
        s_ms, s_launches, s_kms, s_kcnt, s_clocks = timed_steps(eng, STEADY_STEPS, sample_clocks=True)
        ticks_done += STEADY_STEPS * STEP_INTERVAL
        s_local = eng.compute_stats()
        s_stats = psd.allreduce_stats(s_local, device=dev) if world > 1 else s_local
        s_roof, s_roof_sense = rooflines(eng, s_kms, s_kcnt, s_stats, s_clocks, "steady")
        steady = {"value": world * A * N_ROBOTS * STEP_INTERVAL * STEADY_STEPS / (s_ms / 1000.0), "unit": UNIT, "ms_per_step": s_ms / STEADY_STEPS,
                  "warmup_ticks": warm_ticks, "timed_ticks": STEADY_STEPS * STEP_INTERVAL, "gpu_launches": int(s_launches), "clocks": s_clocks,
                  "roofline": s_roof, "roofline_sense": s_roof_sense,
                  "stats": {"mean_percent_completion": psd.mean_completion(s_stats), "cached_contacts_per_arena": s_stats.n_contacts / max(1, s_stats.n_arenas)}}
    else:
        steady = None
    if steady is not None:
        also["steady_state"] = steady

    # ---- no arena of the timed batch may have raised a capacity flag (ps_sync would have failed the run above; this is the count) ----
    err_flags, err_arenas = eng.error_flags()
    if err_flags:
        raise SystemExit("bench.py: %d arena(s) raised capacity flags 0x%x: the batch left the reference's trajectory" % (err_arenas, err_flags))

    # ---- parity of the timed batch: the first PARITY_ARENAS arenas of this rank against the oracle on the same seeds, same ticks ----
    parity_check = None
    n_par = 0 if (args.no_parity or rank != 0) else min(PARITY_ARENAS, A)
    if n_par:
        body = np.zeros((A, 6, eng.NB), np.float32)
        robots = np.zeros((A, N_ROBOTS), abi.ROBOT_DTYPE)
        arenas = np.zeros((A,), abi.ARENA_DTYPE)
        v = abi.ps_state_view()
        v.body = body.ctypes.data_as(C.POINTER(C.c_float))
        v.robot = robots.ctypes.data_as(C.POINTER(abi.ps_robot_state))
        v.arena = arenas.ctypes.data_as(C.POINTER(abi.ps_arena_state))
        eng._check(eng.L.ps_get_state(eng.h, C.byref(v)))
    eng.close()

    # ---- end to end through the C-ABI with host buffers ("e2e"): same W and K as the main line ----
    e2e = None if args.no_e2e else e2e_run(A, lo, args.warmup, args.steps)
    if steady is not None and "same_as" not in steady and not args.no_extras and not args.no_e2e:
        steady["e2e"] = e2e_run(A, lo, 100, STEADY_STEPS)
    elif steady is not None and "same_as" in steady:
        steady["e2e"] = e2e

    if not args.no_extras:
        # ---- the same batch with the sensors on every tick (stepInterval = 1 beside the reference's 3) ----
        eng1 = make_engine(A, lo, step_interval=1)
        for _ in range(9):
            eng1.step(1, sync=False)
        n1 = 15
        t1, *_ = timed_steps(eng1, n1, ticks_per_step=1)
        also["step_interval_1"] = {"value": world * A * N_ROBOTS * n1 / (t1 / 1000.0), "unit": UNIT, "ms_per_tick": t1 / n1, "ticks": n1, "warmup_ticks": 9}
        eng1.close()

        # ---- north_star: "a batched per-tick JNI callback remains for arbitrary Java controllers, reported separately".
        #      One think interval = ps_sense + ps_get_observations (local map, occupancy, pose into pinned host memory) + a host
        #      controller over the whole batch + ps_step_external (commands up, 3 ticks). The host controller here is a vectorised
        #      stand-in (steer towards the nearest perceived puck): the number is the round trip, not the controller. ----
        engx = make_engine(A, lo, controller=abi.PS_CTRL_EXTERNAL)
        obs = engx.new_observations(camera=False)
        pin = {k: torch.empty(getattr(obs, k).nbytes, dtype=torch.uint8).pin_memory() for k in ("occupancy", "localmap", "pose")}
        obs.occupancy = pin["occupancy"].numpy().view(np.uint8).reshape(obs.occupancy.shape)
        obs.localmap = pin["localmap"].numpy().view(abi.LOCALMAP_DTYPE).reshape(obs.localmap.shape)
        obs.pose = pin["pose"].numpy().view(np.float32).reshape(obs.pose.shape)
        d2h_x = obs.occupancy.nbytes + obs.localmap.nbytes + obs.pose.nbytes

        def external_step():
            engx._check(engx.L.ps_sense(engx.h))
            engx.get_observations(obs, camera=False)
            lm = obs.localmap
            has = lm["n_pucks"] > 0
            ang = np.where(has, np.arctan2(lm["puck_y"][..., 0], np.maximum(lm["puck_x"][..., 0], 1e-3)), 0.3).astype(np.float32)
            fwd = np.full((A, N_ROBOTS), 375000.0, np.float32)
            tq = (np.clip(ang, -1.0, 1.0) * 2375000.0).astype(np.float32)
            engx.step_external(fwd, tq, sync=False)

        for _ in range(3):
            external_step()
        engx.sync()
        barrier()
        nx = 10
        t0 = time.perf_counter()
        for _ in range(nx):
            external_step()
        engx.sync()
        tx = max_over_ranks((time.perf_counter() - t0) * 1000.0)
        also["external_controller"] = {
            "value": world * A * N_ROBOTS * STEP_INTERVAL * nx / (tx / 1000.0), "unit": UNIT, "ms_per_step": tx / nx, "steps": nx,
            "d2h_bytes_per_step": int(d2h_x), "h2d_bytes_per_step": int(2 * A * N_ROBOTS * 4),
            "call": "per think step: ps_sense + ps_get_observations(localmap, occupancy, pose; pinned) + host controller (vectorised numpy stand-in) "
                    "+ ps_step_external; host wall clock; controllers.Controller seam (Controller.java:9-31), reported separately from the on-device path"}
        engx.close()

        # ---- FAST mode (north_star: "fixed-iteration order"): the same engine with JBox2D's position-iteration cap at 10 instead
        #      of the 100 RunOffline passes. NOT the reference's configuration and never the headline: labelled, reported beside
        #      it, gated by the ensemble KS test of tools/gpu_fast_mode_gate.py (profiles/r02_fast_mode_gate.json) ----
        engf = make_engine(A, lo, pos_iters=FAST_POS_ITERS)
        for _ in range(100):
            engf.step(STEP_INTERVAL, sync=False)
        nf = 60
        tf, *_ = timed_steps(engf, nf)
        gate = None
        try:
            gate = json.load(open(os.path.join(ROOT, "profiles", "r02_fast_mode_gate.json"))).get("gate")
        except Exception:
            pass
        also["fast_mode"] = {"value": world * A * N_ROBOTS * STEP_INTERVAL * nf / (tf / 1000.0), "unit": UNIT, "ms_per_step": tf / nf, "steps": nf, "warmup": 100,
                             "pos_iters": FAST_POS_ITERS, "parity": "NOT the reference configuration (World.step(1/14, 3, 100)); bit-exact against the oracle run "
                             "with the same cap", "ensemble_ks_gate": gate, "gate_file": "profiles/r02_fast_mode_gate.json"}
        engf.close()

        # ---- BASELINE configs[3]: 65 536 arenas x 16 x 200 sharded over the GPUs of the job (named for 2 / 4 / 8 GPUs) ----
        if world > 1 and 65536 % world == 0:
            A3 = 65536 // world
            eng3 = make_engine(A3, rank * A3)
            for _ in range(3):
                eng3.step(STEP_INTERVAL, sync=False)
            n3 = 6
            t3, *_ = timed_steps(eng3, n3)
            s3 = psd.allreduce_stats(eng3.compute_stats(), device=dev)
            also["config3"] = {"workload": "BASELINE configs[3]: 65536 arenas x 16 robots x 200 pucks over %d GPUs (%d per GPU), NCCL statistics reduce" % (world, A3),
                               "value": 65536 * N_ROBOTS * STEP_INTERVAL * n3 / (t3 / 1000.0), "unit": UNIT, "ms_per_step": t3 / n3, "steps": n3, "warmup": 3,
                               "arenas_reduced": int(s3.n_arenas), "regime": "early (ticks 9-27)"}
            eng3.close()

        # ---- BASELINE configs[4]: 1024 large arenas x 256 robots x 4000 pucks, enclosure scale 6 (sharded like the rest) ----
        if 1024 % world == 0 and not args.no_config4:
            A4 = 1024 // world
            eng4 = make_engine(A4, rank * A4, n_robots=256, n_pucks=4000, enclosure_scale=6.0)
            for _ in range(2):
                eng4.step(STEP_INTERVAL, sync=False)
            n4 = 4
            t4, *_ = timed_steps(eng4, n4)
            also["config4"] = {"workload": "BASELINE configs[4]: 1024 arenas x 256 robots x 4000 pucks, RoundedRectangleEnclosure.scale = 6, CacheCons on device, "
                                           "local map on; %d arenas per GPU" % A4,
                               "value": 1024 * 256 * STEP_INTERVAL * n4 / (t4 / 1000.0), "unit": UNIT, "ms_per_step": t4 / n4, "steps": n4, "warmup": 2,
                               "regime": "early (ticks 6-18)"}
            eng4.close()

    # ---- parity_check, continued: the oracle replays the sampled arenas (clocks are stopped; one arena per host thread) ----
    if n_par:
        g.build_oracle()
        import oracle_binding
        ocfg = abi.config_defaults(n_arenas=n_par, n_robots=N_ROBOTS, n_pucks=N_PUCKS, controller=abi.PS_CTRL_CACHECONS, rng_mode=abi.PS_RNG_JAVA_SHARED)
        orc = oracle_binding.Oracle(ocfg, threads=min(n_par, os.cpu_count() or 1))
        orc.reset(np.arange(lo, lo + n_par))
        t0 = time.perf_counter()
        orc.step(ticks_done)
        ost = orc.get_state()
        ref, got = ost.body, body[:n_par]
        rel = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        per_arena_bad = (rel.reshape(n_par, -1).max(axis=1) > 1e-5) | (robots[:n_par]["ctrl_state"] != ost.robot["ctrl_state"]).any(axis=1) \
            | (arenas[:n_par]["n_contacts"] != ost.arena["n_contacts"]) | (arenas[:n_par]["update_count"] != ost.arena["update_count"])
        parity_check = {"arenas": int(n_par), "ticks": int(ticks_done), "mismatches": int(per_arena_bad.sum()),
                        "bit_exact_arenas": int((got.view(np.uint32).reshape(n_par, -1) == ref.view(np.uint32).reshape(n_par, -1)).all(axis=1).sum()),
                        "max_rel_err": float(rel.max()), "compared": "body poses + velocities (1e-5 relative), controller states, cached-contact counts, tick counters",
                        "error_flag_arenas": int(err_arenas), "oracle_seconds": time.perf_counter() - t0,
                        "what": "arenas %d..%d of the timed batch (rank 0), free-running since the reset, against the CPU oracle on the same seeds" % (lo, lo + n_par - 1)}
        orc.close()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            g.build_oracle()
            cores = os.cpu_count() or 1
            v, ticks, dt, warm = cpu_port_run(cores, STEP_INTERVAL * args.warmup, STEP_INTERVAL * args.steps, 25.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d arenas (one per host thread) x 16 robots x 200 pucks, CacheCons, %d ticks after %d warm-up ticks (%.1f s)" % (cores, ticks, warm, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(A * world, "arenas sharded by index, %d per GPU, one process per GPU" % A),
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": roofline, "roofline_sense": roofline_sense, "parity_check": parity_check, "also": also, "cpu_baseline": cpu,
            "stats": {"mean_percent_completion": psd.mean_completion(total_stats), "arenas": int(total_stats.n_arenas),
                      "robots_carrying": int(total_stats.n_carrying), "cache_points": int(total_stats.n_caches),
                      "cached_contacts_per_arena": total_stats.n_contacts / max(1, total_stats.n_arenas)},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_check and parity_check["mismatches"]:
        raise SystemExit("bench.py: parity_check failed: %r" % (parity_check,))
    return 0


if __name__ == "__main__":
    sys.exit(main())
