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
    ap.add_argument("--no-extras", action="store_true", help="skip the stepInterval = 1 side measurement")
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
    cfg = abi.config_defaults(n_arenas=A, n_robots=N_ROBOTS, n_pucks=N_PUCKS, controller=abi.PS_CTRL_CACHECONS,
                              rng_mode=abi.PS_RNG_JAVA_SHARED)
    eng = engine.Engine(cfg, device=local_rank)
    eng.reset(np.arange(lo, lo + A))
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        eng.step(STEP_INTERVAL, sync=False)
    eng.sync()
    eng.profile(True)
    eng.kernel_ms()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        eng.step(STEP_INTERVAL, sync=False)
    eng.join()
    ev1.record(stream)
    eng.sync()
    local_stats = eng.compute_stats()
    total_stats = psd.allreduce_stats(local_stats, device=torch.device("cuda", local_rank)) if world > 1 else local_stats
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - launches0
    kms, kcnt = eng.kernel_ms()
    eng.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    robot_steps = world * A * N_ROBOTS * STEP_INTERVAL * args.steps
    value = robot_steps / (ms_max / 1000.0)

    # ---- end to end through the C-ABI with host buffers ("e2e") ----
    # Every step uploads the bodies' poses / velocities from pinned memory, steps, and downloads them with the step counters.
    # PS_E2E_HANDLES > 1 splits the arenas over several handles served round-robin, so that one handle's copies overlap the
    # others' stepping; measured in the steady state the copies (2 x 21 MB per step) cost less than what kernels of several
    # handles running side by side lose: 5.64 / 5.37 / 5.22 M robot-steps/s with 1 / 2 / 4 handles. One handle is the default.
    sense_dbg = eng.sense_stats().reshape(-1, 8).astype(np.float64)
    n_groups = eng.num_groups
    import ctypes as C
    halves = []
    old_groups = os.environ.get("PS_GROUPS")
    n_handles = int(os.environ.get("PS_E2E_HANDLES", "1"))
    os.environ["PS_GROUPS"] = str(max(1, n_groups // n_handles))
    h2d = d2h = 0
    for hi in range(n_handles):
        a0, a1 = lo + hi * A // n_handles, lo + (hi + 1) * A // n_handles
        c2 = abi.config_defaults(n_arenas=a1 - a0, n_robots=N_ROBOTS, n_pucks=N_PUCKS, controller=abi.PS_CTRL_CACHECONS,
                                 rng_mode=abi.PS_RNG_JAVA_SHARED)
        e2 = engine.Engine(c2, device=local_rank)
        e2.reset(np.arange(a0, a1))
        st = e2.new_state()
        pinned_body = torch.empty(st.body.shape, dtype=torch.float32).pin_memory()
        pinned_arena = torch.empty(st.arena.nbytes, dtype=torch.uint8).pin_memory()
        body, arena = pinned_body.numpy(), pinned_arena.numpy().view(abi.ARENA_DTYPE)
        vin, vout = abi.ps_state_view(), abi.ps_state_view()
        vin.body = body.ctypes.data_as(C.POINTER(C.c_float))
        vout.body = body.ctypes.data_as(C.POINTER(C.c_float))
        vout.arena = arena.ctypes.data_as(C.POINTER(abi.ps_arena_state))
        e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))
        h2d += body.nbytes
        d2h += body.nbytes + arena.nbytes
        halves.append((e2, vin, vout, pinned_body, pinned_arena))
    if old_groups is None:
        os.environ.pop("PS_GROUPS", None)
    else:
        os.environ["PS_GROUPS"] = old_groups

    def e2e_round(first):
        for e2, vin, vout, _, _ in halves:
            if not first:
                e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))   # D2H: poses + velocities + step counters (synchronous)
            e2._check(e2.L.ps_set_state(e2.h, C.byref(vin)))          # H2D: poses + velocities of every body
            e2._check(e2.L.ps_step(e2.h, STEP_INTERVAL))

    def e2e_drain():
        for e2, vin, vout, _, _ in halves:
            e2._check(e2.L.ps_get_state(e2.h, C.byref(vout)))

    # the same simulated interval as the device-resident run: W warm-up steps after the reset, then K timed steps
    e2e_round(True)
    for _ in range(args.warmup - 1):
        e2e_round(False)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    e2e_round(True)
    for _ in range(args.steps - 1):
        e2e_round(False)
    e2e_drain()                       # the last step's results are on the host when the clock stops
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = robot_steps / (float(t.item()) / 1000.0)
    for e2, *_ in halves:
        e2.close()

    # ---- the same batch with the sensors on every tick (stepInterval = 1 beside the reference's 3) ----
    also = {}
    if not args.no_extras:
        cfg1 = abi.config_defaults(n_arenas=A, n_robots=N_ROBOTS, n_pucks=N_PUCKS, controller=abi.PS_CTRL_CACHECONS,
                                   rng_mode=abi.PS_RNG_JAVA_SHARED, step_interval=1)
        eng1 = engine.Engine(cfg1, device=local_rank)
        eng1.reset(np.arange(lo, lo + A))
        s1 = torch.cuda.ExternalStream(eng1.stream, device=local_rank)
        for _ in range(9):
            eng1.step(1, sync=False)
        eng1.sync()
        n1 = 15
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        for _ in range(n1):
            eng1.step(1, sync=False)
        eng1.join()
        e1.record(s1)
        eng1.sync()
        t1 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        also["step_interval_1"] = {"value": world * A * N_ROBOTS * n1 / (float(t1.item()) / 1000.0), "unit": UNIT,
                                   "ms_per_tick": float(t1.item()) / n1, "ticks": n1, "warmup_ticks": 9}
        eng1.close()

    # ---- roofline of the dominant kernel (k_physics: World.step) ----
    peak, peak_src = measured_peaks()
    c_hat = total_stats.n_touching_points / max(1, total_stats.n_arenas)   # live touching manifold points per arena
    bytes_per_arena_tick = 2.0 * (24.0 * (N_ROBOTS + N_PUCKS) + 24.0 * c_hat + 96.0 * N_ROBOTS)
    # one World.step launch sequence of one arena group = k_phys_pre + island kernels + k_phys_post, each timed by its own
    # event pair on the group's stream (the turn-taking wait between the island phase and k_phys_post is outside all three)
    ticks_per_launch = STEP_INTERVAL
    phys_ms = ticks_per_launch * sum(kms[k] / max(1, kcnt[k]) for k in ("phys_pre", "phys_islands", "phys_post"))
    arenas_per_launch = A / max(1, eng.num_groups)   # every arena group launches its own kernel sequence
    alg_bytes_per_launch = bytes_per_arena_tick * arenas_per_launch * ticks_per_launch
    achieved = alg_bytes_per_launch / (phys_ms / 1000.0) / 1e9 if phys_ms > 0 else 0.0
    top = ("sense", "think", "physics", "other")
    total_k = sum(kms[k] for k in top)
    roofline = {
        "kernel": "World.step pipeline (k_phys_pre + k_island_big | k_island_warp | k_island_batch + k_phys_post) x %d ticks" % ticks_per_launch,
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": None,
        "algorithmic_bytes_per_arena_tick": bytes_per_arena_tick, "touching_points_per_arena": c_hat,
        "kernel_ms_per_launch": phys_ms, "arenas_per_launch": arenas_per_launch, "arena_groups": eng.num_groups,
        "kernel_share_of_step": {k: (kms[k] / total_k if total_k > 0 else 0.0) for k in top},
        "world_step_stage_ms_per_tick": {k: kms[k] / max(1, kcnt[k]) for k in ("phys_pre", "phys_islands", "phys_post")},
        "kernel_ms_total": kms, "kernel_launch_counts": kcnt,
    }
    # ---- shared-memory roofline of the camera stage (algorithmic smem bytes per robot-sense) ----
    p_act = 4556.0
    k_bar = float(sense_dbg[:, 4].mean()) / p_act            # candidate bodies examined per pixel (measured)
    smem_bytes_per_sense = p_act * 9.0 + p_act * k_bar * 12.0 + 3825.0 + 2294.0 * 3.0 + 19040.0
    sense_ms = kms["sense"] / max(1, kcnt["sense"])             # k_sense + k_localmap of one arena group
    robots_per_launch = arenas_per_launch * N_ROBOTS
    smem_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9   # GB/s: 148 SMs x 128 B/clk x SM clock
    smem_achieved = smem_bytes_per_sense * robots_per_launch / (sense_ms / 1000.0) / 1e9 if sense_ms > 0 else 0.0
    roofline_sense = {
        "kernel": "k_sense + k_localmap (camera, occupancy, blobs, local map) per arena group", "bound": "smem",
        "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s", "frac": smem_achieved / smem_peak,
        "peak_source": "148 SMs x 128 B/clk x sampled SM clock", "candidates_per_pixel": k_bar,
        "algorithmic_bytes_per_robot_sense": smem_bytes_per_sense, "kernel_ms_per_launch": sense_ms, "robots_per_launch": robots_per_launch,
        "note": "candidate binning cuts K from the ~12 the survey assumed to the measured value, so the algorithmic byte count (and this fraction) "
                "shrinks with it; the stage is issue/latency bound (profiles/)",
    }
    ncu_traffic = os.path.join(ROOT, "profiles", "k_physics_traffic.json")
    if os.path.exists(ncu_traffic):
        try:
            # per launch like `achieved`: the ncu capture's DRAM bytes per arena-tick x the arenas and ticks of one launch sequence
            per_arena_tick = json.load(open(ncu_traffic)).get("dram_bytes_per_arena_tick")
            roofline["traffic"] = per_arena_tick * arenas_per_launch * ticks_per_launch
            roofline["traffic_bytes_per_arena_tick"] = per_arena_tick
        except Exception:
            pass

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
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": "per step and per handle: ps_set_state(body) + ps_step(3) + ps_get_state(body, arena), pinned host buffers; "
                            "%d handle(s) served round-robin; host wall clock" % n_handles},
            "roofline": roofline, "roofline_sense": roofline_sense, "also": also, "cpu_baseline": cpu,
            "stats": {"mean_percent_completion": psd.mean_completion(total_stats), "arenas": int(total_stats.n_arenas),
                      "robots_carrying": int(total_stats.n_carrying), "cache_points": int(total_stats.n_caches),
                      "cached_contacts_per_arena": total_stats.n_contacts / max(1, total_stats.n_arenas)},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
