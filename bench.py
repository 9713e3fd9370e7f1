#!/usr/bin/env python
"""bench.py — env-steps/s of the batched Memory-Gym step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--env ID] [--envs-per-gpu M] [--impl reference]

A "step" is one pass of the hot path over the whole env batch: one `mg_*_step` kernel launch per GPU
doing kinematics + state machine + in-kernel auto-reset + raster of N x 84x84x3 uint8 observations.
Default workload = BASELINE.json configs[1]: MortarMayhem-v0, 4096 envs per GPU, default options,
seeds 0..N-1, i.i.d. uniform random actions (pre-generated on the device), auto-reset on.

Printed JSON line (rank 0):
  value     whole-job million env-steps/s, actions resident in HBM, device-timed (CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks).
  e2e       same metric through the public Python API with HOST buffers: every step copies the
            actions pinned-host -> device and observations + rewards + dones device -> pinned host.
  roofline  the step kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU oracle (restated reference, PyGame is not installable here) on 1 host core.
`--impl reference` times the reference's CPU path (the oracle port) on all host cores instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "endless-memory-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "million env-steps/sec"
OBS_BYTES = 84 * 84 * 3
GRID_IDS = ("MortarMayhem-Grid-v0", "MortarMayhemB-Grid-v0", "MysteryPath-Grid-v0", "Endless-MysteryPath-v0")  # Discrete(4)
STATE_BYTES = {"mm": 256, "mp": 128, "ss": 1024, "ep": 2048}


def alg_bytes_per_env_step(family, n_act, steps_per_launch=1):
    """Algorithmic HBM bytes of one env-step: obs write + actions read + reward (f64) +
    done + the state record read and written once per launch (a rollout launch keeps it on chip)."""
    return OBS_BYTES + 8 * n_act + 8 + 1 + 2 * STATE_BYTES[family] / steps_per_launch


# ---- CPU arm: the oracle (restated reference) in the shape of the reference's bench.py loop ------------
def _cpu_worker(args):
    env_id, seed, budget_s = args
    import numpy as np
    from oracle import memgym_oracle as O
    env = O.make(env_id)
    rng = np.random.default_rng(seed)
    grid = env_id in GRID_IDS
    steps = 0
    t0 = time.perf_counter()
    env.reset(seed=seed)
    while time.perf_counter() - t0 < budget_s:  # reference bench.py:16-26: reset; random actions until done
        a = int(rng.integers(0, 4)) if grid else rng.integers(0, 3, 2)
        _, _, done, _, _ = env.step(a)
        steps += 1
        if done:
            env.reset()
    return steps, time.perf_counter() - t0


def cpu_throughput(env_id, cores, budget_s):
    """env-steps/s of the oracle on `cores` host processes, each running for ~budget_s seconds."""
    if cores == 1:
        res = [_cpu_worker((env_id, 0, budget_s))]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(env_id, i, budget_s) for i in range(cores)])
    return sum(s for s, _ in res) / max(t for _, t in res)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.times, self.proc, self.t_start = index, [], [], None, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            self.times.append(time.perf_counter())

    def mark_start(self):
        self.t_start = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return None
        t_end = time.perf_counter()
        self.proc.terminate()
        inside = [r for r, t in zip(self.rows, self.times) if self.t_start <= t <= t_end + 0.15]
        self.rows = inside if inside else self.rows[-3:]
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _ref_worker(args):
    """K+W bounded samples ("steps") of `chunk` env-steps each on one core; times the last K."""
    env_id, seed, warm, timed, chunk = args
    import numpy as np
    from oracle import memgym_oracle as O
    env = O.make(env_id)
    rng = np.random.default_rng(seed)
    grid = env_id in GRID_IDS
    env.reset(seed=seed)
    t0 = None
    for i in range(warm + timed):
        if i == warm:
            t0 = time.perf_counter()
        for _ in range(chunk):
            a = int(rng.integers(0, 4)) if grid else rng.integers(0, 3, 2)
            _, _, done, _, _ = env.step(a)
            if done:
                env.reset()
    return timed * chunk, time.perf_counter() - t0


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path, on all host cores.
    PyGame/gymnasium are not installable in this image (no wheel, no network), so the oracle port (the
    restated reference, pinned to trajectories of the unmodified reference) stands in for it. Each
    "step" is a bounded sample: `chunk` env-steps on every core, sized so the whole run takes ~2 min."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    probe = cpu_throughput(args.env, 1, 1.0)  # env-steps/s of one core
    chunk = max(1, int(probe * 100.0 / max(1, args.steps + args.warmup)))
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(args.env, i, args.warmup, args.steps, chunk) for i in range(cores)])
    secs = max(t for _, t in res)
    v = sum(s for s, _ in res) / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "M env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.env}, default options, random actions, reference bench.py loop shape "
                               f"(reset; step until done), {cores} independent single-env processes; "
                               f"one step = {chunk} env-steps per core"},
        "cpu_baseline": {"value": v, "unit": "M env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} samples x {chunk} env-steps x {cores} cores "
                                   "(oracle port; PyGame is not installable in this image)"},
        "e2e": {"value": v, "unit": "M env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100000)
    ap.add_argument("--warmup", type=int, default=256)
    ap.add_argument("--env", default="MortarMayhem-v0")
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--warps-per-cta", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=128, help="steps per mg_*_rollout launch (= slots of the obs buffer)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import memory_gym_b200 as mg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (memory_gym_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    spec = mg.spec(args.env)
    n = args.envs_per_gpu or (4096 if spec.family == "mm" and spec.variant != 2 else 8192)
    total = n * world
    env = mg.make_sharded(args.env, total, rank, world, device=dev, warps_per_cta=args.warps_per_cta)
    K, W = args.steps, args.warmup
    n_act = len(spec.action_nvec)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    T = K + W
    POOL = min(T, args.chunk)  # steps per rollout launch; i.i.d. action sets, cycled
    acts = torch.randint(0, spec.action_nvec[0], (POOL,) + tuple(env._actions_dev.shape), generator=g, device=dev)
    # obs go to a [chunk, N, 84, 84, 3] rollout buffer (step t of a launch -> slot t): every frame has its own
    # address, the buffer (chunk x 86.7 MB) is far larger than the 126 MB L2, so every frame reaches HBM
    ring_n = POOL
    ring = torch.empty((ring_n, n, 84, 84, 3), dtype=torch.uint8, device=dev)
    env.reset(seed=0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # rollout buffers: rewards / dones per step of one action-pool cycle (overwritten cycle after cycle)
    rew_buf = torch.empty((POOL, n), dtype=torch.float64, device=dev)
    done_buf = torch.empty((POOL, n), dtype=torch.bool, device=dev)

    def run_steps(count):
        """`count` env steps through mg_*_step_many in chunks of <= POOL (one launch per step)."""
        left = count
        while left > 0:
            c = min(left, POOL)
            env.rollout(acts[:c], ring[:c], rew_buf[:c], done_buf[:c])
            left -= c

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    run_steps(W)
    if sampler:  # nvidia-smi needs ~1 s to deliver its first sample: keep the GPU busy until it does
        t_spin = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_spin < 5.0:
            run_steps(POOL)
            torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_start()
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = total * K / (ms * 1e-3) / 1e6

    # ---- the same steps as ONE LAUNCH PER STEP (mg_*_step, programmatic dependent launch between them): what a
    # policy in the loop pays per step on the device, reported next to the rollout number
    Ks = min(K, 8 * POOL)
    for _ in range(2):
        env.step_many(acts, ring, rew_buf, done_buf)
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    es0.record()
    left = Ks
    while left > 0:
        c = min(left, POOL)
        env.step_many(acts[:c], ring[:c], rew_buf[:c], done_buf[:c])
        left -= c
    es1.record()
    barrier()
    ms_single = es0.elapsed_time(es1)
    if dist is not None:
        tm = torch.tensor([ms_single], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms_single = float(tm.item())

    # ---- end to end through the public API with host buffers -------------------------------------------
    Ke = args.e2e_steps or min(K, 200)
    Ke = min(Ke, POOL)
    a_host = acts[:Ke].cpu().pin_memory()
    obs_h = torch.empty((2, n, 84, 84, 3), dtype=torch.uint8).pin_memory()
    res_h = torch.empty((2, 9 * n), dtype=torch.uint8).pin_memory()  # reward f64[n] + done u8[n]: env.results
    rew_h = [res_h[b][:8 * n].view(torch.float64) for b in range(2)]
    a_dev = torch.empty_like(env._actions_dev)
    copy_stream = torch.cuda.Stream(device=dev)
    ev_step = [torch.cuda.Event() for _ in range(2)]
    ev_copy = [torch.cuda.Event() for _ in range(2)]
    obs_d = [torch.empty((n, 84, 84, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    res_d = [torch.empty(9 * n, dtype=torch.uint8, device=dev) for _ in range(2)]
    checksum = 0.0

    def e2e_loop(steps):
        nonlocal checksum
        for t in range(steps):
            b = t & 1
            a_dev.copy_(a_host[t % Ke], non_blocking=True)             # H2D actions (pinned)
            torch.cuda.current_stream().wait_event(ev_copy[b])         # slot b's previous D2H has drained
            env.step(a_dev, obs_out=obs_d[b])
            res_d[b].copy_(env.results)                                # reward + done of this step (one buffer)
            ev_step[b].record()
            with torch.cuda.stream(copy_stream):                       # D2H of step t overlaps step t+1
                copy_stream.wait_event(ev_step[b])
                obs_h[b].copy_(obs_d[b], non_blocking=True)
                res_h[b].copy_(res_d[b], non_blocking=True)
                ev_copy[b].record()
            if t >= 1:                                                 # consume step t-1's results on the host
                ev_copy[1 - b].synchronize()
                checksum += float(rew_h[1 - b][0]) + float(obs_h[1 - b][0, 42, 42, 0])
        torch.cuda.synchronize()

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(Ke)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tm = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e_value = total * Ke / e2e_s / 1e6

    # ---- the same loop for a consumer that lives on the GPU (the intended use: the policy reads obs in HBM): actions
    # still come from pinned host memory and reward + done still go back to the host every step; observations stay
    def e2e_device_obs(steps):
        nonlocal checksum
        for t in range(steps):
            b = t & 1
            a_dev.copy_(a_host[t % Ke], non_blocking=True)
            env.step(a_dev, obs_out=obs_d[b])
            res_h[b].copy_(env.results, non_blocking=True)
            ev_copy[b].record()
            if t >= 1:
                ev_copy[1 - b].synchronize()
                checksum += float(rew_h[1 - b][0])
        torch.cuda.synchronize()

    e2e_device_obs(3)
    barrier()
    t0 = time.perf_counter()
    e2e_device_obs(4 * Ke)
    barrier()
    e2e_dev_s = time.perf_counter() - t0
    if dist is not None:
        tm = torch.tensor([e2e_dev_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_dev_s = float(tm.item())
    e2e_dev_value = total * 4 * Ke / e2e_dev_s / 1e6

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        b_alg = alg_bytes_per_env_step(spec.family, n_act, POOL)
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        traffic = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(args.env)
            if cap and cap["envs"] == n and cap["steps_per_launch"] == POOL:
                traffic = cap["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        launches = -(-K // POOL)
        kernel_ms = ms / launches  # launches back to back on one stream: event time / launches = launch duration
        achieved = n * b_alg * K / (ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu and world == 1:  # the CPU baseline is reported by the single-GPU run only
            v = cpu_throughput(args.env, 1, args.cpu_seconds) / 1e6
            cpu = {"value": v, "unit": "M env-steps/s", "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_seconds:.0f} s of single-env stepping (reset; random actions until done) "
                             "with the CPU oracle = restated reference; PyGame is not installable in this image"}
        line = {
            "metric": METRIC, "value": value, "unit": "M env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"{args.env}, {n} envs/GPU x {world} GPU, 84x84x3 uint8 obs, default options, "
                                   "seeds 0..N-1, uniform random actions, auto-reset on",
                       "l2": f"each launch writes {ring_n} steps x {n * OBS_BYTES / 1e6:.1f} MB of obs to distinct slots "
                             f"of a {ring_n * n * OBS_BYTES / 1e9:.1f} GB rollout buffer (>> 126 MB L2), no address rewritten "
                             "within a launch",
                       "steps_per_launch": ring_n,
                       "parallelism": f"env shards x{world}, no collective on the step path",
                       "warps_per_cta": args.warps_per_cta or "default"},
            "e2e": {"value": e2e_value, "unit": "M env-steps/s", "h2d_bytes_per_step": n * n_act * 8,
                    "d2h_bytes_per_step": n * (OBS_BYTES + 8 + 1), "steps": Ke,
                    "note": "actions pinned-host->device and obs+reward+done device->pinned-host every step "
                            "(D2H of step t overlaps step t+1); PCIe-bound",
                    "device_obs": {"value": e2e_dev_value, "unit": "M env-steps/s", "h2d_bytes_per_step": n * n_act * 8,
                                   "d2h_bytes_per_step": n * (8 + 1),
                                   "note": "same public step() loop with the observations consumed in HBM (GPU-resident "
                                           "policy): actions from pinned host memory, reward+done back to the host"}},
            "single_step": {"value": total * Ks / (ms_single * 1e-3) / 1e6, "unit": "M env-steps/s",
                            "us_per_step": ms_single * 1e3 / Ks, "steps": Ks,
                            "note": "one mg_*_step launch per step (policy-in-the-loop shape), device-timed, max over ranks"},
            "gpu_launches": -(-K // POOL),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "mg_env_kernel",
                         "alg_bytes_per_env_step": b_alg, "kernel_ms": kernel_ms, "env_steps_per_launch": n * POOL},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
