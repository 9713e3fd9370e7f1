#!/usr/bin/env python
"""bench.py — env-steps/s of the batched Memory-Gym step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--env ID] [--envs-per-gpu M] [--impl reference]

A "step" is one pass of the hot path over the whole env batch: kinematics + state machine + in-kernel
auto-reset + raster of N x 84x84x3 uint8 observations. Headline workload = BASELINE.json configs[1]:
MortarMayhem-v0, 4096 envs per GPU, default options, seeds 0..N-1, i.i.d. uniform random actions
(pre-generated on the device: a pool of up to 8 distinct K-step blocks, cycled, small enough to stay in L2 like
actions a policy has just written), auto-reset on. The timed launch is one `mg_*_rollout` of K steps (the
reference's bench.py loop shape: pre-sampled random actions, reset when done); the timed region is R such
launches back to back (R sized for >= 200 ms, `config.timed_launches`), entered behind launches that are
already running so that no host latency sits between the two events. Every rank warms up the same way.

Printed JSON line (rank 0):
  value     whole-job million env-steps/s, actions resident in HBM, device-timed (CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks; `per_rank_ms` lists all).
  e2e       same metric through the public Python API with HOST buffers: every step copies the
            actions pinned-host -> device and observations + rewards + dones device -> pinned host.
  roofline  the rollout kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  configs   the other BASELINE.json configs (MysteryPath-v0 8192, SearingSpotlights-v0 8192,
            Endless-MortarMayhem-v0 8192 per GPU = 65 536 on 8 GPUs), each timed the same way with its own
            roofline fraction and single-step figure.
  cpu_baseline  the CPU oracle (restated reference, PyGame is not installable here) on 1 host core; its
            `config1` entry is BASELINE configs[0]: MortarMayhem-Grid-v0, 1 env, the reference bench.py loop.
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
STATE_BYTES = {"mm": 256, "mp": 128, "ss": 1152, "ep": 2048}  # (ep: the record proper; its look-ahead block is not streamed)


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


def gather_ms(dist, dev, ms):
    """(max over ranks, list of every rank's value) of a device-timed duration."""
    import torch
    if dist is None:
        return ms, [ms]
    mine = torch.tensor([ms], dtype=torch.float64, device=dev)
    every = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    vals = [float(t.item()) for t in every]
    return max(vals), vals


def load_traffic():
    """DRAM bytes per env-step of the rollout kernel from the committed `ncu --set full` captures."""
    out = {}
    for name in ("r01_traffic.json", "r02_traffic.json"):
        try:
            for env_id, cap in json.load(open(os.path.join(ROOT, "profiles", name))).items():
                out[env_id] = dict(cap, file=name)
        except (OSError, ValueError):
            pass
    return out


class Timed:
    """One BASELINE config on this rank's GPU: env batch, rollout buffers, and the three device-timed shapes."""

    MIN_REGION_MS = 200.0   # timed region (>= 2 nvidia-smi samples at 100 ms)
    MIN_WARM_S = 0.5        # warm-up: at least this long AND at least two passes over every ring slot, on EVERY rank

    def __init__(self, mg, torch, dist, env_id, n, rank, world, dev, K, chunk, wpc=0):
        self.torch, self.dist, self.dev, self.env_id, self.n, self.world, self.rank = torch, dist, dev, env_id, n, world, rank
        self.spec = mg.spec(env_id)
        self.total = n * world
        self.env = mg.make_sharded(env_id, self.total, rank, world, device=dev, warps_per_cta=wpc)
        self.K = K
        self.L = min(K, chunk)                     # steps per rollout launch
        self.launches_per_block = -(-K // self.L)  # launches that make up K steps
        self.n_act = len(self.spec.action_nvec)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        # uniform random actions: a pool of distinct action blocks, cycled (ONE 20-step block replayed launch after launch
        # makes every env repeat the same 20 actions for ever)
        self.pool = torch.randint(0, self.spec.action_nvec[0], (self.n_blocks(),) + (self.L,) + tuple(self.env._actions_dev.shape),
                                  generator=g, device=dev)
        self.acts = self.pool[0]
        self.blk = 0
        # obs go to a [L, N, 84, 84, 3] rollout buffer (step t of a launch -> slot t): every frame of a launch has
        # its own address and the buffer is far larger than the 126 MB L2, so every frame reaches HBM
        self.ring = torch.empty((self.L, n, 84, 84, 3), dtype=torch.uint8, device=dev)
        self.rew = torch.empty((self.L, n), dtype=torch.float64, device=dev)
        self.done = torch.empty((self.L, n), dtype=torch.bool, device=dev)
        self.env.reset(seed=0)

    ACTION_BLOCKS = 0  # --action-blocks: 0 = up to 8 distinct blocks within ACTION_POOL_BYTES
    ACTION_POOL_BYTES = 12 << 20  # the pool stays L2-resident, like actions a policy has just written (a 21 MB pool is
                                  # evicted by the 1.7 GB of frames of every launch: -3.5 % for MortarMayhem, measured)

    def n_blocks(self):
        block_bytes = self.L * self.n * self.n_act * 8
        return Timed.ACTION_BLOCKS or max(1, min(8, Timed.ACTION_POOL_BYTES // block_bytes))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def block(self, many=False):
        """K consecutive env steps: rollout launches of L steps (or, `many`, one step launch per step)."""
        fn = self.env.step_many if many else self.env.rollout
        left = self.K
        while left > 0:
            c = min(left, self.L)
            fn(self.pool[self.blk % self.pool.shape[0]][:c], self.ring[:c], self.rew[:c], self.done[:c])
            self.blk += 1
            left -= c

    def measure(self, many=False, on_start=None):
        """R blocks of K steps between two CUDA events. Warm-up (same on every rank): >= 2 blocks and >= 0.5 s. The
        start event is recorded behind two blocks that are still running, so the GPU never waits for the host inside
        the timed region. Returns (ms max over ranks, per-rank ms, R)."""
        torch = self.torch
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or time.perf_counter() - t0 < self.MIN_WARM_S:
            self.block(many)
            reps += 1
            if reps % 8 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        # size R from a short probe (device-timed, 4 blocks)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            self.block(many)
        p1.record()
        torch.cuda.synchronize()
        probe_ms = p0.elapsed_time(p1) / 4
        R = max(3, int(self.MIN_REGION_MS / max(probe_ms, 1e-3)) + 1)
        if self.dist is not None:  # the same R everywhere
            r_t = torch.tensor([R], dtype=torch.int64, device=self.dev)
            self.dist.all_reduce(r_t, op=self.dist.ReduceOp.MAX)
            R = int(r_t.item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if on_start:
            on_start()
        self.block(many)
        self.block(many)  # (still running when the timed launches are enqueued)
        e0.record()
        for _ in range(R):
            self.block(many)
        e1.record()
        self.barrier()
        ms, per_rank = gather_ms(self.dist, self.dev, e0.elapsed_time(e1))
        return ms, per_rank, R

    def lookahead_counters(self):
        """Endless-MysteryPath: the statistics words of the look-ahead block behind every record (MgEPSpecRec): resets
        that adopted a path generated ahead, resets that found none ready / a stale one. None for other env types."""
        if self.spec.family != "ep":
            return None
        w = self.env.state[:, 2048:2112].clone().view(self.torch.int32)
        return [int(w[:, 2].sum()), int(w[:, 14].sum()), int(w[:, 15].sum())]

    @staticmethod
    def lookahead_delta(a, b):
        if a is None:
            return None
        adopted, empty, stale = (y - x for x, y in zip(a, b))
        tot = max(1, adopted + empty + stale)
        return {"resets": adopted + empty + stale, "adopted": adopted, "adopted_frac": adopted / tot, "missed_empty": empty,
                "missed_stale": stale}

    def b_alg(self, steps_per_launch):
        return alg_bytes_per_env_step(self.spec.family, self.n_act, steps_per_launch)

    def report(self, peak, traffic_tab, on_start=None):
        """rollout + single-step figures of this config as the dicts the JSON line carries."""
        c0 = self.lookahead_counters()
        ms, per_rank, R = self.measure(False, on_start)
        c1 = self.lookahead_counters()
        done_rate = float(self.done.float().mean())  # auto-resets per env-step in the last timed launch
        ms_s, per_rank_s, R_s = self.measure(True)
        c2 = self.lookahead_counters()
        K, n = self.K, self.n
        steps = K * R
        b = self.b_alg(self.L)
        achieved = n * b * steps / (ms * 1e-3) / 1e9
        cap = traffic_tab.get(self.env_id)
        traffic = traffic_src = None
        if cap and "dram_bytes_per_launch" in cap:
            per_env_step = cap["dram_bytes_per_launch"] / (cap["envs"] * cap["steps_per_launch"])
            traffic = per_env_step * n * self.L
            traffic_src = (f"profiles/{cap['file']}: ncu --set full capture at {cap['envs']} envs x {cap['steps_per_launch']} "
                           f"steps per launch = {per_env_step:.1f} B per env-step, scaled to this launch")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak[0], "unit": "GB/s", "frac": achieved / peak[0],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak[1],
                "kernel": "mg_env_kernel<%sPolicy, true>" % {"mm": "MM", "mp": "MP", "ss": "SS", "ep": "EP"}[self.spec.family],
                "alg_bytes_per_env_step": b, "alg_bytes_per_launch": b * n * self.L,
                "kernel_ms": ms / (R * self.launches_per_block), "steps_per_launch": self.L,
                "env_steps_per_launch": n * self.L}
        single = {"value": self.total * K * R_s / (ms_s * 1e-3) / 1e6, "unit": "M env-steps/s",
                  "us_per_step": ms_s * 1e3 / (K * R_s), "steps": K * R_s,
                  "frac": n * self.b_alg(1) * K * R_s / (ms_s * 1e-3) / 1e9 / peak[0],
                  "note": "one mg_*_step launch per step (policy-in-the-loop shape, programmatic dependent launch between "
                          "them), device-timed, max over ranks"}
        long_run = self.long_launches(peak)
        return {"value": self.total * steps / (ms * 1e-3) / 1e6, "ms": ms, "ms_per_step": ms / steps, "per_rank_ms": per_rank,
                "timed_launches": R * self.launches_per_block, "timed_steps": steps, "roofline": roof, "single_step": single,
                "long_launches": long_run, "resets_per_env_step": done_rate,
                "lookahead": None if c0 is None else {"rollout": self.lookahead_delta(c0, c1), "single_step": self.lookahead_delta(c1, c2),
                                                      "note": "Endless-MysteryPath: auto-resets (warm-up included) that adopted a "
                                                              "path generated ahead by the helper warps"}}

    def long_launches(self, peak, steps_per_launch=128):
        """The same rollout with `steps_per_launch` steps per launch (the shape round 1 reported): what the kernel sustains
        when launch boundaries are rare (Searing Spotlights rebuilds its base buffers, Endless-MysteryPath loses its
        speculative paths at every launch boundary)."""
        torch = self.torch
        if self.K >= steps_per_launch:
            return None
        K0, L0, lpb0 = self.K, self.L, self.launches_per_block
        keep = (self.acts, self.ring, self.rew, self.done, self.pool)
        try:
            self.K = self.L = steps_per_launch
            self.launches_per_block = 1
            g = torch.Generator(device=self.dev).manual_seed(4321 + self.rank)
            self.pool = torch.randint(0, self.spec.action_nvec[0], (self.n_blocks(),) + (self.L,) + tuple(self.env._actions_dev.shape),
                                      generator=g, device=self.dev)
            self.acts = self.pool[0]
            self.ring = torch.empty((self.L, self.n, 84, 84, 3), dtype=torch.uint8, device=self.dev)
            self.rew = torch.empty((self.L, self.n), dtype=torch.float64, device=self.dev)
            self.done = torch.empty((self.L, self.n), dtype=torch.bool, device=self.dev)
            ms, per_rank, R = self.measure(False)
            steps = self.K * R
            b = self.b_alg(self.L)
            achieved = self.n * b * steps / (ms * 1e-3) / 1e9
            return {"steps_per_launch": self.L, "timed_launches": R, "value": self.total * steps / (ms * 1e-3) / 1e6,
                    "unit": "M env-steps/s", "frac": achieved / peak[0], "per_rank_ms": per_rank}
        finally:
            self.K, self.L, self.launches_per_block = K0, L0, lpb0
            self.acts, self.ring, self.rew, self.done, self.pool = keep


def shard_check(mg, torch, dist, dev, rank, world, env_id="MortarMayhem-v0", n_chk=256, T=64, sample=32):
    """Byte equality across shard boundaries ON THE HARDWARE: every rank rolls its shard of a small global batch
    (global actions from one fixed seed) and hashes a sample of its envs; the next rank recomputes exactly those envs on
    ITS GPU from a differently-cut shard (`env_offset`), and the hashes must agree. With one GPU the same check runs
    between two shard cuts on that GPU."""
    spec = mg.spec(env_id)
    total = n_chk * max(world, 2)
    g = torch.Generator(device="cpu").manual_seed(4321)
    acts_all = torch.randint(0, spec.action_nvec[0], (T, total, len(spec.action_nvec)), generator=g).squeeze(-1)

    def run(offset, count):
        env = mg.make(env_id, num_envs=count, device=dev, env_offset=offset)
        env.reset(seed=77)
        a = acts_all[:, offset:offset + count].contiguous().to(dev)
        ring = torch.zeros((T, count, 84, 84, 3), dtype=torch.uint8, device=dev)
        rew = torch.zeros((T, count), dtype=torch.float64, device=dev)
        dn = torch.zeros((T, count), dtype=torch.uint8, device=dev)
        env.rollout(a, ring, rew, dn)
        torch.cuda.synchronize()
        return ring, rew, dn

    def digest(ring, rew, dn, lo, hi):
        parts = []
        for x in (ring[:, lo:hi], rew[:, lo:hi].contiguous().view(torch.int64), dn[:, lo:hi]):
            v = x.reshape(-1).to(torch.int64)
            w = torch.arange(v.numel(), device=dev, dtype=torch.int64) * 2654435761 + 1
            parts.append(int((v * w).sum().item()))  # wraps mod 2^64: a position-weighted checksum
        return parts

    # my shard [rank*n_chk, +n_chk): hash its first `sample` envs
    mine = digest(*run(rank * n_chk, n_chk), 0, sample)
    # the previous rank's sample, recomputed here from a shard that starts in the MIDDLE of its range
    prev = (rank - 1) % max(world, 2)
    redo = digest(*run(prev * n_chk, sample + 7), 0, sample)
    if dist is None:
        other = digest(*run(prev * n_chk, n_chk), 0, sample)
        return "ok" if other == redo and mine != redo else "mismatch"
    t_mine = torch.tensor(mine, dtype=torch.int64, device=dev)
    every = [torch.zeros_like(t_mine) for _ in range(world)]
    dist.all_gather(every, t_mine)
    ok = [int(v) for v in every[prev].tolist()] == redo
    flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=dev)
    dist.all_reduce(flag)
    return "ok" if int(flag.item()) == 0 else "mismatch"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="K: steps per timed block (one rollout launch when K <= --chunk)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--env", default="MortarMayhem-v0")
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--warps-per-cta", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=128, help="max steps per mg_*_rollout launch (= slots of the obs buffer)")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--action-blocks", type=int, default=0, help="distinct blocks of random actions the launches cycle through "
                    "(0: up to 8 blocks within 12 MB; 1: replay one block, the round-1 behaviour)")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the other BASELINE configs)")
    args = ap.parse_args()
    Timed.ACTION_BLOCKS = max(0, args.action_blocks)
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
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()  # (nvidia-smi needs ~1 s for its first sample: started before the env batch is built)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def default_n(env_id):
        sp = mg.spec(env_id)
        return 4096 if sp.family == "mm" and sp.variant != 2 else 8192

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    traffic_tab = load_traffic()

    K, W = args.steps, args.warmup
    n = args.envs_per_gpu or default_n(args.env)
    total = n * world
    main_cfg = Timed(mg, torch, dist, args.env, n, rank, world, dev, K, args.chunk, args.warps_per_cta)
    head = main_cfg.report(peak, traffic_tab, on_start=(sampler.mark_start if sampler else None))
    clocks = sampler.stop() if sampler else None
    env, spec, acts, n_act = main_cfg.env, main_cfg.spec, main_cfg.acts, main_cfg.n_act
    barrier = main_cfg.barrier
    POOL = main_cfg.L

    # ---- end to end through the public API with host buffers -------------------------------------------
    Ke = args.e2e_steps or min(max(K, 60), 200)
    a_host = acts.cpu().pin_memory()
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
            a_dev.copy_(a_host[t % POOL], non_blocking=True)           # H2D actions (pinned)
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

    def wall(fn, steps):
        fn(3)
        barrier()
        t0 = time.perf_counter()
        fn(steps)
        barrier()
        secs = time.perf_counter() - t0
        if dist is not None:
            tm = torch.tensor([secs], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            secs = float(tm.item())
        return total * steps / secs / 1e6

    e2e_raw_value = wall(e2e_loop, Ke)

    # ---- ... and through env.host_pipe() (HostPipe), the public host-buffer API: the same three arrays land in host
    # memory every step, but the observations cross the link delta-coded (mg_obs_delta_encode: only the 16-byte chunks
    # that changed, written by the GPU straight into pinned host memory) and library worker threads patch them into the
    # persistent host array (mg_obs_delta_decode); actions are read, rewards / dones written zero-copy. Two steps in
    # flight (the reference bench.py shape: pre-sampled actions), like the raw loop above.
    # "direct" (default): the host arrays are pinned and the GPU writes every changed chunk to its place in them — the
    # host only waits for the step's event. "patch" (round 2, first half): chunks packed into a pinned stage, patched
    # into a pageable host array by the library's worker threads. Both are timed; `e2e.value` is the default mode.
    a_pool = [a_host[i] for i in range(POOL)]
    pipe_results = {}
    for pipe_mode in ("patch", "direct"):
        pipe = env.host_pipe(depth=2, mode=pipe_mode)

        def e2e_pipe(steps):
            nonlocal checksum
            pipe.submit(a_pool[0])
            for t in range(1, steps):
                pipe.submit(a_pool[t % POOL])
                o, r, d, _, _ = pipe.collect()
                checksum += float(r[0]) + float(o[0, 42, 42, 0])
            pipe.collect()

        pipe_bytes0 = pipe.total_d2h_bytes
        value = wall(e2e_pipe, 10 * Ke)
        pipe_steps = 10 * Ke + 3
        d2h = (pipe.total_d2h_bytes - pipe_bytes0) / pipe_steps
        # the host array must be the device frame, bit for bit (last collected step = env.obs: nothing is in flight now)
        torch.cuda.synchronize()
        pipe_results[pipe_mode] = (value, d2h, bool(torch.equal(pipe.obs, env.obs.cpu())))
        del pipe
    auto_mode = env.host_pipe(depth=1).mode   # what the public default picks on this box (host cores per rank)
    e2e_value, pipe_d2h, _ = pipe_results[auto_mode]
    pipe_ok = pipe_results["direct"][2] and pipe_results["patch"][2]

    # ---- the same loop for a consumer that lives on the GPU (the intended use: the policy reads obs in HBM): actions
    # still come from pinned host memory and reward + done still go back to the host every step; observations stay
    def e2e_device_obs(steps):
        nonlocal checksum
        for t in range(steps):
            b = t & 1
            a_dev.copy_(a_host[t % POOL], non_blocking=True)
            env.step(a_dev, obs_out=obs_d[b])
            res_h[b].copy_(env.results, non_blocking=True)
            ev_copy[b].record()
            if t >= 1:
                ev_copy[1 - b].synchronize()
                checksum += float(rew_h[1 - b][0])
        torch.cuda.synchronize()

    e2e_dev_value = wall(e2e_device_obs, 20 * Ke)

    # ---- ... and with the whole host-driven step captured as ONE CUDA graph (env.capture_step(): actions H2D, the
    # step kernel, reward + done D2H): one graph launch per step instead of three stream operations + a library call
    graphs = [env.capture_step(obs_out=obs_d[b]) for b in range(2)]
    rew_g = [gr.results_host[:8 * n].view(torch.float64) for gr in graphs]

    def e2e_device_obs_graph(steps):
        nonlocal checksum
        for t in range(steps):
            b = t & 1
            graphs[b].actions_host.copy_(a_host[t % POOL])             # the policy's actions land in pinned memory
            graphs[b].replay()
            if t >= 1:
                graphs[1 - b].done_event.synchronize()
                checksum += float(rew_g[1 - b][0])
        torch.cuda.synchronize()

    e2e_graph_value = wall(e2e_device_obs_graph, 20 * Ke)

    # ---- ... and with NO copy operation at all: the step kernel reads the actions straight from pinned host memory and
    # writes reward + done straight into pinned host memory (zero-copy over PCIe), so consecutive step launches run back
    # to back under programmatic dependent launch; the host waits on an event recorded behind each step
    ev_z = [torch.cuda.Event() for _ in range(2)]

    def e2e_device_obs_zero_copy(steps):
        nonlocal checksum
        for t in range(steps):
            b = t & 1
            env.step(a_host[t % POOL], obs_out=obs_d[b], results_out=res_h[b])
            ev_z[b].record()
            if t >= 1:
                ev_z[1 - b].synchronize()
                checksum += float(rew_h[1 - b][0])
        torch.cuda.synchronize()

    e2e_zero_value = wall(e2e_device_obs_zero_copy, 20 * Ke)

    # ---- the other BASELINE.json configs, each timed like the headline ------------------------------------
    configs = []
    if not args.no_configs:
        del main_cfg.ring
        # (the three BASELINE configs, then the two endless env types BASELINE does not list: same shape)
        for env_id in ("MysteryPath-v0", "SearingSpotlights-v0", "Endless-MortarMayhem-v0", "Endless-MysteryPath-v0",
                       "Endless-SearingSpotlights-v0"):
            if env_id == args.env:
                continue
            c = Timed(mg, torch, dist, env_id, default_n(env_id), rank, world, dev, K, args.chunk)
            r = c.report(peak, traffic_tab)
            configs.append({"workload": f"{env_id}, {c.n} envs/GPU x {world} GPU = {c.total} envs, default options, "
                                        "seeds 0..N-1, uniform random actions, auto-reset on",
                            "value": r["value"], "unit": "M env-steps/s", "ms_per_step": r["ms_per_step"],
                            "timed_launches": r["timed_launches"], "per_rank_ms": r["per_rank_ms"],
                            "roofline": r["roofline"], "single_step": r["single_step"], "long_launches": r["long_launches"],
                            "resets_per_env_step": r["resets_per_env_step"],
                            **({"lookahead": r["lookahead"]} if r["lookahead"] else {})})
            del c
            torch.cuda.empty_cache()

    check = shard_check(mg, torch, dist, dev, rank, world)

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:  # the CPU baseline is reported by the single-GPU run only
            v = cpu_throughput(args.env, 1, args.cpu_seconds) / 1e6
            v1 = cpu_throughput("MortarMayhem-Grid-v0", 1, min(args.cpu_seconds, 5.0)) / 1e6
            cpu = {"value": v, "unit": "M env-steps/s", "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_seconds:.0f} s of single-env stepping (reset; random actions until done) "
                             "with the CPU oracle = restated reference; PyGame is not installable in this image",
                   "config1": {"workload": "MortarMayhem-Grid-v0, 1 env, CPU, reference bench.py loop shape (BASELINE "
                                           "configs[0]; no GPU)", "value": v1, "unit": "M env-steps/s", "cores": 1,
                               "kind": "port"}}
        line = {
            "metric": METRIC, "value": head["value"], "unit": "M env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.env}, {n} envs/GPU x {world} GPU, 84x84x3 uint8 obs, default options, "
                                   "seeds 0..N-1, uniform random actions, auto-reset on",
                       "l2": f"each launch writes {POOL} steps x {n * OBS_BYTES / 1e6:.1f} MB of obs to distinct slots "
                             f"of a {POOL * n * OBS_BYTES / 1e9:.1f} GB rollout buffer (>> 126 MB L2), no address rewritten "
                             "within a launch",
                       "steps_per_launch": POOL, "timed_launches": head["timed_launches"],
                       "timed_steps": head["timed_steps"],
                       "timing": "CUDA events around R back-to-back K-step rollout launches (R sized for >= 200 ms), the "
                                 "start event enqueued behind 2 running launches; every rank warms up >= 0.5 s and >= 2 "
                                 "passes over the rollout buffer; ms_per_step = ms / (K * R); max over ranks",
                       "parallelism": f"env shards x{world}, no collective on the step path",
                       "warps_per_cta": args.warps_per_cta or "default"},
            "per_rank_ms": head["per_rank_ms"],
            "e2e": {"value": e2e_value, "unit": "M env-steps/s", "h2d_bytes_per_step": n * n_act * 8,
                    "d2h_bytes_per_step": pipe_d2h, "steps": 10 * Ke, "mode": auto_mode,
                    "host_obs_equals_device": pipe_ok,
                    "note": "env.host_pipe() (default mode 'auto' = '%s' on this box): every step the actions come from pinned "
                            "host memory and obs uint8[N,84,84,3] + reward + done land in HOST arrays; observations cross PCIe "
                            "as a lossless delta (only the 16-byte chunks that changed; d2h_bytes_per_step = measured "
                            "average); 2 steps in flight" % auto_mode,
                    "direct_pipe": {"value": pipe_results["direct"][0], "unit": "M env-steps/s",
                                    "d2h_bytes_per_step": pipe_results["direct"][1],
                                    "note": "mode='direct': the GPU writes every changed chunk to its place in pinned host arrays "
                                            "(one per step in flight); no host work"},
                    "patch_pipe": {"value": pipe_results["patch"][0], "unit": "M env-steps/s",
                                   "d2h_bytes_per_step": pipe_results["patch"][1],
                                   "note": "mode='patch': chunks packed into a pinned stage, patched into ONE pageable host "
                                           "array by the library's worker threads"},
                    "raw_copy": {"value": e2e_raw_value, "unit": "M env-steps/s", "h2d_bytes_per_step": n * n_act * 8,
                                 "d2h_bytes_per_step": n * (OBS_BYTES + 8 + 1), "steps": Ke,
                                 "note": "the same arrays by plain copies: actions pinned-host->device and whole frames + "
                                         "reward + done device->pinned-host every step (D2H of step t overlaps step t+1); "
                                         "PCIe-bound"},
                    "device_obs": {"value": e2e_dev_value, "unit": "M env-steps/s", "h2d_bytes_per_step": n * n_act * 8,
                                   "d2h_bytes_per_step": n * (8 + 1), "steps": 20 * Ke,
                                   "note": "same public step() loop with the observations consumed in HBM (GPU-resident "
                                           "policy): actions from pinned host memory, reward+done back to the host"},
                    "device_obs_graph": {"value": e2e_graph_value, "unit": "M env-steps/s",
                                         "h2d_bytes_per_step": n * n_act * 8, "d2h_bytes_per_step": n * (8 + 1),
                                         "steps": 20 * Ke,
                                         "note": "as device_obs with env.capture_step(): one CUDA-graph launch per step"},
                    "device_obs_zero_copy": {"value": e2e_zero_value, "unit": "M env-steps/s",
                                             "h2d_bytes_per_step": n * n_act * 8, "d2h_bytes_per_step": n * (8 + 1),
                                             "steps": 20 * Ke,
                                             "note": "as device_obs, but the step kernel itself reads the actions from pinned "
                                                     "host memory and writes reward + done into pinned host memory (same "
                                                     "bytes over PCIe, no copy operation between two step launches)"}},
            "single_step": head["single_step"],
            "long_launches": head["long_launches"],
            "resets_per_env_step": head["resets_per_env_step"],
            **({"lookahead": head["lookahead"]} if head["lookahead"] else {}),
            "gpu_launches": head["timed_launches"],
            "roofline": head["roofline"],
            "configs": configs,
            "shard_check": check,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
