"""Where a host-pipe step spends its time (wall clock around the pieces of HostPipe.submit / collect)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import ctypes as C
import torch
import memory_gym_b200 as mg
from memory_gym_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = mg.make("MortarMayhem-v0", num_envs=n, device="cuda:0")
env.reset(seed=0)
g = torch.Generator().manual_seed(0)
acts = [torch.randint(0, 3, (n, 2), generator=g).pin_memory() for _ in range(32)]
for threads in (4, 8, 12, 16):
    pipe = env.host_pipe(threads=threads, depth=2)
    for t in range(40):
        pipe.step(acts[t % 32])
    T = 400
    t_sub = t_wait = t_dec = 0.0
    t0 = time.perf_counter()
    pipe.submit(acts[0])
    for t in range(1, T):
        a = time.perf_counter()
        pipe.submit(acts[t % 32])
        b = time.perf_counter()
        slot = pipe._tail % pipe.depth
        pipe._ev[slot].synchronize()
        c = time.perf_counter()
        L.check(pipe._lib.mg_obs_delta_decode(pipe._stage[slot].data_ptr(), pipe.obs.data_ptr(), n, pipe.threads,
                                              C.byref(pipe._bytes)), "decode")
        d = time.perf_counter()
        pipe._tail += 1
        t_sub += b - a
        t_wait += c - b
        t_dec += d - c
    pipe.collect()
    total = time.perf_counter() - t0
    print(f"threads {threads:2d}: {total / T * 1e6:7.1f} us/step = {n * T / total / 1e6:6.2f} M env-steps/s | submit {t_sub / T * 1e6:6.1f} "
          f"event wait {t_wait / T * 1e6:6.1f} decode {t_dec / T * 1e6:6.1f} us", flush=True)
# GPU side alone: step + encode back to back, no host work in between
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pipe = env.host_pipe(threads=16, depth=2)
torch.cuda.synchronize()
e0.record()
for t in range(200):
    env.step(acts[t % 32], results_out=pipe._res[0])
    pipe._encode(0, False)
e1.record()
torch.cuda.synchronize()
print(f"GPU side (step + encode): {e0.elapsed_time(e1) * 1e3 / 200:.1f} us per step")
