"""Measurement helper: per-env phase timeline of one step kernel in steady state (mg_step_io.trace),
plus write-only / copy bandwidth reference points of the device."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import numpy as np
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
env_id = sys.argv[1] if len(sys.argv) > 1 else "MortarMayhem-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
wpc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
T = int(sys.argv[4]) if len(sys.argv) > 4 else 64


def bw_refs():
    x = torch.empty(2**30, dtype=torch.uint8, device=dev)
    y = torch.empty(2**30, dtype=torch.uint8, device=dev)
    for name, fn, nbytes in (("memset 1 GiB", lambda: x.zero_(), 2**30), ("copy 1 GiB (r+w)", lambda: y.copy_(x), 2**31)):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{name}: {nbytes / best / 1e6:.1f} GB/s")


bw_refs()
env = mg.make(env_id, num_envs=n, device=dev, warps_per_cta=wpc)
env.reset(seed=0)
g = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, env.spec_.action_nvec[0], (T,) + tuple(env._actions_dev.shape), generator=g, device=dev)
ring = torch.empty((6, n, 84, 84, 3), dtype=torch.uint8, device=dev)
rew = torch.empty((T, n), dtype=torch.float64, device=dev)
dn = torch.empty((T, n), dtype=torch.bool, device=dev)
trace = torch.zeros((2, n, 16), dtype=torch.int64, device=dev)
env.step_many(acts, ring, rew, dn)
env._many_ios[T - 2].trace = trace[0].data_ptr()
env._many_ios[T - 1].trace = trace[1].data_ptr()
for _ in range(3):
    env.step_many(acts, ring, rew, dn)
torch.cuda.synchronize()
tr = trace.cpu().numpy().astype(np.int64)
a, b = tr[0], tr[1]
t0 = a[:, 0].min()
print(f"step A: kernel entry min {0} max {(a[:,0].max()-t0)/1e3:.2f} us; last store issued {(a[:,6].max()-t0)/1e3:.2f} us")
print(f"step B: first entry {(b[:,0].min()-t0)/1e3:.2f} us, dependency wait done {(b[:,2].min()-t0)/1e3:.2f}..{(b[:,2].max()-t0)/1e3:.2f} us, "
      f"last store issued {(b[:,6].max()-t0)/1e3:.2f} us")
names = ["entry->prologue", "prologue->depwait", "depwait->record", "record->logic", "logic->composed", "composed->issued"]
for k, nm in enumerate(names):
    d = (b[:, k + 1] - b[:, k]) / 1e3
    print(f"  {nm:20s} mean {d.mean():7.2f} us  p50 {np.median(d):7.2f}  p95 {np.percentile(d, 95):7.2f}  max {d.max():7.2f}")
if b[:, 8].any():
    for nm, i0, i1 in (("emit: update base", 8, 9), ("emit: lit + lut", 9, 10), ("emit: masked copy", 10, 11)):
        d = (b[:, i1] - b[:, i0]) / 1e3
        print(f"  {nm:20s} mean {d.mean():7.2f} us  p50 {np.median(d):7.2f}  p95 {np.percentile(d, 95):7.2f}  max {d.max():7.2f}")
# per-warp chains: envs sharing a warp are n/total_warps apart in issue order
t_issue = np.sort(b[:, 6] - b[:, 2].min()) / 1e3
for q in (1, 10, 25, 50, 75, 90, 99, 100):
    print(f"  {q:3d}% of stores issued by {np.percentile(t_issue, q):7.2f} us after dependency wait")
sm = b[:, 7]
cnt = np.bincount(sm.astype(np.int64))
print("envs per SM: min", cnt[cnt > 0].min(), "max", cnt.max(), "SMs used", (cnt > 0).sum())
# which envs were slow in the logic phase?
st = env.state.cpu().numpy()
ep_len = st[:, 48:52].copy().view(np.int32)[:, 0]
show_pos = st[:, 90:92].copy().view(np.uint16)[:, 0]
show_len = st[:, 92:94].copy().view(np.uint16)[:, 0]
d_logic = (b[:, 4] - b[:, 3]) / 1e3
act = show_pos >= show_len
was_reset = ep_len == 0
for name, m in (("show phase", ~act & ~was_reset), ("act phase", act & ~was_reset), ("auto-reset", was_reset)):
    if m.any():
        print(f"  logic time, {name:10s}: n={m.sum():5d} mean {d_logic[m].mean():7.2f} p50 {np.median(d_logic[m]):7.2f} max {d_logic[m].max():7.2f} us")
