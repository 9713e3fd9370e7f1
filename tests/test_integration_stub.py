"""The ctypes stub printed in INTEGRATION.md §3 is executable documentation: the CPU tier compiles it and checks its
buffer shapes against the header constants, the GPU tier runs it against the package's own env."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert len(blocks) == 1
    return blocks[0]


def test_stub_compiles_and_uses_header_constants():
    from memory_gym_b200 import _lib as L
    src = _stub_source()
    compile(src, "INTEGRATION.md", "exec")
    m = re.search(r"MG_MM_STATE_BYTES, MG_INFO_COLS, MG_FLAG_AUTORESET = (\d+), (\d+), (\d+)", src)
    assert tuple(int(x) for x in m.groups()) == (L.MG_MM_STATE_BYTES, L.MG_INFO_COLS, L.MG_FLAG_AUTORESET)
    assert "(n, MG_INFO_COLS)" in src and "(n, 8)" not in src


@pytest.mark.gpu
def test_stub_runs_and_matches_the_package():
    import torch
    import memory_gym_b200 as mg
    from memory_gym_b200 import _lib as L

    class MortarMayhemEnv:  # stands in for the reference class the stub would live next to
        process_reset_params = staticmethod(mg.spec("MortarMayhem-v0").process_reset_params)

    ns = {"LIBMEMGYM_B200_SO": L.LIB_PATH, "MortarMayhemEnv": MortarMayhemEnv}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    n = 1000  # envs >= 0.8 n finish too: their info rows lie past an [n][8] matrix
    stub = ns["MortarMayhemB200"](n, "cuda:0")
    env = mg.make("MortarMayhem-v0", num_envs=n, device="cuda:0")
    o1, _ = stub.reset(5)
    o2, _ = env.reset(seed=5)
    assert torch.equal(o1, o2)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    guard = torch.zeros(4096, dtype=torch.float64, device="cuda:0")  # (allocated right behind: must stay untouched)
    for _ in range(70):
        a = torch.randint(0, 3, (n, 2), generator=g, device="cuda:0")
        s1 = stub.step(a)
        s2 = env.step(a)
        assert torch.equal(s1[0], s2[0]) and torch.equal(s1[1], s2[1]) and torch.equal(s1[2], s2[2])
        d = s2[2]
        assert torch.equal(s1[4]["episode"][d], s2[4]["episode"][d])
    assert int(guard.abs().sum().item()) == 0
