"""CPU tier: the C-ABI library loads, exports every symbol include/memgym_b200.h declares, and its
host-side pieces (asset raster, SeedSequence/PCG64, option validation) match the oracle / numpy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from memory_gym_b200 import _lib as L
from memory_gym_b200.specs import SPECS
from memory_gym_b200.vec_env import _CudaBackend
from oracle import assets as OA
from oracle import sdl_raster as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host():
    be = _CudaBackend.__new__(_CudaBackend)
    be.lib = L.lib()
    return be


def test_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "memgym_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))
    assert len(names) == 34
    lib = L.lib()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.mg_abi_version() == L.MG_ABI_VERSION


def test_struct_sizes_match_header():
    # sizes the C side asserts (mg_common.h static_asserts) and ctypes mirrors
    assert C.sizeof(L.ChoiceList) == 4 + 4 * L.MG_MAX_CHOICES
    assert C.sizeof(L.MMParams) == 24 + 5 * 68 + 4 + 5 * 8  # 4 bytes padding before the doubles
    assert C.sizeof(L.StepIO) == 8 * 8 + 8 + 8 + 8 + 8  # ... flags + warps_per_cta, vector_obs, trace, debug_obs


def test_seed_sequence_matches_numpy():
    seeds = np.array([0, 1, 2, 100, 12345, 2**31 - 1, 2**32 - 1, 2**32, 2**32 + 5, 2**63 + 11, 2**64 - 1],
                     dtype=np.uint64)
    got = _host().seed(seeds)
    for s, row in zip(seeds, got):
        st = np.random.PCG64(np.random.SeedSequence(int(s))).state["state"]
        want = (st["state"] & (2**64 - 1), st["state"] >> 64, st["inc"] & (2**64 - 1), st["inc"] >> 64)
        assert tuple(int(v) for v in row) == want, int(s)


def test_pcg64_bounded_draws_match_numpy():
    lib = L.lib()
    rs = np.random.default_rng(99)
    for seed in (0, 7, 1234567):
        bg = np.random.PCG64(np.random.SeedSequence(seed))
        gen = np.random.Generator(bg)
        st = bg.state["state"]
        state4 = np.array([st["state"] & (2**64 - 1), st["state"] >> 64, st["inc"] & (2**64 - 1), st["inc"] >> 64],
                          dtype=np.uint64)
        has = np.zeros(2, dtype=np.uint32)
        n = 400
        lo = rs.integers(-5, 5, n).astype(np.int64)
        width = rs.choice([1, 1, 2, 3, 7, 9, 25, 36, 100, 5000, 2**31, 2**32 - 1, 2**32], n).astype(np.int64)
        hi = lo + width
        out = np.zeros(n, dtype=np.int64)
        uni = np.array([0.0025, 0.0075], dtype=np.float64)
        rc = lib.mg_selftest_pcg64_host(state4.ctypes.data, has.ctypes.data, lo.ctypes.data, hi.ctypes.data, n,
                                        out.ctypes.data, uni.ctypes.data)
        assert rc == 0
        want = np.array([int(gen.integers(int(a), int(b))) for a, b in zip(lo, hi)])
        assert np.array_equal(out, want)
        assert uni[0] == gen.uniform(0.0025, 0.0075)
        st2 = bg.state
        assert int(state4[0]) | (int(state4[1]) << 64) == st2["state"]["state"]
        assert int(has[0]) == st2["has_uint32"] and (not has[0] or int(has[1]) == st2["uinteger"])


def _decode(blob):
    hdr = np.frombuffer(blob[:8 * 4], dtype=np.uint32)
    assert hdr[0] == 0x4D474131 and hdr[1] == blob.shape[0]
    entries_off, n_entries = int(hdr[2]), int(hdr[3])
    tmpl_off = [int(v) for v in hdr[4:8]]
    ints = np.frombuffer(blob[32:32 + 8 * 4], dtype=np.int32)
    sprites_raw = np.frombuffer(blob[32 + 8 * 4:32 + 8 * 4 + 24 * 8], dtype=np.uint8).reshape(24, 8)
    dense_off = np.frombuffer(blob[32 + 8 * 4 + 24 * 8:32 + 8 * 4 + 24 * 8 + 24 * 4], dtype=np.uint32)
    p0 = 32 + 8 * 4 + 24 * 8 + 24 * 4
    pal = np.frombuffer(blob[p0:p0 + 32 * 4], dtype=np.uint8).reshape(32, 4)
    entries = np.frombuffer(blob[entries_off:entries_off + 4 * n_entries], dtype=np.uint32)
    sprites = []
    for row in sprites_raw:
        first = int(np.frombuffer(row[:4].tobytes(), dtype=np.uint32)[0])
        count = int(np.frombuffer(row[4:6].tobytes(), dtype=np.uint16)[0])
        sprites.append((first, count, int(row[6]), int(row[7])))
    uints = ints.view(np.uint32)
    return {"agent_dim": int(ints[0]), "agent_radius": int(ints[1]), "icon": (int(ints[2]), int(ints[3])),
            "disc_off": int(uints[4]), "trig_off": int(uints[5]), "stage_bytes": int(uints[6]),
            "dense_off": [int(v) for v in dense_off], "sprites": sprites, "pal": pal, "entries": entries,
            "tmpl_off": tmpl_off, "blob": blob}


def _sprite_surface(dec, slot, key):
    first, count, w, h = dec["sprites"][slot]
    px = np.full((h, w), key, dtype=np.uint32)
    for v in dec["entries"][first:first + count]:
        x, y, pi = int(v) & 255, (int(v) >> 8) & 255, (int(v) >> 16) & 255
        r, g, b = (int(c) for c in dec["pal"][pi][:3])
        px[y, x] = (r << 16) | (g << 8) | b
    return px


@pytest.mark.parametrize("agent_scale", [0.25, 0.3, 0.5, 1.0])
def test_mm_assets_match_oracle_raster(agent_scale):
    spec = SPECS["MortarMayhem-v0"]
    p = spec.process_reset_params({"agent_scale": agent_scale, "arena_size": 4})
    blob = _host().build_assets("mm", spec.to_struct(p))
    dec = _decode(blob)
    want = OA.agent_sprites(agent_scale)
    assert dec["agent_dim"] == want[0].w and dec["agent_radius"] == int(25 * agent_scale)
    for k in range(8):
        assert np.array_equal(_sprite_surface(dec, k, 255), want[k].px), ("agent", k)
    for k, name in enumerate(OA.COMMAND_NAMES):
        icon = OA.command_icon(name, 0.25)
        assert np.array_equal(_sprite_surface(dec, 8 + k, 0), icon.px), name
    assert dec["icon"] == tuple(int(v) for v in OA.command_blit_pos(0.25, 84))
    # arena templates == oracle tiles pasted on the screen, in [x][y][c]
    for slot, red in ((0, False), (1, True)):
        t = blob[dec["tmpl_off"][slot]:dec["tmpl_off"][slot] + 21168].reshape(84, 84, 3)
        scr = R.Surface(84, 84)
        tile = OA.mortar_tile(0.25, red)
        o = 42 - (14 * 4 >> 1)
        for i in range(4):
            for j in range(4):
                R.blit(tile, scr, (o + 14 * i, o + 14 * j))
        assert np.array_equal(t, R.array3d(scr)), slot


def test_mp_assets_match_oracle_raster():
    spec = SPECS["MysteryPath-v0"]
    blob = _host().build_assets("mp", spec.to_struct(spec.process_reset_params(None)))
    dec = _decode(blob)
    assert np.array_equal(_sprite_surface(dec, 17, 0), OA.fall_off_cross(0.25).px)
    assert not blob[dec["tmpl_off"][0]:dec["tmpl_off"][0] + 21168].any()


def _dense_surface(dec, slot, key):
    first, count, w, h = dec["sprites"][slot]
    off = dec["dense_off"][slot]
    idx = dec["blob"][off:off + w * h].reshape(h, w)
    px = np.full((h, w), key, dtype=np.uint32)
    for y in range(h):
        for x in range(w):
            if idx[y, x] != 255:
                r, g, b = (int(c) for c in dec["pal"][idx[y, x]][:3])
                px[y, x] = (r << 16) | (g << 8) | b
    return px


@pytest.mark.parametrize("opts", [None, {"agent_scale": 0.4, "coin_scale": 1.0, "exit_scale": 1.0}])
def test_ss_assets_match_oracle_raster(opts):
    import math
    spec = SPECS["SearingSpotlights-v0"]
    p = spec.process_reset_params(opts)
    blob = _host().build_assets("ss", spec.to_struct(p))
    dec = _decode(blob)
    want = OA.agent_sprites(p["agent_scale"])
    for k in range(8):
        assert np.array_equal(_sprite_surface(dec, k, 255), want[k].px)
        assert np.array_equal(_dense_surface(dec, k, 255), want[k].px)
    for slot, is_open in ((18, False), (19, True)):
        assert np.array_equal(_dense_surface(dec, slot, 255), OA.exit_sprite(p["exit_scale"], is_open).px)
    # coin = what Coin.draw leaves on the coin surface, centred in a (2r+2)^2 sprite
    r = int(10 * p["coin_scale"])
    cs = R.Surface(2 * r + 2, 2 * r + 2)
    R.fill(cs, 255)
    R.circle(cs, (255, 255, 0), (r + 1, r + 1), r)
    R.circle(cs, (255, 165, 0), (r + 1, r + 1), r, int(2 * p["coin_scale"]))
    assert np.array_equal(_dense_surface(dec, 20, 255), cs.px)
    # filled-circle row tables == the oracle primitive, for every radius
    discs = np.frombuffer(blob[dec["disc_off"]:dec["disc_off"] + 4 * 40 * 41].tobytes(), dtype=np.int8).reshape(-1, 2)
    for rad in range(1, 41):
        rows = R.disc_rows(rad)
        for dy in range(-rad, rad):
            lo, hi = discs[rad * (rad - 1) + dy + rad]
            assert (int(lo), int(hi)) == rows[dy], (rad, dy)
        # column spans (second table) describe the same pixel set
        cols = discs[40 * 41 + rad * (rad - 1):40 * 41 + rad * (rad + 1)]
        by_col = {(dx, dy) for dx in range(-rad, rad) for dy in range(int(cols[dx + rad][0]), int(cols[dx + rad][1]) + 1)}
        by_row = {(dx, dy) for dy in range(-rad, rad) for dx in range(rows[dy][0], rows[dy][1] + 1)}
        assert by_col == by_row, rad
    trig = np.frombuffer(blob[dec["trig_off"]:dec["trig_off"] + 720 * 8].tobytes(), dtype=np.float64).reshape(360, 2)
    for a in range(360):
        assert tuple(trig[a]) == R.vec_rotate(1.0, 0.0, float(a)), a
    assert dec["stage_bytes"] % 128 == 0 and dec["stage_bytes"] <= dec["tmpl_off"][0]


def test_option_validation_and_loud_failure_without_gpu():
    import memory_gym_b200 as mg
    import torch
    spec = SPECS["MortarMayhem-v0"]
    with pytest.raises(AssertionError):
        spec.process_reset_params({"not_an_option": 1})
    with pytest.raises(AssertionError):
        spec.process_reset_params({"arena_size": 7})
    with pytest.raises(AssertionError):
        SPECS["Endless-MortarMayhem-v0"].process_reset_params({"allowed_commands": 3})
    be = _host()
    with pytest.raises(L.MemGymError):
        be.check_params("mm", spec.to_struct(spec.process_reset_params({"explosion_delay": [0]})))
    with pytest.raises(L.MemGymError):
        be.check_params("mm", spec.to_struct(spec.process_reset_params({"command_count": [500]})))
    if not torch.cuda.is_available():
        with pytest.raises(L.MemGymError):
            mg.make("MortarMayhem-v0", num_envs=4)  # no CPU fallback
    # all 10 ids the reference registers (memory_gym/__init__.py:13-60) are built
    assert len(mg.registered_ids()) == 10 and not mg.NOT_BUILT
    ess = SPECS["Endless-SearingSpotlights-v0"]
    with pytest.raises(L.MemGymError):  # the reference crashes in step() without show_last_action
        be.check_params("ss", ess.to_struct(ess.process_reset_params({"show_last_action": False})))
    emp = SPECS["Endless-MysteryPath-v0"]
    with pytest.raises(L.MemGymError):
        be.check_params("ep", emp.to_struct(emp.process_reset_params({"stamina_level": 0})))
    assert C.sizeof(L.EPParams) == 8 * 4 + 7 * 8


def test_consumer_step_rejects_bad_pointers_without_a_device():
    """mg_obs_to_policy_bf16 validates its arguments before it touches the device (no compute without a GPU)."""
    lib = L.lib()
    assert lib.mg_obs_to_policy_bf16(None, None, 4, None) == -1
    assert b"NULL" in lib.mg_last_error()
    buf = np.zeros(84 * 84 * 3 * 9 + 64, np.uint8)
    base = buf.ctypes.data + (-buf.ctypes.data) % 16
    assert lib.mg_obs_to_policy_bf16(base + 1, base + 32, 1, None) == -1  # misaligned
    assert b"aligned" in lib.mg_last_error()
    assert lib.mg_obs_to_policy_bf16(base, base + 21168, 0, None) == 0            # empty batch: nothing to do


def test_product_never_touches_the_oracle_or_a_cpu_path():
    """The shipped package (Python host + C/CUDA sources) neither imports nor links anything under oracle/ or
    tests/hostsim, and make() refuses to run without the CUDA library / a CUDA device."""
    pkg = os.path.join(ROOT, "endless-memory-gym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
            assert "memgym_oracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), f
    mk = open(os.path.join(pkg, "csrc", "Makefile")).read()
    assert "oracle" not in mk and "hostsim" not in mk
