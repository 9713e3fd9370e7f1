"""Host half of the observation transport (mg_obs_delta_decode) against a numpy statement of the stage format — the
device encoder is checked against the same statement in the GPU tier (test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

from memory_gym_b200 import _lib as L

CHUNKS = L.MG_OBS_BYTES // 16


def encode_numpy(cur, prev, full=False):
    """Stage bytes for frames `cur` given the frames `prev` the host holds (both uint8[n, 21168])."""
    n = cur.shape[0]
    stage = np.zeros(n * (L.MG_DELTA_HDR_BYTES + L.MG_OBS_BYTES), dtype=np.uint8)
    hdr = stage[:n * L.MG_DELTA_HDR_BYTES].view(np.uint32).reshape(n, L.MG_DELTA_HDR_BYTES // 4)
    pay = stage[n * L.MG_DELTA_HDR_BYTES:].reshape(n, L.MG_OBS_BYTES)
    c16, p16 = cur.reshape(n, CHUNKS, 16), prev.reshape(n, CHUNKS, 16)
    for e in range(n):
        changed = np.ones(CHUNKS, bool) if full else (c16[e] != p16[e]).any(1)
        idx = np.flatnonzero(changed)
        hdr[e, 0] = len(idx)
        for i in idx:
            hdr[e, 1 + i // 32] |= np.uint32(1 << (i % 32))
        pay[e, :16 * len(idx)] = c16[e, idx].reshape(-1)
    return stage


def decode(stage, obs, threads):
    got = C.c_uint64(0)
    rc = L.lib().mg_obs_delta_decode(stage.ctypes.data, obs.ctypes.data, obs.shape[0], threads, C.byref(got))
    return rc, int(got.value)


@pytest.mark.parametrize("n,threads", [(1, 1), (7, 4), (300, 1), (300, 6), (1000, 16)])
def test_decode_roundtrip(n, threads):
    rng = np.random.default_rng(n + threads)
    host = np.zeros((n, L.MG_OBS_BYTES), np.uint8)
    prev = host.copy()
    for step in range(4):
        cur = prev.copy()
        if step == 0:
            cur[:] = rng.integers(0, 256, cur.shape, dtype=np.uint8)
        else:  # sparse changes: single bytes, runs across chunk and mask-word boundaries, the very last chunk
            for e in range(n):
                for _ in range(int(rng.integers(0, 12))):
                    lo = int(rng.integers(0, L.MG_OBS_BYTES))
                    ln = int(rng.integers(1, 700))
                    cur[e, lo:lo + ln] ^= rng.integers(1, 256, len(cur[e, lo:lo + ln]), dtype=np.uint8)
            cur[0, -1] ^= 0x55
            cur[n - 1, 16 * 31:16 * 33] ^= 0xFF
        stage = encode_numpy(cur, prev, full=(step == 2))
        rc, nbytes = decode(stage, host, threads)
        assert rc == 0
        assert np.array_equal(host, cur)
        counts = stage[:n * L.MG_DELTA_HDR_BYTES].view(np.uint32).reshape(n, -1)[:, 0]
        assert nbytes == 16 * int(counts.sum()) + n * L.MG_DELTA_HDR_BYTES
        prev = cur


def test_decode_rejects_malformed_header():
    n = 100
    host = np.zeros((n, L.MG_OBS_BYTES), np.uint8)
    cur = host.copy()
    cur[5, 100] = 1
    stage = encode_numpy(cur, host)
    stage[:n * L.MG_DELTA_HDR_BYTES].view(np.uint32).reshape(n, -1)[5, 0] = 3  # count disagrees with the mask
    rc, _ = decode(stage, host, 4)
    assert rc != 0
    assert b"malformed" in L.lib().mg_last_error()
    assert int(L.lib().mg_obs_delta_stage_bytes(n)) == n * (L.MG_DELTA_HDR_BYTES + L.MG_OBS_BYTES)
