"""Helpers to replay tests/golden/ref_*.npz (trajectories of the unmodified reference)."""
import json
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    case = {k: z[k] for k in z.files}
    case["options"] = json.loads(str(z["options_json"]))
    case["infos"] = json.loads(str(z["infos_json"]))
    case["env_id"] = str(z["env_id"])
    case["seed0"] = int(z["seed0"])
    case["frame_map"] = {(int(e), int(t)): i for i, (e, t) in enumerate(z["frame_idx"])}
    return case


def crc(obs) -> int:
    """CRC of an observation; Dict observations (MortarMayhemB) hash visual then vector bytes."""
    if isinstance(obs, dict):
        return zlib.crc32(np.ascontiguousarray(obs["visual_observation"]).tobytes() +
                          np.ascontiguousarray(obs["vector_observation"], dtype=np.float32).tobytes())
    return zlib.crc32(np.ascontiguousarray(obs).tobytes())


def visual(obs):
    return obs["visual_observation"] if isinstance(obs, dict) else obs


def action_for(case, e, t):
    a = case["action"][e, t]
    return int(a[0]) if a.shape[0] == 1 else a
