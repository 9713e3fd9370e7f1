"""Extract known-answer vectors from the reference's real-PyGame renders (docs/assets/*.gif).

Run in the BUILD container only (needs /root/reference and Pillow):

    python tests/golden/make_kat_from_gifs.py

Writes ``tests/golden/kat_gifs.npz`` (bit-packed masks and small RGB crops; a few kB).  The GIFs are
336x336 frames rendered by real PyGame at SCALE = 1.0, so they pin the *algorithms* of the raster
oracle (oracle/sdl_raster.py) at that scale: thick lines, transform.rotate, filled circle, legacy
thick circle, rounded rect, colour-key blit and the SDL surface-alpha blend.
"""
import os
import sys

import numpy as np
from PIL import Image

ASSETS = "/root/reference/docs/assets"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_gifs.npz")


def frames(name, idx):
    im = Image.open(os.path.join(ASSETS, name))
    out = {}
    for i in idx:
        im.seek(i)
        out[i] = np.array(im.convert("RGB"))
    return out


def n_frames(name):
    return Image.open(os.path.join(ASSETS, name)).n_frames


def main():
    kat = {}
    # --- Mortar Mayhem: command icons (white pixels), 88x88 / 124x124 placed at (124, 124)
    icon_frames = {"left": 0, "right": 8, "down": 16, "left_down": 24, "right_down": 28, "left_up": 32, "stay": 36}
    mm = frames("mortar_mayhem_0.gif", range(n_frames("mortar_mayhem_0.gif")))
    for name, fi in icon_frames.items():
        white = (mm[fi] == 255).all(-1)
        assert white[:124].sum() == 0 and white[:, :124].sum() == 0
        kat["icon_" + name] = np.packbits(white[124:124 + 124, 124:124 + 124])
    # --- agent sprite, rotation 0, 74x74 at (245, 75) in frame 0 (RGB crop)
    kat["agent_rot0_crop"] = mm[0][75:75 + 74, 245:245 + 74].copy()
    # --- show cadence + tile timing
    kat["mm_icon_visible"] = np.array([(mm[i] == 255).all(-1).any() for i in range(len(mm))])
    kat["mm_red_tiles"] = np.array([(mm[i] == np.array([81, 18, 26])).all(-1).any() for i in range(len(mm))])
    # --- Mystery Path: fall-off cross (red), 40x40 centred on the agent
    mp = frames("mystery_path_0.gif", [31, 53, 97])
    for fi, f in mp.items():
        red = (f == np.array([255, 0, 0])).all(-1)
        ys, xs = np.nonzero(red)
        body = (f == np.array([250, 204, 153])).all(-1)
        by, bx = np.nonzero(body)
        kat[f"cross_{fi}"] = np.packbits(red[ys.min():ys.min() + 40, xs.min():xs.min() + 40])
        kat[f"cross_{fi}_origin_minus_body"] = np.array([xs.min() - bx.min(), ys.min() - by.min()])
    # --- Searing Spotlights: closed exit 40x40 at (142, 40) of the ground-truth frame 0
    ssg = frames("searing_spotlights_0_gt.gif", [0])
    kat["exit_closed_crop"] = ssg[0][40:80, 142:182].copy()
    # --- SDL alpha ramp: palette of frames 0..8 of the agent observation
    ss = frames("searing_spotlights_0.gif", range(9))
    for i in range(9):
        kat[f"ss_palette_{i}"] = np.unique(ss[i].reshape(-1, 3), axis=0)
    np.savez_compressed(OUT, **kat)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
