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
    main2()


OUT2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_gifs2.npz")


def main2():
    """Round 2: vectors from the NEWER renders (emm_0 / ess_0 / emp_0 and the *_gt GIFs): pre-rotated agent sprites,
    the 6x6 arena, spotlight discs, and — because those GIFs were rendered with pygame >= 2.1.3 — the new thick-circle
    algorithm (ground-truth ring r=28 w=8, coin outline r=15 w=3) next to the legacy one of the older GIFs."""
    kat = {}
    body = np.array([250, 204, 153])

    def all_frames(name):
        im = Image.open(os.path.join(ASSETS, name))
        for i in range(im.n_frames):
            im.seek(i)
            yield i, np.array(im.convert("RGB"))

    # --- agent sprites (74x74) of the newer, pre-rotated code: the first fully visible, unoccluded frame whose hands
    # sit where rotation k puts them (rotation = which of the four sides shows the hands; diagonals never occur: the
    # demo policies only move along the axes)
    want = {"emm_0.gif": [2, 4, 6], "ess_0_gt.gif": [0]}
    for name, rots in want.items():
        left = set(rots)
        for i, f in all_frames(name):
            if not left:
                break
            b = (f == body).all(-1)
            if b.sum() < 1000:
                continue
            ys, xs = np.nonzero(b)
            if xs.max() - xs.min() != 49 or ys.max() - ys.min() != 49:
                continue
            x0, y0 = xs.min() + 25 - 37, ys.min() + 25 - 37
            if x0 < 0 or y0 < 0 or x0 + 74 > 336 or y0 + 74 > 336:
                continue
            crop = f[y0:y0 + 74, x0:x0 + 74]
            cols = {tuple(c) for c in np.unique(crop.reshape(-1, 3), axis=0).tolist()}
            if (0, 255, 0) in cols or (255, 0, 0) in cols:  # ring / cross on top
                continue
            hand = (crop == np.array([250, 250, 250])).all(-1)
            hy, hx = np.nonzero(hand)
            if len(hy) == 0:
                continue
            side = {0: hy.mean() < 30, 4: hy.mean() > 44, 2: hx.mean() < 30, 6: hx.mean() > 44}
            for k in sorted(left):
                if side[k]:
                    kat[f"agent_new_rot{k}"] = crop.copy()
                    kat[f"agent_new_rot{k}_src"] = np.asarray([i, x0, y0])
                    left.discard(k)
                    break
        assert not left, (name, left)
    # --- ground-truth ring: draw.circle((0,255,0), tile centre, 28, 8)
    for key, name in (("ring_new", "emm_0_gt.gif"), ("ring_legacy", "mortar_mayhem_0_gt.gif")):
        f = frames(name, [0])[0]
        g = (f == np.array([0, 255, 0])).all(-1)
        ys, xs = np.nonzero(g)
        m = g[ys.min():ys.max() + 1, xs.min():xs.max() + 1]
        kat[key + "_shape"] = np.asarray(m.shape)
        kat[key] = np.packbits(m)
    # --- coin: filled (255,255,0) disc r=15 + (255,165,0) outline of width 3
    for key, name in (("coin_new", "ess_0.gif"), ("coin_legacy", "searing_spotlights_0_gt.gif")):
        f = frames(name, [0])[0]
        m = (f == np.array([255, 165, 0])).all(-1) | (f == np.array([255, 255, 0])).all(-1)
        ys, xs = np.nonzero(m)
        kat[key] = f[ys.min():ys.max() + 1, xs.min():xs.max() + 1].copy()
    # --- the 6x6 arena of Endless Mortar Mayhem (frame 0): where the two tile colours are visible
    f = frames("emm_0.gif", [0])[0]
    kat["arena6_fill"] = np.packbits((f == np.array([21, 43, 77])).all(-1))
    kat["arena6_edge"] = np.packbits((f == np.array([29, 60, 107])).all(-1))
    # --- spotlight discs (pygame filled circle) at every radius that shows up whole and unoccluded in the GT render
    from scipy import ndimage
    discs = {}
    for i, f in all_frames("ess_0_gt.gif"):
        if i < 20 or i % 3:
            continue
        lit = (f == 255).all(-1) | (f == np.array([0, 0, 255])).all(-1)
        lit[:16] = False
        if lit.sum() > 336 * 336 // 2:
            continue
        lab, _ = ndimage.label(lit)
        for sl in ndimage.find_objects(lab):
            h, w = sl[0].stop - sl[0].start, sl[1].stop - sl[1].start
            if h != w or h % 2 or sl[0].start <= 16 or sl[1].start <= 0 or sl[0].stop >= 336 or sl[1].stop >= 336:
                continue
            m = lit[sl]
            # a whole disc is symmetric under both mirrors and the transpose (an occluded or merged one is not)
            if h // 2 not in discs and np.array_equal(m, m[::-1]) and np.array_equal(m, m[:, ::-1]) and np.array_equal(m, m.T):
                discs[h // 2] = (np.packbits(m), i)
    for r, (bits, i) in discs.items():
        kat[f"disc_{r}"] = bits
    kat["disc_radii"] = np.asarray(sorted(discs))
    # --- is there a "stay" icon (thick circle r=40 w=8) anywhere in the renders made with pygame >= 2.1.3?
    # An icon frame of the obs GIFs has white pixels only inside [124, 248)^2; "stay" is the only icon that is 88 px
    # wide with a closed ring: both its top row and bottom row carry white pixels and its centre column is empty above
    # the bar. Recorded so the test can state the fact.
    stay_frames = []
    for name in ("emm_0.gif",):
        for i, f in all_frames(name):
            w = (f == 255).all(-1)[124:212, 124:212]
            if w.sum() and w[0].any() and w[-9:].any() and w[:, 0].any() and w[:, 87].any():
                stay_frames.append(i)
    kat["stay_frames_emm"] = np.asarray(stay_frames, dtype=np.int64)
    np.savez_compressed(OUT2, **kat)
    print("wrote", OUT2, os.path.getsize(OUT2), "bytes", "disc radii", sorted(discs), "stay frames", stay_frames)


if __name__ == "__main__":
    sys.exit(main())
