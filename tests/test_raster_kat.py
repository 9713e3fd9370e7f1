"""Known-answer tests: the raster oracle vs real PyGame renders (reference docs/assets/*.gif).

The vectors in tests/golden/kat_gifs.npz were cut out of the reference's GIFs by
tests/golden/make_kat_from_gifs.py (SCALE = 1.0 renders of an older pygame: thick circles use the
pre-2.1.3 algorithm there, hence ``legacy_thick=True``).
"""
import os

import numpy as np
import pytest

from oracle import assets, sdl_raster as R


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "kat_gifs.npz"))


def _mask124(surface):
    m = np.zeros((124, 124), dtype=bool)
    m[:surface.h, :surface.w] = surface.px != 0
    return m


@pytest.mark.parametrize("name", ["left", "right", "down", "left_down", "right_down", "left_up", "stay"])
def test_command_icons_match_pygame(kat, name):
    icon = assets.command_icon(name, 1.0, legacy_thick=True)
    want = np.unpackbits(kat["icon_" + name])[:124 * 124].reshape(124, 124).astype(bool)
    assert (icon.w, icon.h) == ((124, 124) if "_" in name else (88, 88))
    assert np.array_equal(_mask124(icon), want)


def test_new_thick_circle_differs_only_slightly():
    # the >=2.1.3 algorithm (default; parity unpinned) vs the GIF-verified legacy one
    a = assets.command_icon("stay", 0.25, legacy_thick=True)
    b = assets.command_icon("stay", 0.25, legacy_thick=False)
    assert int((a.px != b.px).sum()) == 4


def test_agent_sprite_rot0_matches_pygame(kat):
    sp = assets.agent_sprites(1.0, legacy_thick=True)[0]
    assert (sp.w, sp.h) == (74, 74)
    crop = kat["agent_rot0_crop"]
    rgb = np.stack([(sp.px >> 16) & 255, (sp.px >> 8) & 255, sp.px & 255], -1).astype(np.uint8)
    opaque = sp.px != 255
    assert opaque.sum() == 2288
    assert not ((crop != rgb).any(-1) & opaque).any()
    # everything the GIF shows in sprite colours inside the crop is covered by the sprite
    for col in (assets.BODY, assets.HAND, assets.HAND_OUTLINE):
        assert np.array_equal((crop == np.array(col)).all(-1), (rgb == np.array(col)).all(-1) & opaque)


@pytest.mark.parametrize("fi", [31, 53, 97])
def test_fall_off_cross_matches_pygame(kat, fi):
    cross = assets.fall_off_cross(1.0)
    want = np.unpackbits(kat[f"cross_{fi}"])[:1600].reshape(40, 40).astype(bool)
    assert np.array_equal(cross.px != 0, want)
    # cross top-left = agent centre - 20 (body bbox min = centre - 25)
    assert kat[f"cross_{fi}_origin_minus_body"].tolist() == [5, 5]


def test_exit_rounded_rect_matches_pygame(kat):
    e = assets.exit_sprite(2.0, False)
    rgb = np.stack([(e.px >> 16) & 255, (e.px >> 8) & 255, e.px & 255], -1).astype(np.uint8)
    opaque = e.px != 255
    assert opaque.sum() == 1416
    assert not ((kat["exit_closed_crop"] != rgb).any(-1) & opaque).any()


def test_sdl_surface_alpha_ramp_matches_pygame(kat):
    # black spotlight layer with alpha 42*k over the scene colours of frame 0
    # scene colours under the spotlight layer (the agent is not drawn in the GIF's frame 0)
    base = [tuple(c) for c in kat["ss_palette_0"].tolist()] + [assets.BODY, assets.HAND, assets.HAND_OUTLINE]
    for k in range(1, 8):
        a = min(255, 42 * k)
        dst = R.Surface(len(base), 1)
        for i, c in enumerate(base):
            dst.px[0, i] = R.map_rgb(c)
        src = R.Surface(len(base), 1)
        src.colorkey = R.map_rgb((255, 0, 0))
        src.alpha = a
        R.blit(src, dst, (0, 0))
        got = {((p >> 16) & 255, (p >> 8) & 255, p & 255) for p in dst.px[0].tolist()}
        want = {tuple(c) for c in kat[f"ss_palette_{k}"].tolist()}
        # top bar colours (green health, grey boxes) are not under the spotlight layer
        want -= {(0, 255, 0), (120, 120, 120)}
        got_under = {c for c in got}
        assert want <= got_under, (k, sorted(want - got_under))


def test_mm_show_cadence_and_tile_timing(kat):
    vis = kat["mm_icon_visible"]
    red = kat["mm_red_tiles"]
    assert len(vis) == 275
    # show phase: 10 commands x (3 visible + 1 blank)
    assert vis[:40].tolist() == ([True] * 3 + [False]) * 10
    assert not vis[40:].any()
    on = np.nonzero(red)[0]
    starts = [int(i) for i in on if not red[i - 1]]
    assert starts == [58 + 24 * k for k in range(10)][:len(starts)]
    for s in starts[:-1]:
        assert red[s:s + 6].all() and not red[s + 6]


# ---- round 2: vectors from the newer renders (pygame >= 2.1.3, pre-rotated sprites): tests/golden/kat_gifs2.npz --------
@pytest.fixture(scope="module")
def kat2(golden_dir):
    return np.load(os.path.join(golden_dir, "kat_gifs2.npz"))


def _rgb(px):
    return np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], -1).astype(np.uint8)


@pytest.mark.parametrize("k", [0, 2, 4, 6])
def test_prerotated_agent_sprites_match_pygame(kat2, k):
    """create_character_sprites (character_controller.py:29-75) at SCALE 1: the hand positions after
    Vector2.rotate_ip(360 - 45 k) and the 3-px hand outline (draw.circle width=3 of pygame >= 2.1.3)."""
    sp = assets.agent_sprites(1.0, legacy_thick=False)[k]
    crop = kat2[f"agent_new_rot{k}"]
    rgb, opaque = _rgb(sp.px), sp.px != 255
    assert not ((crop != rgb).any(-1) & opaque).any()
    for col in (assets.BODY, assets.HAND, assets.HAND_OUTLINE):
        assert np.array_equal((crop == np.array(col)).all(-1), (rgb == np.array(col)).all(-1) & opaque)
    others = [j for j in range(8) if j != k]
    assert all(((crop != _rgb(assets.agent_sprites(1.0)[j].px)).any(-1) & (assets.agent_sprites(1.0)[j].px != 255)).any()
               for j in others)  # the vector tells the rotations apart


@pytest.mark.parametrize("key,legacy", [("ring_new", False), ("ring_legacy", True)])
def test_thick_circle_ring_matches_pygame(kat2, key, legacy):
    """draw.circle(r=28, width=8), the ground-truth ring of _build_debug_surface: the >= 2.1.3 algorithm (restated from
    memory in round 1, 'parity unpinned' until now) against emm_0_gt.gif, the legacy one against mortar_mayhem_0_gt.gif."""
    h, w = kat2[key + "_shape"].tolist()
    want = np.unpackbits(kat2[key])[:h * w].reshape(h, w).astype(bool)
    s = R.Surface(64, 64)
    R.circle(s, (0, 255, 0), (32, 32), 28, 8, legacy_thick=legacy)
    ys, xs = np.nonzero(s.px)
    got = (s.px != 0)[ys.min():ys.max() + 1, xs.min():xs.max() + 1]
    assert got.shape == want.shape and np.array_equal(got, want)
    other = R.Surface(64, 64)
    R.circle(other, (0, 255, 0), (32, 32), 28, 8, legacy_thick=not legacy)
    assert int((other.px != s.px).sum()) > 0  # the two algorithms differ at this size: the vector selects one


@pytest.mark.parametrize("key,legacy", [("coin_new", False), ("coin_legacy", True)])
def test_coin_matches_pygame(kat2, key, legacy):
    """Coin.draw (pygame_assets.py:145-152) at SCALE 1: filled disc r=15 + outline of width 3 (thick circle again)."""
    s = R.Surface(40, 40)
    R.circle(s, (255, 255, 0), (20, 20), 15)
    R.circle(s, (255, 165, 0), (20, 20), 15, 3, legacy_thick=legacy)
    ys, xs = np.nonzero(s.px)
    sub = s.px[ys.min():ys.max() + 1, xs.min():xs.max() + 1]
    crop = kat2[key]
    assert crop.shape[:2] == sub.shape
    assert not ((crop != _rgb(sub)).any(-1) & (sub != 0)).any()


def test_arena_6x6_matches_pygame(kat2):
    """MortarArena(size 6) at SCALE 1 (endless_mortar_mayhem.py:100): 56-px tiles with a 4-px frame."""
    tile = assets.mortar_tile(1.0, False)
    full = np.tile(tile.px, (6, 6))
    fill = np.unpackbits(kat2["arena6_fill"])[:336 * 336].reshape(336, 336).astype(bool)
    edge = np.unpackbits(kat2["arena6_edge"])[:336 * 336].reshape(336, 336).astype(bool)
    assert fill.sum() + edge.sum() > 100000
    assert (full[fill] == R.map_rgb((21, 43, 77))).all() and (full[edge] == R.map_rgb((29, 60, 107))).all()


def test_spotlight_discs_match_pygame(kat2):
    """Spotlight.draw (pygame_assets.py:109-117): pygame's filled circle at every radius that appears whole in the
    ground-truth render of Endless Searing Spotlights."""
    radii = kat2["disc_radii"].tolist()
    assert len(radii) >= 8
    for r in radii:
        s = R.Surface(2 * r + 2, 2 * r + 2)
        R.circle(s, (255, 0, 0), (r + 1, r + 1), r)
        ys, xs = np.nonzero(s.px)
        got = (s.px != 0)[ys.min():ys.max() + 1, xs.min():xs.max() + 1]
        want = np.unpackbits(kat2[f"disc_{r}"])[:4 * r * r].reshape(2 * r, 2 * r).astype(bool)
        assert np.array_equal(got, want), r
        rows = R.disc_rows(r)  # the row-span tables the kernels' disc tables are built like
        assert got.tolist() == [[rows[dy][0] <= x <= rows[dy][1] for x in range(-r, r)] for dy in range(-r, r)], r


def test_no_stay_icon_in_the_new_renders(kat2):
    """The 'stay' icon (thick circle r=40 w=8) itself never appears in a GIF rendered with pygame >= 2.1.3; its
    algorithm is pinned through the ring and the coin outline above."""
    assert kat2["stay_frames_emm"].size == 0
