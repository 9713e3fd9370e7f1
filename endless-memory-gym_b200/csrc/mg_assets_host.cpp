// mg_assets_host.cpp — builds the static drawable assets of Memory Gym into one flat blob.
//
// What the reference rebuilds with pygame on every reset (agent sprites, character_controller.py:29-75;
// 25 MortarTile surfaces, pygame_assets.py:316-343) or on every show-phase step (Command icon,
// pygame_assets.py:254-304) is drawn here ONCE per option set with the raster of mg_raster_host.cpp
// and stored as sparse opaque-pixel lists + full-frame background templates. The kernels only copy.
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_raster_host.h"

using mgr::Surface;

static const double SCALE = 0.25;  // mortar_mayhem.py:13 (every env module uses the same constant)

namespace {

struct Builder {
    MgAssetHeader hdr;
    std::vector<uint32_t> entries;
    std::vector<std::vector<uint8_t>> templates;
    std::vector<uint8_t> dense[MG_SPR_COUNT];  // w*h palette indices, 255 = transparent
    std::vector<int8_t> discs;                 // filled-circle row spans, then column spans (SS spotlights)
    bool discs_ok = true;
    std::vector<double> trig;                  // (cos, sin) of integer degrees
    std::vector<uint8_t> rings;                // uint16 start[R + 2] + int8 (dx, dy) pairs of the width-1 circles
    int n_pal = 0;

    Builder() {
        memset(&hdr, 0, sizeof(hdr));
        static const int fixed[][3] = {{0, 0, 0},    {250, 204, 153}, {250, 250, 250}, {50, 50, 50},
                                       {255, 255, 255}, {255, 0, 0},     {55, 55, 55},    {48, 141, 70},
                                       {255, 165, 0},   {0, 0, 255},     {0, 255, 0}};
        for (auto& c : fixed) {
            hdr.palette[n_pal][0] = (uint8_t)c[0];
            hdr.palette[n_pal][1] = (uint8_t)c[1];
            hdr.palette[n_pal][2] = (uint8_t)c[2];
            n_pal++;
        }
    }

    int pal_index(uint32_t px) {
        uint8_t r = (px >> 16) & 255, g = (px >> 8) & 255, b = px & 255;
        for (int i = 0; i < n_pal; ++i)
            if (hdr.palette[i][0] == r && hdr.palette[i][1] == g && hdr.palette[i][2] == b) return i;
        if (n_pal >= MG_MAX_PALETTE) return -1;
        hdr.palette[n_pal][0] = r;
        hdr.palette[n_pal][1] = g;
        hdr.palette[n_pal][2] = b;
        return n_pal++;
    }

    // opaque pixels of `s` (everything != colour key), column-major so neighbouring lanes of a warp
    // touch neighbouring bytes of the [x][y][c] frame
    bool add_sprite(int slot, const Surface& s) {
        if (s.w > 255 || s.h > 255) return false;
        MgSprite& sp = hdr.sprite[slot];
        sp.first = (uint32_t)entries.size();
        sp.w = (uint8_t)s.w;
        sp.h = (uint8_t)s.h;
        uint32_t cnt = 0;
        for (int x = 0; x < s.w; ++x)
            for (int y = 0; y < s.h; ++y) {
                uint32_t v = s.at(x, y);
                if (s.has_key && v == s.key) continue;
                int pi = pal_index(v);
                if (pi < 0) return false;
                entries.push_back((uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)pi << 16));
                cnt++;
            }
        if (cnt > 65535) return false;
        sp.count = (uint16_t)cnt;
        return true;
    }

    // dense bitmap of the same sprite: O(1) "what does this sprite show at (u, v)" for per-pixel composition
    bool add_dense(int slot, const Surface& s) {
        dense[slot].assign((size_t)s.w * s.h, 255);
        for (int y = 0; y < s.h; ++y)
            for (int x = 0; x < s.w; ++x) {
                uint32_t v = s.at(x, y);
                if (s.has_key && v == s.key) continue;
                int pi = pal_index(v);
                if (pi < 0 || pi >= 255) return false;
                dense[slot][(size_t)y * s.w + x] = (uint8_t)pi;
            }
        return true;
    }

    // pygame.draw.circle(filled) row spans for every radius the kernel may draw at run time: derived by
    // running the primitive itself, so the tables cannot drift from it
    void add_disc_tables() {
        discs.clear();
        for (int r = 1; r <= MG_MAX_DISC_R; ++r) {
            Surface s(2 * r + 4, 2 * r + 4);
            int c = r + 2;
            mgr::circle(s, 1, c, c, r, 0, false);
            for (int dy = -r; dy < r; ++dy) {
                int lo = 1, hi = 0;  // empty row: xmin > xmax
                bool any = false;
                for (int x = 0; x < s.w; ++x)
                    if (s.at(x, c + dy)) {
                        if (!any) lo = x - c;
                        hi = x - c;
                        any = true;
                    }
                discs.push_back((int8_t)lo);
                discs.push_back((int8_t)hi);
            }
        }
        // the same discs as COLUMN spans (dx = -r..r-1 -> ymin, ymax): a column of the [x][y][c] frame is
        // contiguous memory, so the kernel fills discs column-wise. Appended after the row tables.
        for (int r = 1; r <= MG_MAX_DISC_R; ++r) {
            Surface s(2 * r + 4, 2 * r + 4);
            int c = r + 2;
            mgr::circle(s, 1, c, c, r, 0, false);
            for (int dx = -r; dx < r; ++dx) {
                int lo = 1, hi = 0, runs = 0;
                bool prev = false;
                for (int y = 0; y < s.h; ++y) {
                    bool on = s.at(c + dx, y) != 0;
                    if (on && !prev) {
                        runs++;
                        lo = y - c;
                    }
                    if (on) hi = y - c;
                    prev = on;
                }
                if (runs > 1) discs_ok = false;  // a column must be one run (convex shape)
                discs.push_back((int8_t)lo);
                discs.push_back((int8_t)hi);
            }
        }
    }

    // pygame.draw.circle(width=1) pixel sets (Spotlight border when black_background, pygame_assets.py:112-113)
    void add_ring_tables() {
        std::vector<uint16_t> start(MG_MAX_DISC_R + 2, 0);
        std::vector<int8_t> px;
        for (int r = 1; r <= MG_MAX_DISC_R; ++r) {
            start[r] = (uint16_t)(px.size() / 2);
            Surface s(2 * r + 4, 2 * r + 4);
            int c = r + 2;
            mgr::circle(s, 1, c, c, r, 1, false);
            for (int x = 0; x < s.w; ++x)
                for (int y = 0; y < s.h; ++y)
                    if (s.at(x, y)) {
                        px.push_back((int8_t)(x - c));
                        px.push_back((int8_t)(y - c));
                    }
        }
        start[MG_MAX_DISC_R + 1] = (uint16_t)(px.size() / 2);
        start[0] = start[1];
        rings.resize(start.size() * 2 + px.size());
        memcpy(rings.data(), start.data(), start.size() * 2);
        memcpy(rings.data() + start.size() * 2, px.data(), px.size());
    }

    // pygame Vector2.rotate for integer degrees: exact quarter turns, else libm sin/cos of angle * pi / 180
    void add_trig_table() {
        trig.resize(720);
        for (int a = 0; a < 360; ++a) {
            double x, y;
            mgr::vec_rotate(1.0, 0.0, (double)a, &x, &y);
            trig[2 * a] = x;
            trig[2 * a + 1] = y;
        }
    }

    void add_template(int slot, const Surface& s) {
        // pygame.surfarray.array3d layout: [x][y][c]
        std::vector<uint8_t> t((size_t)MG_FRAME_BYTES);
        for (int x = 0; x < MG_W; ++x)
            for (int y = 0; y < MG_W; ++y) {
                uint32_t v = s.at(x, y);
                uint8_t* o = &t[((size_t)x * MG_W + y) * 3];
                o[0] = (v >> 16) & 255;
                o[1] = (v >> 8) & 255;
                o[2] = v & 255;
            }
        if ((int)templates.size() <= slot) templates.resize(slot + 1);
        templates[slot] = t;
    }

    long serialize(void* out, size_t cap) {
        size_t off = (sizeof(MgAssetHeader) + 15) & ~(size_t)15;
        hdr.magic = MG_ASSET_MAGIC;
        hdr.entries_off = (uint32_t)off;
        hdr.n_entries = (uint32_t)entries.size();
        off += entries.size() * 4;
        for (int i = 0; i < MG_SPR_COUNT; ++i) {
            hdr.dense_off[i] = 0;
            if (dense[i].empty()) continue;
            off = (off + 3) & ~(size_t)3;
            hdr.dense_off[i] = (uint32_t)off;
            off += dense[i].size();
        }
        hdr.disc_off = 0;
        if (!discs.empty()) {
            off = (off + 3) & ~(size_t)3;
            hdr.disc_off = (uint32_t)off;
            off += discs.size();
        }
        hdr.ring_off = 0;
        if (!rings.empty()) {
            off = (off + 3) & ~(size_t)3;
            hdr.ring_off = (uint32_t)off;
            off += rings.size();
        }
        hdr.trig_off = 0;
        if (!trig.empty()) {
            off = (off + 7) & ~(size_t)7;
            hdr.trig_off = (uint32_t)off;
            off += trig.size() * sizeof(double);
        }
        off = (off + 127) & ~(size_t)127;
        hdr.stage_bytes = (uint32_t)off;
        for (size_t i = 0; i < templates.size() && i < 4; ++i) {
            if (templates[i].empty()) continue;
            hdr.tmpl_off[i] = (uint32_t)off;
            off += MG_FRAME_BYTES;
            off = (off + 127) & ~(size_t)127;
        }
        hdr.total_bytes = (uint32_t)off;
        if (!out || cap < off) return (long)off;  // size query / too small: caller re-allocates
        memset(out, 0, off);
        memcpy(out, &hdr, sizeof(hdr));
        memcpy((char*)out + hdr.entries_off, entries.data(), entries.size() * 4);
        for (int i = 0; i < MG_SPR_COUNT; ++i)
            if (hdr.dense_off[i]) memcpy((char*)out + hdr.dense_off[i], dense[i].data(), dense[i].size());
        if (hdr.disc_off) memcpy((char*)out + hdr.disc_off, discs.data(), discs.size());
        if (hdr.ring_off) memcpy((char*)out + hdr.ring_off, rings.data(), rings.size());
        if (hdr.trig_off) memcpy((char*)out + hdr.trig_off, trig.data(), trig.size() * sizeof(double));
        for (size_t i = 0; i < templates.size() && i < 4; ++i)
            if (!templates[i].empty()) memcpy((char*)out + hdr.tmpl_off[i], templates[i].data(), MG_FRAME_BYTES);
        return (long)off;
    }
};

// character_controller.py:29-75 (create_character_sprites)
bool build_agent(Builder& b, double scale, bool legacy, bool with_dense = false) {
    int radius = (int)(25 * scale), hands_x = (int)(18 * scale), hand_y = (int)(12 * scale);
    int hand_r = (int)(10 * scale), outline = (int)(3 * scale), ext = 14;
    int dim = radius * 2 + hand_r + ext;
    if (dim < 1 || dim > 255) return false;
    int c = dim / 2;
    int lx0 = dim / 2 - hands_x, ly0 = hand_y + ext / 2, rx0 = dim / 2 + hands_x, ry0 = ly0;
    int k = 0;
    for (int ang = 360; ang > 0; ang -= 45, ++k) {
        Surface s(dim, dim);
        mgr::fill(s, 255);
        s.has_key = true;
        s.key = 255;
        mgr::circle(s, mgr::rgb(250, 204, 153), c, c, radius, 0, legacy);
        double lx, ly, rx, ry;
        mgr::vec_rotate((double)(lx0 - c), (double)(ly0 - c), (double)ang, &lx, &ly);
        mgr::vec_rotate((double)(rx0 - c), (double)(ry0 - c), (double)ang, &rx, &ry);
        int lpx = (int)(lx + c), lpy = (int)(ly + c), rpx = (int)(rx + c), rpy = (int)(ry + c);
        mgr::circle(s, mgr::rgb(250, 250, 250), lpx, lpy, hand_r, 0, legacy);
        mgr::circle(s, mgr::rgb(250, 250, 250), rpx, rpy, hand_r, 0, legacy);
        mgr::circle(s, mgr::rgb(50, 50, 50), lpx, lpy, hand_r, outline, legacy);
        mgr::circle(s, mgr::rgb(50, 50, 50), rpx, rpy, hand_r, outline, legacy);
        if (!b.add_sprite(MG_SPR_AGENT0 + k, s)) return false;
        if (with_dense && !b.add_dense(MG_SPR_AGENT0 + k, s)) return false;
    }
    b.hdr.agent_dim = dim;
    b.hdr.agent_radius = radius;
    return true;
}

// pygame_assets.py:254-304 (Command.__init__), always built at SCALE
bool build_icons(Builder& b, bool legacy) {
    static const int angles[9] = {0, 270, 180, 90, -1, 315, 45, 225, 135};  // COMMANDS order; -1 = stay
    double dim = 88 * SCALE;
    int lw = (int)(8 * SCALE);
    uint32_t white = mgr::rgb(255, 255, 255);
    for (int k = 0; k < 9; ++k) {
        Surface s((int)dim, (int)dim);
        s.has_key = true;
        s.key = 0;
        if (angles[k] < 0) {
            double radius = floor(dim / 2) - 4 * SCALE;
            double x = dim - 12 * SCALE, y = floor(dim / 2) - 8 * SCALE;
            mgr::circle(s, white, (int)radius, (int)radius, (int)radius, lw, legacy);
            mgr::line(s, white, 0, (int)y, (int)x, (int)y, lw);
            if (!b.add_sprite(MG_SPR_ICON0 + k, s)) return false;
        } else {
            double x1 = 2 * SCALE, x2 = 80 * SCALE, y1 = 40 * SCALE, y2 = 0;
            mgr::line(s, white, (int)x1, (int)y1, (int)x2, (int)y1, lw);
            mgr::line(s, white, (int)x2, (int)y1, (int)y1, (int)y2, lw);
            mgr::line(s, white, (int)x2, (int)y1, (int)y1, (int)x2, lw);
            Surface r = mgr::rotate(s, (double)angles[k]);
            if (!b.add_sprite(MG_SPR_ICON0 + k, r)) return false;
        }
    }
    // mortar_mayhem.py:363: blit at screen_dim // 2 - command.rect_dim // 2 (un-rotated size)
    b.hdr.icon_x = b.hdr.icon_y = (int)(MG_W / 2 - floor(dim / 2));
    return true;
}

// the ground-truth ring of _build_debug_surface (mortar_mayhem.py:135): pygame.draw.circle(surface, (0, 255, 0), tile
// centre, tile_dim // 2, int(8 * SCALE)); stored as a sprite whose centre pixel is (w >> 1, h >> 1)
bool build_ring(Builder& b, bool legacy) {
    const int r = (int)(56 * SCALE) / 2, w = (int)(8 * SCALE), dim = 2 * r + 2;
    Surface s(dim, dim);
    s.has_key = true;
    s.key = 0;
    mgr::circle(s, mgr::rgb(0, 255, 0), dim >> 1, dim >> 1, r, w, legacy);
    return b.add_sprite(MG_SPR_RING, s);
}

// mystery_path.py:175-182
bool build_cross(Builder& b) {
    double dim = 40 * SCALE;
    Surface s((int)dim, (int)dim);
    s.has_key = true;
    s.key = 0;
    int d = (int)(dim - 1), lw = (int)(12 * SCALE);
    mgr::line(s, mgr::rgb(255, 0, 0), 0, 0, d, d, lw);
    mgr::line(s, mgr::rgb(255, 0, 0), d, 0, 0, d, lw);
    return b.add_sprite(MG_SPR_CROSS, s);
}

// MortarArena drawn onto the 84x84 screen: pygame_assets.py:306-418 + mortar_mayhem.py:93-103
void build_arena(Builder& b, int slot, int size, bool red) {
    Surface scr(MG_W, MG_W);
    int tile = (int)(56 * SCALE);
    int dim = (int)(56 * SCALE * size);
    int origin = MG_W / 2 - (dim >> 1);
    auto unpack = [](uint32_t c) { return mgr::rgb(c & 255, (c >> 8) & 255, (c >> 16) & 255); };
    uint32_t fillc = unpack(red ? MG_RGB_TILE_FILL_ON : MG_RGB_TILE_FILL);
    uint32_t edge = unpack(red ? MG_RGB_TILE_EDGE_ON : MG_RGB_TILE_EDGE);
    Surface t(tile, tile);
    mgr::rect(t, fillc, 0, 0, tile, tile, 0, 0, -1, -1, -1, -1);
    mgr::rect(t, edge, 0, 0, tile, tile, (int)(4 * SCALE), 0, -1, -1, -1, -1);
    for (int i = 0; i < size; ++i)
        for (int j = 0; j < size; ++j)
            for (int y = 0; y < tile; ++y)
                for (int x = 0; x < tile; ++x) {
                    int X = origin + tile * i + x, Y = origin + tile * j + y;
                    if (X >= 0 && X < MG_W && Y >= 0 && Y < MG_W) scr.at(X, Y) = t.at(x, y);
                }
    b.add_template(slot, scr);
}

}  // namespace

extern "C" long mg_mm_build_assets(const mg_mm_params* p, int thick_circle_legacy, void* host_out, size_t capacity) {
    if (!p) return MG_EINVAL;
    int size = p->variant == MG_MM_ENDLESS ? 6 : p->arena_size;
    if (size < 2 || size > 6) return MG_EINVAL;
    Builder b;
    // GridCharacterController is built with SCALE, ignoring agent_scale (mortar_mayhem_grid.py:249)
    double agent_scale = (p->variant == MG_MM_GRID || p->variant == MG_MM_B_GRID) ? SCALE : p->agent_scale;
    if (!build_agent(b, agent_scale, thick_circle_legacy != 0)) return MG_EINVAL;
    if (!build_icons(b, thick_circle_legacy != 0)) return MG_EINVAL;
    if (!build_ring(b, thick_circle_legacy != 0)) return MG_EINVAL;
    build_arena(b, 0, size, false);
    build_arena(b, 1, size, true);
    return b.serialize(host_out, capacity);
}

extern "C" long mg_mp_build_assets(const mg_mp_params* p, void* host_out, size_t capacity) {
    if (!p) return MG_EINVAL;
    Builder b;
    // GridCharacterController(SCALE, ...) ignores agent_scale (mystery_path_grid.py:188)
    if (!build_agent(b, p->variant == MG_MP_GRID ? SCALE : p->agent_scale, false)) return MG_EINVAL;
    if (!build_cross(b)) return MG_EINVAL;
    Surface black(MG_W, MG_W);
    b.add_template(0, black);
    return b.serialize(host_out, capacity);
}

// Endless Mystery Path: the agent sprites + the fall-off cross (endless_mystery_path.py:203-211) on black
extern "C" long mg_ep_build_assets(const mg_ep_params* p, void* host_out, size_t capacity) {
    if (!p) return MG_EINVAL;
    Builder b;
    if (!build_agent(b, p->agent_scale, false)) return MG_EINVAL;
    if (!build_cross(b)) return MG_EINVAL;
    Surface black(MG_W, MG_W);
    b.add_template(0, black);
    return b.serialize(host_out, capacity);
}

namespace {
// Exit.__init__ / Exit.draw (pygame_assets.py:169-205): rounded-top door, 1-px (int(2*scale)) black outline
bool build_exit(Builder& b, double scale) {
    int dim = (int)(20 * scale);
    if (dim < 2 || dim > 255) return false;
    int corner = (int)(10 * scale);
    for (int open = 0; open < 2; ++open) {
        Surface s(dim, dim);
        mgr::fill(s, 255);
        s.has_key = true;
        s.key = 255;
        uint32_t col = open ? mgr::rgb(48, 141, 70) : mgr::rgb(55, 55, 55);
        mgr::rect(s, col, 0, 0, s.w, s.h, 0, 0, corner, corner, 0, 0);
        mgr::rect(s, 0, 0, 0, s.w, s.h, (int)(2 * scale), 0, corner, corner, 0, 0);
        int slot = open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
        if (!b.add_sprite(slot, s) || !b.add_dense(slot, s)) return false;
    }
    return true;
}
}  // namespace

namespace {
// Coin.draw (pygame_assets.py:145-152): yellow disc, then an orange circle of width int(2 * scale) (0 = filled)
bool build_coin(Builder& b, double scale) {
    int r = (int)(10 * scale), w = (int)(2 * scale);
    if (r < 0 || r > MG_MAX_DISC_R) return false;
    int c = r + 1;
    Surface s(2 * c, 2 * c);
    mgr::fill(s, 255);
    s.has_key = true;
    s.key = 255;
    mgr::circle(s, mgr::rgb(255, 255, 0), c, c, r, 0, false);
    mgr::circle(s, mgr::rgb(255, 165, 0), c, c, r, w, false);
    return b.add_sprite(MG_SPR_COIN, s) && b.add_dense(MG_SPR_COIN, s);
}
}  // namespace

extern "C" long mg_ss_build_assets(const mg_ss_params* p, void* host_out, size_t capacity) {
    if (!p) return MG_EINVAL;
    Builder b;
    if (!build_agent(b, p->agent_scale, false, true)) return MG_EINVAL;
    if (p->variant != MG_SS_ENDLESS && !build_exit(b, p->exit_scale)) return MG_EINVAL;
    if (!build_coin(b, p->coin_scale)) return MG_EINVAL;
    b.add_disc_tables();
    if (!b.discs_ok) return MG_EINVAL;
    b.add_trig_table();
    if (p->black_background) b.add_ring_tables();
    Surface black(MG_W, MG_W);
    b.add_template(0, black);
    return b.serialize(host_out, capacity);
}
