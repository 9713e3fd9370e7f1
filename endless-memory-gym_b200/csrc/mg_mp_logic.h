// mg_mp_logic.h — Mystery Path: per-reset procedural path generation (walls + noisy A*) and the
// step state machine for ONE env record.
//
// Reference paths replaced:
//   MysteryPathEnv.reset/step        memory_gym/mystery_path.py:130-200, 202-276
//   MysteryPath.__init__ / Node      memory_gym/pygame_assets.py:438-481, 606-736
//   CharacterController.step         memory_gym/character_controller.py:89-146
// The order and number of RNG draws is part of the contract.
#pragma once
#include <math.h>
#include <string.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"  // kinematics shared with Mortar Mayhem
#include "mg_pcg64.h"

struct MgMPK {
    mg_mp_params p;
    int32_t radius;
    int32_t step_axis, step_diag;
    int32_t pad;
};

#define MG_MP_CELLS 7
#define MG_MP_TILE 12  // screen_dim / grid_dim (mystery_path.py:76-78)
#define MG_MP_NODES (MG_MP_CELLS * MG_MP_CELLS)

MG_HD int mg_mp_id(int i, int j) { return i * MG_MP_CELLS + j; }

// ---- warp-cooperative path generator ------------------------------------------------------------------
// MysteryPath.__init__ (pygame_assets.py:606-736) for the 32 lanes of the warp that owns the env. Every lane
// carries the same register copy of the PCG64 stream and makes the same draws (no broadcast needed). The A*
// bookkeeping lives in REGISTERS, two cells per lane (cell = lane + 32 j): a packed key (insertion order | g cost |
// cell), f, the predecessor. The two list scans of an A* iteration — "head of the open list" and "first open node
// better than the head" — are warp min-reductions over those keys; the set of open / closed / wall cells is a
// 49-bit mask every lane holds.
//
// The reference never updates g_cost when it finds a cheaper way to an open node (:712), so g and f = g + h of a
// node are fixed at insertion; the Euclidean h comes from a per-CTA table of correctly rounded square roots.
// A single-lane context (the CPU test harness) runs the same code with all 49 cells on its one lane.
#define MG_MP_SQRT_N 80  // d2 <= 6^2 + 6^2 = 72
struct MgGenCta {
    double sqrt_d2[MG_MP_SQRT_N];
};
struct MgGenMem {      // per-warp hand-over words of the cooperative control flow (mg_env_logic_coop)
    int32_t action;    // what lane 0 decided for this call
    int32_t flag;      // env-type specific: "needs a segment", "done"
};
MG_HD void mg_gen_cta_init(MgGenCta* t, int tid, int nthreads) {
    for (int i = tid; i < MG_MP_SQRT_N; i += nthreads) t->sqrt_d2[i] = sqrt((double)i);
}

// the n-th (0-based) set bit of a 32-bit mask
MG_HD int mg_nth_set_bit(uint32_t mask, int n) {
#if defined(__CUDA_ARCH__)
    return (int)__fns(mask, 0u, n + 1);
#else
    for (int b = 0; b < 32; ++b)
        if ((mask >> b) & 1u) {
            if (n == 0) return b;
            --n;
        }
    return 32;
#endif
}

// outer-ring cells in the reference's scan order (ascending i * 7 + j): q = 0..23
MG_HD int mg_mp_ring_id(int q) {
    if (q < MG_MP_CELLS) return q;
    if (q < 17) return MG_MP_CELLS * (1 + ((q - 7) >> 1)) + (((q - 7) & 1) ? MG_MP_CELLS - 1 : 0);
    return 42 + (q - 17);
}
// the cell's 8-neighbourhood as a 49-bit mask
MG_HD uint64_t mg_mp_nbr8(int i, int j) {
    const uint64_t cols = (uint64_t)(((7u << j) >> 1) & 0x7Fu);  // columns j-1 .. j+1 inside the grid
    uint64_t m = 0;
    for (int di = -1; di <= 1; ++di)
        if ((unsigned)(i + di) < (unsigned)MG_MP_CELLS) m |= cols << (MG_MP_CELLS * (i + di));
    return m & ~(1ull << mg_mp_id(i, j));
}
// the cell's 4-neighbourhood as a 49-bit mask
MG_HD uint64_t mg_mp_nbr4(int i, int j) {
    uint64_t m = 0;
    if (i < MG_MP_CELLS - 1) m |= 1ull << mg_mp_id(i + 1, j);
    if (i > 0) m |= 1ull << mg_mp_id(i - 1, j);
    if (j < MG_MP_CELLS - 1) m |= 1ull << mg_mp_id(i, j + 1);
    if (j > 0) m |= 1ull << mg_mp_id(i, j - 1);
    return m;
}

MG_HD uint32_t mg_f64_lo(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return (uint32_t)b;
}
MG_HD uint32_t mg_f64_hi(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return (uint32_t)(b >> 32);
}
MG_HD double mg_f64_make(uint32_t lo, uint32_t hi) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

#define MG_GEN_NOKEY 0xFFFFFFFFu
#define MG_GEN_NOPREV 0xFF

// Returns the path as a 49-bit mask (0 if A* ran out of open nodes: the reference raises "No valid path found",
// pygame_assets.py:724). next2 (optional, two words): 2 bits per path cell = where its successor on the path lies
// (0 east = x+1, 1 north = y-1, 2 south = y+1, 3 west = x-1). Called by every lane with identical arguments; the
// results are uniform.
template <class Ctx>
MG_HD uint64_t mg_mp_generate_coop(Ctx& c, MgRngR& r, int sx, int sy, int ex, int ey, const MgGenCta& tab,
                                   uint64_t* next2 = nullptr, uint64_t* wall_out = nullptr) {
    constexpr int n = MG_MP_CELLS, NL = Ctx::kLanes, PER = (MG_MP_NODES + NL - 1) / NL;
    // interior walls (:640-647): 16 draws in scan order
    uint64_t wall = 0;
    MG_NOUNROLL
    for (int id = mg_mp_id(1, 1); id <= mg_mp_id(n - 3, n - 3); ++id) {
        const int j = id % n;
        if (j < 1 || j > n - 3) continue;
        if (mg_rr_integers(r, 0, 100) < 33) wall |= 1ull << id;
    }
    const int start = mg_mp_id(sx, sy), end = mg_mp_id(ex, ey);
    // outer-ring wall candidates (:650-669): not the ends, not next to them, no wall in the 8-neighbourhood
    const uint64_t keep_free = (1ull << start) | (1ull << end) | mg_mp_nbr4(sx, sy) | mg_mp_nbr4(ex, ey);
    uint32_t elig = 0;
    for (int q = c.lane; q < 24; q += NL) {
        const int id = mg_mp_ring_id(q);
        if (!((keep_free >> id) & 1ull) && !(wall & mg_mp_nbr8(id / n, id % n))) elig |= 1u << q;
    }
    elig = c.reduce_or(elig);
    int nring = mg_popc32(elig);
    const int picks = mg_rr_integers(r, 0, 2) ? 8 : 4;  // choice([4, 8])
    MG_NOUNROLL
    for (int q = 0; q < picks && nring > 0; ++q) {
        const int b = mg_nth_set_bit(elig, mg_rr_integers(r, 0, nring));  // choice(outer_nodes); one candidate draws nothing
        wall |= 1ull << mg_mp_ring_id(b);
        elig &= ~(1u << b);
        nring--;
    }
    // noisy A* (:681-724). Per cell of this lane: key = insertion order << 15 | g << 6 | cell (g <= 48 * 8)
    uint32_t key[PER];
    double h[PER], f[PER];
    uint32_t prev[PER];
    MG_UNROLL
    for (int j = 0; j < PER; ++j) {
        const int cell = c.lane + NL * j;
        int d2 = 0;
        if (cell < MG_MP_NODES) {
            const int dx = cell / n - ex, dy = cell % n - ey;
            d2 = dx * dx + dy * dy;
        }
        h[j] = tab.sqrt_d2[d2];
        f[j] = 0.0;
        key[j] = cell == start ? (uint32_t)start : MG_GEN_NOKEY;  // order 0, g 0
        prev[j] = MG_GEN_NOPREV;
    }
    // A cell's key is MG_GEN_NOKEY unless it is on the open list, so the two scans need no membership test; the
    // uniform `open` mask only steers the (uniform) control flow of the relaxation.
    uint64_t open = 1ull << start, closed = 0;
    uint32_t counter = 1, head_key = (uint32_t)start;
    double fh = 0.0;  // f of open[0]
    bool need_head = false;
    uint64_t path = 0;
    int cur = start;
    uint32_t gc = 0;
    // relax the edge cur -> v (one draw per neighbour that is neither closed nor a wall)
    const auto relax = [&](int v) {
        const uint64_t vbit = 1ull << v;
        if ((closed | wall) & vbit) return;
        const uint32_t cost = gc + (uint32_t)mg_rr_integers(r, 1, 9);
        const bool is_open = (open & vbit) != 0;
        MG_UNROLL
        for (int j = 0; j < PER; ++j) {
            if (c.lane + NL * j != v) continue;  // the lane that owns cell v
            if (is_open) {
                // a cheaper way to an open node: g_cost is NOT updated (reference bug, :712), so f stays what it
                // was; only the predecessor changes
                if (cost < ((key[j] >> 6) & 511u)) prev[j] = (uint32_t)cur;
            } else {
                key[j] = (counter << 15) | (cost << 6) | (uint32_t)v;
                f[j] = (double)cost + h[j];
                prev[j] = (uint32_t)cur;
            }
        }
        if (!is_open) {
            counter++;
            open |= vbit;
        }
    };
    for (;;) {
        if (!open) return 0;
        if (need_head) {  // open[0] was taken: the open node inserted first is the new head
            uint32_t best = MG_GEN_NOKEY;
            MG_UNROLL
            for (int j = 0; j < PER; ++j) best = key[j] < best ? key[j] : best;
            head_key = c.reduce_min(best);
            const int head = (int)(head_key & 63u), hj = head / NL;
            double fl = 0.0;
            MG_UNROLL
            for (int j = 0; j < PER; ++j)
                if (j == hj) fl = f[j];
            fh = mg_f64_make(c.shfl(mg_f64_lo(fl), head % NL), c.shfl(mg_f64_hi(fl), head % NL));
            need_head = false;
        }
        // the first node of the list with f < f(open[0]), else open[0] (:685-690)
        uint32_t cand = MG_GEN_NOKEY;
        MG_UNROLL
        for (int j = 0; j < PER; ++j)
            if (f[j] < fh && key[j] < cand) cand = key[j];
        cand = c.reduce_min(cand);
        const uint32_t cur_key = cand == MG_GEN_NOKEY ? head_key : cand;
        cur = (int)(cur_key & 63u);
        gc = (cur_key >> 6) & 511u;
        if (cur_key == head_key) need_head = true;
        if (cur == end) break;
        open &= ~(1ull << cur);
        closed |= 1ull << cur;
        MG_UNROLL
        for (int j = 0; j < PER; ++j)
            if (c.lane + NL * j == cur) key[j] = MG_GEN_NOKEY;  // off the open list
        const int ci = cur / n, cj = cur - n * ci;
        // neighbours in Node.add_neighbors order (+x, -x, +y, -y), pygame_assets.py:465-472
        if (ci < n - 1) relax(cur + n);
        if (ci > 0) relax(cur - n);
        if (cj < n - 1) relax(cur + 1);
        if (cj > 0) relax(cur - 1);
    }
    // walk back from the end node: each predecessor comes from the lane that owns the cell
    for (int at = end; at != MG_GEN_NOPREV;) {
        path |= 1ull << at;
        const int aj = at / NL;
        uint32_t p = MG_GEN_NOPREV;
    MG_UNROLL
        for (int j = 0; j < PER; ++j)
            if (j == aj) p = prev[j];
        at = (int)c.shfl(p, at % NL);
    }
    if (next2) {
        uint64_t n0 = 0, n1 = 0;
    MG_UNROLL
        for (int j = 0; j < PER; ++j) {
            const int b = c.lane + NL * j;
            if (b >= MG_MP_NODES || !((path >> b) & 1ull) || prev[j] == MG_GEN_NOPREV) continue;
            const int a = (int)prev[j];  // a -> b on the path
            const int ax = a / n, ay = a % n, bx = b / n, by = b % n;
            const uint64_t code = bx == ax + 1 ? 0u : (by == ay - 1 ? 1u : (by == ay + 1 ? 2u : 3u));
            if (a >> 5)
                n1 |= code << (2 * (a & 31));
            else
                n0 |= code << (2 * (a & 31));
        }
        next2[0] = (uint64_t)c.reduce_or((uint32_t)n0) | ((uint64_t)c.reduce_or((uint32_t)(n0 >> 32)) << 32);
        next2[1] = (uint64_t)c.reduce_or((uint32_t)n1) | ((uint64_t)c.reduce_or((uint32_t)(n1 >> 32)) << 32);
    }
    if (wall_out) *wall_out = wall;  // MysteryPath.wall_nodes (uniform across the lanes)
    return path;
}

// ---- one-lane-per-env path generator ---------------------------------------------------------------------
// The same MysteryPath.__init__ for ONE thread: the helper warps of the Endless-MysteryPath rollouts run it with a
// different env on every lane (32 paths per warp-instruction instead of one: the cooperative generator above spends
// ~7000 warp-instructions per path, most of them waiting for the stream). The open list is the reference's list
// itself, in insertion order: "open[0]" is list[0], "the first open node with f < f(open[0])" a scan from list[1].
// Per-lane scratch: element i of a lane at base[i * stride] (the helper warps interleave the lanes of a warp in
// shared memory; the host tests pass stride 1).
struct MgLaneMem {
    uint16_t* info;  // per cell: g cost << 6 | predecessor (63 = none); valid for cells that were ever open
    uint8_t* list;   // open list: cells in insertion order
    int stride;      // elements between two consecutive entries of one lane
    const volatile uint32_t* stop;  // optional: stop as soon as *stop >= stop_at (the launch is ending)
    uint32_t stop_at;
    int budget;      // node expansions this call may do before it stops (< 0: no limit; the CPU tests stop mid-search)
    uint64_t not_before;  // ... but not before this time (mg_lane_now_ns; 0: no such limit)
};
// the device's nanosecond timer (the CPU harness has no notion of a launch that ends)
MG_HD uint64_t mg_lane_now_ns() {
#if defined(__CUDA_ARCH__)
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    return ~0ull;
#endif
}
MG_HD bool mg_lane_must_stop(const MgLaneMem& m) {
    return m.stop && *m.stop >= m.stop_at && (m.not_before == 0 || mg_lane_now_ns() >= m.not_before);
}
#define MG_LANE_NOPREV 63u
// the cells within one step (8-neighbourhood) of a set of cells, the set included
MG_HD uint64_t mg_mp_dilate8(uint64_t w) {
    const uint64_t col0 = 0x0040810204081ull;  // j == 0 of every row i (bits 7 i)
    const uint64_t full = (1ull << MG_MP_NODES) - 1ull;
    uint64_t d = w | ((w << 1) & ~col0) | ((w >> 1) & ~(col0 << (MG_MP_CELLS - 1)));
    d |= (d << MG_MP_CELLS) | (d >> MG_MP_CELLS);
    return d & full;
}
// the lane's scratch <-> 19 packed words (uint16 info[49], then uint8 list[49]): a parked search moves through
// registers so that its global loads / stores are independent of each other
#define MG_LANE_BLOB_WORDS 19
MG_HD void mg_lane_pack(const MgLaneMem& m, uint64_t* blob) {
    const int S = m.stride;
    MG_UNROLL
    for (int w = 0; w < 12; ++w)
        blob[w] = (uint64_t)m.info[(4 * w) * S] | ((uint64_t)m.info[(4 * w + 1) * S] << 16) |
                  ((uint64_t)m.info[(4 * w + 2) * S] << 32) | ((uint64_t)m.info[(4 * w + 3) * S] << 48);
    uint64_t v = (uint64_t)m.info[48 * S];
    MG_UNROLL
    for (int k = 0; k < 6; ++k) v |= (uint64_t)m.list[k * S] << (16 + 8 * k);
    blob[12] = v;
    MG_UNROLL
    for (int w = 13; w < MG_LANE_BLOB_WORDS; ++w) {
        v = 0;
        MG_UNROLL
        for (int k = 0; k < 8; ++k) {
            const int i = 6 + 8 * (w - 13) + k;
            if (i < MG_MP_NODES) v |= (uint64_t)m.list[i * S] << (8 * k);
        }
        blob[w] = v;
    }
}
MG_HD void mg_lane_unpack(const uint64_t* blob, const MgLaneMem& m) {
    const int S = m.stride;
    MG_UNROLL
    for (int w = 0; w < 12; ++w) {
        MG_UNROLL
        for (int k = 0; k < 4; ++k) m.info[(4 * w + k) * S] = (uint16_t)(blob[w] >> (16 * k));
    }
    m.info[48 * S] = (uint16_t)blob[12];
    MG_UNROLL
    for (int k = 0; k < 6; ++k) m.list[k * S] = (uint8_t)(blob[12] >> (16 + 8 * k));
    MG_UNROLL
    for (int w = 13; w < MG_LANE_BLOB_WORDS; ++w) {
        MG_UNROLL
        for (int k = 0; k < 8; ++k) {
            const int i = 6 + 8 * (w - 13) + k;
            if (i < MG_MP_NODES) m.list[i * S] = (uint8_t)(blob[w] >> (8 * k));
        }
    }
}

// The search in progress, minus the per-lane scratch: what a stopped lane saves and a resumed lane reloads
struct MgLaneSearch {
    uint64_t wall, open, closed;
    int32_t nopen;
};
// Returns the path mask (0: no path, or stopped early — `stopped` tells which); next2 as in mg_mp_generate_coop.
// resume: `st` and the scratch hold a search stopped earlier (same start / end, the stream `r` as it was then);
// when it stops again they hold what the next call needs.
MG_HD uint64_t mg_mp_generate_lane(MgRngR& r, int sx, int sy, int ex, int ey, const double* sqrt_d2, const MgLaneMem& m,
                                   uint64_t* next2, bool& stopped, MgLaneSearch& st, bool resume = false) {
    constexpr int n = MG_MP_CELLS;
    stopped = false;
    const int S = m.stride;
    const int start = mg_mp_id(sx, sy), end = mg_mp_id(ex, ey);
    uint64_t wall, open, closed;
    int nopen;
    if (resume) {
        wall = st.wall, open = st.open, closed = st.closed, nopen = st.nopen;
    } else {
        wall = 0;
        MG_NOUNROLL
        for (int id = mg_mp_id(1, 1); id <= mg_mp_id(n - 3, n - 3); ++id) {
            const int j = id % n;
            if (j < 1 || j > n - 3) continue;
            if (mg_rr_integers(r, 0, 100) < 33) wall |= 1ull << id;
        }
        const uint64_t bad = (1ull << start) | (1ull << end) | mg_mp_nbr4(sx, sy) | mg_mp_nbr4(ex, ey) | mg_mp_dilate8(wall);
        // the 24 ring cells in scan order: row 0, the two ends of rows 1..5, row 6
        uint32_t taken = (uint32_t)(bad & 0x7Full) | ((uint32_t)((bad >> 42) & 0x7Full) << 17);
        MG_UNROLL
        for (int i = 1; i <= n - 2; ++i) {
            taken |= (uint32_t)((bad >> (n * i)) & 1ull) << (7 + 2 * (i - 1));
            taken |= (uint32_t)((bad >> (n * i + n - 1)) & 1ull) << (8 + 2 * (i - 1));
        }
        uint32_t elig = ~taken & 0xFFFFFFu;
        int nring = mg_popc32(elig);
        const int picks = mg_rr_integers(r, 0, 2) ? 8 : 4;
        MG_NOUNROLL
        for (int q = 0; q < picks && nring > 0; ++q) {
            const int b = mg_nth_set_bit(elig, mg_rr_integers(r, 0, nring));
            wall |= 1ull << mg_mp_ring_id(b);
            elig &= ~(1u << b);
            nring--;
        }
        open = 1ull << start, closed = 0;
        nopen = 1;
        m.list[0] = (uint8_t)start;
        m.info[start * S] = (uint16_t)MG_LANE_NOPREV;  // g = 0
    }
    const auto f_of = [&](int cell, uint32_t g) {
        const int dx = cell / n - ex, dy = cell % n - ey;
        return (double)g + sqrt_d2[dx * dx + dy * dy];
    };
    int cur = start, budget = m.budget;
    MG_NOUNROLL
    for (;;) {
        if (nopen == 0) return 0;
        if (mg_lane_must_stop(m) || budget-- == 0) {
            st.wall = wall, st.open = open, st.closed = closed, st.nopen = nopen;
            stopped = true;
            return 0;
        }
        const int head = m.list[0];
        const double fh = f_of(head, (uint32_t)m.info[head * S] >> 6);
        int pick = 0;
        MG_NOUNROLL
        for (int i = 1; i < nopen; ++i) {
            const int c = m.list[i * S];
            if (f_of(c, (uint32_t)m.info[c * S] >> 6) < fh) {
                pick = i;
                break;
            }
        }
        cur = m.list[pick * S];
        if (cur == end) break;
        const uint32_t gc = (uint32_t)m.info[cur * S] >> 6;
        nopen--;
        MG_NOUNROLL
        for (int i = pick; i < nopen; ++i) m.list[i * S] = m.list[(i + 1) * S];  // list.remove keeps the order
        open &= ~(1ull << cur);
        closed |= 1ull << cur;
        const int ci = cur / n, cj = cur - n * ci;
        // neighbours in Node.add_neighbors order (+x, -x, +y, -y), pygame_assets.py:465-472
        MG_NOUNROLL
        for (int d = 0; d < 4; ++d) {
            int v;
            if (d == 0) {
                if (ci >= n - 1) continue;
                v = cur + n;
            } else if (d == 1) {
                if (ci <= 0) continue;
                v = cur - n;
            } else if (d == 2) {
                if (cj >= n - 1) continue;
                v = cur + 1;
            } else {
                if (cj <= 0) continue;
                v = cur - 1;
            }
            const uint64_t vbit = 1ull << v;
            if ((closed | wall) & vbit) continue;
            const uint32_t cost = gc + (uint32_t)mg_rr_integers(r, 1, 9);
            if (open & vbit) {
                // a cheaper way to an open node: only the predecessor changes (g_cost is not updated, :712)
                const uint32_t old = m.info[v * S];
                if (cost < (old >> 6)) m.info[v * S] = (uint16_t)((old & ~63u) | (uint32_t)cur);
            } else {
                m.info[v * S] = (uint16_t)((cost << 6) | (uint32_t)cur);
                m.list[nopen * S] = (uint8_t)v;
                nopen++;
                open |= vbit;
            }
        }
    }
    uint64_t path = 0, n0 = 0, n1 = 0;
    MG_NOUNROLL
    for (int at = end;;) {
        path |= 1ull << at;
        const uint32_t a = (uint32_t)m.info[at * S] & 63u;
        if (a == MG_LANE_NOPREV) break;
        const int ax = (int)a / n, ay = (int)a % n, bx = at / n, by = at % n;  // a -> at on the path
        const uint64_t code = bx == ax + 1 ? 0u : (by == ay - 1 ? 1u : (by == ay + 1 ? 2u : 3u));
        if (a >> 5)
            n1 |= code << (2 * (a & 31));
        else
            n0 |= code << (2 * (a & 31));
        at = (int)a;
    }
    if (next2) {
        next2[0] = n0;
        next2[1] = n1;
    }
    return path;
}

MG_HD void mg_mp_draw(const MgMPK& k, const MgMPState& s, int cross, MgMPDraw& d) {
    d.f.sx = s.cx;
    d.f.sy = s.cy;
    d.f.srot = s.rot;
    d.f.cross = cross;
    d.f.cross_x = s.cross_x;
    d.f.cross_y = s.cross_y;
    d.show_origin = k.p.show_origin;
    d.show_goal = k.p.show_goal;
    d.ox = s.sx;
    d.oy = s.sy;
    d.gx = s.ex;
    d.gy = s.ey;
}

// reset(): called by every lane of the warp (uniform draws on a register copy of the stream, lane 0 writes)
template <class Ctx>
MG_HD void mg_mp_reset_coop(Ctx& c, const MgMPK& k, MgMPState& s, MgMPDraw& d, MgGenMem& m, const MgGenCta& tab) {
    const int g = MG_MP_CELLS;
    MgRngR r = mg_rr_load(s.rng);
    c.sync();  // every lane has its copy of the stream before lane 0 writes the advanced one back
    const int side = mg_rr_choice(r, k.p.cardinal_origin_choice);
    const int a = mg_rr_integers(r, 0, g);
    const int b = mg_rr_integers(r, 0, g);
    int sx, sy, ex, ey;
    if (side == 0) {
        sx = 0, sy = a, ex = g - 1, ey = b;
    } else if (side == 1) {
        sx = g - 1, sy = a, ex = 0, ey = b;
    } else if (side == 2) {
        sx = a, sy = 0, ex = b, ey = g - 1;
    } else {
        sx = a, sy = g - 1, ex = b, ey = 0;
    }
    uint64_t walls = 0;
    const uint64_t mask = mg_mp_generate_coop(c, r, sx, sy, ex, ey, tab, nullptr, &walls);
    if (c.lane == 0) {
        s.wall_mask = walls;
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        s.sx = (uint8_t)sx;
        s.sy = (uint8_t)sy;
        s.ex = (uint8_t)ex;
        s.ey = (uint8_t)ey;
        s.path_mask = mask;
        s.visited = 0;
        s.rot = 0;
        s.cx = (int16_t)(sx * MG_MP_TILE + k.radius);
        s.cy = (int16_t)(sy * MG_MP_TILE + k.radius);
        s.gx = (int8_t)sx;  // Grid: normalize(start * 12 + int(25 * SCALE)) (mystery_path_grid.py:186-188)
        s.gy = (int8_t)sy;
        s.off_path = 0;
        s.num_fails = 0;
        s.cross_on = 0;
        s.cross_x = 0;
        s.cross_y = 0;
        s.live = mask ? 1 : 0;  // no path: the reference raises; here the env stays finished
        mg_rr_store(r, s.rng);
        mg_mp_draw(k, s, 0, d);
    }
    c.sync();
}

MG_HD bool mg_mp_step_env(const MgMPK& k, MgMPState& s, int a0, int a1, double& reward_out, double* info,
                          MgMPDraw& d) {
    double reward = 0.0;
    bool done = false;
    int success = 0;
    int cx = s.cx, cy = s.cy, rot = s.rot;
    if (k.p.variant == MG_MP_GRID) {
        // GridCharacterController.step (character_controller.py:177-210): 1 / 2 rotate by +-90, 3 forward;
        // cell centres 12 i + 6 (MysteryPath.to_grid, pygame_assets.py:764-778)
        int gx = s.gx, gy = s.gy;
        if (s.off_path) {  // reset_position(start) then agent.step(0) (mystery_path_grid.py:218-221)
            gx = s.sx;
            gy = s.sy;
            a0 = 0;
        }
        if (a0 == 1) rot = (rot + 2) & 7;
        if (a0 == 2) rot = (rot + 6) & 7;
        if (a0 == 3) {
            int face = rot >> 1, last = MG_MP_CELLS - 1;
            if (face == 0 && gy > 0)
                gy--;
            else if (face == 3 && gx < last)
                gx++;
            else if (face == 2 && gy < last)
                gy++;
            else if (face == 1 && gx > 0)
                gx--;
        }
        s.gx = (int8_t)gx;
        s.gy = (int8_t)gy;
        cx = MG_MP_TILE * gx + MG_MP_TILE / 2;
        cy = MG_MP_TILE * gy + MG_MP_TILE / 2;
    } else {
        if (s.off_path) {  // respawn on the origin tile, then a no-op controller step (mystery_path.py:216-220)
            cx = s.sx * MG_MP_TILE + k.radius;
            cy = s.sy * MG_MP_TILE + k.radius;
            a0 = 0;
            a1 = 0;
        }
        mg_move_clamped(cx, cy, rot, a0, a1, k.step_axis, k.step_diag, k.radius, 0, 0, MG_W, MG_W);
    }
    s.cx = (int16_t)cx;
    s.cy = (int16_t)cy;
    s.rot = (uint8_t)rot;
    int px = cx / MG_MP_TILE, py = cy / MG_MP_TILE;  // centre is clamped to [radius, 84 - radius]: non-negative
    if (px == s.ex && py == s.ey) {
        reward += k.p.reward_goal;
        done = true;
        success = 1;
    } else {
        int id = mg_mp_id(px, py);
        bool on_path = px < MG_MP_CELLS && py < MG_MP_CELLS && ((s.path_mask >> id) & 1);
        if (on_path) {
            bool is_start = (px == s.sx && py == s.sy);
            if (!((s.visited >> id) & 1) && !is_start) {
                reward += k.p.reward_path_progress;
                s.visited |= 1ull << id;
            }
            s.cross_on = 0;
            s.off_path = 0;
        } else {
            reward += k.p.reward_fall_off;
            s.num_fails++;
            if (k.p.visual_feedback) s.cross_on = 1;
            s.off_path = 1;
        }
        // fall_off_rect.center = agent centre; the cross is 10x10 -> top-left = centre - 5
        s.cross_x = (int16_t)(cx - 5);
        s.cross_y = (int16_t)(cy - 5);
    }
    reward += k.p.reward_step;
    s.t++;
    if (s.t == k.p.max_steps) done = true;
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    mg_mp_draw(k, s, s.cross_on, d);
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            info[MG_INFO_SUCCESS] = (double)success;
            info[MG_INFO_A] = (double)s.num_fails;
            info[MG_INFO_B] = 0.0;
            info[MG_INFO_GT0] = 0.0;
            info[MG_INFO_GT1] = 0.0;
            info[7] = info[8] = info[9] = 0.0;
        }
    }
    return done;
}
