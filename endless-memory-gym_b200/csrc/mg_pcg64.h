// mg_pcg64.h — numpy.random.Generator(PCG64) draw semantics, host + device.
//
// Replaces the third-party RNG behind every `self.np_random.*` call of the reference
// (gymnasium 0.29 `Env.np_random` = numpy `Generator(PCG64(SeedSequence(seed)))`), e.g.
// memory_gym/mortar_mayhem.py:174,238,247-248,262-263. Semantics restated from numpy's published
// `pcg64.h` (setseq 128, XSL-RR output), `_bounded_integers.pyx` (buffered 32-bit Lemire
// rejection for ranges < 2^32, zero-width ranges draw nothing) and `distributions.c`
// (`random_standard_uniform` = (next64 >> 11) * 2^-53).
#pragma once
#include "../../include/memgym_b200.h"
#include "mg_common.h"

typedef unsigned __int128 mg_u128;

MG_HD mg_u128 mg_pcg_mult() {
    return ((mg_u128)0x2360ED051FC65DA4ull << 64) | (mg_u128)0x4385DF649FCCF645ull;
}

MG_HD_NOINLINE uint64_t mg_next64(MgRng& r) {
    mg_u128 s = ((mg_u128)r.state_hi << 64) | r.state_lo;
    mg_u128 inc = ((mg_u128)r.inc_hi << 64) | r.inc_lo;
    s = s * mg_pcg_mult() + inc;
    r.state_lo = (uint64_t)s;
    r.state_hi = (uint64_t)(s >> 64);
    uint64_t x = r.state_hi ^ r.state_lo;
    unsigned rot = (unsigned)(r.state_hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

MG_HD uint32_t mg_next32(MgRng& r) {
    if (r.has_buf) {
        r.has_buf = 0;
        return r.buf;
    }
    uint64_t v = mg_next64(r);
    r.has_buf = 1;
    r.buf = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// Generator.integers(lo, hi) for int64 bounds with hi - lo <= 2^32 (all call sites of the reference)
MG_HD_NOINLINE int64_t mg_integers(MgRng& r, int64_t lo, int64_t hi) {
    uint64_t rng = (uint64_t)(hi - 1 - lo);
    if (rng == 0) return lo;
    if (rng == 0xFFFFFFFFull) return lo + (int64_t)mg_next32(r);
    uint32_t excl = (uint32_t)rng + 1u;
    uint64_t m = (uint64_t)mg_next32(r) * excl;
    uint32_t left = (uint32_t)m;
    if (left < excl) {
        uint32_t thr = (0xFFFFFFFFu - (uint32_t)rng) % excl;
        while (left < thr) {
            m = (uint64_t)mg_next32(r) * excl;
            left = (uint32_t)m;
        }
    }
    return lo + (int64_t)(m >> 32);
}

// Generator.uniform(a, b) = a + (b - a) * next_double; no FMA contraction (numpy computes in two steps)
MG_HD double mg_uniform(MgRng& r, double a, double b) {
    double u = (double)(mg_next64(r) >> 11) * (1.0 / 9007199254740992.0);
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, __dmul_rn(b - a, u));
#else
    volatile double prod = (b - a) * u;
    return a + prod;
#endif
}

// ---- register-resident copy of the stream ---------------------------------------------------------
// For code that EVERY lane of a warp runs redundantly (the cooperative resets: each lane holds the same state,
// makes the same draws, lane 0 writes the stream back). Fully inlined: the state never leaves registers.
struct MgRngR {
    uint64_t lo, hi, ilo, ihi;
    uint32_t buf, has;
};
MG_HD MgRngR mg_rr_load(const MgRng& r) { return MgRngR{r.state_lo, r.state_hi, r.inc_lo, r.inc_hi, r.buf, r.has_buf}; }
MG_HD void mg_rr_store(const MgRngR& q, MgRng& r) {
    r.state_lo = q.lo, r.state_hi = q.hi, r.buf = q.buf, r.has_buf = q.has;
}
MG_HD uint64_t mg_rr_next64(MgRngR& r) {
    mg_u128 s = (((mg_u128)r.hi << 64) | r.lo) * mg_pcg_mult() + (((mg_u128)r.ihi << 64) | r.ilo);
    r.lo = (uint64_t)s;
    r.hi = (uint64_t)(s >> 64);
    uint64_t x = r.hi ^ r.lo;
    unsigned rot = (unsigned)(r.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
MG_HD uint32_t mg_rr_next32(MgRngR& r) {
    if (r.has) {
        r.has = 0;
        return r.buf;
    }
    uint64_t v = mg_rr_next64(r);
    r.has = 1;
    r.buf = (uint32_t)(v >> 32);
    return (uint32_t)v;
}
// Generator.integers(lo, hi), 0 < hi - lo < 2^32 (same draw semantics as mg_integers). Everything stays inline:
// an out-of-line rejection path would take the stream's address and force it into local memory.
MG_HD int mg_rr_integers(MgRngR& r, int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - 1 - lo);
    if (rng == 0) return lo;
    const uint32_t excl = rng + 1u;
    uint64_t m = (uint64_t)mg_rr_next32(r) * excl;
    if ((uint32_t)m < excl) {  // Lemire rejection: probability < range / 2^32
        const uint32_t thr = (0xFFFFFFFFu - rng) % excl;
        while ((uint32_t)m < thr) m = (uint64_t)mg_rr_next32(r) * excl;
    }
    return lo + (int)(m >> 32);
}
MG_HD int mg_rr_choice(MgRngR& r, const mg_choice_list& c) { return c.v[mg_rr_integers(r, 0, c.n)]; }
