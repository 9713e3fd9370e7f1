// mg_seed_host.cpp — numpy.random.SeedSequence(seed) -> PCG64 initial state, on the host.
//
// Replaces gymnasium.utils.seeding.np_random(seed) (gymnasium 0.29, third-party, setup.py:36), the
// call behind `super().reset(seed=seed)` (memory_gym/mortar_mayhem.py:217). Restated from numpy's
// published bit_generator.pyx (SeedSequence: hashmix/mix over a 4-word pool, generate_state) and
// pcg64.pyx / pcg64.h (pcg_setseq_128_srandom_r).
#include <stdint.h>

#include "../../include/memgym_b200.h"
#include "mg_pcg64.h"

namespace {
const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;

struct Hasher {
    uint32_t c;
    uint32_t operator()(uint32_t v) {
        v ^= c;
        c *= MULT_A;
        v *= c;
        v ^= v >> 16;
        return v;
    }
};
inline uint32_t mix(uint32_t x, uint32_t y) {
    uint32_t r = MIX_L * x - MIX_R * y;
    return r ^ (r >> 16);
}

void seed_one(uint64_t seed, uint64_t* out) {
    uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_ent = entropy[1] ? 2 : 1;  // ints are split into the minimal number of 32-bit words (0 -> [0])
    uint32_t pool[4];
    Hasher h{INIT_A};
    for (int i = 0; i < 4; ++i) pool[i] = h(i < n_ent ? entropy[i] : 0u);
    for (int s = 0; s < 4; ++s)
        for (int d = 0; d < 4; ++d)
            if (s != d) pool[d] = mix(pool[d], h(pool[s]));
    // generate_state(4, uint64) = 8 x 32-bit words
    uint32_t w[8], hc = INIT_B;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pool[i & 3];
        v ^= hc;
        hc *= MULT_B;
        v *= hc;
        v ^= v >> 16;
        w[i] = v;
    }
    uint64_t val[4];
    for (int i = 0; i < 4; ++i) val[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    // pcg64_set_seed: initstate = (val[0] << 64) | val[1]; initseq = (val[2] << 64) | val[3]
    mg_u128 initstate = ((mg_u128)val[0] << 64) | val[1];
    mg_u128 initseq = ((mg_u128)val[2] << 64) | val[3];
    mg_u128 inc = (initseq << 1) | 1u;
    mg_u128 st = 0;
    st = st * mg_pcg_mult() + inc;
    st += initstate;
    st = st * mg_pcg_mult() + inc;
    out[0] = (uint64_t)st;
    out[1] = (uint64_t)(st >> 64);
    out[2] = (uint64_t)inc;
    out[3] = (uint64_t)(inc >> 64);
}
}  // namespace

extern "C" int mg_seed_pcg64(const uint64_t* seeds, int n, uint64_t* out) {
    if (!seeds || !out || n < 0) return MG_EINVAL;
    for (int i = 0; i < n; ++i) seed_one(seeds[i], out + 4 * (size_t)i);
    return 0;
}

extern "C" int mg_selftest_pcg64_host(uint64_t* state4, uint32_t* has_buf2, const int64_t* lo, const int64_t* hi,
                                      int n_draws, int64_t* out, double* out_uniform) {
    if (!state4 || !has_buf2) return MG_EINVAL;
    MgRng r;
    r.state_lo = state4[0];
    r.state_hi = state4[1];
    r.inc_lo = state4[2];
    r.inc_hi = state4[3];
    r.has_buf = has_buf2[0];
    r.buf = has_buf2[1];
    for (int i = 0; i < n_draws; ++i) out[i] = mg_integers(r, lo[i], hi[i]);
    if (out_uniform) *out_uniform = mg_uniform(r, out_uniform[0], out_uniform[1]);
    state4[0] = r.state_lo;
    state4[1] = r.state_hi;
    state4[2] = r.inc_lo;
    state4[3] = r.inc_hi;
    has_buf2[0] = r.has_buf;
    has_buf2[1] = r.buf;
    return 0;
}
