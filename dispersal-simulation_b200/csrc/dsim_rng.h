/*
 * dsim_rng.h — the keyed random stream of the B200 step loop (host + device).
 *
 * Replaces the unseeded thread-local `fastrand` draws of the reference
 * (lib.rs:859,1271,1278-1279,1296,1310,1318,1326,1334).  Every draw is
 *     Philox4x32-10(key = (seed_lo, seed_hi), counter = (step, entity_lo, entity_hi, tag<<24 | sub))
 * so results do not depend on thread, launch geometry or GPU count.
 */
#ifndef DSIM_RNG_H
#define DSIM_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define DSIM_HD __host__ __device__ __forceinline__
#else
#define DSIM_HD inline
#endif

enum dsim_purpose : uint32_t {
    DSIM_P_MEM = 1,     /* remembered-patch sampling, lib.rs:859 (top byte of the draw); entity = family id */
    DSIM_P_NOISE = 2,   /* resource_uncertainty, lib.rs:1317; entity = family id               */
    DSIM_P_SHUF = 3,    /* shuffle key, lib.rs:1333; entity = family id                        */
    DSIM_P_FIGHT = 4,   /* fight_is_deadly x2, lib.rs:1219,1226; entity = newcomer family id   */
    DSIM_P_PIVOT = 5,   /* select_pivot_family, lib.rs:1295; entity = node id                  */
    DSIM_P_MUT = 6,     /* mutation clock and bit, lib.rs:1270-1285; entity = family id        */
    DSIM_P_REGROW = 7,  /* recover_resources_proportion, lib.rs:1309; entity = node id         */
    DSIM_P_MEM2 = 8,    /* low 32 bits of the remembered-patch draw, lib.rs:859; entity = family id */
    DSIM_P_MEMT = 9,    /* skip sampling over the tail of the history, lib.rs:859; entity = family id */
    /* workload generation only (dsim_synth) */
    DSIM_P_SYN_EDGE = 32, DSIM_P_SYN_POPCAP = 33, DSIM_P_SYN_FAMILY = 34, DSIM_P_SYN_CULTURE = 35,
    DSIM_P_SYN_HIST = 36, DSIM_P_SYN_SITE = 37
};

struct dsim_u4 { uint32_t x, y, z, w; };

DSIM_HD uint32_t dsim_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

DSIM_HD dsim_u4 dsim_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = dsim_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = dsim_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    dsim_u4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

struct dsim_stream { uint32_t k0, k1; };

DSIM_HD dsim_u4 dsim_draw(dsim_stream s, uint32_t step, uint64_t entity, uint32_t tag, uint32_t sub) {
    return dsim_philox(s.k0, s.k1, step, (uint32_t)entity, (uint32_t)(entity >> 32), (tag << 24) | (sub & 0xFFFFFFu));
}

/* uniform on [0,1) with 53 random bits: (u64 >> 11) * 2^-53 */
DSIM_HD double dsim_u53(uint32_t lo, uint32_t hi) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

DSIM_HD uint32_t dsim_word(const dsim_u4 &o, uint32_t i) { return i == 0 ? o.x : i == 1 ? o.y : i == 2 ? o.z : o.w; }

#endif
