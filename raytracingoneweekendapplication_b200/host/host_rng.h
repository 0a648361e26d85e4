// host_rng.h — the host-side random source used while a scene is being CONSTRUCTED
// (random sphere positions, Perlin tables ...).  The reference draws these from libc
// rand() (rtweekend.h:26-29: rand() / (RAND_MAX + 1.0), RAND_MAX = 2^31-1 on glibc).
// glibc's rand() takes a global lock, so both this mirror and the oracle
// driver (which interposes rand() with rtb200::host_rand31) use this thread-local
// generator instead: same [0, 2^31) integer range, same seeding call (srand-like),
// so a scene built against either header set from the same seed is identical.
//
// Device-side sampling does NOT use this; it uses counter-based Philox (csrc/).
#ifndef RTB200_HOST_RNG_H
#define RTB200_HOST_RNG_H
#include <cstdint>

namespace rtb200 {

struct host_rng_state {
    uint64_t s = 0x853c49e6748fea9bULL;
    uint64_t inc = 0xda3e39cb94b95bdbULL;
};

inline host_rng_state& host_rng() {
    thread_local host_rng_state st;
    return st;
}

// PCG-XSH-RR 64/32 (O'Neill), top 31 bits returned.
inline int host_rand31() {
    host_rng_state& g = host_rng();
    uint64_t old = g.s;
    g.s = old * 6364136223846793005ULL + g.inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    uint32_t r = (xs >> rot) | (xs << ((32u - rot) & 31u));
    return (int)(r >> 1);
}

inline void host_srand(uint64_t seed, uint64_t stream = 0) {
    host_rng_state& g = host_rng();
    g.s = 0;
    g.inc = ((0xda3e39cb94b95bdbULL + 2 * stream) << 1u) | 1u;
    (void)host_rand31();
    g.s += seed * 0x9E3779B97F4A7C15ULL + 0x853c49e6748fea9bULL;
    (void)host_rand31();
}

}  // namespace rtb200
#endif
