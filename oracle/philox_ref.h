// philox_ref.h — TEST INFRASTRUCTURE (oracle/).  Philox4x32-10 and the dimension
// assignment of the device path (csrc/rt_device.cuh "Rng"), restated for the CPU checker so
// that the restatement and the GPU trace the SAME sample sequence: counter =
// (pixel, sample, bounce<<8 | stream, 0), key = seed; 24-bit uniforms.
//   stream 0        camera: x = jitter x, y = jitter y, z = ray time
//   stream 1..15    defocus-disk rejection attempts, two candidates per block
//   stream 16       scatter: x,y,z = cube point of random_unit_vector, w = dielectric test
//   stream 32 + k   media 4k..4k+3: one free-flight uniform each
#pragma once
#include <cstdint>

namespace oracle {

struct U4 {
    double x, y, z, w;
};

inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

struct Rng {
    uint32_t pixel = 0, sample = 0, k0 = 0, k1 = 0;
    U4 draw(uint32_t bounce, uint32_t stream) const {
        uint32_t c[4] = {pixel, sample, (bounce << 8) | stream, 0u};
        philox4x32_10(c, k0, k1);
        const double s = 1.0 / 16777216.0;
        return U4{(c[0] >> 8) * s, (c[1] >> 8) * s, (c[2] >> 8) * s, (c[3] >> 8) * s};
    }
};

constexpr uint32_t RS_CAMERA = 0, RS_DEFOCUS = 1, RS_SCATTER = 16, RS_MEDIUM = 32;

}  // namespace oracle
