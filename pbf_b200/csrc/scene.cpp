// scene.cpp -- seeded initial state (host side).
// Follows Simulation::ResetParticleBuffer (reference src/Simulation.cpp:206-246): blocks of particles on a 0.94
// lattice, loop order x, z, y (y innermost) so ids match the reference, +-0.005 jitter per axis, zero velocity,
// w = 0.  The reference draws the jitter from rand() seeded with time(NULL) (:215, not reproducible); here it is a
// counter-based hash of (seed, 3*id + axis) so that every implementation starts from the same state.
#include <cstdint>

#include "../../include/pbf_c.h"

static inline uint32_t hash32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

static inline float unit_float(uint32_t seed, uint32_t counter) {   // [0,1) with 24 random bits
    return (float)(hash32(seed ^ (counter * 0x9E3779B9u)) >> 8) * (1.0f / 16777216.0f);
}

extern "C" int pbf_scene_dam_break(int nx, int ny, int nz, const float origin[3], float spacing, int mirror_xz,
                                   uint32_t seed, uint32_t id0, float *pos4, float *vel4) {
    if (nx < 0 || ny < 0 || nz < 0 || !origin || !pos4) return PBF_ERR_INVALID;
    const float sx = mirror_xz ? -1.0f : 1.0f;
    uint32_t id = id0;
    float *p = pos4, *v = vel4;
    for (int x = 0; x < nx; x++)
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++, id++, p += 4) {
                // volatile stores keep the mul/add pairs unfused whatever the host compiler's contraction default is
                volatile float ax = spacing * (sx * (float)x), ay = spacing * (float)y, az = spacing * (sx * (float)z);
                volatile float jx = 0.01f * (unit_float(seed, 3u * id) - 0.5f);
                volatile float jy = 0.01f * (unit_float(seed, 3u * id + 1u) - 0.5f);
                volatile float jz = 0.01f * (unit_float(seed, 3u * id + 2u) - 0.5f);
                volatile float bx = origin[0] + ax, by = origin[1] + ay, bz = origin[2] + az;
                p[0] = bx + jx;
                p[1] = by + jy;
                p[2] = bz + jz;
                p[3] = 0.0f;
                if (v) {
                    v[0] = v[1] = v[2] = v[3] = 0.0f;
                    v += 4;
                }
            }
    return PBF_OK;
}
