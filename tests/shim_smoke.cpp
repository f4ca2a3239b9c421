// Drives the C++ shim classes (include/pbf/*.h) the way the reference's main loop drives Simulation, headless, and
// prints a checksum of the particle state after a few frames and writes the state itself to argv[2] (both compared with the
// Python path by tests/test_shim.py).
#include <cstdio>
#include <vector>

#include "pbf/Simulation.h"

int main(int argc, char **argv) {
    int frames = argc > 1 ? atoi(argv[1]) : 3;
    try {
        Simulation sim(12345);
        SPH &sph = sim.GetSPH();
        sph.SetNumSolverIterations(3);
        sim.OnKeyUp(Simulation::KEY_V);       // vorticity on
        sim.OnKeyUp(Simulation::KEY_SPACE);   // running
        for (int f = 0; f < frames; f++) sim.Frame();
        sph.RunStaged();                      // one more step through RadixSort / NeighbourCellFinder shims
        const float eye[3] = {64.0f, 10.0f, -20.0f}, dir[3] = {-0.3f, 0.0f, 1.0f};
        const int picked = sim.OnMouseDown(eye, dir);   // highlight toggle along a ray into the first block
        const int again = sim.OnMouseDown(eye, dir);    // the same particle again: toggled back
        printf("SHIM pick=%d again=%d\n", picked >= 0, picked == again);
        unsigned n = sim.GetNumberOfParticles();
        std::vector<float> pos(4 * (size_t)n), vel(4 * (size_t)n);
        pbf_shim::check(pbf_download_state(sph.GetHandle(), pos.data(), vel.data(), nullptr), "download");
        double sp = 0, sv = 0;
        for (size_t i = 0; i < pos.size(); i++) { sp += pos[i]; sv += vel[i] * vel[i]; }
        printf("SHIM n=%u frames=%d sum_pos=%.6f sum_v2=%.6f rest_density=%.3f iters=%u\n", n, frames + 1, sp, sv,
               sph.GetRestDensity(), sph.GetNumSolverIterations());
        if (argc > 2) {                       // the whole state, for a particle-by-particle comparison (tests/test_shim.py)
            FILE *f = fopen(argv[2], "wb");
            if (!f || fwrite(pos.data(), sizeof(float), pos.size(), f) != pos.size() ||
                fwrite(vel.data(), sizeof(float), vel.size(), f) != vel.size() || fclose(f) != 0) {
                printf("SHIM_ERROR cannot write %s\n", argv[2]);
                return 3;
            }
        }
        try {
            SPH bad(1000);                    // not a multiple of 512: must throw like the reference's constructors do
            printf("SHIM_ERROR no exception\n");
            return 2;
        } catch (const std::runtime_error &e) {
            printf("SHIM expected error: %s\n", e.what());
        }
    } catch (const std::exception &e) {
        printf("SHIM_ERROR %s\n", e.what());
        return 1;
    }
    return 0;
}
