/* Simulation.h -- headless counterpart of the reference's Simulation (src/Simulation.h:40-315) restricted to what
 * touches the simulation path: the constructor's `sph(GetNumberOfParticles())` (src/Simulation.cpp:25-28),
 * ResetParticleBuffer (:206-273), GetNumberOfParticles (:200-204), Frame()'s `if (running) sph.Run()` (:464-465) and the
 * keys that drive SPH (S step :348-350, Space run, V vorticity :299-301, T timing :308-318, Tab reset :303-305, F external
 * force :280-282/:360-362, highlight picking :160-195 as a ray cast).  Rendering, camera and GUI text stay with the reference's own Simulation when
 * the shim SPH is dropped into its tree; this class is for headless use and the tests. */
#ifndef PBF_SHIM_SIMULATION_H
#define PBF_SHIM_SIMULATION_H

#include <vector>

#include "SPH.h"

class Simulation {
public:
    enum Key { KEY_SPACE = 32, KEY_F = 70, KEY_S = 83, KEY_T = 84, KEY_V = 86, KEY_TAB = 258 };   /* GLFW key codes */

    explicit Simulation(uint32_t seed = 12345) : sph(GetNumberOfParticles()), running(false), seed_(seed) { ResetParticleBuffer(); }

    /* must be a multiple of 512 (src/Simulation.cpp:200-204) */
    unsigned int GetNumberOfParticles(void) const { return 32 * 32 * 32 * 2; }

    /* two 32^3 blocks, the second mirrored in x and z (src/Simulation.cpp:216-246); jitter from a seeded hash */
    void ResetParticleBuffer(void) {
        const unsigned int half = 32 * 32 * 32;
        std::vector<float> pos(8 * (size_t)half), vel(8 * (size_t)half);
        const float o1[3] = {32.5f, 0.5f, 32.5f}, o2[3] = {32.5f + 63.0f, 0.5f, 32.5f + 63.0f};
        pbf_scene_dam_break(32, 32, 32, o1, 0.94f, 0, seed_, 0, pos.data(), vel.data());
        pbf_scene_dam_break(32, 32, 32, o2, 0.94f, 1, seed_, half, pos.data() + 4 * (size_t)half, vel.data() + 4 * (size_t)half);
        pbf_shim::check(pbf_upload_state(sph.GetHandle(), pos.data(), vel.data(), 2 * half), "Simulation::ResetParticleBuffer");
    }

    bool Frame(void) {
        if (running) sph.Run();
        return true;
    }
    void OnKeyDown(int key) { if (key == KEY_F) sph.SetExternalForce(true); }
    void OnKeyUp(int key) {
        switch (key) {
        case KEY_TAB: ResetParticleBuffer(); break;
        case KEY_V: sph.SetVorticityConfinementEnabled(!sph.IsVorticityConfinementEnabled()); break;
        case KEY_T: sph.OutputTiming(); break;
        case KEY_S: sph.Run(); break;
        case KEY_SPACE: running = !running; break;
        case KEY_F: sph.SetExternalForce(false); break;
        }
    }
    /* Simulation::OnMouseDown with H held (src/Simulation.cpp:160-195): pick the particle under the cursor and toggle its
     * highlight word.  The reference finds it by rendering an id buffer (Selection::GetParticle); headless, the caller
     * passes the cursor's ray in grid units and the library casts it (pbf_pick_particle, sphere radius 0.5 grid units =
     * the 0.1 render units the reference draws).  Returns the particle id or -1. */
    int OnMouseDown(const float origin[3], const float direction[3]) {
        int32_t id = -1;
        pbf_shim::check(pbf_pick_particle(sph.GetHandle(), origin, direction, 0.5f, &id), "Simulation::OnMouseDown");
        if (id >= 0) pbf_shim::check(pbf_toggle_highlight(sph.GetHandle(), (uint32_t)id), "Simulation::OnMouseDown");
        return id;
    }
    SPH &GetSPH(void) { return sph; }

private:
    SPH sph;
    bool running;
    uint32_t seed_;
};

#endif
