/* SPH.h -- shim with the public interface of the reference's SPH class (src/SPH.h:33-447); every method forwards to
 * the C ABI of libpbf_b200 (include/pbf_c.h).  Drop it in place of the reference's src/SPH.{h,cpp}, RadixSort.{h,cpp},
 * NeighbourCellFinder.{h,cpp} and the shaders/sph, shaders/radixsort, shaders/neighbourcellfinder directories.
 *
 * With PBF_WITH_GL the constructor creates the position / velocity / highlight GL buffers exactly as src/SPH.cpp:96-133
 * does and hands their names to pbf_register_gl_buffers; from then on every pbf_step maps them, steps on the mapped
 * memory and unmaps, so the renderer keeps reading the buffers it got from GetPositionBuffer() (src/Simulation.cpp:78-79,
 * 473-479).  The shim itself needs no CUDA header.
 * Headless, the handle owns the buffers and GetPositionDevice() etc. expose the device pointers. */
#ifndef PBF_SHIM_SPH_H
#define PBF_SHIM_SPH_H

#include <iostream>

#include "NeighbourCellFinder.h"
#include "RadixSort.h"
#include "shim_common.h"

class SPH {
public:
    SPH(const GLuint &_numparticles, const glm::ivec3 &gridsize = glm::ivec3(128, 64, 128))
        : numparticles(_numparticles), radixsort(512, _numparticles >> 9, gridsize),
          neighbourcellfinder(_numparticles, gridsize), handle(nullptr) {
        pbf_config cfg;
        cfg.num_particles = numparticles;
        cfg.capacity = 0;
        cfg.grid[0] = gridsize.x; cfg.grid[1] = gridsize.y; cfg.grid[2] = gridsize.z;
        cfg.wall[0] = 16.0f; cfg.wall[1] = 0.0f; cfg.wall[2] = 16.0f;   /* shaders/sph/updatepos.glsl:98 */
        cfg.ref_quirks = 1;
        cfg.device = -1;
        cfg.use_graph = 1;
        pbf_shim::check(pbf_create(&cfg, &handle), "SPH::SPH");
        pbf_shim::check(pbf_get_params(handle, &params), "SPH::SPH");   /* defaults of src/SPH.cpp:26, :137-144 */
        radixsort.Attach(handle);
        neighbourcellfinder.Attach(handle);
#ifdef PBF_WITH_GL
        glGenBuffers(3, buffers);
        const GLsizeiptr sizes[3] = {GLsizeiptr(4 * sizeof(float)) * numparticles, GLsizeiptr(4 * sizeof(float)) * numparticles,
                                     GLsizeiptr(sizeof(GLuint)) * numparticles};
        for (int i = 0; i < 3; i++) {
            glBindBuffer(GL_SHADER_STORAGE_BUFFER, buffers[i]);
            glBufferData(GL_SHADER_STORAGE_BUFFER, sizes[i], NULL, GL_DYNAMIC_COPY);
            if (i == 2) glClearBufferData(GL_SHADER_STORAGE_BUFFER, GL_R32UI, GL_RED_INTEGER, GL_UNSIGNED_INT, NULL);
        }
        pbf_shim::check(pbf_register_gl_buffers(handle, buffers[0], buffers[1], buffers[2]), "SPH::SPH");
#endif
    }
    ~SPH(void) {
        pbf_destroy(handle);   /* unregisters the GL buffers first */
#ifdef PBF_WITH_GL
        glDeleteBuffers(3, buffers);
#endif
    }
    SPH(const SPH &) = delete;
    SPH &operator=(const SPH &) = delete;

#ifdef PBF_WITH_GL
    GLuint GetPositionBuffer(void) const { return buffers[0]; }
    GLuint GetVelocityBuffer(void) const { return buffers[1]; }
    GLuint GetHighlightBuffer(void) const { return buffers[2]; }
#else
    /* headless: pseudo names; the data is reached through GetPositionDevice() / pbf_download_state */
    GLuint GetPositionBuffer(void) const { return 1; }
    GLuint GetVelocityBuffer(void) const { return 2; }
    GLuint GetHighlightBuffer(void) const { return 3; }
#endif
    float *GetPositionDevice(void) const { float *p; pbf_device_buffers(handle, &p, nullptr, nullptr); return p; }
    float *GetVelocityDevice(void) const { float *p; pbf_device_buffers(handle, nullptr, &p, nullptr); return p; }
    uint32_t *GetHighlightDevice(void) const { uint32_t *p; pbf_device_buffers(handle, nullptr, nullptr, &p); return p; }
    pbf_handle GetHandle(void) const { return handle; }

    float GetRestDensity(void) const { return 1.0f / params.one_over_rho_0; }
    void SetRestDensity(const float &rho) { params.one_over_rho_0 = 1.0f / rho; UploadSPHParams(); }
    const float &GetCFMEpsilon(void) const { return params.epsilon; }
    void SetCFMEpsilon(const float &epsilon) { params.epsilon = epsilon; UploadSPHParams(); }
    const float &GetGravity(void) const { return params.gravity; }
    void SetGravity(const float &gravity) { params.gravity = gravity; UploadSPHParams(); }
    const float &GetTimestep(void) const { return params.timestep; }
    void SetTimestep(const float &timestep) { params.timestep = timestep; UploadSPHParams(); }
    static float Wpoly6(const float &r, const float &h) { return pbf_wpoly6(r, h); }
    const float &GetTensileInstabilityK(void) const { return params.tensile_instability_k; }
    void SetTensileInstabilityK(const float &k) { params.tensile_instability_k = k; UploadSPHParams(); }
    const float &GetTensileInstabilityScale(void) const { return params.tensile_instability_scale; }
    void SetTensileInstabilityScale(const float &v) { params.tensile_instability_scale = v; UploadSPHParams(); }
    const float &GetXSPHViscosity(void) const { return params.xsph_viscosity_c; }
    void SetXSPHViscosity(const float &v) { params.xsph_viscosity_c = v; UploadSPHParams(); }
    const float &GetVorticityEpsilon(void) const { return params.vorticity_epsilon; }
    void SetVorticityEpsilon(const float &epsilon) { params.vorticity_epsilon = epsilon; UploadSPHParams(); }
    GLuint GetNumSolverIterations(void) const { return (GLuint)params.num_solver_iterations; }
    void SetNumSolverIterations(const GLuint &iter) { params.num_solver_iterations = (int32_t)iter; UploadSPHParams(); }
    bool IsVorticityConfinementEnabled(void) const { return params.vorticity_confinement != 0; }
    void SetVorticityConfinementEnabled(const bool &flag) { params.vorticity_confinement = flag ? 1 : 0; UploadSPHParams(); }
    void SetExternalForce(bool state) { params.external_force = state ? 1 : 0; UploadSPHParams(); }

    /* SPH::Run (src/SPH.cpp:246-334) */
    void Run(void) { pbf_shim::check(pbf_step(handle, 1), "SPH::Run"); }   /* maps / unmaps registered GL buffers itself */
    /* opt-in corrections of reference defects (include/pbf_c.h, pbf_options); not part of the reference's interface */
    void SetDensitySelfTerm(bool on) { pbf_options o; pbf_get_options(handle, &o); o.density_self_term = on ? 1 : 0; pbf_shim::check(pbf_set_options(handle, &o), "SPH::SetDensitySelfTerm"); }
    void SetWallRestitution(float e) { pbf_options o; pbf_get_options(handle, &o); o.wall_restitution = e; pbf_shim::check(pbf_set_options(handle, &o), "SPH::SetWallRestitution"); }
    void SetFullSupportSearch(bool on) { pbf_options o; pbf_get_options(handle, &o); o.full_support = on ? 1 : 0; pbf_shim::check(pbf_set_options(handle, &o), "SPH::SetFullSupportSearch"); }
    /* verification mode: one summation order on every code path (slab runs then equal the single-domain run bit for bit) */
    void SetCanonicalOrder(bool on) { pbf_shim::check(pbf_set_canonical_order(handle, on ? 1 : 0), "SPH::SetCanonicalOrder"); }
    /* dump / resume (no counterpart in the reference): by-id buffers, parameters, step counter */
    void SaveState(const std::string &path) const { pbf_shim::check(pbf_save_state(handle, path.c_str()), "SPH::SaveState"); }
    void LoadState(const std::string &path) {
        pbf_shim::check(pbf_load_state(handle, path.c_str()), "SPH::LoadState");
        pbf_shim::check(pbf_get_params(handle, &params), "SPH::LoadState");
    }
    /* the same step through the stage entry points, in the order and with the member objects of src/SPH.cpp:246-334 */
    void RunStaged(void) {
        pbf_shim::check(pbf_predict(handle), "predictpos");
        radixsort.Run();
        neighbourcellfinder.FindNeighbourCells(radixsort.GetBuffer());
        pbf_shim::check(pbf_highlight(handle), "highlight");
        for (int it = 0; it < params.num_solver_iterations; it++) {
            pbf_shim::check(pbf_calc_lambda(handle), "calclambda");
            pbf_shim::check(pbf_update_positions(handle), "updatepos");
        }
        pbf_shim::check(pbf_finalize(handle), "update");
        if (params.vorticity_confinement) pbf_shim::check(pbf_vorticity(handle), "vorticity");
    }
    /* SPH::OutputTiming (src/SPH.cpp:218-240): same labels; timing has to be enabled first (EnableTiming) */
    void EnableTiming(bool on) { pbf_enable_timing(handle, on ? 1 : 0); }
    void OutputTiming(void) {
        float ms[5];
        if (pbf_get_timings(handle, ms) != 0) return;   /* the reference prints nothing for queries that never ran */
        std::cout << "Position prediction: " << ms[0] << " ms" << std::endl;
        std::cout << "Sorting: " << ms[1] << " ms" << std::endl;
        std::cout << "Neighbour cell search: " << ms[2] << " ms" << std::endl;
        std::cout << "Solver: " << ms[3] << " ms" << std::endl;
        std::cout << "Vorticity confinement: " << ms[4] << " ms" << std::endl;
    }

private:
    void UploadSPHParams(void) { pbf_shim::check(pbf_set_params(handle, &params), "SPH::UploadSPHParams"); }

    const GLuint numparticles;
    RadixSort radixsort;
    NeighbourCellFinder neighbourcellfinder;
    pbf_handle handle;
    pbf_params params;
#ifdef PBF_WITH_GL
    GLuint buffers[3];
#endif
};

#endif
