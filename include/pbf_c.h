/*
 * pbf_c.h -- C ABI of libpbf_b200.so: the B200-native replacement of ekpyron/pbf's per-timestep PBF
 * simulation (SPH::Run and the classes it owns).  Plain pointers and sizes only; no C++ or torch types.
 *
 * The reference has no FFI layer: its boundary is the C++ class API of SPH (src/SPH.h:33-447),
 * RadixSort (src/RadixSort.h:33-131) and NeighbourCellFinder (src/NeighbourCellFinder.h:35-115), called
 * only by Simulation (src/Simulation.cpp).  Each entry point below cites the reference interface it
 * replaces; the headers under include/pbf/ hold shim classes with the reference's class names and method signatures
 * that forward to this ABI.
 *
 * Conventions
 *  - every function returns 0 on success and a negative pbf_status otherwise; pbf_last_error() returns a
 *    message for the calling thread.  No exception crosses the ABI (the reference throws
 *    std::runtime_error / std::logic_error, src/ShaderProgram.cpp:44-116, src/RadixSort.cpp:41-42; the
 *    shim classes rethrow).
 *  - a handle is bound to one CUDA device and one stream; calls are asynchronous on that stream unless
 *    they take or return HOST pointers (those synchronise).  Not thread safe per handle (the reference is
 *    single threaded on the GL thread, src/main.cpp:337-344).
 *  - there is no CPU fallback: every call fails with PBF_ERR_CUDA when no sm_100 device is usable.
 *  - particle state visible to the caller is indexed by persistent particle id, N x float4 {x,y,z,0}
 *    (src/SPH.cpp:119-133; consumed by src/PointSprite.cpp:64-73 at stride 16), highlight N x uint32.
 */
#ifndef PBF_C_H
#define PBF_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

typedef struct pbf_sim *pbf_handle;

typedef enum {
    PBF_OK = 0,
    PBF_ERR_INVALID = -1,   /* bad argument (e.g. N not a multiple of 512: src/Simulation.cpp:202)   */
    PBF_ERR_CUDA = -2,      /* CUDA runtime error, message in pbf_last_error                         */
    PBF_ERR_NCCL = -3,
    PBF_ERR_CAPACITY = -4,  /* slab overflow: more local + halo particles than the handle was sized for */
    PBF_ERR_STATE = -5      /* call order violated (e.g. stage entry point before pbf_predict)       */
} pbf_status;

/* Constructor arguments of SPH (src/SPH.h:40 `SPH(numparticles, gridsize = (128,64,128))`) plus what the
 * reference hard-codes: wall offsets (shaders/sph/updatepos.glsl:98 `vec3 wall = (16,0,16)`), and the
 * restatement policy switch for findcells.glsl:39-43.  h = 2.0 is a constant of the
 * kernels as in the reference (src/SPH.cpp:58). */
typedef struct {
    uint32_t num_particles;     /* local particle count at creation (multiple of 512)                 */
    uint32_t capacity;          /* max particles this handle may hold incl. halo (0 = num_particles)  */
    int32_t grid[3];            /* GRID_SIZE                                                          */
    float wall[3];              /* wall offsets                                                        */
    int32_t ref_quirks;         /* 1 = reproduce findcells thread-0 quirk (parity), 0 = corrected      */
    int32_t device;             /* CUDA device ordinal, -1 = current                                   */
    int32_t use_graph;          /* 1 = replay the step as a CUDA graph                                 */
} pbf_config;

/* sphparams_t (src/SPH.h:252-285), same order, + the non-UBO state of SPH: num_solveriterations
 * (src/SPH.h:189-199), vorticityconfinement (:213-222), extforce uniform (src/SPH.cpp:242-244). */
typedef struct {
    float one_over_rho_0;
    float epsilon;
    float gravity;
    float timestep;
    float tensile_instability_k;
    float tensile_instability_scale;
    float xsph_viscosity_c;
    float vorticity_epsilon;
    int32_t num_solver_iterations;
    int32_t vorticity_confinement;
    int32_t external_force;
} pbf_params;

const char *pbf_last_error(void);
int pbf_version(void);

/* SPH::Wpoly6 (src/SPH.cpp:159-164), host side; used for tensile_instability_scale = 1/Wpoly6(0.2, 2). */
float pbf_wpoly6(float r, float h);
/* defaults of SPH::SPH (src/SPH.cpp:26, :25, :137-144): K = 5, vorticity off, extforce off */
void pbf_default_params(pbf_params *p);
/* RadixSort's numbits/pass count (src/RadixSort.cpp:24-30, :44, :127): number of low key bits sorted */
int pbf_sort_bits(const int32_t grid[3]);
/* onesweep passes libpbf_b200 runs for that many key bits (the reference runs pbf_sort_bits / 2 two-bit passes) */
int pbf_sort_passes(const int32_t grid[3]);

/* SPH::SPH / SPH::~SPH (src/SPH.cpp:24-156) */
int pbf_create(const pbf_config *cfg, pbf_handle *out);
int pbf_destroy(pbf_handle h);

/* setters/getters of SPH (src/SPH.cpp:166-216, src/SPH.h:57-222) */
int pbf_set_params(pbf_handle h, const pbf_params *p);
int pbf_get_params(pbf_handle h, pbf_params *p);

/* Opt-in corrections of known defects of the reference; all off by default, so that the default
 * behaviour -- and every parity test -- is the reference's.
 *   density_self_term  calclambda.glsl skips j == i (foreachneighbour.glsl:9), so rho_i lacks W(0) = 0.1958, unlike the PBF
 *                      paper; 1 adds it back (rest densities tuned for the reference, e.g. rho_0 = 1 on a 0.94 lattice,
 *                      then need retuning).
 *   wall_restitution   updatepos.glsl:98-100 clamps positions to the walls and update.glsl derives the velocity from the
 *                      clamped position, which leaves a particle pressed against a wall with whatever normal velocity
 *                      the clamp implies.  >= 0: a particle on a wall moving outwards gets v_n <- -e * v_n in update
 *                      (0 = sticks, 1 = elastic).  < 0: off.
 *   full_support       the kernels have support h = 2 cells, but FOR_EACH_NEIGHBOUR only visits the 27 cells around a
 *                      particle (neighbourcells.glsl:37-47), so every sum is truncated at about one cell; 1 visits all
 *                      5 x 5 x 5 cells the support reaches (25 rows of five cells).  Several times the work per sweep; the
 *                      neighbour runs walk from global memory, not through the tiled shared-memory path. */
typedef struct {
    int32_t density_self_term;
    float wall_restitution;
    int32_t full_support;
} pbf_options;
int pbf_set_options(pbf_handle h, const pbf_options *o);
int pbf_get_options(pbf_handle h, pbf_options *o);
/* Verification mode, off by default: one order of every floating-point sum on every code path (runs walked in row order,
 * pairs counted from a run's first candidate, the particles of a cell in ascending global id order on a slab rank), so that a
 * slab decomposition (pbf_slab_*, peer-memory transport) reproduces the single-domain run BIT FOR BIT when both sides
 * switch it on -- the reference's "stable sort from id order" carried across GPU counts.  Results stay within
 * the parity tolerances of the default mode; a few per cent slower on disordered scenes. */
int pbf_set_canonical_order(pbf_handle h, int on);

/* Simulation::ResetParticleBuffer's upload (src/Simulation.cpp:249-272): HOST arrays of N float4 by id.
 * vel may be NULL (zero), highlight is cleared.  pbf_download_state copies back (any pointer may be NULL). */
int pbf_upload_state(pbf_handle h, const float *pos4, const float *vel4, uint32_t n);
int pbf_download_state(pbf_handle h, float *pos4, float *vel4, uint32_t *highlight);
/* SPH::GetPositionBuffer / GetVelocityBuffer / GetHighlightBuffer (src/SPH.h:49, :181, :205): DEVICE
 * pointers of the by-id buffers; callers may write them between steps (src/Simulation.cpp:160-195). */
int pbf_device_buffers(pbf_handle h, float **pos4, float **vel4, uint32_t **highlight);
/* Use caller-owned DEVICE memory (e.g. a CUDA-GL interop mapping of the renderer's buffers) for the by-id
 * state instead of the handle's own allocations; NULL restores the internal buffer. */
int pbf_bind_device_buffers(pbf_handle h, float *pos4, float *vel4, uint32_t *highlight);
uint32_t pbf_num_particles(pbf_handle h);
/* CUDA-GL interop: run on the renderer's own buffer objects.  Replaces SPH's ownership of positionbuffer /
 * velocitybuffer / highlightbuffer (src/SPH.cpp:96-133, names handed out by src/SPH.h:49, :181, :205 and bound as
 * vertex attributes by src/PointSprite.cpp:64-84): the three GL buffer names (N x float4, N x float4, N x uint32) are
 * registered once, with the GL context current on the calling thread; afterwards every pbf_step maps them on the
 * handle's stream, steps on the mapped memory and unmaps, so GL draws issued after pbf_step returns see the new
 * positions.  Fails with PBF_ERR_CUDA when no GL context is current.  pbf_unregister_gl_buffers (also done by
 * pbf_destroy) returns to the handle's own buffers. */
int pbf_register_gl_buffers(pbf_handle h, unsigned int pos, unsigned int vel, unsigned int highlight);
int pbf_unregister_gl_buffers(pbf_handle h);
/* The same ownership protocol for any other owner of the three by-id buffers (a Vulkan / EGL renderer importing external
 * memory, a host application with its own allocator): instead of GL names the caller registers two callbacks.  Every entry
 * point that reads or writes particle state calls map(user, stream, &pos4, &vel4, &highlight) first -- `stream` is the
 * handle's cudaStream_t; the callback returns 0 and three DEVICE pointers (N x float4, N x float4, N x uint32, 16-byte
 * aligned) that stay valid until unmap -- enqueues its work on that stream, and calls unmap(user, stream) on every way
 * out.  Nothing is touched outside a map/unmap bracket; pbf_device_buffers and pbf_step_host are refused while
 * registered.  pbf_register_gl_buffers is this protocol with cudaGraphicsMapResources / cudaGraphicsUnmapResources. */
typedef int (*pbf_map_fn)(void *user, void *stream, float **pos4, float **vel4, uint32_t **highlight);
typedef int (*pbf_unmap_fn)(void *user, void *stream);
int pbf_register_external_buffers(pbf_handle h, pbf_map_fn map, pbf_unmap_fn unmap, void *user);
int pbf_unregister_external_buffers(pbf_handle h);

/* SPH::Run (src/SPH.cpp:246-334), nsteps times. */
int pbf_step(pbf_handle h, int nsteps);
/* One step from HOST buffers and back (h2d pos+vel, step, d2h pos+vel): the end-to-end call. */
int pbf_step_host(pbf_handle h, float *pos4, float *vel4, int nsteps);
int pbf_sync(pbf_handle h);

/* Stage-level entry points, in the order SPH::Run issues them; used by the RadixSort /
 * NeighbourCellFinder shims and by the parity tests.
 *   pbf_predict          predictpos.glsl (src/SPH.cpp:247-261) + cell keys + clearhighlight
 *   pbf_sort             RadixSort::Run (src/RadixSort.cpp:124-133)
 *   pbf_build_cells      NeighbourCellFinder::FindNeighbourCells (src/NeighbourCellFinder.cpp:113-148)
 *   pbf_highlight        highlight.glsl (src/SPH.cpp:288-296)
 *   pbf_calc_lambda      calclambda.glsl (src/SPH.cpp:304-307)
 *   pbf_update_positions updatepos.glsl (src/SPH.cpp:308-310)
 *   pbf_finalize         update.glsl (src/SPH.cpp:318-324)
 *   pbf_vorticity        vorticity.glsl (src/SPH.cpp:325-331) */
int pbf_predict(pbf_handle h);
int pbf_sort(pbf_handle h);
int pbf_build_cells(pbf_handle h);
int pbf_highlight(pbf_handle h);
int pbf_calc_lambda(pbf_handle h);
int pbf_update_positions(pbf_handle h);
int pbf_finalize(pbf_handle h);
int pbf_vorticity(pbf_handle h);

/* Standalone RadixSort (src/RadixSort.h:33-131): stable sort of n (key, value) pairs on the low `bits`
 * bits of the key, DEVICE pointers, on the handle's stream (n <= capacity). */
int pbf_sort_pairs(pbf_handle h, const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                   uint32_t *vals_out, uint32_t n, int bits);

/* Debug read-back into HOST arrays (any pointer may be NULL).
 *   keys[N]     full cell key of each sorted slot (bit 31: outside the cell images)
 *   perm[N]     particle id at each sorted slot
 *   records[4N] sorted {pos, id-bits} records = RadixSort::GetBuffer (src/RadixSort.h:45-53)
 *   start/end   dense gx*gy*gz tables = gridtexture / gridendtexture (src/NeighbourCellFinder.cpp:59-90)
 *   runs        per particle 9 x {start,count} = the neighbour buffer of neighbourcells.glsl:62-90, unpacked */
int pbf_get_predicted(pbf_handle h, float *records, uint32_t *keys);
int pbf_get_sorted(pbf_handle h, uint32_t *keys, uint32_t *perm, float *records);
int pbf_get_cell_ranges(pbf_handle h, int32_t *start, int32_t *end);
int pbf_get_neighbour_runs(pbf_handle h, int32_t *run_start, int32_t *run_count);
int pbf_get_lambda(pbf_handle h, float *lambda);
int pbf_get_vorticity(pbf_handle h, float *vorticity);

/* SPH::OutputTiming (src/SPH.cpp:218-240): last step's five phases in ms: predict, sort, neighbour cells,
 * solver, vorticity.  Timing needs pbf_enable_timing(h, 1), which runs steps outside the CUDA graph. */
int pbf_enable_timing(pbf_handle h, int on);
int pbf_get_timings(pbf_handle h, float ms[5]);
/* finer than the reference's queries: mean duration of one calclambda and one updatepos launch (src/SPH.cpp:304-310)
 * over the solver iterations of the last timed step -- what bench.py's roofline is computed from (the last iteration's
 * updatepos launch, which also carries update.glsl, is left out of its mean) */
int pbf_get_solver_kernel_timings(pbf_handle h, float *lambda_ms, float *delta_p_ms);

/* Selection.  pbf_pick_particle replaces Selection::GetParticle (src/Selection.cpp:55-85), which
 * renders every particle as a sphere into an id buffer and reads the pixel under the cursor: here the id of the nearest
 * sphere (radius in grid units; the reference draws 0.1 render units = 0.5 grid units) hit by the ray origin +
 * t * direction, t >= 0, or -1.  The caller un-projects the cursor (the reference's camera lives with the renderer).
 * pbf_toggle_highlight is the highlight-word update of Simulation::OnMouseDown (src/Simulation.cpp:160-195):
 * flag > 0 -> 0, else 1, on the device, stream ordered. */
int pbf_pick_particle(pbf_handle h, const float origin[3], const float direction[3], float radius, int32_t *id);
int pbf_toggle_highlight(pbf_handle h, uint32_t id);

/* Aggregates the north star's long-run criterion needs (not in the reference): mean |rho_i/rho_0 - 1| at
 * the current positions (one extra density sweep) and sum 0.5 |v|^2. */
int pbf_get_diagnostics(pbf_handle h, double *density_error, double *kinetic_energy);

/* Tiles (128 consecutive sorted particles) of the last pbf_build_cells / step, and how many of them run the
 * shared-memory tiled sweep path; the rest walk their neighbour runs from global memory ("sweeps").
 * why (may be NULL): [0] tiled, [1] general path; [2..7] tiles whose nine ranges hold <= 9, 10, 11, 12, 13, more
 * records per particle of the tile (what sizes the shared-memory image). */
int pbf_get_tile_stats(pbf_handle h, uint32_t *tiles, uint32_t *tiled, uint32_t why[8]);

/* how many kernels the handle has launched (graph replays count their kernel nodes) */
uint64_t pbf_kernel_launches(pbf_handle h);
/* stream the handle launches on (a cudaStream_t), for callers that time with CUDA events */
void *pbf_stream(pbf_handle h);

/* Seeded restatement of Simulation::ResetParticleBuffer's block fill (src/Simulation.cpp:216-230): HOST
 * arrays, nx*ny*nz particles, loop order x,z,y, ids from id0. */
int pbf_scene_dam_break(int nx, int ny, int nz, const float origin[3], float spacing, int mirror_xz,
                        uint32_t seed, uint32_t id0, float *pos4, float *vel4);

/* ---- state files (dump / resume; no counterpart in the reference, whose state dies with the process).
 * A file holds what SPH::Run is a function of: the three by-id buffers of src/SPH.cpp:106-133, sphparams_t and the
 * solver switches (src/SPH.h:189-285), the constructor arguments and a step counter; a resumed run continues bit for
 * bit.  128-byte header + N x float4 + N x float4 + N x uint32, little endian, FNV-1a-64 checksum of the payload.
 * The pbf_state_file_* functions work on HOST arrays and need no device. */
typedef struct {
    uint32_t num_particles;
    int32_t grid[3];
    float wall[3];
    int32_t ref_quirks;
    pbf_params params;
    uint64_t steps;             /* SPH::Run calls completed when the file was written */
    pbf_options options;
} pbf_state_info;
int pbf_state_file_write(const char *path, const pbf_state_info *info, const float *pos4, const float *vel4,
                         const uint32_t *highlight);       /* vel4 / highlight may be NULL (zeros) */
int pbf_state_file_info(const char *path, pbf_state_info *info);
int pbf_state_file_read(const char *path, pbf_state_info *info, float *pos4, float *vel4, uint32_t *highlight,
                        uint32_t capacity);                /* arrays hold `capacity` particles; verifies the checksum */
/* the handle's current state -> file; file -> handle (N and grid must match the handle; parameters and the step
 * counter are taken from the file) */
int pbf_save_state(pbf_handle h, const char *path);
int pbf_load_state(pbf_handle h, const char *path);
uint64_t pbf_step_count(pbf_handle h);

/* ---- slab decomposition (multi GPU).  No counterpart in the reference (single GPU): this is
 * the north star's slab runtime.  One handle per rank owns the particles whose cell layer z lies in [z_lo, z_hi); the
 * handle must have been created with grid z = (z_hi - z_lo) + 2 (one ghost layer per face), positions stay global.
 * pbf_slab_step runs SPH::Run with particle migration and halo exchange over NCCL send/recv on the handle's stream. */
/* rank 0: create the 128-byte NCCL unique id; the host runtime broadcasts it (torch.distributed / MPI / files) */
int pbf_slab_unique_id(void *out128);
int pbf_slab_init(pbf_handle h, const void *id128, int rank, int nranks, int z_lo, int z_hi, int gz_global,
                  uint32_t halo_capacity);
/* Peer-memory halo refresh (optional, after pbf_slab_init): the per-iteration lambda / position / |omega| halos go as
 * plain stores into the neighbour's mailbox over NVLink with a release flag, instead of ncclSend/ncclRecv pairs.
 * pbf_slab_p2p_handle returns this rank's 64-byte CUDA IPC handle; the host runtime gives every rank its neighbours'
 * handles (NULL where there is none).  Migration and ghost records keep going through NCCL. */
int pbf_slab_p2p_handle(pbf_handle h, void *out64);
int pbf_slab_p2p_connect(pbf_handle h, const void *lo64, const void *hi64);
/* "virtual ranks": n handles of ONE process on one device, stepped in lock step with device copies instead of
 * NCCL (tests on a single GPU).  Rank r owns layers [z_planes[r], z_planes[r+1]). */
int pbf_slab_init_group(pbf_handle *hs, int n, const int32_t *z_planes, int gz_global, uint32_t halo_capacity);
/* local particles of the slab (HOST arrays by slot) with their global ids; download returns the current owners */
int pbf_slab_upload(pbf_handle h, const float *pos4, const float *vel4, const uint32_t *gid, uint32_t n);
int pbf_slab_download(pbf_handle h, float *pos4, float *vel4, uint32_t *gid, uint32_t *n);
/* highlight words of the local particles in the same slot order (selection bits travel with migrating particles and with
 * the ghosts, so highlight.glsl's marks cross slab planes); pbf_toggle_highlight on a slab handle takes a SLOT */
int pbf_slab_download_highlight(pbf_handle h, uint32_t *highlight);
/* with a virtual group, stepping any member steps the whole group */
int pbf_slab_step(pbf_handle h, int nsteps);
/* One rank's end-to-end call: n_in local particles from HOST arrays (pinned for speed) in, nsteps, the particles the rank
 * owns afterwards back out into the same arrays (*n_out of them; the arrays hold `capacity`) */
int pbf_slab_step_host(pbf_handle h, float *pos4, float *vel4, uint32_t *gid, uint32_t n_in, uint32_t capacity,
                       uint32_t *n_out, int nsteps);
/* Load balancing: this rank's particle count per GLOBAL cell layer (gz_global counters, HOST array), and new planes for
 * this rank from the next step on.  Every rank must be given consistent planes before any of them steps; the particles that
 * fall outside travel through the next step's migration, so move a plane by a layer or two at a time.  The handle's grid
 * depth (pbf_config.grid[2]) bounds the window: create slab handles a few layers deeper than (z_hi - z_lo) + 2. */
int pbf_slab_layer_counts(pbf_handle h, uint32_t *counts);
int pbf_slab_set_planes(pbf_handle h, int z_lo, int z_hi);
/* with PBF_SLAB_PHASES=1 in the environment (direct launches): device ms of the last step's five phases on this rank:
 * predict + migration out, arrivals + boundary + ghosts out, ghosts in + sort + cells, solver incl. halos, vorticity */
int pbf_slab_phase_times(pbf_handle h, float ms[5]);
/* out: local particles, ghosts from z-, ghosts from z+, boundary sent to z-, to z+, migrated away (total),
 * exchanges (total), bytes sent (total) */
int pbf_slab_stats(pbf_handle h, uint64_t out[8]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PBF_C_H */
