// api.cu -- the C ABI of include/pbf_c.h: handle life cycle, step orchestration (SPH::Run, reference
// src/SPH.cpp:246-334), CUDA-graph replay, phase timing (SPH::OutputTiming, src/SPH.cpp:218-240) and debug read-back.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "pbf_internal.cuh"

static thread_local std::string g_err;
void pbf_set_error(const std::string &msg) { g_err = msg; }

namespace {

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

int bitlength(uint64_t v) {   // count_sortbits, src/RadixSort.cpp:24-30
    int r = 1;
    while (v >>= 1) r++;
    return r;
}

template <class T>
cudaError_t dalloc(T **p, size_t count) {
    return cudaMalloc((void **)p, count * sizeof(T) + 16);
}

void invalidate_graph(pbf_sim *s) {
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    s->graph_exec = nullptr;
    s->graph = nullptr;
    s->graph_valid = false;
}

// SPH::Run, src/SPH.cpp:246-334.  ev != nullptr records the five phase boundaries of the reference's timer queries.
// pos_ready (optional) is recorded once the by-id positions are final (after K10), i.e. before the vorticity kernels.
int enqueue_step(pbf_sim *s, bool with_events, cudaEvent_t pos_ready = nullptr) {
    int k = 0;
    cudaStream_t st = s->stream;
    if (with_events) cudaEventRecord(s->ev[0], st);
    // [predict] :247-261   (+ reset of last step's cell starts, + histograms, + clearhighlight)
    cudaMemsetAsync(s->flags, 0, sizeof(u32), st);
    k += launch_predict(s);            // also resets the cell-table entries the previous step wrote
    if (with_events) cudaEventRecord(s->ev[1], st);
    // [sort] :263-268
    k += launch_sort_scan(s);
    k += launch_sort_passes(s);
    if (with_events) cudaEventRecord(s->ev[2], st);
    // [neighbour cells] :270-275
    k += launch_reorder_cells(s);
    if (with_events) cudaEventRecord(s->ev[3], st);
    // [solver] :277-313
    k += launch_highlight(s);
    const int K = s->params.num_solver_iterations;
    if (with_events && (int)s->ev_solver.size() < 2 * K + 1) {
        const size_t have = s->ev_solver.size();
        s->ev_solver.resize(2 * K + 1, nullptr);
        for (size_t i = have; i < s->ev_solver.size(); i++) cudaEventCreate(&s->ev_solver[i]);
    }
    // K10 (update.glsl) runs in the epilogue of the LAST delta-p sweep (sweeps.cu); PBF_SEPARATE_UPDATE=1 keeps it apart
    const bool fuse_update = K > 0 && s->fuse_update;
    for (int it = 0; it < K; it++) {
        if (with_events) cudaEventRecord(s->ev_solver[2 * it], st);
        k += launch_lambda(s);
        if (with_events) cudaEventRecord(s->ev_solver[2 * it + 1], st);
        k += (fuse_update && it == K - 1) ? launch_delta_p_update(s) : launch_delta_p(s);
    }
    if (with_events) { cudaEventRecord(s->ev_solver[2 * K], st); s->ev_solver_iters = K; }
    if (with_events) cudaEventRecord(s->ev[4], st);
    // [vorticity] :315-333
    if (!fuse_update) k += launch_update(s);
    if (pos_ready) cudaEventRecord(pos_ready, st);
    if (s->params.vorticity_confinement) k += launch_vorticity(s);
    if (with_events) cudaEventRecord(s->ev[5], st);
    return k;
}

int check_handle(pbf_handle h) {
    if (!h) return fail(PBF_ERR_INVALID, "null handle");
    return PBF_OK;
}

}  // namespace

extern "C" {

const char *pbf_last_error(void) { return g_err.c_str(); }
int pbf_version(void) { return 100; }

float pbf_wpoly6(float r, float h) {   // SPH::Wpoly6, src/SPH.cpp:159-164
    if (r > h) return 0.0f;
    float tmp = h * h - r * r;
    return 1.56668147106f * tmp * tmp * tmp / (h * h * h * h * h * h * h * h * h);
}

void pbf_default_params(pbf_params *p) {   // src/SPH.cpp:25-26, :137-144
    p->one_over_rho_0 = 1.0f;
    p->epsilon = 5.0f;
    p->gravity = 10.0f;
    p->timestep = 0.016f;
    p->tensile_instability_k = 0.1f;
    p->tensile_instability_scale = 1.0f / pbf_wpoly6(0.2f, 2.0f);
    p->xsph_viscosity_c = 0.01f;
    p->vorticity_epsilon = 5.0f;
    p->num_solver_iterations = 5;
    p->vorticity_confinement = 0;
    p->external_force = 0;
}

int pbf_sort_bits(const int32_t grid[3]) {   // src/RadixSort.cpp:44, :127
    int numbits = bitlength((uint64_t)grid[0] * (uint64_t)grid[1] * (uint64_t)grid[2] - 1);
    return 2 * ((numbits + 1) >> 1);
}

// Key bits the simulation's own sort has to look at.  The reference sorts on pbf_sort_bits low bits of the hash (an even
// number, it sorts two bits per pass).  The hash of a clamped cell never exceeds ncell + gx*gz + gx (a cell on the x = gx,
// y = gy or z = gz plane), so every masked key is below 2^live and the bits from `live` up are zero: leaving them
// out gives the same stable permutation.  512 x 256 x 514 (a slab rank's window): 28 -> 27 bits = three 9-bit passes, not four.
static int live_sort_bits(const int32_t grid[3]) {
    const int bits = pbf_sort_bits(grid);
    const uint64_t ncell = (uint64_t)grid[0] * (uint64_t)grid[1] * (uint64_t)grid[2];
    const uint64_t kmax = ncell + (uint64_t)grid[0] * (uint64_t)grid[2] + (uint64_t)grid[0];
    const uint64_t top = (1ull << bits) - 1;
    const int live = bitlength(kmax < top ? kmax : top);
    return live < bits ? live : bits;
}

int pbf_sort_passes(const int32_t grid[3]) { return make_sort_plan(live_sort_bits(grid)).passes; }

int pbf_create(const pbf_config *cfg, pbf_handle *out) {
    if (!cfg || !out) return fail(PBF_ERR_INVALID, "pbf_create: null argument");
    *out = nullptr;
    if (cfg->num_particles == 0 || (cfg->num_particles & 511u))   // src/Simulation.cpp:202, src/SPH.cpp:25 (N >> 9 blocks)
        return fail(PBF_ERR_INVALID, "pbf_create: num_particles must be a non-zero multiple of 512");
    if (cfg->grid[0] < 1 || cfg->grid[1] < 1 || cfg->grid[2] < 1)
        return fail(PBF_ERR_INVALID, "pbf_create: grid dimensions must be positive");
    const uint64_t ncell = (uint64_t)cfg->grid[0] * cfg->grid[1] * cfg->grid[2];
    const int sortbits = pbf_sort_bits(cfg->grid);
    if (ncell > (1ull << 30) || sortbits > 30) return fail(PBF_ERR_INVALID, "pbf_create: grid larger than 2^30 cells");
    const int bx = bitlength((uint64_t)cfg->grid[0] + 3), bz = bitlength((uint64_t)cfg->grid[2] + 3),
              by = bitlength((uint64_t)cfg->grid[1] + 3);
    if (bx + by + bz > 32) return fail(PBF_ERR_INVALID, "pbf_create: grid extents do not pack into a 32-bit home cell");
    u32 cap = cfg->capacity ? cfg->capacity : cfg->num_particles;
    if (cap < cfg->num_particles) return fail(PBF_ERR_INVALID, "pbf_create: capacity < num_particles");
    if (cap >= (1u << 30)) return fail(PBF_ERR_INVALID, "pbf_create: capacity must be < 2^30");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PBF_ERR_CUDA, std::string("pbf_create: no CUDA device (") + cudaGetErrorString(e) +
                                      "); libpbf_b200 has no CPU fallback");
    int dev = cfg->device;
    if (dev < 0) PBF_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PBF_ERR_INVALID, "pbf_create: device ordinal out of range");
    cudaDeviceProp prop;
    PBF_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(PBF_ERR_CUDA, std::string("pbf_create: device '") + prop.name + "' is not sm_100 class; this library is built for sm_100a only");

    pbf_sim *s = new (std::nothrow) pbf_sim();   // value-initialised: every plain member starts at zero
    if (!s) return fail(PBF_ERR_INVALID, "pbf_create: out of host memory");
    s->cfg = *cfg;
    s->device = dev;
    s->sm_count = prop.multiProcessorCount;
    s->n = cfg->num_particles;
    s->cap = cap;
    s->ncell = (size_t)ncell;
    s->grid.gx = cfg->grid[0]; s->grid.gy = cfg->grid[1]; s->grid.gz = cfg->grid[2];
    s->grid.gxgz = cfg->grid[0] * cfg->grid[2];
    for (int a = 0; a < 3; a++) {
        s->grid.wlo[a] = 0.0f + cfg->wall[a];                      // updatepos.glsl:100
        s->grid.whi[a] = (float)cfg->grid[a] - cfg->wall[a];
    }
    s->grid.ref_quirks = cfg->ref_quirks;
    s->grid.bx = bx; s->grid.bz = bz;
    s->grid.zoff = 0; s->grid.gz_global = cfg->grid[2];
    s->plan = make_sort_plan(live_sort_bits(cfg->grid));
    const char *su = getenv("PBF_SEPARATE_UPDATE");
    s->fuse_update = !(su && su[0] == '1');
    const char *gs = getenv("PBF_GENERAL_SWEEPS");
    s->tiled_sweeps = !(gs && gs[0] == '1');
    pbf_default_params(&s->params);
    s->options.density_self_term = 0;
    s->options.wall_restitution = -1.0f;
    s->options.full_support = 0;

    DeviceGuard guard(dev);
#define ALLOC(ptr, count)                                                                                       \
    do {                                                                                                        \
        cudaError_t e2 = dalloc(&(ptr), (count));                                                              \
        if (e2 != cudaSuccess) {                                                                                \
            std::string m = std::string("pbf_create: cudaMalloc " #ptr ": ") + cudaGetErrorString(e2);        \
            pbf_destroy(s);                                                                                     \
            return fail(PBF_ERR_CUDA, m);                                                                       \
        }                                                                                                       \
    } while (0)
    e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; return fail(PBF_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    ALLOC(s->pos_own, cap); ALLOC(s->vel_own, cap); ALLOC(s->hl_own, cap);
    ALLOC(s->pred, cap); ALLOC(s->keys, cap);
    ALLOC(s->ktmp[0], cap); ALLOC(s->ktmp[1], cap); ALLOC(s->vtmp[0], cap); ALLOC(s->vtmp[1], cap);
    ALLOC(s->skey, cap); ALLOC(s->perm, cap); ALLOC(s->home, cap);
    s->max_tiles = sort_max_tiles(cap);
    ALLOC(s->hist, 8 * PBF_RADIX); ALLOC(s->gbase, 8 * PBF_RADIX);   // second halves: pbf_sort_pairs
    ALLOC(s->tile_counter, 4);
    ALLOC(s->status, (size_t)4 * s->max_tiles * PBF_RADIX);
    ALLOC(s->cells, s->ncell); ALLOC(s->runs3, s->ncell);
    ALLOC(s->bufA, cap); ALLOC(s->bufB, cap); ALLOC(s->svel, cap); ALLOC(s->vprime, cap); ALLOC(s->omega, cap);
    ALLOC(s->flags, 4); ALLOC(s->diag, 2); ALLOC(s->dn, DN_WORDS);
    ALLOC(s->tile_desc, plan_desc_ints(cap)); ALLOC(s->tile_runs, plan_run_words(cap));
#undef ALLOC
    if (sort_init() != 0) {
        pbf_destroy(s);
        return fail(PBF_ERR_CUDA, "pbf_create: the sort kernel needs 56 KB of opt-in shared memory per block");
    }
    if (sweeps_init() != 0) {
        pbf_destroy(s);
        return fail(PBF_ERR_CUDA, "pbf_create: the sweep kernels need 90 KB of opt-in shared memory per block");
    }
    s->pos = s->pos_own; s->vel = s->vel_own; s->hl = s->hl_own;
    cudaMemsetAsync(s->pos_own, 0, (size_t)cap * 16, s->stream);
    cudaMemsetAsync(s->vel_own, 0, (size_t)cap * 16, s->stream);
    cudaMemsetAsync(s->hl_own, 0, (size_t)cap * 4, s->stream);
    cudaMemsetAsync(s->hist, 0, 8 * PBF_RADIX * 4, s->stream);
    cudaMemsetAsync(s->tile_counter, 0, 16, s->stream);
    cudaMemsetAsync(s->flags, 0, 16, s->stream);
    cudaMemsetAsync(s->dn, 0, DN_WORDS * sizeof(u32), s->stream);
    for (float4 *b : {s->pred, s->bufA, s->bufB, s->svel, s->vprime, s->omega})   // pair loads may touch one slot of padding
        cudaMemsetAsync(b, 0, (size_t)cap * 16 + 16, s->stream);
    launch_fill_tables(s);   // start = -1 (gridtexture clear, src/NeighbourCellFinder.cpp:116-126), end = 0, runs empty
    for (int i = 0; i < 6; i++) cudaEventCreate(&s->ev[i]);
    e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess) e = cudaGetLastError();   // a rejected memset / launch above is not sticky: ask explicitly
    if (e != cudaSuccess) {
        std::string m = std::string("pbf_create: ") + cudaGetErrorString(e);
        pbf_destroy(s);
        return fail(PBF_ERR_CUDA, m);
    }
    *out = s;
    return PBF_OK;
}

int pbf_destroy(pbf_handle s) {
    if (!s) return PBF_OK;
    DeviceGuard guard(s->device);
    const bool shared_stream = slab_borrows_stream(s);   // virtual ranks > 0 borrow rank 0's stream
    if (shared_stream) { cudaDeviceSynchronize(); s->stream = nullptr; }
    if (s->stream) cudaStreamSynchronize(s->stream);
    slab_free(s);
    if (s->gl_registered && !s->ext_map)
        for (int i = 0; i < 3; i++) cudaGraphicsUnregisterResource(s->gl_res[i]);
    invalidate_graph(s);
    void *ptrs[] = {s->pos_own, s->vel_own, s->hl_own, s->pred, s->keys, s->ktmp[0], s->ktmp[1], s->vtmp[0], s->vtmp[1],
                    s->skey, s->perm, s->home, s->hist, s->gbase, s->tile_counter, s->status, s->cells, s->runs3, s->bufA, s->bufB,
                    s->svel, s->vprime, s->omega, s->flags, s->diag, s->tile_desc, s->tile_runs, s->dn};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 6; i++)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (cudaEvent_t e : s->ev_solver)
        if (e) cudaEventDestroy(e);
    if (s->copy_stream) { cudaStreamDestroy(s->copy_stream); cudaEventDestroy(s->ev_pos); cudaEventDestroy(s->ev_copied); }
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return PBF_OK;
}

int pbf_set_params(pbf_handle s, const pbf_params *p) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!p) return fail(PBF_ERR_INVALID, "pbf_set_params: null");
    if (p->num_solver_iterations < 0) return fail(PBF_ERR_INVALID, "pbf_set_params: negative iteration count");
    if (memcmp(&s->params, p, sizeof(*p)) != 0) {
        s->params = *p;
        invalidate_graph(s);   // kernel arguments are baked into the captured graph
    }
    return PBF_OK;
}

int pbf_get_params(pbf_handle s, pbf_params *p) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!p) return fail(PBF_ERR_INVALID, "pbf_get_params: null");
    *p = s->params;
    return PBF_OK;
}

int pbf_set_options(pbf_handle s, const pbf_options *o) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!o) return fail(PBF_ERR_INVALID, "pbf_set_options: null");
    if (o->wall_restitution > 1.0f) return fail(PBF_ERR_INVALID, "pbf_set_options: wall_restitution must be <= 1 (negative = off)");
    if (memcmp(&s->options, o, sizeof(*o)) != 0) {
        s->options = *o;
        invalidate_graph(s);   // kernel arguments are baked into the captured graph
    }
    return PBF_OK;
}

// Verification mode: one order of every floating-point sum on every code path, so that a slab decomposition reproduces the
// single-domain run bit for bit (SURVEY.md 8e "bit-exactness across GPU counts"): a particle's nine runs are walked in row
// order on the tiled and on the general path alike (the plan does not sort them by length), pairs are counted from a run's
// first candidate, one thread per particle in the first vorticity sweep, and a slab rank orders the particles of every cell
// by global id after its sort (slab.cu, k_cell_order; needs the device-count step).  Costs a few per cent on a disordered
// scene; results stay within the parity tolerances of the default mode.
int pbf_set_canonical_order(pbf_handle s, int on) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->canonical != (on != 0)) {
        s->canonical = on != 0;
        invalidate_graph(s);
    }
    return PBF_OK;
}

int pbf_get_options(pbf_handle s, pbf_options *o) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!o) return fail(PBF_ERR_INVALID, "pbf_get_options: null");
    *o = s->options;
    return PBF_OK;
}

uint32_t pbf_num_particles(pbf_handle s) { return s ? s->n : 0; }

int pbf_upload_state(pbf_handle s, const float *pos4, const float *vel4, uint32_t n) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!pos4) return fail(PBF_ERR_INVALID, "pbf_upload_state: null positions");
    if (n != s->n) return fail(PBF_ERR_INVALID, "pbf_upload_state: n differs from the handle's particle count");
    DeviceGuard guard(s->device);
    GlScope gl(s);
    if (gl.rc) return gl.rc;
    PBF_CUDA(cudaMemcpyAsync(s->pos, pos4, (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    if (vel4) PBF_CUDA(cudaMemcpyAsync(s->vel, vel4, (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    else PBF_CUDA(cudaMemsetAsync(s->vel, 0, (size_t)n * 16, s->stream));
    PBF_CUDA(cudaMemsetAsync(s->hl, 0, (size_t)n * 4, s->stream));   // src/Simulation.cpp:271-272
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    s->stage = 0;
    return PBF_OK;
}

int pbf_download_state(pbf_handle s, float *pos4, float *vel4, uint32_t *highlight) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    DeviceGuard guard(s->device);
    GlScope gl(s);
    if (gl.rc) return gl.rc;
    if (pos4) PBF_CUDA(cudaMemcpyAsync(pos4, s->pos, (size_t)s->n * 16, cudaMemcpyDeviceToHost, s->stream));
    if (vel4) PBF_CUDA(cudaMemcpyAsync(vel4, s->vel, (size_t)s->n * 16, cudaMemcpyDeviceToHost, s->stream));
    if (highlight) PBF_CUDA(cudaMemcpyAsync(highlight, s->hl, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

int pbf_device_buffers(pbf_handle s, float **pos4, float **vel4, uint32_t **highlight) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->gl_registered)   // a mapped address is only valid between map and unmap: the renderer reaches the data through GL
        return fail(PBF_ERR_STATE, "pbf_device_buffers: the state lives in registered (GL / external) buffers, whose addresses are only valid while mapped");
    if (pos4) *pos4 = (float *)s->pos;
    if (vel4) *vel4 = (float *)s->vel;
    if (highlight) *highlight = s->hl;
    return PBF_OK;
}

int pbf_bind_device_buffers(pbf_handle s, float *pos4, float *vel4, uint32_t *highlight) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    float4 *np = pos4 ? (float4 *)pos4 : s->pos_own, *nv = vel4 ? (float4 *)vel4 : s->vel_own;
    u32 *nh = highlight ? highlight : s->hl_own;
    if (((uintptr_t)np | (uintptr_t)nv) & 15u) return fail(PBF_ERR_INVALID, "pbf_bind_device_buffers: buffers must be 16-byte aligned");
    if (np != s->pos || nv != s->vel || nh != s->hl) {
        s->pos = np; s->vel = nv; s->hl = nh;
        invalidate_graph(s);
    }
    return PBF_OK;
}

}  // extern "C"

// ---- CUDA-GL interop -------------------------------------------------------------------------------
// The renderer owns the three GL buffer objects (src/SPH.cpp:96-133 creates them, src/PointSprite.cpp:64-84 reads them as
// vertex attributes).  They are registered once; every pbf_step maps them on the handle's stream, runs on the mapped
// pointers and unmaps, which orders GL's later reads after the step.  cudaGraphicsGLRegisterBuffer is declared by hand
// (cuda_gl_interop.h drags in <GL/gl.h>, which a headless build does not have); it lives in libcudart like the rest.
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource **resource, unsigned int buffer,
                                                     unsigned int flags);

namespace {

int gl_map(pbf_sim *s) {
    if (s->ext_map) {   // caller-owned buffers behind callbacks (pbf_register_external_buffers)
        float *p = nullptr, *v = nullptr;
        uint32_t *h = nullptr;
        const int rc = s->ext_map(s->ext_user, (void *)s->stream, &p, &v, &h);
        if (rc != 0 || !p || !v || !h) {
            if (rc == 0) s->ext_unmap(s->ext_user, (void *)s->stream);
            return fail(PBF_ERR_STATE, "external buffer map callback failed (code " + std::to_string(rc) + ")");
        }
        const int br = pbf_bind_device_buffers(s, p, v, h);
        if (br) s->ext_unmap(s->ext_user, (void *)s->stream);
        return br;
    }
    cudaError_t e = cudaGraphicsMapResources(3, s->gl_res, s->stream);
    if (e != cudaSuccess) return fail(PBF_ERR_CUDA, std::string("cudaGraphicsMapResources: ") + cudaGetErrorString(e));
    void *ptr[3] = {nullptr, nullptr, nullptr};
    size_t bytes[3] = {0, 0, 0};
    for (int i = 0; i < 3 && e == cudaSuccess; i++) e = cudaGraphicsResourceGetMappedPointer(&ptr[i], &bytes[i], s->gl_res[i]);
    if (e == cudaSuccess && (bytes[0] < (size_t)s->n * 16 || bytes[1] < (size_t)s->n * 16 || bytes[2] < (size_t)s->n * 4)) {
        cudaGraphicsUnmapResources(3, s->gl_res, s->stream);
        return fail(PBF_ERR_INVALID, "GL buffers are smaller than N x float4 / N x uint32");
    }
    if (e != cudaSuccess) {
        cudaGraphicsUnmapResources(3, s->gl_res, s->stream);
        return fail(PBF_ERR_CUDA, std::string("cudaGraphicsResourceGetMappedPointer: ") + cudaGetErrorString(e));
    }
    // mapped addresses are normally the same from one map to the next, so the captured graph survives
    return pbf_bind_device_buffers(s, (float *)ptr[0], (float *)ptr[1], (uint32_t *)ptr[2]);
}

}  // namespace

GlScope::GlScope(pbf_sim *sim) : s(sim), rc(PBF_OK), owner(false) {
    if (!s || !s->gl_registered || s->gl_mapped) return;
    rc = gl_map(s);
    if (rc == PBF_OK) { s->gl_mapped = true; owner = true; }
}

GlScope::~GlScope() {
    if (!owner) return;
    if (s->ext_map) s->ext_unmap(s->ext_user, (void *)s->stream);
    else cudaGraphicsUnmapResources(3, s->gl_res, s->stream);
    s->gl_mapped = false;
}

extern "C" int pbf_register_gl_buffers(pbf_handle s, unsigned int pos, unsigned int vel, unsigned int highlight) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->gl_registered) return fail(PBF_ERR_STATE, "pbf_register_gl_buffers: buffers already registered");
    if (s->slab) return fail(PBF_ERR_STATE, "pbf_register_gl_buffers: not available on slab handles");
    DeviceGuard guard(s->device);
    const unsigned int names[3] = {pos, vel, highlight};
    for (int i = 0; i < 3; i++) {
        cudaError_t e = cudaGraphicsGLRegisterBuffer(&s->gl_res[i], names[i], 0u /* cudaGraphicsRegisterFlagsNone */);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; j++) cudaGraphicsUnregisterResource(s->gl_res[j]);
            (void)cudaGetLastError();
            return fail(PBF_ERR_CUDA, std::string("cudaGraphicsGLRegisterBuffer (is a GL context current on this thread?): ") +
                                          cudaGetErrorString(e));
        }
    }
    s->gl_registered = true;
    return PBF_OK;
}

extern "C" int pbf_unregister_gl_buffers(pbf_handle s) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!s->gl_registered) return PBF_OK;
    DeviceGuard guard(s->device);
    cudaStreamSynchronize(s->stream);
    if (!s->ext_map)
        for (int i = 0; i < 3; i++) cudaGraphicsUnregisterResource(s->gl_res[i]);
    s->ext_map = nullptr; s->ext_unmap = nullptr; s->ext_user = nullptr;
    s->gl_registered = false;
    return pbf_bind_device_buffers(s, nullptr, nullptr, nullptr);
}

extern "C" int pbf_register_external_buffers(pbf_handle s, pbf_map_fn map, pbf_unmap_fn unmap, void *user) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!map || !unmap) return fail(PBF_ERR_INVALID, "pbf_register_external_buffers: null callback");
    if (s->gl_registered) return fail(PBF_ERR_STATE, "pbf_register_external_buffers: buffers already registered");
    if (s->slab) return fail(PBF_ERR_STATE, "pbf_register_external_buffers: not available on slab handles");
    s->ext_map = map; s->ext_unmap = unmap; s->ext_user = user;
    s->gl_registered = true;
    return PBF_OK;
}

extern "C" int pbf_unregister_external_buffers(pbf_handle s) { return pbf_unregister_gl_buffers(s); }

extern "C" {

int pbf_step(pbf_handle s, int nsteps) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (nsteps < 0) return fail(PBF_ERR_INVALID, "pbf_step: negative step count");
    DeviceGuard guard(s->device);
    if (s->slab) return fail(PBF_ERR_STATE, "pbf_step: this handle is one slab of a decomposed domain; use pbf_slab_step");
    GlScope gl(s);   // GL gets its buffers back on every way out
    if (gl.rc) return gl.rc;
    const bool use_graph = s->cfg.use_graph && !s->timing;
    for (int i = 0; i < nsteps; i++) {
        if (use_graph) {
            // the very first step has no previous cell table to reset, so its launch list differs: run it directly
            if (!s->graph_valid && s->n_prev_sorted == s->n) {
                PBF_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
                int k = enqueue_step(s, false);
                cudaError_t e = cudaStreamEndCapture(s->stream, &s->graph);
                if (e != cudaSuccess) { s->graph = nullptr; return fail(PBF_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e)); }
                PBF_CUDA(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
                s->graph_kernels = (u32)k;
                s->graph_valid = true;
            }
            if (s->graph_valid) {
                PBF_CUDA(cudaGraphLaunch(s->graph_exec, s->stream));
                s->launches += s->graph_kernels;
                s->steps++;
                continue;
            }
        }
        s->launches += (uint64_t)enqueue_step(s, s->timing);
        s->steps++;
        s->ev_valid = s->timing;
        PBF_CUDA(cudaGetLastError());
    }
    s->stage = 0;
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_step_host(pbf_handle s, float *pos4, float *vel4, int nsteps) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!pos4 || !vel4) return fail(PBF_ERR_INVALID, "pbf_step_host: null buffer");
    if (nsteps < 0) return fail(PBF_ERR_INVALID, "pbf_step_host: negative step count");
    if (s->gl_registered) return fail(PBF_ERR_STATE, "pbf_step_host: the state lives in registered (GL / external) buffers; use pbf_step");
    DeviceGuard guard(s->device);
    const size_t bytes = (size_t)s->n * 16;
    if (!s->copy_stream) {
        PBF_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        PBF_CUDA(cudaEventCreateWithFlags(&s->ev_pos, cudaEventDisableTiming));
        PBF_CUDA(cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
    }
    PBF_CUDA(cudaMemcpyAsync(s->pos, pos4, bytes, cudaMemcpyHostToDevice, s->stream));
    PBF_CUDA(cudaMemcpyAsync(s->vel, vel4, bytes, cudaMemcpyHostToDevice, s->stream));
    if (nsteps == 0) {
        PBF_CUDA(cudaStreamSynchronize(s->stream));
        return PBF_OK;
    }
    if (nsteps > 1) {
        int r = pbf_step(s, nsteps - 1);
        if (r) return r;
    }
    // Last step by direct launches: the positions are final after K10 (update.glsl), so their copy back to the host
    // runs on a second stream underneath the vorticity kernels (compute bound, the copy engine is idle); the
    // velocities follow on the main stream.  Without vorticity both arrays are final at the same point.
    s->launches += (uint64_t)enqueue_step(s, false, s->ev_pos);
    s->steps++;
    s->ev_valid = false;
    s->stage = 0;
    PBF_CUDA(cudaGetLastError());
    PBF_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_pos, 0));
    PBF_CUDA(cudaMemcpyAsync(pos4, s->pos, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    PBF_CUDA(cudaEventRecord(s->ev_copied, s->copy_stream));
    PBF_CUDA(cudaMemcpyAsync(vel4, s->vel, bytes, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamWaitEvent(s->stream, s->ev_copied, 0));   // later work on the handle's stream may rewrite pos
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

int pbf_sync(pbf_handle s) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    DeviceGuard guard(s->device);
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

// ---- stage-level entry points -------------------------------------------------------------------------------------
#define STAGE_PROLOGUE(need, name)                                                                       \
    if (check_handle(s)) return PBF_ERR_INVALID;                                                         \
    if (s->stage < (need)) return fail(PBF_ERR_STATE, name ": called before the stages it depends on"); \
    DeviceGuard guard(s->device);                                                                        \
    GlScope gl(s);                                                                                       \
    if (gl.rc) return gl.rc;

int pbf_predict(pbf_handle s) {
    STAGE_PROLOGUE(0, "pbf_predict");
    cudaMemsetAsync(s->flags, 0, sizeof(u32), s->stream);
    s->launches += launch_predict(s);   // also resets the cell-table entries the previous step wrote
    s->n_prev_sorted = 0;   // table is clean until pbf_build_cells refills it
    PBF_CUDA(cudaGetLastError());
    s->stage = 1;
    return PBF_OK;
}

int pbf_sort(pbf_handle s) {
    STAGE_PROLOGUE(1, "pbf_sort");
    // RadixSort::Run sorts the records predictpos.glsl has just written; the digit histograms of those keys are consumed
    // by the scan, so a second Run without a new pbf_predict has nothing to sort from
    if (s->stage != 1) return fail(PBF_ERR_STATE, "pbf_sort: already sorted; call pbf_predict first");
    s->launches += launch_sort_scan(s);
    s->launches += launch_sort_passes(s);
    PBF_CUDA(cudaGetLastError());
    s->stage = 2;
    return PBF_OK;
}

int pbf_build_cells(pbf_handle s) {
    STAGE_PROLOGUE(2, "pbf_build_cells");
    s->launches += launch_reorder_cells(s);
    PBF_CUDA(cudaGetLastError());
    s->stage = 3;
    return PBF_OK;
}

int pbf_highlight(pbf_handle s) {
    STAGE_PROLOGUE(3, "pbf_highlight");
    s->launches += launch_highlight(s);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_calc_lambda(pbf_handle s) {
    STAGE_PROLOGUE(3, "pbf_calc_lambda");
    s->launches += launch_lambda(s);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_update_positions(pbf_handle s) {
    STAGE_PROLOGUE(3, "pbf_update_positions");
    s->launches += launch_delta_p(s);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_finalize(pbf_handle s) {
    STAGE_PROLOGUE(3, "pbf_finalize");
    s->launches += launch_update(s);
    PBF_CUDA(cudaGetLastError());
    s->stage = 4;
    return PBF_OK;
}

int pbf_vorticity(pbf_handle s) {
    STAGE_PROLOGUE(4, "pbf_vorticity");
    if (!s->params.vorticity_confinement)
        return fail(PBF_ERR_STATE, "pbf_vorticity: vorticity_confinement was off in pbf_finalize (sorted velocities missing)");
    s->launches += launch_vorticity(s);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_sort_pairs(pbf_handle s, const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                   uint32_t *vals_out, uint32_t n, int bits) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!keys_in || !keys_out || !vals_out) return fail(PBF_ERR_INVALID, "pbf_sort_pairs: null buffer");
    if (n > s->cap) return fail(PBF_ERR_CAPACITY, "pbf_sort_pairs: n exceeds the handle's capacity");
    if (bits < 1 || bits > 32) return fail(PBF_ERR_INVALID, "pbf_sort_pairs: bits must be in [1,32]");
    DeviceGuard guard(s->device);
    s->launches += launch_sort_pairs(s, keys_in, vals_in, keys_out, vals_out, n, bits);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

// ---- debug read-back --------------------------------------------------------------------------------------------------
int pbf_get_predicted(pbf_handle s, float *records, uint32_t *keys) {
    STAGE_PROLOGUE(1, "pbf_get_predicted");
    if (records) PBF_CUDA(cudaMemcpyAsync(records, s->pred, (size_t)s->n * 16, cudaMemcpyDeviceToHost, s->stream));
    if (keys) PBF_CUDA(cudaMemcpyAsync(keys, s->keys, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

int pbf_get_sorted(pbf_handle s, uint32_t *keys, uint32_t *perm, float *records) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->n_prev_sorted != s->n && s->stage < 2) return fail(PBF_ERR_STATE, "pbf_get_sorted: nothing sorted yet");
    DeviceGuard guard(s->device);
    if (keys) PBF_CUDA(cudaMemcpyAsync(keys, s->skey, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->stream));
    if (perm) PBF_CUDA(cudaMemcpyAsync(perm, s->perm, (size_t)s->n * 4, cudaMemcpyDeviceToHost, s->stream));
    if (records) {
        if (s->n_prev_sorted != s->n) return fail(PBF_ERR_STATE, "pbf_get_sorted: records need pbf_build_cells");
        // compose {pos, id} = RadixSort::GetBuffer's record layout (predictpos.glsl:3-6) in scratch (bufB is free between sweeps
        // only after lambda has been consumed; use vprime, which only the vorticity kernels touch)
        s->launches += launch_compose_records(s, s->vprime);
        PBF_CUDA(cudaMemcpyAsync(records, s->vprime, (size_t)s->n * 16, cudaMemcpyDeviceToHost, s->stream));
    }
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

int pbf_get_cell_ranges(pbf_handle s, int32_t *start, int32_t *end) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    DeviceGuard guard(s->device);
    std::vector<int2> tmp(s->ncell);
    PBF_CUDA(cudaMemcpyAsync(tmp.data(), s->cells, s->ncell * sizeof(int2), cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t c = 0; c < s->ncell; c++) {
        if (start) start[c] = tmp[c].x;
        if (end) end[c] = tmp[c].y;
    }
    return PBF_OK;
}

int pbf_get_neighbour_runs(pbf_handle s, int32_t *run_start, int32_t *run_count) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->n_prev_sorted != s->n) return fail(PBF_ERR_STATE, "pbf_get_neighbour_runs: needs pbf_build_cells");
    if (!run_start || !run_count) return fail(PBF_ERR_INVALID, "pbf_get_neighbour_runs: null buffer");
    DeviceGuard guard(s->device);
    int *rs = nullptr, *rc = nullptr;
    PBF_CUDA(cudaMalloc(&rs, (size_t)s->n * 36));
    cudaError_t e = cudaMalloc(&rc, (size_t)s->n * 36);
    if (e != cudaSuccess) { cudaFree(rs); return fail(PBF_ERR_CUDA, cudaGetErrorString(e)); }
    s->launches += launch_neighbour_runs(s, rs, rc);
    cudaMemcpyAsync(run_start, rs, (size_t)s->n * 36, cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(run_count, rc, (size_t)s->n * 36, cudaMemcpyDeviceToHost, s->stream);
    e = cudaStreamSynchronize(s->stream);
    cudaFree(rs); cudaFree(rc);
    if (e != cudaSuccess) return fail(PBF_ERR_CUDA, cudaGetErrorString(e));
    return PBF_OK;
}

static int get_w(pbf_handle s, float *out, const char *what) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!out) return fail(PBF_ERR_INVALID, std::string(what) + ": null buffer");
    DeviceGuard guard(s->device);
    std::vector<float4> tmp(s->n);
    PBF_CUDA(cudaMemcpyAsync(tmp.data(), s->bufB, (size_t)s->n * 16, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    for (u32 i = 0; i < s->n; i++) out[i] = tmp[i].w;
    return PBF_OK;
}
// lambda of the last pbf_calc_lambda / |omega| of the last vorticity sweep, by sorted slot (both live in bufB.w)
int pbf_get_lambda(pbf_handle s, float *lambda) { return get_w(s, lambda, "pbf_get_lambda"); }
int pbf_get_vorticity(pbf_handle s, float *vorticity) { return get_w(s, vorticity, "pbf_get_vorticity"); }

int pbf_enable_timing(pbf_handle s, int on) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    s->timing = on != 0;
    return PBF_OK;
}

int pbf_get_timings(pbf_handle s, float ms[5]) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!ms) return fail(PBF_ERR_INVALID, "pbf_get_timings: null");
    if (!s->ev_valid) return fail(PBF_ERR_STATE, "pbf_get_timings: no timed step (pbf_enable_timing first)");
    DeviceGuard guard(s->device);
    PBF_CUDA(cudaEventSynchronize(s->ev[5]));
    for (int i = 0; i < 5; i++) PBF_CUDA(cudaEventElapsedTime(&ms[i], s->ev[i], s->ev[i + 1]));
    return PBF_OK;
}

int pbf_get_solver_kernel_timings(pbf_handle s, float *lambda_ms, float *delta_p_ms) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!s->ev_valid || s->ev_solver_iters <= 0) return fail(PBF_ERR_STATE, "pbf_get_solver_kernel_timings: no timed step with solver iterations");
    DeviceGuard guard(s->device);
    const int K = s->ev_solver_iters;
    PBF_CUDA(cudaEventSynchronize(s->ev_solver[2 * K]));
    // the last iteration's updatepos launch also runs update.glsl in its epilogue (sweeps.cu): it is left out of the mean
    // unless it is the only one
    const int Kd = s->fuse_update && K > 1 ? K - 1 : K;
    double a = 0.0, b = 0.0;
    for (int it = 0; it < K; it++) {
        float x, y;
        PBF_CUDA(cudaEventElapsedTime(&x, s->ev_solver[2 * it], s->ev_solver[2 * it + 1]));
        PBF_CUDA(cudaEventElapsedTime(&y, s->ev_solver[2 * it + 1], s->ev_solver[2 * it + 2]));
        a += x;
        if (it < Kd) b += y;
    }
    if (lambda_ms) *lambda_ms = (float)(a / K);
    if (delta_p_ms) *delta_p_ms = (float)(b / Kd);
    return PBF_OK;
}

int pbf_pick_particle(pbf_handle s, const float origin[3], const float direction[3], float radius, int32_t *id) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (!origin || !direction || !id) return fail(PBF_ERR_INVALID, "pbf_pick_particle: null argument");
    const float len = sqrtf(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
    if (!(len > 0.0f) || !(radius > 0.0f)) return fail(PBF_ERR_INVALID, "pbf_pick_particle: zero direction or radius");
    const float d[3] = {direction[0] / len, direction[1] / len, direction[2] / len};
    DeviceGuard guard(s->device);
    GlScope gl(s);
    if (gl.rc) return gl.rc;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(s->diag);   // 16 bytes of scratch, stream ordered
    PBF_CUDA(cudaMemsetAsync(best, 0xff, sizeof(*best), s->stream));
    s->launches += launch_pick(s, origin, d, radius, best);
    unsigned long long host = 0;
    PBF_CUDA(cudaMemcpyAsync(&host, best, sizeof(host), cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    *id = host == ~0ull ? -1 : (int32_t)(host & 0xffffffffull);
    return PBF_OK;
}

int pbf_toggle_highlight(pbf_handle s, uint32_t id) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (id >= s->n) return fail(PBF_ERR_INVALID, "pbf_toggle_highlight: particle id out of range");
    DeviceGuard guard(s->device);
    GlScope gl(s);
    if (gl.rc) return gl.rc;
    s->launches += launch_toggle_highlight(s, id);
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

int pbf_get_diagnostics(pbf_handle s, double *density_error, double *kinetic_energy) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    DeviceGuard guard(s->device);
    GlScope gl(s);
    if (gl.rc) return gl.rc;
    PBF_CUDA(cudaMemsetAsync(s->diag, 0, 2 * sizeof(double), s->stream));
    if (density_error) {
        if (s->n_prev_sorted != s->n) return fail(PBF_ERR_STATE, "pbf_get_diagnostics: density needs a completed step");
        s->launches += launch_density_diag(s);
    }
    if (kinetic_energy) s->launches += launch_kinetic_diag(s);
    double host[2];
    PBF_CUDA(cudaMemcpyAsync(host, s->diag, sizeof(host), cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    if (density_error) *density_error = host[0] / (double)s->n;
    if (kinetic_energy) *kinetic_energy = host[1];
    return PBF_OK;
}

int pbf_get_tile_stats(pbf_handle s, uint32_t *tiles, uint32_t *tiled, uint32_t why[8]) {
    if (check_handle(s)) return PBF_ERR_INVALID;
    if (s->n_prev_sorted != s->n) return fail(PBF_ERR_STATE, "pbf_get_tile_stats: needs pbf_build_cells");
    DeviceGuard guard(s->device);
    const size_t nt = plan_desc_ints(s->n);
    std::vector<int> tmp(nt);
    PBF_CUDA(cudaMemcpyAsync(tmp.data(), s->tile_desc, nt * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    const size_t stride = plan_desc_ints(1);
    uint32_t a = 0, b = 0;
    if (why) memset(why, 0, 8 * sizeof(uint32_t));
    for (size_t t = 0; t < nt; t += stride) {
        const int mode = tmp[t] & 0xff;          // a slab rank's descriptors carry PBF_TILE_* flags above the mode
        a++;
        b += mode != 0;
        if (why) {
            why[mode != 0 ? 0 : 1]++;
            // staged records of the tile (desc[1]) per particle of a full tile, in six classes: <= 9, 10, 11, 12, 13, more
            const int per = (tmp[t + 1] + (int)plan_tile_size() - 1) / (int)plan_tile_size();
            why[2 + (per <= 9 ? 0 : (per > 13 ? 5 : per - 9))]++;
        }
    }
    if (tiles) *tiles = a;
    if (tiled) *tiled = b;
    return PBF_OK;
}

uint64_t pbf_kernel_launches(pbf_handle s) { return s ? s->launches : 0; }
void *pbf_stream(pbf_handle s) { return s ? (void *)s->stream : nullptr; }

}  // extern "C"
