// ref_harness.cpp -- runs the reference's OWN compute shaders on the CPU: oracle/_ref/libpbf_ref.so.
//
// TEST INFRASTRUCTURE ONLY.  Used by tests/ to pin oracle/pbf_oracle.c (and through it the CUDA path) against the
// reference's source text.  Nothing under pbf_b200/ links, loads or calls this.
//
// The shader bodies are the files of /root/reference/shaders/{sph,radixsort,neighbourcellfinder}, included below after the
// lexical translation of oracle/ref_translate.py (generated into oracle/_ref/gen/, never committed) and compiled against
// oracle/glsl_compat.h.  What this file adds is the HOST side the shaders need, restated from the reference's C++ (which
// is GL calls and cannot be compiled here -- no GL/GLFW/glm headers in the image):
//   * the constants every shader gets prepended (src/SPH.cpp:28-60, src/RadixSort.cpp:35-39, src/NeighbourCellFinder.cpp:27-31),
//   * buffer / texture / image bindings and dispatch sizes of SPH::Run (src/SPH.cpp:246-334), RadixSort::Run / SortBits
//     (src/RadixSort.cpp:124-200, block-sum hierarchy :67-85) and NeighbourCellFinder::FindNeighbourCells
//     (src/NeighbourCellFinder.cpp:113-148),
//   * a work-group scheduler: invocations of one group run as ucontext fibers that yield at barrier(), groups run one after
//     the other.
// GL leaves the order of invocations open and two of the shaders race by design (updatepos.glsl:53-55 updates positions in
// place while neighbours read them; vorticity.glsl:62-72 reads |omega| of other work groups after a group-local barrier and
// velocities that other invocations overwrite).  `order` selects the schedule for those two:
//   0  Jacobi: every invocation sees the buffer as it was before the dispatch (the oracle's and the CUDA path's policy);
//      done by saving each invocation's output and putting the old value back before the next one runs;
//   1  as dispatched: invocations in ascending order, writes land at once -- one schedule a GPU could produce.
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <vector>

#include "glsl_compat.h"

GlState gl_state;
uvec3 gl_GlobalInvocationID, gl_WorkGroupID;
uint gl_LocalInvocationIndex;

// ---- what the host prepends to every shader -------------------------------------------------------------------------
static vec3 g_GRID_SIZE;
static ivec3 g_GRID_HASHWEIGHTS;
struct SphParams {   // sphparams_t, src/SPH.h:252-285 = the std140 block of src/SPH.cpp:46-56
    float one_over_rho_0, epsilon, gravity, timestep, tensile_instability_k, tensile_instability_scale, xsph_viscosity_c,
        vorticity_epsilon;
};
static SphParams g_params;

#define GRID_HEADER                                              \
    static const vec3 &GRID_SIZE = g_GRID_SIZE;                 \
    static const ivec3 &GRID_HASHWEIGHTS = g_GRID_HASHWEIGHTS;
// src/SPH.cpp:28-60 (the uniform-block branch) ends with `const float h = 2.0;` and `#define BLOCKSIZE 256`
#define SPH_HEADER                                                                   \
    GRID_HEADER                                                                      \
    static const float &one_over_rho_0 = g_params.one_over_rho_0;                    \
    static const float &epsilon = g_params.epsilon;                                  \
    static const float &gravity = g_params.gravity;                                  \
    static const float &timestep = g_params.timestep;                                \
    static const float &tensile_instability_k = g_params.tensile_instability_k;      \
    static const float &tensile_instability_scale = g_params.tensile_instability_scale; \
    static const float &xsph_viscosity_c = g_params.xsph_viscosity_c;                \
    static const float &vorticity_epsilon = g_params.vorticity_epsilon;              \
    static const float h = 2.0f;

#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wunused-function"
#pragma GCC diagnostic ignored "-Wsign-compare"
#pragma GCC diagnostic ignored "-Wshadow"

#define BLOCKSIZE 256
namespace sh_predictpos { SPH_HEADER
#include "foreachneighbour.inc"
#include "predictpos.inc"
}
namespace sh_calclambda { SPH_HEADER
#include "foreachneighbour.inc"
#include "calclambda.inc"
}
namespace sh_updatepos { SPH_HEADER
#include "foreachneighbour.inc"
#include "updatepos.inc"
}
namespace sh_update { SPH_HEADER
#include "foreachneighbour.inc"
#include "update.inc"
}
namespace sh_vorticity { SPH_HEADER
#include "foreachneighbour.inc"
#include "vorticity.inc"
}
namespace sh_highlight { SPH_HEADER
#include "foreachneighbour.inc"
#include "highlight.inc"
}
namespace sh_clearhighlight { SPH_HEADER
#include "foreachneighbour.inc"
#include "clearhighlight.inc"
}
namespace sh_findcells { GRID_HEADER      // src/NeighbourCellFinder.cpp:27-31: GRID_SIZE, GRID_HASHWEIGHTS, BLOCKSIZE 256
#include "findcells.inc"
}
namespace sh_neighbourcells { GRID_HEADER
#include "neighbourcells.inc"
}
#undef BLOCKSIZE
#define BLOCKSIZE 512                     // src/RadixSort.cpp:35-39 with blocksize = 512 (src/SPH.cpp:25)
#define HALFBLOCKSIZE 256
namespace sh_counting { GRID_HEADER
#include "counting.inc"
}
namespace sh_blockscan { GRID_HEADER
#include "blockscan.inc"
}
namespace sh_addblocksum { GRID_HEADER
#include "addblocksum.inc"
}
namespace sh_globalsort { GRID_HEADER
#include "globalsort.inc"
}
#undef BLOCKSIZE
#undef HALFBLOCKSIZE

// ---- work-group scheduler ----------------------------------------------------------------------------------------------
namespace {

enum BarrierMode { BAR_FIBER, BAR_ABORT, BAR_NOOP };
BarrierMode g_barrier_mode = BAR_FIBER;
jmp_buf g_abort;

// Fibers: on x86-64 a six-register stack switch (swapcontext costs two sigprocmask system calls per switch, and the 2-bit
// radix sort of 64K particles crosses ~7M barriers); ucontext elsewhere.
#if defined(__x86_64__)
extern "C" void pbf_fiber_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl pbf_fiber_switch\n"
    ".type pbf_fiber_switch,@function\n"
    "pbf_fiber_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size pbf_fiber_switch,.-pbf_fiber_switch\n");
struct Fiber {
    void *sp;
    bool done;
};
void *g_sched_sp = nullptr;
#else
struct Fiber {
    ucontext_t ctx;
    bool done;
};
ucontext_t g_sched;
#endif
Fiber *g_cur = nullptr;
void (*g_main)(void) = nullptr;
std::vector<Fiber> g_fibers;
std::vector<char> g_stacks;
constexpr size_t STACK = 64 * 1024;

void to_scheduler(void) {
#if defined(__x86_64__)
    pbf_fiber_switch(&g_cur->sp, g_sched_sp);
#else
    swapcontext(&g_cur->ctx, &g_sched);
#endif
}

void trampoline(void) {
    g_main();
    g_cur->done = true;
    to_scheduler();
    abort();   // a finished fiber is never resumed
}

void fiber_init(Fiber &f, char *stack) {
    f.done = false;
#if defined(__x86_64__)
    // the frame pbf_fiber_switch pops: six callee-saved registers, then `ret` into trampoline with the stack aligned as
    // after a call (return-address slot 16-byte aligned, so %rsp = 8 mod 16 on entry)
    uintptr_t top = ((uintptr_t)stack + STACK) & ~(uintptr_t)15;
    void **sp = (void **)(top - 16);
    *sp = (void *)trampoline;
    for (int i = 0; i < 6; i++) *--sp = nullptr;
    f.sp = sp;
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stack;
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, trampoline, 0);
#endif
}

void fiber_resume(Fiber &f) {
    g_cur = &f;
#if defined(__x86_64__)
    pbf_fiber_switch(&g_sched_sp, f.sp);
#else
    swapcontext(&g_sched, &f.ctx);
#endif
}

void set_ids(uint group, uint lid, uint local_size) {
    gl_WorkGroupID = uvec3{group, 0, 0};
    gl_LocalInvocationIndex = lid;
    gl_GlobalInvocationID = uvec3{group * local_size + lid, 0, 0};
}

// glDispatchCompute(groups, 1, 1) of a shader whose local size is local_size
void dispatch(void (*shader_main)(void), uint groups, uint local_size, bool has_barrier) {
    if (!has_barrier) {
        for (uint g = 0; g < groups; g++)
            for (uint l = 0; l < local_size; l++) { set_ids(g, l, local_size); shader_main(); }
        return;
    }
    g_barrier_mode = BAR_FIBER;
    g_main = shader_main;
    if (g_fibers.size() < local_size) { g_fibers.resize(local_size); g_stacks.resize((size_t)local_size * STACK + 64); }
    for (uint g = 0; g < groups; g++) {
        for (uint l = 0; l < local_size; l++) fiber_init(g_fibers[l], g_stacks.data() + (size_t)l * STACK);
        uint live = local_size;
        while (live) {                                   // one round = every invocation runs to its next barrier
            for (uint l = 0; l < local_size; l++) {
                Fiber &f = g_fibers[l];
                if (f.done) continue;
                set_ids(g, l, local_size);
                fiber_resume(f);
                if (f.done) live--;
            }
        }
    }
    g_cur = nullptr;
}

}  // namespace

void barrier(void) {
    if (g_barrier_mode == BAR_NOOP) return;
    if (g_barrier_mode == BAR_ABORT) longjmp(g_abort, 1);
    to_scheduler();
}

// ---- the objects SPH owns ---------------------------------------------------------------------------------------------
struct RefSim {
    uint n;
    int gx, gy, gz;
    // SPH (src/SPH.cpp:96-133)
    std::vector<float> position, velocity, lambda, vorticity;
    std::vector<uint> highlight;
    bool extforce = false;
    // RadixSort(512, n >> 9, grid) (src/RadixSort.cpp:32-109)
    uint blocksize, numblocks, numbits;
    std::vector<float> sortbuf[2];         // buffer / result, swapped after every pass
    int cur = 0;                           // sortbuf[cur] = RadixSort::buffer
    std::vector<uint> prefixsums;
    std::vector<std::vector<uint>> blocksums;
    // NeighbourCellFinder (src/NeighbourCellFinder.cpp:24-100)
    std::vector<int> gridtexture, gridendtexture, neighbourcells;
};

namespace {

unsigned count_sortbits(uint64_t v) {   // src/RadixSort.cpp:24-30
    unsigned r = 1;
    while (v >>= 1) r++;
    return r;
}

uint32_t intpow(uint32_t x, uint32_t y) {   // src/RadixSort.cpp:141-152
    uint32_t r = 1;
    while (y) {
        if (y & 1) r *= x;
        y >>= 1;
        x *= x;
    }
    return r;
}

void bind_tex_buffer(int unit, void *data, int comps, long n) { gl_state.tex[unit] = GlUnit{data, comps, n, {0, 0, 0}}; }
void bind_img_buffer(int unit, void *data, int comps, long n) { gl_state.img[unit] = GlUnit{data, comps, n, {0, 0, 0}}; }
void bind_tex_3d(int unit, RefSim *s, std::vector<int> &t) { gl_state.tex[unit] = GlUnit{t.data(), 1, (long)t.size(), {s->gx, s->gy, s->gz}}; }
void bind_img_3d(int unit, RefSim *s, std::vector<int> &t) { gl_state.img[unit] = GlUnit{t.data(), 1, (long)t.size(), {s->gx, s->gy, s->gz}}; }

void set_grid(RefSim *s) {
    g_GRID_SIZE = vec3((float)s->gx, (float)s->gy, (float)s->gz);
    g_GRID_HASHWEIGHTS = ivec3(1, s->gx * s->gz, s->gx);
}

// RadixSort::SortBits, src/RadixSort.cpp:154-200
void sort_bits(RefSim *s, int bits) {
    sh_counting::bitshift = bits;
    sh_globalsort::bitshift = bits;
    float *buffer = s->sortbuf[s->cur].data(), *result = s->sortbuf[s->cur ^ 1].data();
    gl_state.ssbo[0] = buffer; gl_state.ssbo[1] = s->prefixsums.data(); gl_state.ssbo[2] = s->blocksums.front().data();
    gl_state.ssbo[3] = result;
    dispatch(sh_counting::shader_main, s->numblocks, sh_counting::local_size_x, true);
    uint32_t numblocksums = (4 * s->numblocks) / s->blocksize;
    for (int i = 0; i < (int)s->blocksums.size() - 1; i++) {
        gl_state.ssbo[0] = s->blocksums[i].data(); gl_state.ssbo[1] = s->blocksums[i + 1].data();
        dispatch(sh_blockscan::shader_main, numblocksums > 0 ? numblocksums : 1, sh_blockscan::local_size_x, true);
        numblocksums /= s->blocksize;
    }
    for (int i = (int)s->blocksums.size() - 3; i >= 0; i--) {
        uint32_t nb = (4 * s->numblocks) / intpow(s->blocksize, i + 1);
        gl_state.ssbo[0] = s->blocksums[i].data(); gl_state.ssbo[1] = s->blocksums[i + 1].data();
        dispatch(sh_addblocksum::shader_main, nb > 0 ? nb : 1, sh_addblocksum::local_size_x, false);
    }
    gl_state.ssbo[0] = buffer; gl_state.ssbo[1] = s->prefixsums.data();
    dispatch(sh_globalsort::shader_main, s->numblocks, sh_globalsort::local_size_x, false);
}

void bind_solver(RefSim *s) {   // src/SPH.cpp:280-286
    gl_state.ssbo[1] = s->sortbuf[s->cur].data();
    bind_tex_buffer(2, s->neighbourcells.data(), 4, 3L * s->n);
    bind_tex_buffer(3, s->lambda.data(), 1, s->n);
}

}  // namespace

extern "C" {

// SPH::SPH(numparticles, gridsize), src/SPH.cpp:24-151
RefSim *ref_create(uint n, int gx, int gy, int gz) {
    if (n == 0 || (n & 511u)) return nullptr;             // numblocks = n >> 9 blocks of 512 (src/SPH.cpp:25)
    RefSim *s = new RefSim();
    s->n = n; s->gx = gx; s->gy = gy; s->gz = gz;
    s->position.assign(4 * (size_t)n, 0.0f); s->velocity.assign(4 * (size_t)n, 0.0f);
    s->lambda.assign(n, 0.0f); s->vorticity.assign(n, 0.0f); s->highlight.assign(n, 0u);
    s->blocksize = 512; s->numblocks = n >> 9;
    s->numbits = count_sortbits((uint64_t)gx * (uint64_t)gy * (uint64_t)gz - 1);           // src/RadixSort.cpp:44
    s->sortbuf[0].assign(4 * (size_t)n, 0.0f); s->sortbuf[1].assign(4 * (size_t)n, 0.0f);
    s->prefixsums.assign(n, 0u);
    uint32_t numblocksums = 4 * s->numblocks;                                                 // src/RadixSort.cpp:66-85
    const uint32_t bs = s->blocksize;
    int levels = (int)ceil(log((double)(((numblocksums + bs - 1) / bs) * bs)) / log((double)bs));
    levels++;
    s->blocksums.resize(levels);
    for (int i = 0; i < levels; i++) {
        numblocksums = ((numblocksums + bs - 1) / bs) * bs;
        if (numblocksums < 1) numblocksums = 1;
        s->blocksums[i].assign(numblocksums, 0u);                                             // cleared once (:103-108)
        numblocksums /= bs;
    }
    const size_t ncell = (size_t)gx * gy * gz;
    s->gridtexture.assign(ncell, -1);
    s->gridendtexture.assign(ncell, 0);           // never initialised nor cleared by the reference: 0 here
    s->neighbourcells.assign(12 * (size_t)n, 0);
    g_params.one_over_rho_0 = 1.0f; g_params.epsilon = 5.0f; g_params.gravity = 10.0f; g_params.timestep = 0.016f;   // :137-144
    g_params.tensile_instability_k = 0.1f;
    {
        const float r = 0.2f, hh = 2.0f, tmp = hh * hh - r * r;                               // SPH::Wpoly6, :159-164
        g_params.tensile_instability_scale = 1.0f / (1.56668147106f * tmp * tmp * tmp / (hh * hh * hh * hh * hh * hh * hh * hh * hh));
    }
    g_params.xsph_viscosity_c = 0.01f; g_params.vorticity_epsilon = 5.0f;
    gl_state.oob_fetch = 0;
    return s;
}
void ref_destroy(RefSim *s) { delete s; }

void ref_set_params(RefSim *, const float p[8]) { memcpy(&g_params, p, sizeof(g_params)); }
void ref_get_params(RefSim *, float p[8]) { memcpy(p, &g_params, sizeof(g_params)); }
void ref_set_extforce(RefSim *s, int on) { s->extforce = on != 0; }           // SPH::SetExternalForce, :242-244
void ref_set_oob_fetch(RefSim *, int v) { gl_state.oob_fetch = v; }
void ref_upload(RefSim *s, const float *pos4, const float *vel4, const uint *hl) {
    memcpy(s->position.data(), pos4, 16 * (size_t)s->n);
    memcpy(s->velocity.data(), vel4, 16 * (size_t)s->n);
    if (hl) memcpy(s->highlight.data(), hl, 4 * (size_t)s->n);
    else std::fill(s->highlight.begin(), s->highlight.end(), 0u);
}
void ref_download(RefSim *s, float *pos4, float *vel4, uint *hl) {
    if (pos4) memcpy(pos4, s->position.data(), 16 * (size_t)s->n);
    if (vel4) memcpy(vel4, s->velocity.data(), 16 * (size_t)s->n);
    if (hl) memcpy(hl, s->highlight.data(), 4 * (size_t)s->n);
}

// [predictpos] src/SPH.cpp:247-261
void ref_predict(RefSim *s) {
    set_grid(s);
    gl_state.ssbo[1] = s->sortbuf[s->cur].data();
    bind_tex_buffer(0, s->position.data(), 4, s->n);
    bind_tex_buffer(1, s->velocity.data(), 4, s->n);
    sh_predictpos::extforce = s->extforce;
    dispatch(sh_predictpos::shader_main, s->n >> 8, sh_predictpos::local_size_x, false);
}

// RadixSort::Run, src/RadixSort.cpp:124-133
void ref_sort(RefSim *s) {
    set_grid(s);
    const uint nb = s->numblocks;
    sh_counting::blocksumoffsets = uvec4(0, nb, nb * 2, nb * 3);              // src/RadixSort.cpp:92-96
    sh_globalsort::blocksumoffsets = uvec4(0, nb, nb * 2, nb * 3);
    for (uint i = 0; i < (s->numbits + 1) >> 1; i++) {
        sort_bits(s, 2 * (int)i);
        s->cur ^= 1;                                                           // std::swap(result, buffer)
    }
}

// NeighbourCellFinder::FindNeighbourCells, src/NeighbourCellFinder.cpp:113-148
void ref_find_cells(RefSim *s) {
    set_grid(s);
    std::fill(s->gridtexture.begin(), s->gridtexture.end(), -1);               // glClearTexImage(gridtexture, -1)
    gl_state.ssbo[0] = s->sortbuf[s->cur].data();
    bind_img_3d(0, s, s->gridtexture);
    bind_img_3d(1, s, s->gridendtexture);
    dispatch(sh_findcells::shader_main, s->n >> 8, sh_findcells::local_size_x, false);
}
// Restatement policy, NOT reference behaviour: the reference never writes the `end` of the highest-key
// occupied cell in the step that creates it (stale or uninitialised memory); the oracle and the CUDA path define it as N.
void ref_policy_define_last_end(RefSim *s) {
    const float *p = s->sortbuf[s->cur].data() + 4 * (size_t)(s->n - 1);
    const vec3 c = clamp(vec3(p[0], p[1], p[2]), vec3(0, 0, 0), vec3((float)s->gx, (float)s->gy, (float)s->gz));
    const ivec3 g(c);
    if (g.x < s->gx && g.y < s->gy && g.z < s->gz) s->gridendtexture[((size_t)g.z * s->gy + g.y) * s->gx + g.x] = (int)s->n;
}
void ref_neighbour_cells(RefSim *s) {
    set_grid(s);
    gl_state.ssbo[0] = s->sortbuf[s->cur].data();
    bind_tex_3d(0, s, s->gridtexture);
    bind_tex_3d(1, s, s->gridendtexture);
    bind_img_buffer(0, s->neighbourcells.data(), 4, 3L * s->n);
    dispatch(sh_neighbourcells::shader_main, s->n >> 8, sh_neighbourcells::local_size_x, false);
}

// particle highlighting, src/SPH.cpp:287-296
void ref_highlight(RefSim *s) {
    set_grid(s);
    bind_solver(s);
    bind_img_buffer(0, s->highlight.data(), 1, s->n);
    dispatch(sh_clearhighlight::shader_main, s->n >> 8, sh_clearhighlight::local_size_x, false);
    dispatch(sh_highlight::shader_main, s->n >> 8, sh_highlight::local_size_x, false);
}

void ref_calclambda(RefSim *s) {   // src/SPH.cpp:304-307
    set_grid(s);
    bind_solver(s);
    bind_img_buffer(0, s->lambda.data(), 1, s->n);
    dispatch(sh_calclambda::shader_main, s->n >> 8, sh_calclambda::local_size_x, false);
}

void ref_updatepos(RefSim *s, int order) {   // src/SPH.cpp:308-310
    set_grid(s);
    bind_solver(s);
    if (order == 1) {
        dispatch(sh_updatepos::shader_main, s->n >> 8, sh_updatepos::local_size_x, false);
        return;
    }
    float *rec = s->sortbuf[s->cur].data();
    std::vector<float> out(3 * (size_t)s->n);
    const uint ls = sh_updatepos::local_size_x;
    for (uint i = 0; i < s->n; i++) {
        float *p = rec + 4 * (size_t)i;
        const float old[3] = {p[0], p[1], p[2]};
        set_ids(i / ls, i % ls, ls);
        sh_updatepos::shader_main();
        for (int a = 0; a < 3; a++) { out[3 * (size_t)i + a] = p[a]; p[a] = old[a]; }
    }
    for (uint i = 0; i < s->n; i++)
        for (int a = 0; a < 3; a++) rec[4 * (size_t)i + a] = out[3 * (size_t)i + a];
}

void ref_update(RefSim *s) {   // src/SPH.cpp:318-324
    set_grid(s);
    gl_state.ssbo[1] = s->sortbuf[s->cur].data();
    bind_img_buffer(0, s->position.data(), 4, s->n);
    bind_img_buffer(1, s->velocity.data(), 4, s->n);
    dispatch(sh_update::shader_main, s->n >> 8, sh_update::local_size_x, false);
}

void ref_vorticity(RefSim *s, int order) {   // src/SPH.cpp:325-331
    set_grid(s);
    bind_solver(s);
    gl_state.ssbo[3] = s->vorticity.data();
    bind_img_buffer(1, s->velocity.data(), 4, s->n);
    const uint ls = sh_vorticity::local_size_x;
    if (order == 1) {          // as dispatched: work groups in order, barrier() honoured inside each group
        dispatch(sh_vorticity::shader_main, s->n >> 8, ls, true);
        return;
    }
    // Jacobi.  Pass A: every invocation up to its barrier() -> vorticities[] of all particles from the old velocities.
    g_barrier_mode = BAR_ABORT;
    for (uint i = 0; i < s->n; i++) {
        set_ids(i / ls, i % ls, ls);
        if (setjmp(g_abort) == 0) sh_vorticity::shader_main();
    }
    // Pass B: the whole shader per invocation (its first half recomputes the same value), velocity written to a copy.
    g_barrier_mode = BAR_NOOP;
    const float *rec = s->sortbuf[s->cur].data();
    std::vector<float> out(s->velocity);
    for (uint i = 0; i < s->n; i++) {
        const int id = floatBitsToInt(rec[4 * (size_t)i + 3]);
        float *v = s->velocity.data() + 4 * (size_t)id;
        const float old[4] = {v[0], v[1], v[2], v[3]};
        set_ids(i / ls, i % ls, ls);
        sh_vorticity::shader_main();
        for (int a = 0; a < 4; a++) { out[4 * (size_t)id + a] = v[a]; v[a] = old[a]; }
    }
    s->velocity.swap(out);
    g_barrier_mode = BAR_FIBER;
}

// SPH::Run, src/SPH.cpp:246-334
void ref_step(RefSim *s, int iterations, int vorticityconfinement, int order, int define_last_end) {
    ref_predict(s);
    ref_sort(s);
    ref_find_cells(s);
    if (define_last_end) ref_policy_define_last_end(s);
    ref_neighbour_cells(s);
    ref_highlight(s);
    for (int it = 0; it < iterations; it++) {
        ref_calclambda(s);
        ref_updatepos(s, order);
    }
    ref_update(s);
    if (vorticityconfinement) ref_vorticity(s, order);
}

// ---- read-back ------------------------------------------------------------------------------------------------------------
uint ref_numbits(RefSim *s) { return s->numbits; }
int ref_blocksum_levels(RefSim *s) { return (int)s->blocksums.size(); }
void ref_get_records(RefSim *s, float *rec4) { memcpy(rec4, s->sortbuf[s->cur].data(), 16 * (size_t)s->n); }   // RadixSort::GetBuffer
void ref_set_records(RefSim *s, const float *rec4) { memcpy(s->sortbuf[s->cur].data(), rec4, 16 * (size_t)s->n); }
void ref_get_lambda(RefSim *s, float *l) { memcpy(l, s->lambda.data(), 4 * (size_t)s->n); }
void ref_get_vorticity(RefSim *s, float *w) { memcpy(w, s->vorticity.data(), 4 * (size_t)s->n); }
void ref_get_neighbours(RefSim *s, int *words12) { memcpy(words12, s->neighbourcells.data(), 48 * (size_t)s->n); }
// gridtexture / gridendtexture in the oracle's dense order (index x + z*gx + y*gx*gz = the cell hash)
void ref_get_grid(RefSim *s, int *start, int *end) {
    for (int y = 0; y < s->gy; y++)
        for (int z = 0; z < s->gz; z++)
            for (int x = 0; x < s->gx; x++) {
                const size_t img = ((size_t)z * s->gy + y) * s->gx + x, hash = (size_t)x + (size_t)z * s->gx + (size_t)y * s->gx * s->gz;
                if (start) start[hash] = s->gridtexture[img];
                if (end) end[hash] = s->gridendtexture[img];
            }
}

}  // extern "C"
