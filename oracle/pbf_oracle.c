/*
 * pbf_oracle.c -- CPU restatement of ekpyron/pbf's per-timestep PBF simulation (SPH::Run).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product library (pbf_b200/) may include, link or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * as the checker or as the reported CPU baseline -- never as the thing shipped.
 *
 * PARITY PINNED by oracle/_ref: the reference has no tests or golden vectors of its own, and its GLSL cannot run here
 * (no GL), but its compute shaders are plain C-like code: oracle/ref_harness.cpp compiles the files of
 * /root/reference/shaders/{sph,radixsort,neighbourcellfinder} VERBATIM with g++ behind oracle/glsl_compat.h (work-group
 * barriers as fibers) and runs them through SPH::Run's own dispatch sequence.  tests/test_oracle_ref.py checks this file
 * against that library stage by stage and over whole runs -- bit for bit, floats included -- and against the golden files
 * it minted (tests/golden/ref_*.npz: BASELINE configs[0] over 100 steps, the reference's own scene).  What remains a
 * policy rather than reference behaviour is listed below ((ii), (iii) second sentence, (iv)).
 *
 * All arithmetic is IEEE binary32 in the source order of the GLSL, compiled with -ffp-contract=off.
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Restatement policies:
 *  (i)   no FMA contraction, no reassociation;
 *  (ii)  updatepos / vorticity use Jacobi semantics (all reads see the pre-kernel values); the reference
 *        updates in place and is order dependent (shaders/sph/updatepos.glsl:53-55, vorticity.glsl:70-71);
 *  (iii) ref_quirks=1 reproduces findcells.glsl:39-43 exactly (sorted particle 0 never writes its cell's
 *        start; start[(0,0,0)] = 0 is always written); ref_quirks=0 writes start[cell(0)] = 0 instead.
 *        In both modes end[cell(N-1)] = N (the reference leaves it stale: undefined behaviour);
 *  (iv)  out-of-grid texel fetches read "empty" (start = -1);
 *  (v)   the cell hash is evaluated in integers.  The reference's sort shaders compute it as uint(dot(ivec3, ivec3))
 *        (counting.glsl:53-57, globalsort.glsl:50-55), and GLSL's dot is a FLOAT operation: identical while the hash is
 *        < 2^24, rounded beyond (grids of more than 2^24 cells: the low bits of x are lost for large y, neighbouring
 *        cells share a sort key and the sort is no longer a sort by cell).  ref_quirks bit 1 (value 2 or 3) reproduces
 *        that literally -- x*1 + y*(gx*gz) + z*gx in binary32, left to right -- so that the oracle can be pinned to the
 *        compiled shaders on such grids too (tests/test_oracle_ref.py); the product implements the integer hash.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int gx, gy, gz;        /* GRID_SIZE (src/SPH.h:40 default 128,64,128)                        */
    float wall_x, wall_y, wall_z; /* wall offsets, shaders/sph/updatepos.glsl:98 = (16,0,16)     */
    int ref_quirks;        /* bit 0: policy (iii); bit 1: policy (v), the sort key through a float dot */
} ora_grid;

typedef struct {           /* src/SPH.h:252-285 sphparams_t, same order                          */
    float one_over_rho_0, epsilon, gravity, timestep;
    float tensile_instability_k, tensile_instability_scale, xsph_viscosity_c, vorticity_epsilon;
} ora_params;

#define H_ 2.0f            /* src/SPH.cpp:58 "const float h = 2.0;"                              */
#define KEY_NOCELL 0x80000000u

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ora_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- SPH::Wpoly6, src/SPH.cpp:159-164 (host) == shaders/sph/calclambda.glsl:41-47 ------------------- */
float ora_wpoly6(float r, float h) {
    if (r > h) return 0.0f;
    float tmp = h * h - r * r;
    return 1.56668147106f * tmp * tmp * tmp / (h * h * h * h * h * h * h * h * h);
}

/* default parameters, src/SPH.cpp:137-144 */
void ora_default_params(ora_params *p) {
    p->one_over_rho_0 = 1.0f;
    p->epsilon = 5.0f;
    p->gravity = 10.0f;
    p->timestep = 0.016f;
    p->tensile_instability_k = 0.1f;
    p->tensile_instability_scale = 1.0f / ora_wpoly6(0.2f, 2.0f);
    p->xsph_viscosity_c = 0.01f;
    p->vorticity_epsilon = 5.0f;
}

/* shaders/sph/calclambda.glsl:57-64 gradWspiky */
static inline void grad_wspiky(const float r[3], float out[3]) {
    float l = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (l > H_ || l == 0.0f) { out[0] = out[1] = out[2] = 0.0f; return; }
    float tmp = H_ - l;
    float num = (-3.0f * 4.774648292756860f * tmp * tmp);
    float den = (l * H_ * H_ * H_ * H_ * H_ * H_);
    out[0] = num * r[0] / den;
    out[1] = num * r[1] / den;
    out[2] = num * r[2] / den;
}

/* ---- initial state: Simulation::ResetParticleBuffer, src/Simulation.cpp:206-246 ---------------------
 * One block of nx*ny*nz particles, loop order x, z, y (y innermost), id increasing; position =
 * origin + sign*spacing*(x,y,z) + 0.01*(u-0.5) per axis.  The reference seeds rand() with time(NULL)
 * (:215, non reproducible); u comes from a counter based hash of (seed, 3*id+axis) instead.          */
static inline uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static inline float jitter01(uint32_t seed, uint32_t ctr) {
    uint32_t h = mix32(seed ^ (ctr * 0x9E3779B9u));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
void ora_dam_break(int nx, int ny, int nz, float ox, float oy, float oz, float spacing, int mirror_xz,
                   uint32_t seed, uint32_t id0, float *pos4, float *vel4) {
    uint32_t id = id0;
    for (int x = 0; x < nx; x++)
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++) {
                float fx = mirror_xz ? (float)(-x) : (float)x;
                float fz = mirror_xz ? (float)(-z) : (float)z;
                float px = ox + spacing * fx, py = oy + spacing * (float)y, pz = oz + spacing * fz;
                px += 0.01f * (jitter01(seed, 3u * id + 0u) - 0.5f);
                py += 0.01f * (jitter01(seed, 3u * id + 1u) - 0.5f);
                pz += 0.01f * (jitter01(seed, 3u * id + 2u) - 0.5f);
                float *p = pos4 + 4 * (size_t)(id - id0), *v = vel4 + 4 * (size_t)(id - id0);
                p[0] = px; p[1] = py; p[2] = pz; p[3] = 0.0f;
                v[0] = v[1] = v[2] = v[3] = 0.0f;
                id++;
            }
}

/* ---- K1: shaders/sph/predictpos.glsl:18-38 -------------------------------------------------------- */
void ora_predict(int n, const float *pos4, const float *vel4, const ora_params *P, const ora_grid *G,
                 int extforce, float *rec4 /* {pos, id-as-int-bits} */) {
#pragma omp parallel for schedule(static)
    for (int id = 0; id < n; id++) {
        float p[3] = {pos4[4 * id], pos4[4 * id + 1], pos4[4 * id + 2]};
        float v[3] = {vel4[4 * id], vel4[4 * id + 1], vel4[4 * id + 2]};
        if (extforce && p[2] > (float)G->gz / 2.0f) {          /* :27-28 */
            float f = 2.0f * P->gravity;
            v[0] += f * 0.0f * P->timestep;
            v[1] += f * 0.0f * P->timestep;
            v[2] += f * -1.0f * P->timestep;
        }
        v[0] += P->gravity * 0.0f * P->timestep;               /* :32 */
        v[1] += P->gravity * -1.0f * P->timestep;
        v[2] += P->gravity * 0.0f * P->timestep;
        p[0] += P->timestep * v[0];                            /* :34 */
        p[1] += P->timestep * v[1];
        p[2] += P->timestep * v[2];
        rec4[4 * id] = p[0]; rec4[4 * id + 1] = p[1]; rec4[4 * id + 2] = p[2];
        int32_t iid = id;
        memcpy(&rec4[4 * id + 3], &iid, 4);                    /* :37 */
    }
}

/* ---- cell of a position: shaders/radixsort/counting.glsl:53-57, findcells.glsl:45 -------------------
 * ivec3(clamp(pos, 0, GRID_SIZE)): clamp is inclusive, so a coordinate can equal the grid size.       */
static inline void cell_clamped(const float *p, const ora_grid *G, int c[3]) {
    float gs[3] = {(float)G->gx, (float)G->gy, (float)G->gz};
    for (int a = 0; a < 3; a++) {
        float v = p[a];
        /* GLSL clamp = min(max(x, lo), hi); NaN is not produced on this path */
        v = v < 0.0f ? 0.0f : v;
        v = v > gs[a] ? gs[a] : v;
        c[a] = (int)v; /* truncation */
    }
}
/* full hash = dot(cell, (1, gx*gz, gx)) (src/RadixSort.cpp:36-37); bit 31 flags a cell outside the images */
uint32_t ora_key_of(const float *p, const ora_grid *G) {
    int c[3];
    cell_clamped(p, G, c);
    uint32_t k = (uint32_t)c[0] + (uint32_t)c[2] * (uint32_t)G->gx + (uint32_t)c[1] * (uint32_t)G->gx * (uint32_t)G->gz;
    if (c[0] >= G->gx || c[1] >= G->gy || c[2] >= G->gz) k |= KEY_NOCELL;
    return k;
}
/* policy (v): the sort shaders' GetHash, uint(dot(vec3(cell), vec3(1, gx*gz, gx))) evaluated in binary32 */
static uint32_t sort_key_float_dot(const float *p, const ora_grid *G) {
    int c[3];
    cell_clamped(p, G, c);
    volatile float t = (float)c[0] * 1.0f;                    /* volatile: every operation rounds to binary32 */
    t = t + (float)c[1] * (float)(G->gx * G->gz);
    t = t + (float)c[2] * (float)G->gx;
    return (uint32_t)t;
}

void ora_keys(int n, const float *rec4, const ora_grid *G, uint32_t *keys) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) keys[i] = ora_key_of(rec4 + 4 * (size_t)i, G);
}

/* number of key bits actually sorted: src/RadixSort.cpp:24-30 count_sortbits, :44 numbits, :127 passes */
int ora_sortbits(int gx, int gy, int gz) {
    uint64_t v = (uint64_t)gx * (uint64_t)gy * (uint64_t)gz - 1;
    unsigned r = 1;
    while (v >>= 1) r++;
    return 2 * (int)((r + 1) >> 1);
}

/* ---- RadixSort::Run, src/RadixSort.cpp:124-133 + counting/globalsort shaders ------------------------
 * LSD, 2 bits per pass, stable (globalsort.glsl:62-64: position = digit base + blocks before + prefix
 * inside the block, i.e. index order is preserved inside a digit).  The hash is recomputed from the
 * record's position in every pass as the shaders do.  perm_out[i] = record id that landed at slot i. */
void ora_sort(int n, const float *rec_in, const ora_grid *G, float *rec_out, uint32_t *key_out /* may be NULL */) {
    int passes = ora_sortbits(G->gx, G->gy, G->gz) / 2;
    float *a = (float *)malloc((size_t)n * 16), *b = (float *)malloc((size_t)n * 16);
    uint32_t *ka = (uint32_t *)malloc((size_t)n * 4), *kb = (uint32_t *)malloc((size_t)n * 4);
    memcpy(a, rec_in, (size_t)n * 16);
    ora_keys(n, a, G, ka);
    if (G->ref_quirks & 2) {                                  /* policy (v): sort by the float-dot hash of the shaders */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) ka[i] = sort_key_float_dot(a + 4 * (size_t)i, G);
    }
    int nt = ora_num_threads();
    size_t *hist = (size_t *)malloc(sizeof(size_t) * 4 * (size_t)nt);
    for (int pass = 0; pass < passes; pass++) {
        int shift = 2 * pass;
        memset(hist, 0, sizeof(size_t) * 4 * (size_t)nt);
#pragma omp parallel num_threads(nt)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            size_t lo = (size_t)n * t / nt, hi = (size_t)n * (t + 1) / nt;
            size_t *h = hist + 4 * t;
            for (size_t i = lo; i < hi; i++) h[(ka[i] >> shift) & 3u]++;
#pragma omp barrier
#pragma omp single
            {
                size_t run = 0;
                for (int d = 0; d < 4; d++)
                    for (int tt = 0; tt < nt; tt++) { size_t c = hist[4 * tt + d]; hist[4 * tt + d] = run; run += c; }
            }
            for (size_t i = lo; i < hi; i++) {
                size_t dst = h[(ka[i] >> shift) & 3u]++;
                memcpy(b + 4 * dst, a + 4 * i, 16);
                kb[dst] = ka[i];
            }
        }
        float *tf = a; a = b; b = tf;
        uint32_t *tk = ka; ka = kb; kb = tk;
    }
    memcpy(rec_out, a, (size_t)n * 16);
    if (key_out) memcpy(key_out, ka, (size_t)n * 4);
    free(a); free(b); free(ka); free(kb); free(hist);
}

/* ---- K6: shaders/neighbourcellfinder/findcells.glsl:34-53, clear src/NeighbourCellFinder.cpp:116-126 --
 * start/end are dense gx*gy*gz tables indexed x + z*gx + y*gx*gz (image coordinate (x,y,z)).
 * `end` is NOT cleared (the reference never clears it).  Stores outside the image are dropped.      */
static inline int64_t cell_index(const int c[3], const ora_grid *G) {
    if (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[0] >= G->gx || c[1] >= G->gy || c[2] >= G->gz) return -1;
    return (int64_t)c[0] + (int64_t)c[2] * G->gx + (int64_t)c[1] * G->gx * (int64_t)G->gz;
}
void ora_findcells(int n, const float *rec4, const ora_grid *G, int32_t *start, int32_t *end) {
    size_t ncell = (size_t)G->gx * G->gy * G->gz;
    memset(start, 0xFF, ncell * 4);                                   /* clear to -1 */
    if (n <= 0) return;
    int c0[3];
    if (G->ref_quirks & 1) {
        start[0] = 0;                                                 /* :39-43, thread 0 */
    } else {
        cell_clamped(rec4, G, c0);
        int64_t ci = cell_index(c0, G);
        if (ci >= 0) start[ci] = 0;
    }
    for (int gid = 1; gid < n; gid++) {                               /* serial: write order is irrelevant */
        int a[3], b[3];
        cell_clamped(rec4 + 4 * (size_t)gid, G, a);
        cell_clamped(rec4 + 4 * (size_t)(gid - 1), G, b);
        if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2]) {
            int64_t ia = cell_index(a, G), ib = cell_index(b, G);
            if (ia >= 0) start[ia] = gid;
            if (ib >= 0) end[ib] = gid;
        }
    }
    /* policy (iii): define the end of the highest-key occupied cell */
    cell_clamped(rec4 + 4 * (size_t)(n - 1), G, c0);
    int64_t cl = cell_index(c0, G);
    if (cl >= 0) end[cl] = n;
}

/* ---- K7: shaders/neighbourcellfinder/neighbourcells.glsl:52-91 ----------------------------------------
 * 9 runs per particle in the order of gridoffsets (:37-47); run = first existing start among cells
 * x-1,x,x+1 and the sum of their sizes.  The reference packs `cell + (entries<<24)` (:84); here start and
 * count are kept as two int32 (run_start[9n], run_count[9n]); ora_pack_run gives the packed word.      */
static const int GRIDOFF[9][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 0}, {0, 1}, {1, -1}, {1, 0}, {1, 1}}; /* (dy,dz) */
void ora_neighbourcells(int n, const float *rec4, const ora_grid *G, const int32_t *start, const int32_t *end,
                        int32_t *run_start, int32_t *run_count) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float *p = rec4 + 4 * (size_t)i;
        int g[3] = {(int)p[0], (int)p[1], (int)p[2]};               /* :57 ivec3(pos), NOT clamped */
        for (int o = 0; o < 9; o++) {
            int entries = 0, cell = -1;
            for (int j = -1; j <= 1; j++) {
                int c3[3] = {g[0] + j, g[1] + GRIDOFF[o][0], g[2] + GRIDOFF[o][1]};
                int64_t ci = cell_index(c3, G);
                int c = ci >= 0 ? start[ci] : -1;                     /* policy (iv) */
                if (cell == -1) cell = c;
                if (c != -1) entries += end[ci] - c;
            }
            run_start[9 * (size_t)i + o] = cell;
            run_count[9 * (size_t)i + o] = cell == -1 ? 0 : entries;
        }
    }
}
int32_t ora_pack_run(int32_t cell, int32_t entries) { return cell + (int32_t)((uint32_t)entries << 24); }

/* FOR_EACH_NEIGHBOUR, shaders/sph/foreachneighbour.glsl:1-10: runs in order, ascending index, self skipped */
#define FOR_EACH_NEIGHBOUR(i, j)                                                     \
    for (int o_ = 0; o_ < 9; o_++) {                                                 \
        int s_ = run_start[9 * (size_t)(i) + o_], e_ = s_ + run_count[9 * (size_t)(i) + o_]; \
        for (int j = s_; j < e_; j++) {                                              \
            if (j != (i)) {
#define END_FOR_EACH_NEIGHBOUR }}}

/* ---- K8: shaders/sph/calclambda.glsl:66-103 ------------------------------------------------------- */
void ora_calclambda(int n, const float *rec4, const int32_t *run_start, const int32_t *run_count,
                    const ora_params *P, float *lambda, float *rho_out /* may be NULL */) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {
        const float *pi = rec4 + 4 * (size_t)i;
        float sum_k_grad_Ci = 0.0f, rho = 0.0f;
        float grad_pi_Ci[3] = {0.0f, 0.0f, 0.0f};
        FOR_EACH_NEIGHBOUR(i, j)
            const float *pj = rec4 + 4 * (size_t)j;
            float d[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);   /* distance() */
            rho += ora_wpoly6(len, H_);
            float g[3];
            grad_wspiky(d, g);
            g[0] *= P->one_over_rho_0; g[1] *= P->one_over_rho_0; g[2] *= P->one_over_rho_0;
            sum_k_grad_Ci += g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
            grad_pi_Ci[0] += g[0]; grad_pi_Ci[1] += g[1]; grad_pi_Ci[2] += g[2];
        END_FOR_EACH_NEIGHBOUR
        sum_k_grad_Ci += grad_pi_Ci[0] * grad_pi_Ci[0] + grad_pi_Ci[1] * grad_pi_Ci[1] + grad_pi_Ci[2] * grad_pi_Ci[2];
        float C_i = rho * P->one_over_rho_0 - 1.0f;
        lambda[i] = -C_i / (sum_k_grad_Ci + P->epsilon);
        if (rho_out) rho_out[i] = rho;
    }
}

/* ---- K9: shaders/sph/updatepos.glsl:43-105, Jacobi (policy ii): reads rec_in, writes rec_out ---------- */
void ora_updatepos(int n, const float *rec_in, const int32_t *run_start, const int32_t *run_count,
                   const float *lambda, const ora_params *P, const ora_grid *G, float *rec_out) {
    float lo[3] = {0.0f + G->wall_x, 0.0f + G->wall_y, 0.0f + G->wall_z};
    float hi[3] = {(float)G->gx - G->wall_x, (float)G->gy - G->wall_y, (float)G->gz - G->wall_z};
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {
        const float *pi = rec_in + 4 * (size_t)i;
        float deltap[3] = {0.0f, 0.0f, 0.0f};
        float lam = lambda[i];
        FOR_EACH_NEIGHBOUR(i, j)
            const float *pj = rec_in + 4 * (size_t)j;
            float d[3] = {pi[0] - pj[0], pi[1] - pj[1], pi[2] - pj[2]};
            float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            float scorr = P->tensile_instability_scale * ora_wpoly6(len, H_);   /* :57-60 */
            scorr *= scorr;
            scorr *= scorr;
            scorr = -P->tensile_instability_k * scorr;
            float lam_j = lambda[j];
            float g[3];
            grad_wspiky(d, g);
            float f = lam + lam_j + scorr;
            deltap[0] += f * g[0]; deltap[1] += f * g[1]; deltap[2] += f * g[2];   /* :65 */
        END_FOR_EACH_NEIGHBOUR
        float *po = rec_out + 4 * (size_t)i;
        for (int a = 0; a < 3; a++) {
            float v = pi[a] + P->one_over_rho_0 * deltap[a];                     /* :85 */
            v = v < lo[a] ? lo[a] : v;                                             /* :100 clamp */
            v = v > hi[a] ? hi[a] : v;
            po[a] = v;
        }
        po[3] = pi[3];                                                             /* id untouched */
    }
}

/* ---- K10: shaders/sph/update.glsl:16-28 ------------------------------------------------------------ */
void ora_update(int n, const float *rec4, const ora_params *P, float *pos4, float *vel4) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float *r = rec4 + 4 * (size_t)i;
        int32_t id;
        memcpy(&id, r + 3, 4);
        float *p = pos4 + 4 * (size_t)id, *v = vel4 + 4 * (size_t)id;
        float vx = (r[0] - p[0]) / P->timestep, vy = (r[1] - p[1]) / P->timestep, vz = (r[2] - p[2]) / P->timestep;
        p[0] = r[0]; p[1] = r[1]; p[2] = r[2]; p[3] = 0.0f;
        v[0] = vx; v[1] = vy; v[2] = vz; v[3] = 0.0f;
    }
}

/* ---- K11: shaders/sph/vorticity.glsl:34-86, Jacobi two-phase (policy ii) ----------------------------- */
void ora_vorticity(int n, const float *rec4, const int32_t *run_start, const int32_t *run_count,
                   const ora_params *P, float *vel4 /* by id, in/out */, float *vort_mag /* n, sorted index */) {
    float *vnew = (float *)malloc((size_t)n * 12), *omega = (float *)malloc((size_t)n * 12);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {                                                 /* sweep 1, :41-60 */
        const float *ki = rec4 + 4 * (size_t)i;
        int32_t id; memcpy(&id, ki + 3, 4);
        float vel[3] = {vel4[4 * (size_t)id], vel4[4 * (size_t)id + 1], vel4[4 * (size_t)id + 2]};
        float v[3] = {0, 0, 0}, w[3] = {0, 0, 0};
        FOR_EACH_NEIGHBOUR(i, j)
            const float *kj = rec4 + 4 * (size_t)j;
            int32_t idj; memcpy(&idj, kj + 3, 4);
            float vij[3] = {vel4[4 * (size_t)idj] - vel[0], vel4[4 * (size_t)idj + 1] - vel[1], vel4[4 * (size_t)idj + 2] - vel[2]};
            float pij[3] = {ki[0] - kj[0], ki[1] - kj[1], ki[2] - kj[2]};
            float tmp = ora_wpoly6(sqrtf(pij[0] * pij[0] + pij[1] * pij[1] + pij[2] * pij[2]), H_);
            v[0] += vij[0] * tmp; v[1] += vij[1] * tmp; v[2] += vij[2] * tmp;
            float g[3];
            grad_wspiky(pij, g);
            w[0] += vij[1] * g[2] - g[1] * vij[2];                                /* cross(v_ij, g) */
            w[1] += vij[2] * g[0] - g[2] * vij[0];
            w[2] += vij[0] * g[1] - g[0] * vij[1];
        END_FOR_EACH_NEIGHBOUR
        for (int a = 0; a < 3; a++) { vnew[3 * (size_t)i + a] = vel[a] + P->xsph_viscosity_c * v[a]; omega[3 * (size_t)i + a] = w[a]; }
        vort_mag[i] = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);            /* :60 */
    }
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {                                                 /* sweep 2, :65-85 */
        const float *ki = rec4 + 4 * (size_t)i;
        int32_t id; memcpy(&id, ki + 3, 4);
        float gv[3] = {0, 0, 0};
        FOR_EACH_NEIGHBOUR(i, j)
            const float *kj = rec4 + 4 * (size_t)j;
            float pij[3] = {ki[0] - kj[0], ki[1] - kj[1], ki[2] - kj[2]};
            float g[3];
            grad_wspiky(pij, g);
            gv[0] += vort_mag[j] * g[0]; gv[1] += vort_mag[j] * g[1]; gv[2] += vort_mag[j] * g[2];
        END_FOR_EACH_NEIGHBOUR
        float l = sqrtf(gv[0] * gv[0] + gv[1] * gv[1] + gv[2] * gv[2]);
        if (l > 0.0f) { gv[0] /= l; gv[1] /= l; gv[2] /= l; }
        const float *w = omega + 3 * (size_t)i;
        float c[3] = {gv[1] * w[2] - w[1] * gv[2], gv[2] * w[0] - w[2] * gv[0], gv[0] * w[1] - w[0] * gv[1]}; /* cross(N, w) */
        float s = P->timestep * P->vorticity_epsilon;
        float *vo = vel4 + 4 * (size_t)id;
        vo[0] = vnew[3 * (size_t)i] + s * c[0];
        vo[1] = vnew[3 * (size_t)i + 1] + s * c[1];
        vo[2] = vnew[3 * (size_t)i + 2] + s * c[2];
        vo[3] = 0.0f;
    }
    free(vnew); free(omega);
}

/* ---- K12: shaders/sph/clearhighlight.glsl:4-7, highlight.glsl:17-30 ---------------------------------- */
void ora_highlight(int n, const float *rec4, const int32_t *run_start, const int32_t *run_count, uint32_t *hl) {
    for (int i = 0; i < n; i++) hl[i] &= 1u;
    for (int i = 0; i < n; i++) {
        int32_t id; memcpy(&id, rec4 + 4 * (size_t)i + 3, 4);
        if ((hl[id] & 1u) == 1u) {
            FOR_EACH_NEIGHBOUR(i, j)
                int32_t idj; memcpy(&idj, rec4 + 4 * (size_t)j + 3, 4);
                hl[idj] |= 2u;
            END_FOR_EACH_NEIGHBOUR
        }
    }
}

/* ---- diagnostics (not in the reference; north star's long-run traces) -------------------------------- */
double ora_kinetic_energy(int n, const float *vel4) {
    double e = 0.0;
#pragma omp parallel for reduction(+ : e)
    for (int i = 0; i < n; i++) {
        const float *v = vel4 + 4 * (size_t)i;
        e += 0.5 * ((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
    }
    return e;
}
double ora_density_error(int n, const float *rho, const ora_params *P) {
    double e = 0.0;
#pragma omp parallel for reduction(+ : e)
    for (int i = 0; i < n; i++) e += fabs((double)(rho[i] * P->one_over_rho_0) - 1.0);
    return n ? e / n : 0.0;
}

/* ---- SPH::Run, src/SPH.cpp:246-334: one full step with persistent scratch --------------------------- */
typedef struct {
    int n; ora_grid G;
    float *rec, *rec_sorted, *rec_tmp, *lambda, *vort, *rho;
    int32_t *start, *end, *run_start, *run_count;
    uint32_t *skey;
} ora_sim;

ora_sim *ora_sim_create(int n, const ora_grid *G) {
    ora_sim *s = (ora_sim *)calloc(1, sizeof(ora_sim));
    size_t ncell = (size_t)G->gx * G->gy * G->gz;
    s->n = n; s->G = *G;
    s->rec = (float *)malloc((size_t)n * 16); s->rec_sorted = (float *)malloc((size_t)n * 16);
    s->rec_tmp = (float *)malloc((size_t)n * 16);
    s->lambda = (float *)calloc(n, 4); s->vort = (float *)calloc(n, 4); s->rho = (float *)calloc(n, 4);
    s->start = (int32_t *)malloc(ncell * 4); s->end = (int32_t *)calloc(ncell, 4);
    s->run_start = (int32_t *)malloc((size_t)n * 36); s->run_count = (int32_t *)malloc((size_t)n * 36);
    s->skey = (uint32_t *)malloc((size_t)n * 4);
    return s;
}
void ora_sim_destroy(ora_sim *s) {
    free(s->rec); free(s->rec_sorted); free(s->rec_tmp); free(s->lambda); free(s->vort); free(s->rho);
    free(s->start); free(s->end); free(s->run_start); free(s->run_count); free(s->skey); free(s);
}
/* accessors for the tests */
float *ora_sim_sorted(ora_sim *s) { return s->rec_sorted; }
float *ora_sim_predicted(ora_sim *s) { return s->rec; }
float *ora_sim_lambda(ora_sim *s) { return s->lambda; }
float *ora_sim_rho(ora_sim *s) { return s->rho; }
float *ora_sim_vorticity(ora_sim *s) { return s->vort; }
int32_t *ora_sim_start(ora_sim *s) { return s->start; }
int32_t *ora_sim_end(ora_sim *s) { return s->end; }
int32_t *ora_sim_run_start(ora_sim *s) { return s->run_start; }
int32_t *ora_sim_run_count(ora_sim *s) { return s->run_count; }
uint32_t *ora_sim_skey(ora_sim *s) { return s->skey; }

void ora_sim_step(ora_sim *s, const ora_params *P, int iterations, int vorticity, int extforce,
                  float *pos4, float *vel4, uint32_t *hl /* may be NULL */) {
    int n = s->n;
    ora_predict(n, pos4, vel4, P, &s->G, extforce, s->rec);                       /* SPH.cpp:247-261 */
    ora_sort(n, s->rec, &s->G, s->rec_sorted, s->skey);                           /* :263-268 */
    ora_findcells(n, s->rec_sorted, &s->G, s->start, s->end);                     /* :270-275 */
    ora_neighbourcells(n, s->rec_sorted, &s->G, s->start, s->end, s->run_start, s->run_count);
    if (hl) ora_highlight(n, s->rec_sorted, s->run_start, s->run_count, hl);     /* :288-296 */
    for (int it = 0; it < iterations; it++) {                                     /* :303-311 */
        ora_calclambda(n, s->rec_sorted, s->run_start, s->run_count, P, s->lambda, s->rho);
        ora_updatepos(n, s->rec_sorted, s->run_start, s->run_count, s->lambda, P, &s->G, s->rec_tmp);
        float *t = s->rec_sorted; s->rec_sorted = s->rec_tmp; s->rec_tmp = t;
    }
    ora_update(n, s->rec_sorted, P, pos4, vel4);                                  /* :318-324 */
    if (vorticity) ora_vorticity(n, s->rec_sorted, s->run_start, s->run_count, P, vel4, s->vort); /* :325-331 */
}
