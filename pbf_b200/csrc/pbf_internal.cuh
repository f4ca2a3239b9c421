// pbf_internal.cuh -- handle layout, kernel-side parameter blocks and launch prototypes shared by the
// translation units of libpbf_b200.so.  sm_100a only; there is no CPU path in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pbf_c.h"

typedef uint32_t u32;

#define PBF_KEY_NOCELL 0x80000000u   // bit 31 of a cell key: the clamped cell lies outside the cell images
#define PBF_RADIX 512   // digit slots of a sort pass (up to 9 bits per pass; sort.cu)

// what the reference injects into every shader as GLSL header constants (src/SPH.cpp:28-60)
struct GridInfo {
    int gx, gy, gz;        // GRID_SIZE
    int gxgz;              // GRID_HASHWEIGHTS.y = gx*gz; hash = x + z*gx + y*gx*gz
    float wlo[3], whi[3];  // clamp bounds of updatepos.glsl:98-100: wall, GRID_SIZE - wall
    int ref_quirks;
    int bx, bz;            // bit widths of the packed home cell (x | z << bx | y << (bx+bz))
    // slab window (multi GPU): this handle's cell images cover global cell layers [zoff, zoff + gz) of a domain that is
    // gz_global layers deep; a single-GPU handle has zoff = 0, gz_global = gz.  Positions are always global.
    int zoff, gz_global;
};

// the std140 SPHParameters block (src/SPH.cpp:46-56) + the extforce uniform (predictpos.glsl:13)
struct SimParams {
    float one_over_rho_0, epsilon, gravity, timestep;
    float tensile_k, tensile_scale, xsph_c, vort_eps;
    int extforce;
    int self_term;          // pbf_options: rho_i includes W(0)
    float restitution;      // pbf_options: < 0 off
    int canonical;          // pbf_set_canonical_order: one order of every floating-point sum on every code path
};

// Fused halo push of the slab runtime (slab.cu): the sweep that produces a halo quantity (lambda, new position, |omega|)
// stores it for its boundary particles straight into the neighbour GPU's mailbox over NVLink, and the block that
// finishes last publishes the refresh number.  All zero on a single-GPU handle.
struct HaloPush {
    const u32 *map;                  // per sorted slot: 0 = not a boundary particle, k+1 = k-th of the z- face, 0x80000000|(k+1) of z+
    char *data[2];                   // the z- / z+ neighbour's mailbox slot for this refresh
    unsigned long long *flag[2];
    unsigned long long seq;
    u32 *done;                       // counter of the blocks that have pushed
    const u32 *expect;               // how many blocks (tiles) hold boundary particles at all: the last of them publishes
    u32 count[2];                    // boundary particles per face (a flag is only published for a non-empty face)
    // device-side counts (slab.cu, one graph per step): the sequence number is 256 * dn[DN_STEP] + e and the per-face counts
    // are dn[DN_HALO_N], dn[DN_HALO_N + 1]; seq and count[] above are then unused, and only the tiles whose descriptor carries
    // PBF_TILE_PUSH look at `map` at all
    const u32 *dn;
    u32 e;
};

// Fused halo pull (slab ranks with device-side counts): the sweep that CONSUMES a halo quantity fetches it itself.  Its
// first blocks wait for the neighbours' refresh number e, copy the values of the ghost particles from this rank's mailbox
// into the sweep's input array and publish `ready`; tiles that touch the halo layers (PBF_TILE_WAIT in their descriptor) wait
// for `ready` before they stage anything, every other tile -- 98 % of them -- starts at once.  The exchange costs no kernel of its own and
// the interior tiles hide its latency and the skew between the ranks.  e = 0: this sweep pulls nothing.
#define PBF_PULL_BLOCKS 32
// flags k_halo_index (slab.cu) ORs into the mode word of a tile descriptor after k_plan wrote it (sweeps.cu, TileCtx::flags)
#define PBF_TILE_PUSH 0x100u         // the tile holds boundary particles
#define PBF_TILE_WAIT 0x200u         // the tile holds particles of the ghost or boundary layers: its candidates may be ghosts
struct HaloPull {
    u32 *dn;                         // DN_STEP, DN_HALO_N + 2 / + 3 = ghosts from z- / z+; DN_OVERFLOW bit 4 = a wait timed out
    const char *data[2];             // this rank's mailbox slot of refresh e: from z- / from z+
    const unsigned long long *flag[2];
    const u32 *ghost_sorted;         // sorted slot of every ghost, z- ghosts first
    float4 *buf;                     // the sweep's input array
    unsigned long long *ready;       // PBF_PULL_BLOCKS local flags
    u32 e;
    int wide;                        // 16-byte records (positions) or 4-byte values into .w (lambda, |omega|)
};

// A particle count as kernels take it.  Single-domain handles know their count on the host: p is null and n is the value.
// A slab rank's count changes every step ON THE DEVICE (migration, ghosts) and the host never waits for it: p points at the
// device-side value and n is only the bound the grid was sized for.  Kernels loop over their tiles / elements up to the
// device value, so a bound that is too small costs time, never particles.
struct NRef {
    u32 n;
    const u32 *p;
};
#ifdef __CUDACC__
__device__ __forceinline__ u32 nref(const NRef &r) { return r.p ? *r.p : r.n; }
#endif

// slots of pbf_sim::dn, the device-side counters of a slab rank (slab.cu)
enum { DN_TOTAL = 0, DN_LOCAL = 1, DN_PREV = 2, DN_STAY = 3, DN_LEAVE = 4 /* lo, hi */, DN_BND = 6 /* lo, hi */,
       DN_ARRIVE = 8 /* lo, hi */, DN_GHOST = 10 /* lo, hi */, DN_MOVERS = 12, DN_HOLES = 13, DN_OVERFLOW = 14, DN_STEP = 15,
       DN_HALO_N = 16 /* boundary + ghost totals for the halo kernels: push lo, push hi, pull lo, pull hi */, DN_WORDS = 24 };

// slab ranks: k_predict also decides which particles left the slab (what k_mark_leavers does as a pass of its own)
struct LeaveArgs {
    int z_lo, z_hi;
    bool has_lo, has_hi;
    u32 *btag, *leave_lo, *leave_hi, *cnt;     // cnt[0..1] leavers lo / hi, cnt[2..3] boundary-layer particles lo / hi
    u32 cap;
    u32 *bnd_lo, *bnd_hi;                      // non-null: also list the particles of the boundary layers z_lo and z_hi - 1
};

struct SortPlan {
    int bits;        // low key bits that take part in the sort = 2*ceil(numbits/2) (src/RadixSort.cpp:127)
    int passes;      // onesweep passes of up to 9 bits
    int shift[4];
    u32 mask[4];
};

struct pbf_sim {
    pbf_config cfg;
    pbf_params params;
    pbf_options options;
    int device;
    int sm_count;
    cudaStream_t stream;
    u32 n;           // particles held (slab rank with device-side counts: the bound its grids are sized for)
    u32 *dn;         // device-side counters (DN_*); n_dev / n_prev_dev point into it on a slab rank, null otherwise
    const u32 *n_dev, *n_prev_dev;
    u32 cap;
    GridInfo grid;
    SortPlan plan;
    size_t ncell;

    // by-id state (SPH::positionbuffer / velocitybuffer / highlightbuffer, src/SPH.cpp:106-133)
    float4 *pos_own, *vel_own; u32 *hl_own;
    float4 *pos, *vel; u32 *hl;          // bound (own or caller's)

    // predicted records by id {p*, id} (RadixSort::buffer as written by predictpos.glsl:37) and their keys
    float4 *pred; u32 *keys;
    // sort scratch
    u32 *ktmp[2], *vtmp[2];
    u32 *skey, *perm;                     // sorted cell keys, particle id per sorted slot
    u32 *home;                            // packed unclamped cell of the predicted position per sorted slot
    u32 *hist, *gbase, *tile_counter, *status;
    u32 max_tiles;
    // cell table {start,end} per cell (gridtexture / gridendtexture, src/NeighbourCellFinder.cpp:59-90)
    int2 *cells;
    int2 *runs3;                          // per cell: cells x-1,x,x+1 merged {start,count} (neighbourcells.glsl:62-84)
    u32 n_prev_sorted;                    // slots of skey that describe the current table contents
    // plan of the tiled sweeps (sweeps.cu): per 256-particle tile the nine sorted-index ranges that hold all its
    // candidates, per particle its nine neighbour runs relative to the tile's shared-memory image
    int *tile_desc; u32 *tile_runs;
    bool canonical;                       // pbf_set_canonical_order (verification mode)
    bool fuse_update;                     // update.glsl in the epilogue of the last delta-p sweep (env PBF_SEPARATE_UPDATE=1: own kernel)
    bool tiled_sweeps;                    // false (env PBF_GENERAL_SWEEPS=1, debugging): every tile takes the general path
    // solver state in sorted order
    float4 *bufA;                         // {x,y,z,-}   positions (Jacobi ping)
    float4 *bufB;                         // {x,y,z,lambda} resp. {x,y,z,|omega|}
    float4 *svel, *vprime, *omega;
    u32 *flags;                           // [0] = any highlight bit0 set this step
    double *diag;                         // [0] density error sum, [1] kinetic energy

    // graph replay of the whole step
    cudaGraph_t graph; cudaGraphExec_t graph_exec; bool graph_valid; u32 graph_kernels;
    bool timing; cudaEvent_t ev[6]; bool ev_valid;
    std::vector<cudaEvent_t> ev_solver; int ev_solver_iters;   // timing mode: one event before every solver kernel
    uint64_t launches;
    bool gl_registered; cudaGraphicsResource_t gl_res[3];   // renderer-owned GL buffers (pbf_register_gl_buffers)
    bool gl_mapped;                       // inside a GlScope: pos/vel/hl point at the mapped GL buffers
    // the same protocol for any other owner of the by-id buffers (pbf_register_external_buffers): gl_registered is set,
    // gl_res stay null and map / unmap go through the caller's callbacks
    pbf_map_fn ext_map; pbf_unmap_fn ext_unmap; void *ext_user;
    uint64_t steps;                       // completed SPH::Run calls since creation / the last state load (checkpoints)
    cudaStream_t copy_stream; cudaEvent_t ev_pos, ev_copied;   // pbf_step_host: position read-back under the vorticity kernels
    int stage;                            // 0 idle, 1 predicted, 2 sorted, 3 cells built
    struct pbf_slab_state *slab;          // non-null once pbf_slab_init has run (slab.cu)
};

inline NRef nref_total(const pbf_sim *s) { return NRef{s->n, s->n_dev}; }
inline NRef nref_prev(const pbf_sim *s) { return NRef{s->n_prev_sorted, s->n_prev_dev}; }

struct DeviceGuard {   // every entry point runs on the handle's device and restores the caller's
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// With GL buffers registered (pbf_register_gl_buffers) the by-id state lives in the renderer's buffer objects, which CUDA may
// only touch between map and unmap.  Every entry point that reads or writes pos / vel / hl opens one of these: it maps
// the three resources on the handle's stream and binds the mapped pointers; the destructor unmaps, which orders GL's
// later reads after the work enqueued in between.  Without registered buffers it does nothing.  Nestable.
struct GlScope {
    pbf_sim *s;
    int rc;          // PBF_OK or the error of the map; check before touching the buffers
    bool owner;
    explicit GlScope(pbf_sim *sim);
    ~GlScope();
    GlScope(const GlScope &) = delete;
    GlScope &operator=(const GlScope &) = delete;
};

// error plumbing (api.cu)
void pbf_set_error(const std::string &msg);
#define PBF_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            pbf_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                    \
            return PBF_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// ---- launchers (each returns the number of kernels it enqueued; errors surface at the next PBF_CUDA) ----
// sim_kernels.cu
int launch_fill_tables(pbf_sim *s);
int launch_unclear_cells(pbf_sim *s);
int launch_predict(pbf_sim *s);
int launch_predict_range(pbf_sim *s, u32 first, NRef count, bool with_hist);
int launch_predict_slab(pbf_sim *s, NRef n_local, const LeaveArgs &la);   // predict + table reset + leaver lists
int launch_keys_only(pbf_sim *s, u32 first, u32 count);
int launch_reorder_cells(pbf_sim *s);
int launch_highlight(pbf_sim *s);
int launch_lambda(pbf_sim *s, const HaloPush *push = nullptr, const HaloPull *pull = nullptr);
int launch_delta_p(pbf_sim *s, const HaloPush *push = nullptr, const HaloPull *pull = nullptr);
int launch_delta_p_update(pbf_sim *s, const HaloPush *push = nullptr, const HaloPull *pull = nullptr);   // sweeps.cu: last iteration, K10 fused into the epilogue
int launch_update(pbf_sim *s);
int launch_vorticity(pbf_sim *s);
int launch_vorticity_a(pbf_sim *s, const HaloPush *push = nullptr);
int launch_vorticity_b(pbf_sim *s, const HaloPull *pull = nullptr);
int launch_density_diag(pbf_sim *s);
int launch_kinetic_diag(pbf_sim *s);
int launch_compose_records(pbf_sim *s, float4 *out);
int launch_neighbour_runs(pbf_sim *s, int *run_start, int *run_count);
int launch_pick(pbf_sim *s, const float origin[3], const float dir[3], float radius, unsigned long long *best);
int launch_toggle_highlight(pbf_sim *s, u32 id);
// sweeps.cu
int sweeps_init(void);                   // opt-in shared-memory sizes of the sweep kernels (once per device)
size_t plan_desc_ints(u32 cap);
size_t plan_run_words(u32 cap);
u32 plan_tile_size(void);
u32 plan_desc_stride(void);      // ints per tile descriptor; the mode word is the first                // particles per tile
int launch_plan(pbf_sim *s);
SimParams sim_params(const pbf_sim *s);
// slab.cu
void slab_free(pbf_sim *s);
bool slab_borrows_stream(const pbf_sim *s);
// sort.cu
SortPlan make_sort_plan(int bits);
int sort_init(void);                     // opt-in shared-memory size of the onesweep kernel (once per device)
u32 sort_max_tiles(u32 cap);
int launch_sort_scan(pbf_sim *s);
int launch_sort_passes(pbf_sim *s);
int launch_sort_hist(pbf_sim *s, const u32 *keys, NRef n);
int launch_sort_pairs(pbf_sim *s, const u32 *kin, const u32 *vin, u32 *kout, u32 *vout, u32 n, int bits);
