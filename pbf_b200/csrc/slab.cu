// slab.cu -- multi-GPU runtime: 1-D slab decomposition along z with one-cell-layer halos and particle migration.
//
// The reference is single GPU; this is the north star's "slab-decomposed across the 8xB200 box, with
// halo particle exchange and migration over NVLink via NCCL send/recv between solver iterations".  Facts it rests on:
//  * the neighbour stencil is exactly +-1 cell (neighbourcells.glsl:37-50), so ONE ghost layer per slab face reproduces
//    the single-GPU arithmetic;
//  * the cell key is y-major with x minor (GRID_HASHWEIGHTS = (1, gx*gz, gx), src/SPH.cpp:31), so cutting along z keeps
//    every 3-cell x-window contiguous and the merged-run machinery of sim_kernels.cu works unchanged inside a slab.
//
// A rank owns the global cell layers [z_lo, z_hi); its cell tables cover [z_lo-1, z_hi+1) (GridInfo::zoff).  Per step:
//   predict -> migrate (particles whose predicted cell left the slab move to the +-1 neighbour, 64 B records)
//           -> ghosts  (owners send the predicted + old position of their boundary-layer particles, 32 B records;
//                       the receiver appends them after its local particles and sorts them with everything else)
//           -> sort, cells -> K x [lambda, halo lambda (4 B), delta-p, halo position (16 B)]
//           -> update -> vorticity A, halo |omega| (4 B), vorticity B.
// Ghost slots are computed like any other particle and then overwritten by the owner's values, so every kernel of the
// single-GPU path is reused as is.
//
// Two implementations of the step share the kernels:
//  * device-side counts (default with the peer-memory transport, "slab_step_dev" below): migration and ghost records are
//    stored straight into the neighbour's inbox over NVLink (CUDA IPC), every count lives in pbf_sim::dn, kernels loop up to
//    the device value (NRef), the halo refreshes run inside the sweeps (HaloPush / HaloPull, sweeps.cu) and the whole step
//    replays as ONE CUDA graph -- the host never waits for the device;
//  * count read-back ("slab_step", the round-1 step, kept as the baseline and for the NCCL transport): two host
//    synchronisations per step size the transfers and the launches.
//
// Transport: peer-memory stores with flags (pbf_slab_p2p_connect), NCCL send/recv on the handle's stream (one process per
// GPU, communicator bootstrapped from a unique id the host runtime broadcasts), or -- for tests on one GPU -- "virtual
// ranks": several handles of one process stepped in lock step on one stream, mailboxes reached through plain pointers.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "pbf_internal.cuh"

#define LEAVER 0xffffffffu

struct MigRec {   // 64 B
    float4 pos, vel, pred;
    u32 gid, hl, pad0, pad1;
};
struct GhostRec {   // 32 B
    float4 pred, old;
};

struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};

struct pbf_slab_state {
    int rank, nranks, z_lo, z_hi;
    bool has[2];                 // neighbour on the z- / z+ side
    u32 halo_cap;
    ncclComm_t comm;
    pbf_sim **group;             // virtual ranks (same process, same stream); null with NCCL
    int group_size;
    u32 n_local, n_ghost[2], n_bnd[2];
    u32 *gid, *btag;
    u32 *list[4];                // leave lo/hi, boundary lo/hi: slot indices
    u32 *movers, *holes;
    u32 *counters, *h_counters;  // [0..3] list counts, [4] movers, [5] holes, [8..9] incoming (from lo, from hi)
    char *send[2], *recv[2];
    u32 *send_idx[2], *ghost_sorted;
    uint64_t migrated, exchanges, bytes_sent;
    // peer-memory halo refresh (NVLink P2P stores + flags instead of NCCL send/recv): my mailbox, written by my neighbours,
    // and where I write in theirs
    bool p2p;
    char *mbox;                         // [2 sides: from lo, from hi][MB_SLOTS][halo_cap x 16 B], then the flags
    char *peer_data[2];                 // base of the area I fill in the lo / hi neighbour's mailbox
    unsigned long long *peer_flag[2];
    void *ipc_base[2];                  // opened IPC mappings (null for virtual ranks)
    u32 *push_done;                     // last-block counter of k_halo_push / of the fused push
    u32 *push_tiles;                    // [0] sweep tiles that hold boundary particles, [1..] per-tile "seen" flags
    u32 max_tiles;
    u32 *push_map;                      // per sorted slot: which boundary particle of which face (HaloPush::map)
    bool pushed;                        // the sweep just launched has pushed refresh number xseq already
    bool fused;                         // PBF_SLAB_FUSED=1: the producing sweep pushes its halo itself (default: a separate push kernel)
    unsigned long long xseq;            // halo refreshes so far: the same number on every rank
    // ---- step without host round trips (peer-memory transport, csrc/slab.cu "device-side counts") -------------------------
    bool devcount;                      // counts live in pbf_sim::dn; the host only knows bounds
    char *peer_inbox[2];                // base of the record area I fill in the lo / hi neighbour's mailbox
    u32 *rec_done;                      // last-block counters of the record pushes [4]
    u32 *h_ring;                        // pinned: RING x DN_WORDS counters read back after every step, never waited for
    cudaEvent_t ring_ev[4];
    uint64_t ring_step[4];              // step each ring slot belongs to (0 = empty)
    uint64_t step;                      // steps enqueued so far
    u32 bound_local;                    // grid bound for the local particles (pbf_sim::n is the bound incl. ghosts)
    bool bounds_exact;                  // bounds were set from exact host knowledge (upload), not from a read-back
    cudaGraph_t graph; cudaGraphExec_t graph_exec;   // the captured step of this rank (rank 0: of the whole virtual group)
    uint64_t graph_key;
    u32 graph_kernels;
    bool use_graph;
    bool ov_push, ov_pull;              // PBF_SLAB_OVERLAP=2 / 3 (experiments): only the push / only the pull inside the sweeps
    bool overlap;                       // halo refreshes inside the sweeps (fused push + fused pull, sweeps.cu); PBF_SLAB_OVERLAP=0:
                                        // a push and a pull kernel per refresh
    unsigned long long *pull_ready;     // PBF_PULL_BLOCKS flags of the fused pull (HaloPull::ready)
    bool phases;                        // PBF_SLAB_PHASES=1: direct launches with an event at every phase boundary
    cudaEvent_t ph_ev[8];
    bool ph_valid;
    // PBF_SLAB_TRACE=1 (with PBF_SLAB_PHASES=1, debugging): an event after every sweep / refresh of the solver and vorticity
    // phases of rank 0; the intervals of the last step are printed to stderr by pbf_slab_phase_times
    bool trace;
    cudaEvent_t tr_ev[96];
    const char *tr_label[96];
    int tr_n;
};

constexpr int MB_SLOTS = 4;   // 2 would do: a rank pushes refresh e+2 only after it received e+1, which its neighbour
                              // pushed after consuming e (push e+1 follows pull e in stream order)
constexpr int RING = 4;
inline size_t mbox_slot_bytes(const pbf_slab_state *b) { return (size_t)b->halo_cap * 16; }
inline size_t mbox_flags_offset(const pbf_slab_state *b) { return 2 * (size_t)MB_SLOTS * mbox_slot_bytes(b); }
// After the halo slots and their flags: the record inbox.  Per side (0 = written by my lo neighbour, 1 = by my hi
// neighbour) and step parity one area: migration records, ghost records, then a 64-byte header
//   {u32 migrants; u32 ghosts; u64 migrants_ready; u64 ghosts_ready}   (ready = the step number, release/acquire).
struct InboxHeader {
    u32 n_mig, n_ghost;
    unsigned long long mig_ready, ghost_ready;
    unsigned long long pad[5];
};
static_assert(sizeof(InboxHeader) == 64, "inbox header is one 64-byte line");
inline size_t inbox_offset(const pbf_slab_state *b) { return (mbox_flags_offset(b) + 2 * MB_SLOTS * sizeof(unsigned long long) + 255) & ~(size_t)255; }
inline size_t inbox_area_bytes(const pbf_slab_state *b) { return (size_t)b->halo_cap * (sizeof(MigRec) + sizeof(GhostRec)) + sizeof(InboxHeader); }
inline size_t mbox_bytes(const pbf_slab_state *b) { return inbox_offset(b) + 4 * inbox_area_bytes(b); }

namespace {

NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.lib) return PBF_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        pbf_set_error(std::string("slab: cannot load libnccl.so.2: ") + dlerror());
        return PBF_ERR_NCCL;
    }
#define SYM(field, name)                                                          \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                                 \
    if (!g_nccl.field) { pbf_set_error("slab: libnccl lacks " name); return PBF_ERR_NCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return PBF_OK;
}

#define PBF_NCCL(call)                                                                          \
    do {                                                                                        \
        ncclResult_t r_ = (call);                                                               \
        if (r_ != ncclSuccess) {                                                                \
            pbf_set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r_));               \
            return PBF_ERR_NCCL;                                                                \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ int global_layer(float z, const GridInfo &g) {
    return (int)fminf(fmaxf(z, 0.0f), (float)g.gz_global);
}

// which local particles left the slab with their predicted position (predictpos.glsl:34)
__global__ void __launch_bounds__(256)
k_mark_leavers(NRef nr, const float4 *__restrict__ pred, GridInfo g, int z_lo, int z_hi, bool has_lo, bool has_hi,
               u32 *__restrict__ btag, u32 *__restrict__ leave_lo, u32 *__restrict__ leave_hi, u32 *__restrict__ cnt,
               u32 cap) {
    const u32 n = nref(nr);
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const int cz = global_layer(pred[s].z, g);
    u32 tag = 0;
    if (has_lo && cz < z_lo) {
        u32 k = atomicAdd(&cnt[0], 1u);
        if (k < cap) leave_lo[k] = s;
        tag = LEAVER;
    } else if (has_hi && cz >= z_hi) {
        u32 k = atomicAdd(&cnt[1], 1u);
        if (k < cap) leave_hi[k] = s;
        tag = LEAVER;
    }
    btag[s] = tag;
    }
}

__global__ void __launch_bounds__(256)
k_pack_migrants(const u32 *__restrict__ cnt, u32 cap, const u32 *__restrict__ list, const float4 *__restrict__ pos,
                const float4 *__restrict__ vel, const float4 *__restrict__ pred, const u32 *__restrict__ gid,
                const u32 *__restrict__ hl, MigRec *__restrict__ out) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= min(*cnt, cap)) return;
    u32 s = list[k];
    MigRec r;
    r.pos = pos[s]; r.vel = vel[s]; r.pred = pred[s]; r.gid = gid[s]; r.hl = hl[s]; r.pad0 = r.pad1 = 0;
    out[k] = r;
}

// compaction after the leavers are gone: stayers of the tail [n_stay, n) fill the holes below n_stay
__global__ void __launch_bounds__(256)
k_find_movers(u32 n_stay, u32 n, const u32 *__restrict__ btag, u32 *__restrict__ movers, u32 *__restrict__ cnt) {
    u32 s = n_stay + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (btag[s] != LEAVER) movers[atomicAdd(&cnt[4], 1u)] = s;
}
__global__ void __launch_bounds__(256)
k_find_holes(u32 n_stay, u32 count, const u32 *__restrict__ list, u32 *__restrict__ holes, u32 *__restrict__ cnt) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    u32 s = list[k];
    if (s < n_stay) holes[atomicAdd(&cnt[5], 1u)] = s;
}
__global__ void __launch_bounds__(256)
k_fill_holes(const u32 *__restrict__ cnt, const u32 *__restrict__ holes, const u32 *__restrict__ movers, float4 *pos,
             float4 *vel, float4 *pred, u32 *gid, u32 *hl, u32 *keys) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt[5]) return;
    u32 d = holes[k], s = movers[k];
    pos[d] = pos[s]; vel[d] = vel[s]; gid[d] = gid[s]; hl[d] = hl[s]; keys[d] = keys[s];
    float4 p = pred[s];
    p.w = __int_as_float((int)d);
    pred[d] = p;
}

__device__ __forceinline__ u32 window_key(float x, float y, float z, const GridInfo &g) {   // = cell_key of sim_kernels.cu
    int cx = (int)fminf(fmaxf(x, 0.0f), (float)g.gx);
    int cy = (int)fminf(fmaxf(y, 0.0f), (float)g.gy);
    const int czg = global_layer(z, g);
    const int cz = min(max(czg - g.zoff, 0), g.gz);
    u32 k = (u32)cx + (u32)cz * (u32)g.gx + (u32)cy * (u32)g.gxgz;
    if (cx >= g.gx || cy >= g.gy || cz >= g.gz || czg >= g.gz_global || czg < g.zoff) k |= PBF_KEY_NOCELL;
    return k;
}

__global__ void __launch_bounds__(256)
k_unpack_migrants(u32 count, u32 base, const MigRec *__restrict__ in, float4 *pos, float4 *vel, float4 *pred, u32 *gid,
                  u32 *hl, u32 *keys, GridInfo g) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    MigRec r = in[k];
    u32 s = base + k;
    pos[s] = r.pos; vel[s] = r.vel; gid[s] = r.gid; hl[s] = r.hl;
    r.pred.w = __int_as_float((int)s);
    pred[s] = r.pred;
    keys[s] = window_key(r.pred.x, r.pred.y, r.pred.z, g);
}

// boundary layers z_lo and z_hi-1: their particles are the neighbours' ghosts
__global__ void __launch_bounds__(256)
k_mark_boundary(NRef nr, const float4 *__restrict__ pred, GridInfo g, int z_lo, int z_hi, bool has_lo, bool has_hi,
                u32 *__restrict__ btag, u32 *__restrict__ bnd_lo, u32 *__restrict__ bnd_hi, u32 *__restrict__ cnt, u32 cap) {
    const u32 n = nref(nr);
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const int cz = global_layer(pred[s].z, g);
    u32 tag = 0;
    if (has_lo && cz == z_lo) {
        u32 k = atomicAdd(&cnt[2], 1u);
        if (k < cap) { bnd_lo[k] = s; tag = k + 1u; }
    } else if (has_hi && cz == z_hi - 1) {
        u32 k = atomicAdd(&cnt[3], 1u);
        if (k < cap) { bnd_hi[k] = s; tag = 0x80000000u | (k + 1u); }
    }
    btag[s] = tag;
    }
}

__global__ void __launch_bounds__(256)
k_pack_ghosts(u32 count, const u32 *__restrict__ list, const float4 *__restrict__ pred, const float4 *__restrict__ pos,
              const u32 *__restrict__ hl, GhostRec *__restrict__ out) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    u32 s = list[k];
    GhostRec r;
    r.pred = pred[s];
    r.old = pos[s];
    // the selection bit travels with the ghost (position.w is 0 by contract): highlight.glsl:17-30 marks every neighbour of
    // a selected particle, and a neighbour may live on the other side of the plane.  The stencil is symmetric, so it is
    // enough that the OWNER of each particle sees the selected ghosts; marks put on ghost slots are dropped.
    r.old.w = __uint_as_float(hl[s] & 1u);
    out[k] = r;
}

__global__ void __launch_bounds__(256)
k_unpack_ghosts(u32 count, u32 base, const GhostRec *__restrict__ in, float4 *pos, float4 *vel, float4 *pred, u32 *hl,
                u32 *keys, u32 *flags, GridInfo g) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    GhostRec r = in[k];
    u32 s = base + k;
    const u32 selected = __float_as_uint(r.old.w) & 1u;
    r.old.w = 0.0f;
    pos[s] = r.old;                          // so that k_update derives the ghost's velocity like any other particle's
    vel[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    hl[s] = selected;
    if (selected) flags[0] = 1u;             // k_highlight has work even if no local particle is selected
    r.pred.w = __int_as_float((int)s);
    pred[s] = r.pred;
    keys[s] = window_key(r.pred.x, r.pred.y, r.pred.z, g);
}

// after the sort: where did the boundary particles (to pack) and the ghosts (to overwrite) land?
__global__ void __launch_bounds__(256)
k_halo_index(NRef nr, NRef nloc, const u32 *__restrict__ skey, const u32 *__restrict__ perm, const u32 *__restrict__ btag,
             u32 *__restrict__ send_lo, u32 *__restrict__ send_hi, u32 *__restrict__ ghost_sorted, u32 *__restrict__ push_map,
             u32 *__restrict__ push_tiles, u32 tile_size, GridInfo g, int *__restrict__ desc, u32 desc_stride) {
    const u32 n = nref(nr), n_local = nref(nloc);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 kraw = skey[i];
    const u32 k = kraw & ~PBF_KEY_NOCELL;
    const int cz = (int)((k % (u32)g.gxgz) / (u32)g.gx);
    u32 t = 0;
    // interior layers hold neither ghosts nor boundary particles.  The layer decoded from the hash is only exact for keys
    // with a cell: a particle clamped to x = gx hashes like (0, z + 1), so keys without a cell always take the slow path
    const bool edge = (kraw & PBF_KEY_NOCELL) != 0u || !(cz > 1 && cz < g.gz - 2);
    const u32 id = edge ? perm[i] : 0u;
    if (edge && id < n_local) t = btag[id];
    if (t == LEAVER) t = 0;        // only after a capacity overflow (a leaver that could not be sent): reported by the host
    push_map[i] = t;
    if (!edge) continue;
    // desc (refreshes inside the sweeps): flags in the mode word of the tile's descriptor, which every sweep block loads first
    // anyway -- PBF_TILE_WAIT: the tile holds particles of the ghost or boundary layers, so its candidates may be ghosts and it
    // has to wait for the fused pull; PBF_TILE_PUSH: it holds boundary particles and pushes them
    int *tf = desc ? desc + (size_t)(i / tile_size) * desc_stride : nullptr;
    if (tf && ((u32)*tf & PBF_TILE_WAIT) == 0u) atomicOr(tf, (int)PBF_TILE_WAIT);
    if (id >= n_local) { ghost_sorted[id - n_local] = i; continue; }
    if (t == 0) continue;
    if (tf) { if (((u32)atomicOr(tf, (int)PBF_TILE_PUSH) & PBF_TILE_PUSH) == 0u) atomicAdd(&push_tiles[0], 1u); }
    else if (atomicExch(&push_tiles[1 + i / tile_size], 1u) == 0u) atomicAdd(&push_tiles[0], 1u);   // first of its tile
    if (t & 0x80000000u) send_hi[(t & 0x7fffffffu) - 1u] = i;
    else send_lo[t - 1u] = i;
    }
}

__global__ void __launch_bounds__(256)
k_gather_w(u32 n, const u32 *__restrict__ idx, const float4 *__restrict__ buf, float *__restrict__ out) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = buf[idx[k]].w;
}
__global__ void __launch_bounds__(256)
k_scatter_w(u32 n, const u32 *__restrict__ idx, const float *__restrict__ in, float4 *__restrict__ buf) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) buf[idx[k]].w = in[k];
}
__global__ void __launch_bounds__(256)
k_gather_p(u32 n, const u32 *__restrict__ idx, const float4 *__restrict__ buf, float4 *__restrict__ out) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = buf[idx[k]];
}
__global__ void __launch_bounds__(256)
k_scatter_p(u32 n, const u32 *__restrict__ idx, const float4 *__restrict__ in, float4 *__restrict__ buf) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) buf[idx[k]] = in[k];
}

// ---- peer-memory halo refresh ----------------------------------------------------------------------------------------------
// Push: every boundary particle's value goes straight into the neighbour's mailbox over NVLink (plain stores to peer
// memory); the block that finishes last publishes the refresh number in the neighbour's flag (release, system scope).
// Pull: the neighbour's kernel waits for that number (acquire, system scope) and copies mailbox -> ghost slots.
struct HaloSide {
    u32 n;
    const u32 *idx;                 // push: sorted slots of my boundary particles; pull: unused
    char *data;                     // push: the neighbour's mailbox slot; pull: my mailbox slot
    unsigned long long *flag;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
k_halo_push(HaloSide lo, HaloSide hi, const float4 *__restrict__ buf, int wide, unsigned long long seq, u32 *done) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < lo.n + hi.n) {
        const HaloSide &h = k < lo.n ? lo : hi;
        if (k >= lo.n) k -= lo.n;
        const float4 v = buf[h.idx[k]];
        if (wide) reinterpret_cast<float4 *>(h.data)[k] = v;
        else reinterpret_cast<float *>(h.data)[k] = v.w;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        __threadfence_system();
        if (lo.n) st_release_sys(lo.flag, seq);
        if (hi.n) st_release_sys(hi.flag, seq);
    }
}

__global__ void __launch_bounds__(256)
k_halo_pull(HaloSide lo, HaloSide hi, const u32 *__restrict__ ghost_sorted, float4 *__restrict__ buf, int wide,
            unsigned long long seq) {
    if (threadIdx.x == 0) {
        if (lo.n) while (ld_acquire_sys(lo.flag) < seq) __nanosleep(64);
        if (hi.n) while (ld_acquire_sys(hi.flag) < seq) __nanosleep(64);
    }
    __syncthreads();
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lo.n + hi.n) return;
    const u32 i = ghost_sorted[k];                 // ghosts from lo first, then from hi
    const char *src = k < lo.n ? lo.data : hi.data;
    const u32 j = k < lo.n ? k : k - lo.n;
    // the mailbox is written by another GPU: volatile loads bypass L1, which may still hold the slot's previous contents
    if (wide) {
        const volatile float *q = reinterpret_cast<const volatile float *>(src) + 4 * (size_t)j;
        buf[i] = make_float4(q[0], q[1], q[2], q[3]);
    } else {
        buf[i].w = reinterpret_cast<const volatile float *>(src)[j];
    }
}

// ======================================================================================================================
// Device-side counts: the slab step without a single host round trip (peer-memory transport only).
//
// Which particles leave, how many arrive and how many ghosts a rank holds is decided on the device every step; the old
// path read those counts back twice per step to size its NCCL transfers and its launches (2 x cudaStreamSynchronize,
// ~50 direct launches: all of the 1 -> 2 GPU loss).  Here
//   * migration and ghost records go as plain stores into the neighbour's record inbox over NVLink, followed by the count
//     and a release flag that carries the step number; the receiver's unpack kernel spins on the flag (acquire);
//   * every count lives in pbf_sim::dn; kernels get it as NRef (device pointer + the bound the grid was sized for) and
//     loop over their tiles / elements, so a stale bound costs time, never particles;
//   * nothing in the step depends on host knowledge any more, so the whole step is captured as ONE CUDA graph and
//     replayed; the host reads the counters back asynchronously (pinned ring) only to keep the bounds tight.
// ======================================================================================================================
__device__ __forceinline__ InboxHeader *inbox_header(char *area, u32 cap) {
    return reinterpret_cast<InboxHeader *>(area + (size_t)cap * (sizeof(MigRec) + sizeof(GhostRec)));
}
__device__ __forceinline__ MigRec *inbox_mig(char *area) { return reinterpret_cast<MigRec *>(area); }
__device__ __forceinline__ GhostRec *inbox_ghost(char *area, u32 cap) { return reinterpret_cast<GhostRec *>(area + (size_t)cap * sizeof(MigRec)); }
__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) { return *reinterpret_cast<const volatile u32 *>(p); }

// first kernel of a step: the step number (flags carry it) and the stay count of the coming compaction are reset
__global__ void k_step_begin(u32 *dn) {
    if (threadIdx.x == 0) { dn[DN_STEP] += 1u; }
}

// after k_mark_leavers: how many stay.  Leavers beyond the record capacity cannot be sent: overflow (reported by the host)
__global__ void k_counts_leave(u32 *dn, const u32 *__restrict__ cnt, u32 cap, bool has_lo, bool has_hi) {
    if (threadIdx.x != 0) return;
    u32 lo = has_lo ? cnt[0] : 0u, hi = has_hi ? cnt[1] : 0u;
    if (lo > cap || hi > cap) { dn[DN_OVERFLOW] |= 1u; lo = min(lo, cap); hi = min(hi, cap); }
    dn[DN_LEAVE] = lo; dn[DN_LEAVE + 1] = hi;
    dn[DN_STAY] = dn[DN_LOCAL] - lo - hi;
    dn[20] += lo + hi;                                   // particles migrated away so far (pbf_slab_stats)
}

// pack the leavers of one side straight into the neighbour's inbox (its area for this step's parity); the block that
// finishes last publishes the count and the step number
__global__ void __launch_bounds__(256)
k_push_migrants(const u32 *__restrict__ dn, int side, const u32 *__restrict__ list, const float4 *__restrict__ pos,
                const float4 *__restrict__ vel, const float4 *__restrict__ pred, const u32 *__restrict__ gid,
                const u32 *__restrict__ hl, char *peer_area0, size_t area_bytes, u32 cap, u32 *done) {
    const u32 step = dn[DN_STEP], count = dn[DN_LEAVE + side];
    char *area = peer_area0 + (size_t)(step & 1u) * area_bytes;
    MigRec *out = inbox_mig(area);
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const u32 s = list[k];
        MigRec r;
        r.pos = pos[s]; r.vel = vel[s]; r.pred = pred[s]; r.gid = gid[s]; r.hl = hl[s]; r.pad0 = r.pad1 = 0;
        out[k] = r;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        InboxHeader *h = inbox_header(area, cap);
        h->n_mig = count;
        __threadfence_system();
        st_release_sys(&h->mig_ready, (unsigned long long)step);
    }
}

__global__ void __launch_bounds__(256)
k_find_movers_dev(const u32 *__restrict__ dn, const u32 *__restrict__ btag, u32 *__restrict__ movers, u32 *__restrict__ cnt) {
    const u32 n_stay = dn[DN_STAY], n = dn[DN_LOCAL];
    for (u32 s = n_stay + blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
        if (btag[s] != LEAVER) movers[atomicAdd(&cnt[4], 1u)] = s;
}
__global__ void __launch_bounds__(256)
k_find_holes_dev(const u32 *__restrict__ dn, int side, const u32 *__restrict__ list, u32 *__restrict__ holes, u32 *__restrict__ cnt) {
    const u32 n_stay = dn[DN_STAY], count = dn[DN_LEAVE + side];
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const u32 s = list[k];
        if (s < n_stay) holes[atomicAdd(&cnt[5], 1u)] = s;
    }
}
__global__ void __launch_bounds__(256)
k_fill_holes_dev(const u32 *__restrict__ cnt, const u32 *__restrict__ holes, const u32 *__restrict__ movers, float4 *pos,
                 float4 *vel, float4 *pred, u32 *gid, u32 *hl, u32 *keys, u32 *btag, u32 *bnd_lo, u32 *bnd_hi) {
    // holes and movers pair up one to one -- except after a capacity overflow (leavers that could not be listed), which the
    // host reports; until then nothing may be read or written out of bounds
    const u32 count = min(cnt[5], cnt[4]);
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const u32 d = holes[k], s = movers[k];
        pos[d] = pos[s]; vel[d] = vel[s]; gid[d] = gid[s]; hl[d] = hl[s]; keys[d] = keys[s];
        const u32 tag = btag[s];                       // a boundary-layer particle keeps its place in the list k_predict made
        btag[d] = tag;
        if (tag) {
            if (tag & 0x80000000u) bnd_hi[(tag & 0x7fffffffu) - 1u] = d;
            else bnd_lo[tag - 1u] = d;
        }
        float4 p = pred[s];
        p.w = __int_as_float((int)d);
        pred[d] = p;
    }
}

// wait for both neighbours' migrants of this step and append them after the stayers (lo first, then hi)
__global__ void __launch_bounds__(256)
k_pull_migrants(const u32 *__restrict__ dn, char *my_area_lo0, char *my_area_hi0, size_t area_bytes, u32 cap, u32 part_cap,
                bool has_lo, bool has_hi, float4 *pos, float4 *vel, float4 *pred, u32 *gid, u32 *hl, u32 *keys, GridInfo g,
                int z_lo, int z_hi, u32 *btag, u32 *bnd_lo, u32 *bnd_hi, u32 *cnt) {
    const u32 step = dn[DN_STEP];
    char *alo = my_area_lo0 + (size_t)(step & 1u) * area_bytes, *ahi = my_area_hi0 + (size_t)(step & 1u) * area_bytes;
    if (threadIdx.x == 0) {
        if (has_lo) while (ld_acquire_sys(&inbox_header(alo, cap)->mig_ready) < step) __nanosleep(64);
        if (has_hi) while (ld_acquire_sys(&inbox_header(ahi, cap)->mig_ready) < step) __nanosleep(64);
    }
    __syncthreads();
    const u32 in_lo = has_lo ? ld_volatile_u32(&inbox_header(alo, cap)->n_mig) : 0u;
    const u32 in_hi = has_hi ? ld_volatile_u32(&inbox_header(ahi, cap)->n_mig) : 0u;
    const u32 base = dn[DN_STAY];
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < in_lo + in_hi; k += gridDim.x * blockDim.x) {
        const u32 s = base + k;
        if (s >= part_cap) continue;                                   // overflow: k_counts_arrive reports it
        const MigRec *src = k < in_lo ? inbox_mig(alo) + k : inbox_mig(ahi) + (k - in_lo);
        // the records were written by another GPU: ld.cv bypasses L1, which may hold the area's previous contents
        const uint4 *q = reinterpret_cast<const uint4 *>(src);
        uint4 a = __ldcv(q), b4 = __ldcv(q + 1), c = __ldcv(q + 2), d = __ldcv(q + 3);
        MigRec r;
        r.pos = *reinterpret_cast<float4 *>(&a); r.vel = *reinterpret_cast<float4 *>(&b4); r.pred = *reinterpret_cast<float4 *>(&c);
        r.gid = d.x; r.hl = d.y;
        pos[s] = r.pos; vel[s] = r.vel; gid[s] = r.gid; hl[s] = r.hl;
        r.pred.w = __int_as_float((int)s);
        pred[s] = r.pred;
        keys[s] = window_key(r.pred.x, r.pred.y, r.pred.z, g);
        // an arrival that landed in one of my boundary layers is a ghost-to-be of the neighbour on that side
        const int cz = global_layer(r.pred.z, g);
        u32 tag = 0;
        if (has_lo && cz == z_lo) {
            const u32 q2 = atomicAdd(&cnt[2], 1u);
            if (q2 < cap) { bnd_lo[q2] = s; tag = q2 + 1u; }
        } else if (has_hi && cz == z_hi - 1) {
            const u32 q2 = atomicAdd(&cnt[3], 1u);
            if (q2 < cap) { bnd_hi[q2] = s; tag = 0x80000000u | (q2 + 1u); }
        }
        btag[s] = tag;
    }
}

// after the arrivals are in: the new local count
__global__ void k_counts_arrive(u32 *dn, char *my_area_lo0, char *my_area_hi0, size_t area_bytes, u32 cap, u32 part_cap,
                                bool has_lo, bool has_hi) {
    if (threadIdx.x != 0) return;
    const u32 step = dn[DN_STEP];
    char *alo = my_area_lo0 + (size_t)(step & 1u) * area_bytes, *ahi = my_area_hi0 + (size_t)(step & 1u) * area_bytes;
    const u32 in_lo = has_lo ? ld_volatile_u32(&inbox_header(alo, cap)->n_mig) : 0u;
    const u32 in_hi = has_hi ? ld_volatile_u32(&inbox_header(ahi, cap)->n_mig) : 0u;
    dn[DN_ARRIVE] = in_lo; dn[DN_ARRIVE + 1] = in_hi;
    u32 n = dn[DN_STAY] + in_lo + in_hi;
    if (n > part_cap) { dn[DN_OVERFLOW] |= 2u; n = part_cap; }
    dn[DN_LOCAL] = n;
}

// after k_mark_boundary: the boundary counts (what the neighbours will hold as ghosts)
__global__ void k_counts_boundary(u32 *dn, const u32 *__restrict__ cnt, u32 cap, bool has_lo, bool has_hi) {
    if (threadIdx.x != 0) return;
    u32 lo = has_lo ? cnt[2] : 0u, hi = has_hi ? cnt[3] : 0u;
    if (lo > cap || hi > cap) { dn[DN_OVERFLOW] |= 4u; lo = min(lo, cap); hi = min(hi, cap); }
    dn[DN_BND] = lo; dn[DN_BND + 1] = hi;
}

__global__ void __launch_bounds__(256)
k_push_ghosts(const u32 *__restrict__ dn, int side, const u32 *__restrict__ list, const float4 *__restrict__ pred,
              const float4 *__restrict__ pos, const u32 *__restrict__ hl, const u32 *__restrict__ gid, char *peer_area0,
              size_t area_bytes, u32 cap, u32 *done) {
    const u32 step = dn[DN_STEP], count = dn[DN_BND + side];
    char *area = peer_area0 + (size_t)(step & 1u) * area_bytes;
    GhostRec *out = inbox_ghost(area, cap);
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const u32 s = list[k];
        GhostRec r;
        r.pred = pred[s];
        r.pred.w = __uint_as_float(gid[s]);              // the receiver files the ghost under its own slot; it keeps the global id
        r.old = pos[s];                                  // for the canonical order of a cell's particles (k_cell_order)
        r.old.w = __uint_as_float(hl[s] & 1u);           // the selection bit travels with the ghost (see k_pack_ghosts)
        out[k] = r;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        InboxHeader *h = inbox_header(area, cap);
        h->n_ghost = count;
        __threadfence_system();
        st_release_sys(&h->ghost_ready, (unsigned long long)step);
    }
}

__global__ void __launch_bounds__(256)
k_pull_ghosts(const u32 *__restrict__ dn, char *my_area_lo0, char *my_area_hi0, size_t area_bytes, u32 cap, u32 part_cap,
              bool has_lo, bool has_hi, float4 *pos, float4 *vel, float4 *pred, u32 *hl, u32 *keys, u32 *flags, u32 *gid,
              GridInfo g) {
    const u32 step = dn[DN_STEP];
    char *alo = my_area_lo0 + (size_t)(step & 1u) * area_bytes, *ahi = my_area_hi0 + (size_t)(step & 1u) * area_bytes;
    if (threadIdx.x == 0) {
        if (has_lo) while (ld_acquire_sys(&inbox_header(alo, cap)->ghost_ready) < step) __nanosleep(64);
        if (has_hi) while (ld_acquire_sys(&inbox_header(ahi, cap)->ghost_ready) < step) __nanosleep(64);
    }
    __syncthreads();
    const u32 g_lo = has_lo ? ld_volatile_u32(&inbox_header(alo, cap)->n_ghost) : 0u;
    const u32 g_hi = has_hi ? ld_volatile_u32(&inbox_header(ahi, cap)->n_ghost) : 0u;
    const u32 base = dn[DN_LOCAL];
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < g_lo + g_hi; k += gridDim.x * blockDim.x) {
        const u32 s = base + k;
        if (s >= part_cap) continue;
        const GhostRec *src = k < g_lo ? inbox_ghost(alo, cap) + k : inbox_ghost(ahi, cap) + (k - g_lo);
        const uint4 *q = reinterpret_cast<const uint4 *>(src);
        uint4 a = __ldcv(q), b4 = __ldcv(q + 1);
        float4 prd = *reinterpret_cast<float4 *>(&a), old = *reinterpret_cast<float4 *>(&b4);
        const u32 selected = __float_as_uint(old.w) & 1u;
        old.w = 0.0f;
        pos[s] = old;
        vel[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        hl[s] = selected;
        if (selected) flags[0] = 1u;
        gid[s] = __float_as_uint(prd.w);
        prd.w = __int_as_float((int)s);
        pred[s] = prd;
        keys[s] = window_key(prd.x, prd.y, prd.z, g);
    }
}

// Canonical order (pbf_set_canonical_order): the reference orders the particles of a cell by ascending id (a stable sort
// from id order).  A rank's slots are not in id order -- migrants fill holes, ghosts sit at the end -- so after the sort every
// cell's members are put in ascending GLOBAL id order: the order a single-domain run has.  One thread per sorted slot: its
// cell's segment is found by looking left and right (a cell holds a handful of particles), its place by counting smaller ids.
__global__ void __launch_bounds__(256)
k_cell_order(NRef nr, const u32 *__restrict__ skey, const u32 *__restrict__ perm, const u32 *__restrict__ gid, u32 *__restrict__ out) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 k = skey[i], id = perm[i], g = gid[id];
        u32 lo = i, hi = i + 1, less = 0;
        while (lo > 0 && skey[lo - 1] == k) { lo--; less += gid[perm[lo]] < g ? 1u : 0u; }
        while (hi < n && skey[hi] == k) { less += gid[perm[hi]] < g ? 1u : 0u; hi++; }
        out[lo + less] = id;
    }
}

// after the ghosts are in: the count every later kernel of the step runs on, and the sizes of the halo refreshes
__global__ void k_counts_ghosts(u32 *dn, char *my_area_lo0, char *my_area_hi0, size_t area_bytes, u32 cap, u32 part_cap,
                                bool has_lo, bool has_hi) {
    if (threadIdx.x != 0) return;
    const u32 step = dn[DN_STEP];
    char *alo = my_area_lo0 + (size_t)(step & 1u) * area_bytes, *ahi = my_area_hi0 + (size_t)(step & 1u) * area_bytes;
    const u32 g_lo = has_lo ? ld_volatile_u32(&inbox_header(alo, cap)->n_ghost) : 0u;
    const u32 g_hi = has_hi ? ld_volatile_u32(&inbox_header(ahi, cap)->n_ghost) : 0u;
    dn[DN_GHOST] = g_lo; dn[DN_GHOST + 1] = g_hi;
    u32 n = dn[DN_LOCAL] + g_lo + g_hi;
    if (n > part_cap) { dn[DN_OVERFLOW] |= 8u; n = part_cap; }
    dn[DN_TOTAL] = n;
    dn[DN_HALO_N] = dn[DN_BND]; dn[DN_HALO_N + 1] = dn[DN_BND + 1];
    dn[DN_HALO_N + 2] = g_lo; dn[DN_HALO_N + 3] = g_hi;
}

// After the last position refresh: a ghost's sorted velocity from its owner's final position and the owner's old position
// (which k_pull_ghosts stored by slot), exactly what update.glsl computes for the particle on its own rank.
__global__ void __launch_bounds__(256)
k_ghost_velocity(const u32 *__restrict__ dn, const u32 *__restrict__ ghost_sorted, const float4 *__restrict__ A,
                 const u32 *__restrict__ perm, const float4 *__restrict__ pos, float4 *__restrict__ svel, float dt) {
    const u32 ng = dn[DN_GHOST] + dn[DN_GHOST + 1];
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < ng; k += gridDim.x * blockDim.x) {
        const u32 i = ghost_sorted[k];
        const float4 p = A[i], o = pos[perm[i]];
        svel[i] = make_float4(__fdiv_rn(__fsub_rn(p.x, o.x), dt), __fdiv_rn(__fsub_rn(p.y, o.y), dt), __fdiv_rn(__fsub_rn(p.z, o.z), dt), 0.0f);
    }
}

// particles per global cell layer (by the positions the next predict starts from): what the host runtime balances on
__global__ void __launch_bounds__(256)
k_layer_counts(NRef nr, const float4 *__restrict__ pos, GridInfo g, u32 *__restrict__ counts) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&counts[min(global_layer(pos[i].z, g), g.gz_global - 1)], 1u);
}

// last kernel of a step: what the next step's table reset runs over
__global__ void k_step_end(u32 *dn) {
    if (threadIdx.x == 0) dn[DN_PREV] = dn[DN_TOTAL];
}

// halo refresh with device-side sizes: as k_halo_push / k_halo_pull, counts from dn, sequence number = 256 * step + e
struct HaloSideDev {
    const u32 *idx;
    char *data;
    unsigned long long *flag;
};
__global__ void __launch_bounds__(256)
k_halo_push_dev(const u32 *__restrict__ dn, HaloSideDev lo, HaloSideDev hi, const float4 *__restrict__ buf, int wide, u32 e, u32 *done) {
    const u32 nlo = dn[DN_HALO_N], nhi = dn[DN_HALO_N + 1];
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < nlo + nhi; k += gridDim.x * blockDim.x) {
        const HaloSideDev &h = k < nlo ? lo : hi;
        const u32 j = k < nlo ? k : k - nlo;
        const float4 v = buf[h.idx[j]];
        if (wide) reinterpret_cast<float4 *>(h.data)[j] = v;
        else reinterpret_cast<float *>(h.data)[j] = v.w;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
        *done = 0u;
        __threadfence_system();
        const unsigned long long seq = 256ull * dn[DN_STEP] + e;
        if (nlo) st_release_sys(lo.flag, seq);
        if (nhi) st_release_sys(hi.flag, seq);
    }
}
__global__ void __launch_bounds__(256)
k_halo_pull_dev(const u32 *__restrict__ dn, HaloSideDev lo, HaloSideDev hi, const u32 *__restrict__ ghost_sorted,
                float4 *__restrict__ buf, int wide, u32 e) {
    const u32 nlo = dn[DN_HALO_N + 2], nhi = dn[DN_HALO_N + 3];
    const unsigned long long seq = 256ull * dn[DN_STEP] + e;
    if (threadIdx.x == 0) {
        if (nlo) while (ld_acquire_sys(lo.flag) < seq) __nanosleep(64);
        if (nhi) while (ld_acquire_sys(hi.flag) < seq) __nanosleep(64);
    }
    __syncthreads();
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < nlo + nhi; k += gridDim.x * blockDim.x) {
        const u32 i = ghost_sorted[k];
        const char *src = k < nlo ? lo.data : hi.data;
        const u32 j = k < nlo ? k : k - nlo;
        if (wide) {
            const volatile float *q = reinterpret_cast<const volatile float *>(src) + 4 * (size_t)j;
            buf[i] = make_float4(q[0], q[1], q[2], q[3]);
        } else {
            buf[i].w = reinterpret_cast<const volatile float *>(src)[j];
        }
    }
}

inline int nb(u32 n) { return (int)((n + 255) / 256); }

// ---- transport -------------------------------------------------------------------------------------------------------
// every rank of `grp` has packed send[0] (to z-) / send[1] (to z+); deliver into the neighbours' recv[1] / recv[0]
int exchange(pbf_sim **grp, int ng, const size_t sbytes[][2], const size_t rbytes[][2]) {
    if (grp[0]->slab->group) {                     // virtual ranks: one stream, plain copies
        for (int r = 0; r < ng; r++) {
            pbf_slab_state *b = grp[r]->slab;
            if (b->has[0] && rbytes[r][0])
                PBF_CUDA(cudaMemcpyAsync(b->recv[0], grp[r - 1]->slab->send[1], rbytes[r][0], cudaMemcpyDeviceToDevice, grp[r]->stream));
            if (b->has[1] && rbytes[r][1])
                PBF_CUDA(cudaMemcpyAsync(b->recv[1], grp[r + 1]->slab->send[0], rbytes[r][1], cudaMemcpyDeviceToDevice, grp[r]->stream));
            b->exchanges++;
            b->bytes_sent += sbytes[r][0] + sbytes[r][1];
        }
        return PBF_OK;
    }
    pbf_sim *s = grp[0];
    pbf_slab_state *b = s->slab;
    PBF_NCCL(g_nccl.GroupStart());
    for (int side = 0; side < 2; side++) {
        if (!b->has[side]) continue;
        const int peer = b->rank + (side ? 1 : -1);
        if (sbytes[0][side]) PBF_NCCL(g_nccl.Send(b->send[side], sbytes[0][side], ncclChar, peer, b->comm, s->stream));
        if (rbytes[0][side]) PBF_NCCL(g_nccl.Recv(b->recv[side], rbytes[0][side], ncclChar, peer, b->comm, s->stream));
    }
    PBF_NCCL(g_nccl.GroupEnd());
    b->exchanges++;
    b->bytes_sent += sbytes[0][0] + sbytes[0][1];
    return PBF_OK;
}

// host-visible counts: out[r] = {to lo, to hi}, in[r] = {from lo, from hi}; `which` selects counters [0,1] or [2,3]
int exchange_counts(pbf_sim **grp, int ng, int which, u32 out[][2], u32 in[][2]) {
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        if (!b->group) {
            PBF_NCCL(g_nccl.GroupStart());
            for (int side = 0; side < 2; side++) {
                if (!b->has[side]) continue;
                const int peer = b->rank + (side ? 1 : -1);
                PBF_NCCL(g_nccl.Send(b->counters + 2 * which + side, 1, ncclUint32, peer, b->comm, s->stream));
                PBF_NCCL(g_nccl.Recv(b->counters + 8 + side, 1, ncclUint32, peer, b->comm, s->stream));
            }
            PBF_NCCL(g_nccl.GroupEnd());
        }
        PBF_CUDA(cudaMemcpyAsync(b->h_counters + 16 * 0, b->counters, 16 * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    }
    for (int r = 0; r < ng; r++) PBF_CUDA(cudaStreamSynchronize(grp[r]->stream));
    for (int r = 0; r < ng; r++) {
        pbf_slab_state *b = grp[r]->slab;
        for (int side = 0; side < 2; side++) {
            out[r][side] = b->has[side] ? b->h_counters[2 * which + side] : 0;
            if (out[r][side] > b->halo_cap) {
                pbf_set_error("slab: halo capacity exceeded (raise halo_capacity in pbf_slab_init)");
                return PBF_ERR_CAPACITY;
            }
        }
    }
    for (int r = 0; r < ng; r++) {
        pbf_slab_state *b = grp[r]->slab;
        if (b->group) {
            in[r][0] = b->has[0] ? out[r - 1][1] : 0;
            in[r][1] = b->has[1] ? out[r + 1][0] : 0;
        } else {
            in[r][0] = b->has[0] ? b->h_counters[8] : 0;
            in[r][1] = b->has[1] ? b->h_counters[9] : 0;
        }
        if (in[r][0] > b->halo_cap || in[r][1] > b->halo_cap) {
            pbf_set_error("slab: incoming halo exceeds capacity");
            return PBF_ERR_CAPACITY;
        }
    }
    return PBF_OK;
}

// refresh one 4-byte (.w of bufB) or 16-byte (bufA) quantity of every ghost from its owner
// Fused push: fills the HaloPush block of the sweep that is about to produce the next halo quantity of rank `s` (its
// epilogue stores the boundary particles' values into the neighbours' mailboxes; halo_refresh then only pulls).
// false: not applicable -- NCCL transport, or PBF_SLAB_FUSED=0 -- and halo_refresh pushes with its own kernel.
bool make_push(pbf_sim *s, HaloPush *hp) {
    pbf_slab_state *b = s->slab;
    if (!b->p2p || !b->fused) return false;
    const unsigned long long seq = ++b->xseq;
    const size_t slot = (size_t)(seq % MB_SLOTS);
    memset(hp, 0, sizeof(*hp));
    for (int k = 0; k < 2; k++) {
        hp->count[k] = b->has[k] ? b->n_bnd[k] : 0;
        hp->data[k] = b->peer_data[k] ? b->peer_data[k] + slot * mbox_slot_bytes(b) : nullptr;
        hp->flag[k] = b->peer_flag[k] ? b->peer_flag[k] + slot : nullptr;
    }
    hp->seq = seq;
    hp->done = b->push_done;
    hp->expect = b->push_tiles;
    hp->map = hp->count[0] + hp->count[1] ? b->push_map : nullptr;     // nothing to push: the sweep runs as on one GPU
    b->pushed = true;
    return true;
}

int halo_refresh_p2p(pbf_sim **grp, int ng, bool wide) {
    const size_t esz = wide ? 16 : 4;
    // all pushes first: with virtual ranks every kernel is on one stream, and a pull waits for its neighbours' pushes
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        b->exchanges++;
        b->bytes_sent += ((size_t)(b->has[0] ? b->n_bnd[0] : 0) + (b->has[1] ? b->n_bnd[1] : 0)) * esz;
        if (b->pushed) { b->pushed = false; continue; }       // the producing sweep has pushed this refresh itself
        const unsigned long long seq = ++b->xseq;
        const size_t slot = (size_t)(seq % MB_SLOTS);
        HaloSide side[2];
        for (int k = 0; k < 2; k++) {
            side[k].n = b->has[k] ? b->n_bnd[k] : 0;
            side[k].idx = b->send_idx[k];
            side[k].data = b->peer_data[k] ? b->peer_data[k] + slot * mbox_slot_bytes(b) : nullptr;
            side[k].flag = b->peer_flag[k] ? b->peer_flag[k] + slot : nullptr;
        }
        const u32 n = side[0].n + side[1].n;
        if (n) {
            k_halo_push<<<nb(n), 256, 0, s->stream>>>(side[0], side[1], wide ? s->bufA : s->bufB, wide ? 1 : 0, seq, b->push_done);
            s->launches++;
        }
    }
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        const unsigned long long seq = b->xseq;
        const size_t slot = (size_t)(seq % MB_SLOTS);
        HaloSide side[2];
        for (int k = 0; k < 2; k++) {
            side[k].n = b->has[k] ? b->n_ghost[k] : 0;
            side[k].idx = nullptr;
            side[k].data = b->mbox + ((size_t)k * MB_SLOTS + slot) * mbox_slot_bytes(b);
            side[k].flag = reinterpret_cast<unsigned long long *>(b->mbox + mbox_flags_offset(b)) + k * MB_SLOTS + slot;
        }
        const u32 n = side[0].n + side[1].n;
        if (n) {
            k_halo_pull<<<nb(n), 256, 0, s->stream>>>(side[0], side[1], b->ghost_sorted, wide ? s->bufA : s->bufB, wide ? 1 : 0, seq);
            s->launches++;
        }
    }
    return PBF_OK;
}

int halo_refresh(pbf_sim **grp, int ng, bool wide) {
    if (grp[0]->slab->p2p) return halo_refresh_p2p(grp, ng, wide);
    const size_t esz = wide ? 16 : 4;
    std::vector<size_t> sb(2 * ng), rb(2 * ng);
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        for (int side = 0; side < 2; side++) {
            const u32 n = b->n_bnd[side];
            sb[2 * r + side] = n * esz;
            rb[2 * r + side] = b->n_ghost[side] * esz;
            if (!n) continue;
            if (wide) k_gather_p<<<nb(n), 256, 0, s->stream>>>(n, b->send_idx[side], s->bufA, (float4 *)b->send[side]);
            else k_gather_w<<<nb(n), 256, 0, s->stream>>>(n, b->send_idx[side], s->bufB, (float *)b->send[side]);
            s->launches++;
        }
    }
    int rc = exchange(grp, ng, (const size_t(*)[2])sb.data(), (const size_t(*)[2])rb.data());
    if (rc) return rc;
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        u32 off = 0;
        for (int side = 0; side < 2; side++) {
            const u32 n = b->n_ghost[side];
            if (n) {
                if (wide) k_scatter_p<<<nb(n), 256, 0, s->stream>>>(n, b->ghost_sorted + off, (const float4 *)b->recv[side], s->bufA);
                else k_scatter_w<<<nb(n), 256, 0, s->stream>>>(n, b->ghost_sorted + off, (const float *)b->recv[side], s->bufB);
                s->launches++;
            }
            off += n;
        }
    }
    return PBF_OK;
}

int slab_step(pbf_sim **grp, int ng) {
    std::vector<u32> out(2 * ng), in(2 * ng);
    u32(*O)[2] = (u32(*)[2])out.data();
    u32(*I)[2] = (u32(*)[2])in.data();
    std::vector<size_t> sb(2 * ng), rb(2 * ng);
    int rc;
    // ---- predict + who leaves -----------------------------------------------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        cudaMemsetAsync(s->flags, 0, sizeof(u32), s->stream);
        cudaMemsetAsync(b->counters, 0, 16 * sizeof(u32), s->stream);
        s->launches += launch_unclear_cells(s);
        s->n = b->n_local;
        s->launches += launch_predict_range(s, 0, NRef{b->n_local, nullptr}, false);
        if (b->n_local) {
            k_mark_leavers<<<nb(b->n_local), 256, 0, s->stream>>>(NRef{b->n_local, nullptr}, s->pred, s->grid, b->z_lo, b->z_hi, b->has[0],
                                                                   b->has[1], b->btag, b->list[0], b->list[1], b->counters,
                                                                   b->halo_cap);
            for (int side = 0; side < 2; side++)
                if (b->has[side])
                    k_pack_migrants<<<nb(b->halo_cap), 256, 0, s->stream>>>(b->counters + side, b->halo_cap, b->list[side], s->pos,
                                                                             s->vel, s->pred, b->gid, s->hl, (MigRec *)b->send[side]);
            s->launches += 3;
        }
    }
    if ((rc = exchange_counts(grp, ng, 0, O, I))) return rc;
    for (int r = 0; r < ng; r++)
        for (int side = 0; side < 2; side++) {
            sb[2 * r + side] = (size_t)O[r][side] * sizeof(MigRec);
            rb[2 * r + side] = (size_t)I[r][side] * sizeof(MigRec);
        }
    if ((rc = exchange(grp, ng, (const size_t(*)[2])sb.data(), (const size_t(*)[2])rb.data()))) return rc;
    // ---- compact, append arrivals, mark the boundary layers ---------------------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        const u32 n_leave = O[r][0] + O[r][1], n_arrive = I[r][0] + I[r][1];
        const u32 n_stay = b->n_local - n_leave;
        if (n_stay + n_arrive > s->cap) {
            pbf_set_error("slab: particle capacity exceeded after migration");
            return PBF_ERR_CAPACITY;
        }
        if (n_leave) {
            k_find_movers<<<nb(n_leave), 256, 0, s->stream>>>(n_stay, b->n_local, b->btag, b->movers, b->counters);
            for (int side = 0; side < 2; side++)
                if (O[r][side])
                    k_find_holes<<<nb(O[r][side]), 256, 0, s->stream>>>(n_stay, O[r][side], b->list[side], b->holes, b->counters);
            k_fill_holes<<<nb(n_leave), 256, 0, s->stream>>>(b->counters, b->holes, b->movers, s->pos, s->vel, s->pred, b->gid,
                                                             s->hl, s->keys);
            s->launches += 4;
        }
        u32 base = n_stay;
        for (int side = 0; side < 2; side++) {
            if (I[r][side]) {
                k_unpack_migrants<<<nb(I[r][side]), 256, 0, s->stream>>>(I[r][side], base, (const MigRec *)b->recv[side], s->pos,
                                                                          s->vel, s->pred, b->gid, s->hl, s->keys, s->grid);
                s->launches++;
            }
            base += I[r][side];
        }
        b->n_local = n_stay + n_arrive;
        b->migrated += n_leave;
        if (b->n_local) {
            k_mark_boundary<<<nb(b->n_local), 256, 0, s->stream>>>(NRef{b->n_local, nullptr}, s->pred, s->grid, b->z_lo, b->z_hi, b->has[0],
                                                                    b->has[1], b->btag, b->list[2], b->list[3], b->counters,
                                                                    b->halo_cap);
            s->launches++;
        }
    }
    if ((rc = exchange_counts(grp, ng, 1, O, I))) return rc;
    // ---- ghosts -----------------------------------------------------------------------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        for (int side = 0; side < 2; side++) {
            b->n_bnd[side] = O[r][side];
            b->n_ghost[side] = I[r][side];
            sb[2 * r + side] = (size_t)O[r][side] * sizeof(GhostRec);
            rb[2 * r + side] = (size_t)I[r][side] * sizeof(GhostRec);
            if (O[r][side]) {
                k_pack_ghosts<<<nb(O[r][side]), 256, 0, s->stream>>>(O[r][side], b->list[2 + side], s->pred, s->pos, s->hl,
                                                                      (GhostRec *)b->send[side]);
                s->launches++;
            }
        }
        if (b->n_local + I[r][0] + I[r][1] > s->cap) {
            pbf_set_error("slab: particle capacity exceeded by the halo");
            return PBF_ERR_CAPACITY;
        }
    }
    if ((rc = exchange(grp, ng, (const size_t(*)[2])sb.data(), (const size_t(*)[2])rb.data()))) return rc;
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        u32 base = b->n_local;
        for (int side = 0; side < 2; side++) {
            if (b->n_ghost[side]) {
                k_unpack_ghosts<<<nb(b->n_ghost[side]), 256, 0, s->stream>>>(b->n_ghost[side], base, (const GhostRec *)b->recv[side],
                                                                              s->pos, s->vel, s->pred, s->hl, s->keys, s->flags, s->grid);
                s->launches++;
            }
            base += b->n_ghost[side];
        }
        // ---- sort + cells over local + ghost particles ------------------------------------------------------------------
        s->n = base;
        s->launches += launch_sort_hist(s, s->keys, NRef{s->n, nullptr});
        s->launches += launch_sort_scan(s);
        s->launches += launch_sort_passes(s);
        s->launches += launch_reorder_cells(s);
        if (b->n_ghost[0] + b->n_ghost[1] + b->n_bnd[0] + b->n_bnd[1]) {
            cudaMemsetAsync(b->push_tiles, 0, (size_t)(1 + b->max_tiles) * 4, s->stream);
            k_halo_index<<<nb(s->n), 256, 0, s->stream>>>(NRef{s->n, nullptr}, NRef{b->n_local, nullptr}, s->skey, s->perm, b->btag, b->send_idx[0],
                                                          b->send_idx[1], b->ghost_sorted, b->push_map, b->push_tiles,
                                                          plan_tile_size(), s->grid, nullptr, 0u);
            s->launches++;
        }
        s->launches += launch_highlight(s);
    }
    // ---- solver: K x [lambda, halo lambda, delta-p, halo positions] ---------------------------------------------------------
    const int K = grp[0]->params.num_solver_iterations;
    HaloPush hp;
    for (int it = 0; it < K; it++) {
        for (int r = 0; r < ng; r++) grp[r]->launches += launch_lambda(grp[r], make_push(grp[r], &hp) ? &hp : nullptr);
        if ((rc = halo_refresh(grp, ng, false))) return rc;
        for (int r = 0; r < ng; r++) grp[r]->launches += launch_delta_p(grp[r], make_push(grp[r], &hp) ? &hp : nullptr);
        if ((rc = halo_refresh(grp, ng, true))) return rc;
    }
    // ---- update, vorticity ----------------------------------------------------------------------------------------------------
    for (int r = 0; r < ng; r++) grp[r]->launches += launch_update(grp[r]);
    if (grp[0]->params.vorticity_confinement) {
        for (int r = 0; r < ng; r++) {
            pbf_sim *s = grp[r];
            launch_vorticity_a(s, make_push(s, &hp) ? &hp : nullptr);
            s->launches++;
        }
        if ((rc = halo_refresh(grp, ng, false))) return rc;     // |omega| lives in bufB.w like lambda did
        for (int r = 0; r < ng; r++) {
            launch_vorticity_b(grp[r]);
            grp[r]->launches++;
        }
    }
    for (int r = 0; r < ng; r++) {
        grp[r]->n = grp[r]->slab->n_local;
        grp[r]->stage = 0;
        PBF_CUDA(cudaGetLastError());
    }
    return PBF_OK;
}

// ---- host side of the device-count step -----------------------------------------------------------------------------------
inline char *inbox_area(char *mbox, const pbf_slab_state *b, int side /* 0: filled by the lo neighbour */) {
    return mbox + inbox_offset(b) + (size_t)side * 2 * inbox_area_bytes(b);          // parity 0; parity 1 follows
}

constexpr int REC_BLOCKS = 128;      // grid of the record / halo kernels (grid-stride over device-side counts)
inline u32 round_up(u32 v, u32 q) { return (v + q - 1) / q * q; }

// Bounds from exact knowledge (upload): local count + slack, ghosts unknown -> the most the inbox can deliver.
// PBF_SLAB_STARVE_BOUNDS=1 (tests): bounds a third of what is needed, so that every kernel has to loop over its count
bool starve_bounds() {
    static const int v = [] { const char *e = getenv("PBF_SLAB_STARVE_BOUNDS"); return e && e[0] == '1' ? 1 : 0; }();
    return v != 0;
}

void set_bounds_exact(pbf_sim *s, u32 n_local) {
    pbf_slab_state *b = s->slab;
    b->bound_local = min(s->cap, round_up(n_local + n_local / 64 + 4096, 4096));
    s->n = min(s->cap, round_up(b->bound_local + 2 * b->halo_cap, 4096));
    if (starve_bounds()) { b->bound_local = max(512u, n_local / 3); s->n = max(512u, n_local / 3); }
    b->bounds_exact = true;
}

// Non-blocking: the newest counters the device has finished writing back tighten (or raise) the grid bounds.  The kernels
// loop over the device-side counts, so a bound that lags only costs time.  Returns an error once a step has overflowed.
int refresh_bounds(pbf_sim *s) {
    pbf_slab_state *b = s->slab;
    int best = -1;
    for (int k = 0; k < RING; k++)
        if (b->ring_step[k] && (best < 0 || b->ring_step[k] > b->ring_step[best]) && cudaEventQuery(b->ring_ev[k]) == cudaSuccess) best = k;
    (void)cudaGetLastError();            // cudaErrorNotReady of a pending event is not an error
    if (best < 0) return PBF_OK;
    const u32 *h = b->h_ring + (size_t)best * DN_WORDS;
    if (h[DN_OVERFLOW]) {
        pbf_set_error("slab: capacity exceeded on the device (bit 0: leavers, 1: arrivals, 2: boundary layer, 3: ghosts > capacity; bit 4: a halo refresh never arrived): " +
                      std::to_string(h[DN_OVERFLOW]) + "; raise the particle / halo capacity");
        return PBF_ERR_CAPACITY;
    }
    // the counters are at most RING steps old; headroom = that many steps of arrivals at twice the last observed rate.  A
    // bound only changes (and the graph is only captured again) when the need outgrows it or falls well below it.
    const u32 arrive = h[DN_ARRIVE] + h[DN_ARRIVE + 1], ghosts = h[DN_GHOST] + h[DN_GHOST + 1];
    u32 need_local = h[DN_LOCAL] + (RING + 1) * 2 * arrive + 4096;
    u32 need_total = need_local + ghosts + ghosts / 4 + 4096;
    u32 want_local = need_local > b->bound_local || b->bound_local > need_local + need_local / 16 + 32768
                         ? round_up(need_local + need_local / 64, 8192) : b->bound_local;
    u32 want_total = need_total > s->n || s->n > need_total + need_total / 16 + 32768 ? round_up(need_total + need_total / 64, 8192) : s->n;
    want_local = min(want_local, s->cap); want_total = min(want_total, s->cap);
    if (starve_bounds()) { want_local = max(512u, h[DN_LOCAL] / 3); want_total = want_local; b->bound_local = want_local; s->n = want_total; }
    b->bound_local = want_local;
    s->n = want_total;
    b->bounds_exact = false;
    b->n_local = h[DN_LOCAL]; b->n_ghost[0] = h[DN_GHOST]; b->n_ghost[1] = h[DN_GHOST + 1];
    b->n_bnd[0] = h[DN_BND]; b->n_bnd[1] = h[DN_BND + 1];
    return PBF_OK;
}

void halo_push_dev(pbf_sim *s, bool wide, u32 e, cudaStream_t st) {
    pbf_slab_state *b = s->slab;
    const size_t slot = e % MB_SLOTS;
    HaloSideDev side[2];
    for (int k = 0; k < 2; k++) {
        side[k].idx = b->send_idx[k];
        side[k].data = b->peer_data[k] ? b->peer_data[k] + slot * mbox_slot_bytes(b) : nullptr;
        side[k].flag = b->peer_flag[k] ? b->peer_flag[k] + slot : nullptr;
    }
    k_halo_push_dev<<<REC_BLOCKS, 256, 0, st>>>(s->dn, side[0], side[1], wide ? s->bufA : s->bufB, wide ? 1 : 0, e, b->push_done);
    s->launches++;
}

void halo_pull_dev(pbf_sim *s, bool wide, u32 e) {
    pbf_slab_state *b = s->slab;
    const size_t slot = e % MB_SLOTS;
    HaloSideDev side[2];
    for (int k = 0; k < 2; k++) {
        side[k].idx = nullptr;
        side[k].data = b->mbox + ((size_t)k * MB_SLOTS + slot) * mbox_slot_bytes(b);
        side[k].flag = reinterpret_cast<unsigned long long *>(b->mbox + mbox_flags_offset(b)) + k * MB_SLOTS + slot;
    }
    k_halo_pull_dev<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, side[0], side[1], b->ghost_sorted, wide ? s->bufA : s->bufB, wide ? 1 : 0, e);
    s->launches++;
}

// Refreshes inside the sweeps need the plan's tile descriptors (no full-support search) and update.glsl in the last delta-p
bool overlap_on(const pbf_sim *s) { return s->slab->overlap && s->fuse_update && !s->options.full_support; }

// Parameter blocks of the refreshes that run inside the sweeps (device-side counts): refresh number e of the step
void fused_push(pbf_sim *s, u32 e, HaloPush *hp) {
    pbf_slab_state *b = s->slab;
    const size_t slot = e % MB_SLOTS;
    memset(hp, 0, sizeof(*hp));
    for (int k = 0; k < 2; k++) {
        hp->data[k] = b->peer_data[k] ? b->peer_data[k] + slot * mbox_slot_bytes(b) : nullptr;
        hp->flag[k] = b->peer_flag[k] ? b->peer_flag[k] + slot : nullptr;
    }
    hp->map = b->push_map;
    hp->done = b->push_done;
    hp->expect = b->push_tiles;
    hp->dn = s->dn;
    hp->e = e;
}

void fused_pull(pbf_sim *s, u32 e, bool wide, float4 *buf, HaloPull *pl) {
    pbf_slab_state *b = s->slab;
    const size_t slot = e % MB_SLOTS;
    memset(pl, 0, sizeof(*pl));
    for (int k = 0; k < 2; k++) {
        pl->data[k] = b->mbox + ((size_t)k * MB_SLOTS + slot) * mbox_slot_bytes(b);
        pl->flag[k] = reinterpret_cast<unsigned long long *>(b->mbox + mbox_flags_offset(b)) + k * MB_SLOTS + slot;
    }
    pl->dn = s->dn;
    pl->ghost_sorted = b->ghost_sorted;
    pl->buf = buf;
    pl->ready = b->pull_ready;
    pl->e = e;
    pl->wide = wide ? 1 : 0;
}

// A sweep that produces a halo quantity, and the refresh of that quantity (exchange number e of the step): sweep, push,
// pull.  Measured and NOT adopted (2 x 8M particles on 2 B200, splash scene): running the tiles that hold boundary particles
// first and pushing from a second stream under the interior tiles -- 5.09-5.14 ms per step against 5.00-5.06 ms: the
// tile-selection logic costs every sweep block ~3 % (also on a single GPU: 4.39 -> 4.49 ms), more than the ~30 us per
// exchange the overlap hides.
template <class Launch, class Mark>
void sweep_and_refresh(pbf_sim **grp, int ng, bool wide, u32 e, Launch launch, Mark tmark, const char *name) {
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        b->exchanges++;
        s->launches += launch(s);
        if (ng == 1) tmark(name);
        if (b->has[0] || b->has[1]) halo_push_dev(s, wide, e, s->stream);
    }
    tmark(ng == 1 ? "push" : name);
    for (int r = 0; r < ng; r++)            // all pushes are enqueued before any pull (virtual ranks share one stream)
        if (grp[r]->slab->has[0] || grp[r]->slab->has[1]) halo_pull_dev(grp[r], wide, e);
    tmark("pull");
}

// One step of every rank of `grp`, enqueued without touching the host-side state the device decides (capturable).
int enqueue_slab_step_dev(pbf_sim **grp, int ng) {
    auto mark = [&](int k) {          // phase boundaries of rank 0 (pbf_slab_phase_times)
        pbf_slab_state *b = grp[0]->slab;
        if (b->phases) { cudaEventRecord(b->ph_ev[k], grp[0]->stream); b->ph_valid = true; }
    };
    auto tmark = [&](const char *label) {
        pbf_slab_state *b = grp[0]->slab;
        if (b->phases && b->trace && b->tr_n < 96) { b->tr_label[b->tr_n] = label; cudaEventRecord(b->tr_ev[b->tr_n++], grp[0]->stream); }
    };
    grp[0]->slab->tr_n = 0;
    mark(0);
    tmark("t0");
    // ---- predict, who leaves, leavers into the neighbours' inboxes, compaction ------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        const NRef nloc = {b->bound_local, s->dn + DN_LOCAL};
        cudaMemsetAsync(s->flags, 0, sizeof(u32), s->stream);
        cudaMemsetAsync(b->counters, 0, 16 * sizeof(u32), s->stream);
        k_step_begin<<<1, 32, 0, s->stream>>>(s->dn);
        // one pass over the local particles: predict, reset of the previous step's table entries, who leaves
        const LeaveArgs la = {b->z_lo, b->z_hi, b->has[0], b->has[1], b->btag, b->list[0], b->list[1], b->counters, b->halo_cap,
                              b->list[2], b->list[3]};
        s->launches += 1 + launch_predict_slab(s, nloc, la);
        if (ng == 1) tmark("predict");
        k_counts_leave<<<1, 32, 0, s->stream>>>(s->dn, b->counters, b->halo_cap, b->has[0], b->has[1]);
        s->launches += 1;
        for (int side = 0; side < 2; side++)
            if (b->has[side]) {
                k_push_migrants<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, side, b->list[side], s->pos, s->vel, s->pred, b->gid, s->hl,
                                                                   b->peer_inbox[side], inbox_area_bytes(b), b->halo_cap,
                                                                   b->rec_done + side);
                s->launches++;
            }
        if (b->has[0] || b->has[1]) {
            k_find_movers_dev<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, b->btag, b->movers, b->counters);
            for (int side = 0; side < 2; side++)
                if (b->has[side]) { k_find_holes_dev<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, side, b->list[side], b->holes, b->counters); s->launches++; }
            k_fill_holes_dev<<<REC_BLOCKS, 256, 0, s->stream>>>(b->counters, b->holes, b->movers, s->pos, s->vel, s->pred, b->gid, s->hl, s->keys,
                                                                b->btag, b->list[2], b->list[3]);
            s->launches += 2;
        }
    }
    mark(1);
    tmark("migrate_out");
    // ---- arrivals, boundary layers, ghosts into the neighbours' inboxes --------------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        const NRef nloc = {b->bound_local, s->dn + DN_LOCAL};
        char *alo = inbox_area(b->mbox, b, 0), *ahi = inbox_area(b->mbox, b, 1);
        if (b->has[0] || b->has[1]) {
            k_pull_migrants<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, alo, ahi, inbox_area_bytes(b), b->halo_cap, s->cap, b->has[0], b->has[1],
                                                               s->pos, s->vel, s->pred, b->gid, s->hl, s->keys, s->grid,
                                                               b->z_lo, b->z_hi, b->btag, b->list[2], b->list[3], b->counters);
            s->launches++;
        }
        // the boundary-layer lists were made on the way: by k_predict for the particles that stayed (slots patched by the
        // compaction), by k_pull_migrants for the arrivals -- no pass of its own over the local particles
        k_counts_arrive<<<1, 32, 0, s->stream>>>(s->dn, alo, ahi, inbox_area_bytes(b), b->halo_cap, s->cap, b->has[0], b->has[1]);
        k_counts_boundary<<<1, 32, 0, s->stream>>>(s->dn, b->counters, b->halo_cap, b->has[0], b->has[1]);
        s->launches += 2;
        if (ng == 1) tmark("arrivals");
        for (int side = 0; side < 2; side++)
            if (b->has[side]) {
                k_push_ghosts<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, side, b->list[2 + side], s->pred, s->pos, s->hl, b->gid, b->peer_inbox[side],
                                                                 inbox_area_bytes(b), b->halo_cap, b->rec_done + 2 + side);
                s->launches++;
            }
    }
    mark(2);
    tmark("ghosts_out");
    // ---- ghosts in, sort + cells over local + ghost particles --------------------------------------------------------------
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        char *alo = inbox_area(b->mbox, b, 0), *ahi = inbox_area(b->mbox, b, 1);
        if (b->has[0] || b->has[1]) {
            k_pull_ghosts<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, alo, ahi, inbox_area_bytes(b), b->halo_cap, s->cap, b->has[0], b->has[1],
                                                             s->pos, s->vel, s->pred, s->hl, s->keys, s->flags, b->gid, s->grid);
            s->launches++;
        }
        k_counts_ghosts<<<1, 32, 0, s->stream>>>(s->dn, alo, ahi, inbox_area_bytes(b), b->halo_cap, s->cap, b->has[0], b->has[1]);
        s->launches++;
        if (ng == 1) tmark("ghosts_in");
        s->launches += launch_sort_hist(s, s->keys, nref_total(s));
        s->launches += launch_sort_scan(s);
        s->launches += launch_sort_passes(s);
        if (s->canonical) {                 // cells in ascending global id order (scratch: the sort's ping-pong buffer is free)
            k_cell_order<<<nb(s->n), 256, 0, s->stream>>>(nref_total(s), s->skey, s->perm, b->gid, s->vtmp[0]);
            cudaMemcpyAsync(s->perm, s->vtmp[0], (size_t)s->n * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream);
            s->launches++;
        }
        if (ng == 1) tmark("sort");
        s->launches += launch_reorder_cells(s);
        if (ng == 1) tmark("cells");
        if (b->has[0] || b->has[1]) {
            cudaMemsetAsync(b->push_tiles, 0, (size_t)(1 + b->max_tiles) * 4, s->stream);
            k_halo_index<<<nb(s->n), 256, 0, s->stream>>>(nref_total(s), NRef{b->bound_local, s->dn + DN_LOCAL}, s->skey, s->perm, b->btag,
                                                          b->send_idx[0], b->send_idx[1], b->ghost_sorted, b->push_map, b->push_tiles,
                                                          plan_tile_size(), s->grid, overlap_on(s) ? s->tile_desc : nullptr, plan_desc_stride());
            s->launches++;
        }
        if (ng == 1) tmark("halo_index");
        s->launches += launch_highlight(s);
    }
    mark(3);
    tmark("highlight");
    // ---- solver, update, vorticity ------------------------------------------------------------------------------------------
    const int K = grp[0]->params.num_solver_iterations;
    u32 e = 0;
    // update.glsl runs in the epilogue of the last delta-p sweep as on a single GPU.  A ghost's own sweep result is
    // meaningless (half its neighbourhood is missing) and is replaced by its owner's position right after; with vorticity
    // on, its sorted velocity -- which the neighbours' vorticity sweeps read -- is then derived again from that position.
    const bool vort = grp[0]->params.vorticity_confinement != 0;
    if (overlap_on(grp[0])) {
        // Refreshes inside the sweeps: the sweep that produces a halo quantity pushes it from its boundary tiles, the sweep that
        // consumes it pulls it with its first blocks while every tile that does not touch the halo layers is already running
        // (HaloPush / HaloPull, sweeps.cu).  Only the positions after the LAST delta-p are pulled by a kernel of their own:
        // the ghosts' sorted velocities are derived from them before the vorticity sweep starts.
        auto neighbours = [](pbf_sim *s) { return s->slab->has[0] || s->slab->has[1]; };
        // PBF_SLAB_OVERLAP=2 / 3 (experiments): only the push / only the pull runs inside the sweeps, the other as a kernel
        const bool fpush = grp[0]->slab->ov_push, fpull = grp[0]->slab->ov_pull;
        // One sweep of every rank.  kind 0 lambda, 1 delta-p, 2 delta-p + update, 3 vorticity A, 4 vorticity B;  pull_in: it
        // consumes refresh e_in (wide_in: positions into bufA, else values into bufB.w); push_: it produces refresh e_out
        auto sweep = [&](int kind, bool pull_in, u32 e_in, bool wide_in, bool push_, u32 e_out, bool pull_kernel_after, const char *name) {
            for (int r = 0; r < ng; r++) {
                pbf_sim *s = grp[r];
                HaloPush hp; HaloPull pl;
                const bool nb_ = neighbours(s), dopush = nb_ && push_, dopull = nb_ && pull_in && fpull;
                if (dopush) { fused_push(s, e_out, &hp); s->slab->exchanges++; }
                if (dopull) fused_pull(s, e_in, wide_in, wide_in ? s->bufA : s->bufB, &pl);
                const HaloPush *php = dopush && fpush ? &hp : nullptr;
                const HaloPull *ppl = dopull ? &pl : nullptr;
                switch (kind) {
                case 0: s->launches += launch_lambda(s, php, ppl); break;
                case 1: s->launches += launch_delta_p(s, php, ppl); break;
                case 2: s->launches += launch_delta_p_update(s, php, ppl); break;
                case 3: s->launches += launch_vorticity_a(s, php); break;
                default: s->launches += launch_vorticity_b(s, ppl); break;
                }
                if (ng == 1) tmark(name);
                if (dopush && !fpush) halo_push_dev(s, kind == 1 || kind == 2, e_out, s->stream);
            }
            tmark(ng == 1 ? "push" : name);
            if (push_ && pull_kernel_after) {
                for (int r = 0; r < ng; r++)
                    if (neighbours(grp[r])) halo_pull_dev(grp[r], kind == 1 || kind == 2, e_out);
                tmark("pull");
            }
        };
        for (int it = 0; it < K; it++) {
            ++e;
            sweep(0, it > 0, e - 1, true, true, e, !fpull, "lambda");
            ++e;
            const bool last = it == K - 1;
            // the positions of the last iteration are only read by the vorticity sweeps, and pulled by a kernel (below)
            if (last) sweep(2, true, e - 1, false, vort, e, false, "delta_p_update");
            else sweep(1, true, e - 1, false, true, e, !fpull, "delta_p");
        }
        mark(4);
        tmark("-");
        if (vort) {
            for (int r = 0; r < ng; r++) {
                pbf_sim *s = grp[r];
                pbf_slab_state *b = s->slab;
                if (!neighbours(s)) continue;
                if (K > 0) {
                    halo_pull_dev(s, true, e);
                    k_ghost_velocity<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, b->ghost_sorted, s->bufA, s->perm, s->pos, s->svel,
                                                                        sim_params(s).timestep);
                    s->launches++;
                }
            }
            tmark("pull+ghost_velocity");
            ++e;
            sweep(3, false, 0, false, true, e, !fpull, "vorticity_a");
            sweep(4, true, e, false, false, 0, false, "vorticity_b");
        }
    } else {
    for (int it = 0; it < K; it++) {
        sweep_and_refresh(grp, ng, false, ++e, [](pbf_sim *s) { return launch_lambda(s, nullptr); }, tmark, "lambda");
        const bool last = it == K - 1 && grp[0]->fuse_update;
        sweep_and_refresh(grp, ng, true, ++e, [last](pbf_sim *s) { return last ? launch_delta_p_update(s) : launch_delta_p(s, nullptr); },
                          tmark, last ? "delta_p_update" : "delta_p");
    }
    mark(4);
    tmark("-");
    if (K > 0 && grp[0]->fuse_update) {
        for (int r = 0; r < ng; r++) {
            pbf_sim *s = grp[r];
            pbf_slab_state *b = s->slab;
            if (vort && (b->has[0] || b->has[1])) {
                k_ghost_velocity<<<REC_BLOCKS, 256, 0, s->stream>>>(s->dn, b->ghost_sorted, s->bufA, s->perm, s->pos, s->svel,
                                                                    sim_params(s).timestep);
                s->launches++;
            }
        }
    } else {
        for (int r = 0; r < ng; r++) grp[r]->launches += launch_update(grp[r]);
    }
    if (grp[0]->params.vorticity_confinement) {
        sweep_and_refresh(grp, ng, false, ++e, [](pbf_sim *s) { return launch_vorticity_a(s, nullptr); }, tmark, "vorticity_a");
        for (int r = 0; r < ng; r++) grp[r]->launches += launch_vorticity_b(grp[r]);
        tmark("vorticity_b");
    }
    }
    for (int r = 0; r < ng; r++) {
        k_step_end<<<1, 32, 0, grp[r]->stream>>>(grp[r]->dn);
        grp[r]->launches++;
    }
    mark(5);
    return PBF_OK;
}

uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
    return h;
}

int slab_step_dev(pbf_sim **grp, int ng) {
    pbf_slab_state *b0 = grp[0]->slab;
    if (grp[0]->params.num_solver_iterations > 60) { pbf_set_error("slab: more than 60 solver iterations per step"); return PBF_ERR_INVALID; }
    uint64_t key = 1469598103934665603ull;
    for (int r = 0; r < ng; r++) {
        int rc = refresh_bounds(grp[r]);
        if (rc) return rc;
        key = fnv(key, &grp[r]->n, 4); key = fnv(key, &grp[r]->slab->bound_local, 4);
        key = fnv(key, &grp[r]->slab->z_lo, 4); key = fnv(key, &grp[r]->slab->z_hi, 4);
        key = fnv(key, &grp[r]->params, sizeof(pbf_params)); key = fnv(key, &grp[r]->options, sizeof(pbf_options));
        key = fnv(key, &grp[r]->canonical, sizeof(bool));
    }
    cudaStream_t st = grp[0]->stream;
    const bool graph = b0->use_graph && !grp[0]->timing && !b0->phases;
    if (graph) {
        if (!b0->graph_exec || b0->graph_key != key) {
            if (b0->graph_exec) { cudaGraphExecDestroy(b0->graph_exec); cudaGraphDestroy(b0->graph); b0->graph_exec = nullptr; b0->graph = nullptr; }
            std::vector<uint64_t> before(ng);
            for (int r = 0; r < ng; r++) before[r] = grp[r]->launches;
            PBF_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            enqueue_slab_step_dev(grp, ng);
            cudaError_t e = cudaStreamEndCapture(st, &b0->graph);
            if (e != cudaSuccess) { b0->graph = nullptr; pbf_set_error(std::string("slab graph capture: ") + cudaGetErrorString(e)); return PBF_ERR_CUDA; }
            PBF_CUDA(cudaGraphInstantiate(&b0->graph_exec, b0->graph, 0));
            for (int r = 0; r < ng; r++) { grp[r]->slab->graph_kernels = (u32)(grp[r]->launches - before[r]); grp[r]->launches = before[r]; }
            b0->graph_key = key;
        }
        PBF_CUDA(cudaGraphLaunch(b0->graph_exec, st));
        for (int r = 0; r < ng; r++) grp[r]->launches += grp[r]->slab->graph_kernels;
    } else {
        enqueue_slab_step_dev(grp, ng);
    }
    for (int r = 0; r < ng; r++) {
        pbf_sim *s = grp[r];
        pbf_slab_state *b = s->slab;
        b->step++;
        const int k = (int)(b->step % RING);
        PBF_CUDA(cudaMemcpyAsync(b->h_ring + (size_t)k * DN_WORDS, s->dn, DN_WORDS * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
        PBF_CUDA(cudaEventRecord(b->ring_ev[k], s->stream));
        b->ring_step[k] = b->step;
        s->stage = 0;
        PBF_CUDA(cudaGetLastError());
    }
    return PBF_OK;
}

// exact counts on the host (download, statistics): waits for the stream
int sync_counts(pbf_sim *s) {
    pbf_slab_state *b = s->slab;
    if (!b->devcount) return PBF_OK;
    u32 h[DN_WORDS];
    PBF_CUDA(cudaMemcpyAsync(h, s->dn, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    if (h[DN_OVERFLOW]) { pbf_set_error("slab: capacity exceeded on the device; raise the particle / halo capacity"); return PBF_ERR_CAPACITY; }
    b->n_local = h[DN_LOCAL]; b->n_ghost[0] = h[DN_GHOST]; b->n_ghost[1] = h[DN_GHOST + 1];
    b->n_bnd[0] = h[DN_BND]; b->n_bnd[1] = h[DN_BND + 1];
    b->migrated = h[20];
    return PBF_OK;
}

int slab_alloc(pbf_sim *s, int rank, int nranks, int z_lo, int z_hi, int gz_global, u32 halo_cap) {
    if (s->slab) { pbf_set_error("slab: already initialised"); return PBF_ERR_STATE; }
    if (z_hi - z_lo < 2) { pbf_set_error("slab: a slab must own at least 2 cell layers"); return PBF_ERR_INVALID; }
    if (z_hi - z_lo + 2 > s->cfg.grid[2]) {
        pbf_set_error("slab: the handle's grid z extent must be at least (z_hi - z_lo) + 2 ghost layers (more leaves room for pbf_slab_set_planes)");
        return PBF_ERR_INVALID;
    }
    if (rank < 0 || rank >= nranks || halo_cap == 0) { pbf_set_error("slab: bad rank or halo capacity"); return PBF_ERR_INVALID; }
    pbf_slab_state *b = new pbf_slab_state();
    memset((void *)b, 0, sizeof(*b));
    b->rank = rank; b->nranks = nranks; b->z_lo = z_lo; b->z_hi = z_hi;
    b->has[0] = rank > 0; b->has[1] = rank + 1 < nranks;
    b->halo_cap = halo_cap;
    b->n_local = s->n;
    s->grid.zoff = z_lo - 1;
    s->grid.gz = z_hi - z_lo + 2;           // the window in use; the tables (and the key stride gx * gz) keep the allocated depth
    s->grid.gz_global = gz_global;
    s->grid.ref_quirks = 0;     // the lowest-key-cell quirk is a single-domain artefact (findcells.glsl:39-43)
    s->grid.whi[2] = (float)gz_global - s->cfg.wall[2];
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&b->gid, (size_t)s->cap * 4); A((void **)&b->btag, (size_t)s->cap * 4);
    for (int i = 0; i < 4; i++) A((void **)&b->list[i], (size_t)halo_cap * 4);
    A((void **)&b->movers, (size_t)2 * halo_cap * 4); A((void **)&b->holes, (size_t)2 * halo_cap * 4);
    A((void **)&b->counters, 16 * 4);
    for (int i = 0; i < 2; i++) {
        A((void **)&b->send[i], (size_t)halo_cap * sizeof(MigRec));
        A((void **)&b->recv[i], (size_t)halo_cap * sizeof(MigRec));
        A((void **)&b->send_idx[i], (size_t)halo_cap * 4);
    }
    A((void **)&b->ghost_sorted, (size_t)2 * halo_cap * 4);
    A((void **)&b->mbox, mbox_bytes(b)); A((void **)&b->push_done, 16); A((void **)&b->push_map, (size_t)s->cap * 4);
    b->max_tiles = (s->cap + plan_tile_size() - 1) / plan_tile_size();
    A((void **)&b->push_tiles, (size_t)(1 + b->max_tiles) * 4);
    A((void **)&b->rec_done, 16);
    A((void **)&b->pull_ready, PBF_PULL_BLOCKS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&b->h_counters, 16 * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&b->h_ring, (size_t)RING * DN_WORDS * 4);
    for (int k = 0; k < RING && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&b->ring_ev[k], cudaEventDisableTiming);
    if (e != cudaSuccess) { pbf_set_error(std::string("slab: allocation failed: ") + cudaGetErrorString(e)); return PBF_ERR_CUDA; }
    cudaMemsetAsync(b->gid, 0, (size_t)s->cap * 4, s->stream);
    cudaMemsetAsync(b->mbox, 0, mbox_bytes(b), s->stream);
    cudaMemsetAsync(b->push_done, 0, 16, s->stream);
    cudaMemsetAsync(b->rec_done, 0, 16, s->stream);
    if (b->pull_ready) cudaMemsetAsync(b->pull_ready, 0, PBF_PULL_BLOCKS * sizeof(unsigned long long), s->stream);
    cudaMemsetAsync(s->dn, 0, DN_WORDS * sizeof(u32), s->stream);
    {
        const char *g = getenv("PBF_SLAB_GRAPH");
        b->use_graph = !(g && g[0] == '0');
        const char *ov = getenv("PBF_SLAB_OVERLAP");
        b->overlap = !(ov && ov[0] == '0');
        b->ov_push = !(ov && ov[0] == '3');
        b->ov_pull = !(ov && ov[0] == '2');
        const char *ph = getenv("PBF_SLAB_PHASES");
        b->phases = ph && ph[0] == '1';
        if (b->phases)
            for (int k = 0; k < 8; k++) cudaEventCreate(&b->ph_ev[k]);
        const char *tr = getenv("PBF_SLAB_TRACE");
        b->trace = b->phases && tr && tr[0] == '1';
        if (b->trace)
            for (int k = 0; k < 96; k++) cudaEventCreate(&b->tr_ev[k]);
    }
    cudaStreamSynchronize(s->stream);      // the mailbox flags are zero before any neighbour can see them
    s->slab = b;
    if (s->graph_valid) { cudaGraphExecDestroy(s->graph_exec); cudaGraphDestroy(s->graph); s->graph_valid = false; s->graph = nullptr; s->graph_exec = nullptr; }
    return PBF_OK;
}

// my lo neighbour's mailbox: I am its hi neighbour, so I fill its "from hi" half (index 1), and vice versa
void p2p_attach(pbf_slab_state *b, int side, char *peer_mbox) {
    const int theirs = side == 0 ? 1 : 0;
    b->peer_data[side] = peer_mbox + (size_t)theirs * MB_SLOTS * mbox_slot_bytes(b);
    b->peer_flag[side] = reinterpret_cast<unsigned long long *>(peer_mbox + mbox_flags_offset(b)) + theirs * MB_SLOTS;
    b->peer_inbox[side] = inbox_area(peer_mbox, b, theirs);
}

// Device-side counts need the peer-memory transport (records go into the neighbour's inbox) and the separate push kernels;
// PBF_SLAB_DEVCOUNT=0 keeps the step that reads its counts back twice (the baseline it is measured against).
void enable_devcount(pbf_sim *s) {
    pbf_slab_state *b = s->slab;
    const char *dc = getenv("PBF_SLAB_DEVCOUNT");
    b->devcount = b->p2p && !b->fused && !(dc && dc[0] == '0');
    if (!b->devcount) return;
    s->n_dev = s->dn + DN_TOTAL;
    s->n_prev_dev = s->dn + DN_PREV;
    const u32 n = b->n_local;
    cudaMemcpyAsync(s->dn + DN_LOCAL, &n, 4, cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(s->dn + DN_TOTAL, &n, 4, cudaMemcpyHostToDevice, s->stream);
    cudaStreamSynchronize(s->stream);
    set_bounds_exact(s, n);
}

}  // namespace

bool slab_borrows_stream(const pbf_sim *s) { return s->slab && s->slab->group && s->slab->rank != 0; }

void slab_free(pbf_sim *s) {
    pbf_slab_state *b = s->slab;
    if (!b) return;
    if (b->comm && g_nccl.lib) g_nccl.CommDestroy(b->comm);
    for (int k = 0; k < 2; k++)
        if (b->ipc_base[k]) cudaIpcCloseMemHandle(b->ipc_base[k]);
    if (b->mbox) cudaFree(b->mbox);
    if (b->push_done) cudaFree(b->push_done);
    if (b->push_map) cudaFree(b->push_map);
    if (b->push_tiles) cudaFree(b->push_tiles);
    if (b->rec_done) cudaFree(b->rec_done);
    if (b->pull_ready) cudaFree(b->pull_ready);
    if (b->h_ring) cudaFreeHost(b->h_ring);
    for (int k = 0; k < RING; k++)
        if (b->ring_ev[k]) cudaEventDestroy(b->ring_ev[k]);
    if (b->graph_exec) { cudaGraphExecDestroy(b->graph_exec); cudaGraphDestroy(b->graph); }
    void *ptrs[] = {b->gid, b->btag, b->list[0], b->list[1], b->list[2], b->list[3], b->movers, b->holes, b->counters, b->send[0],
                    b->send[1], b->recv[0], b->recv[1], b->send_idx[0], b->send_idx[1], b->ghost_sorted};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (b->h_counters) cudaFreeHost(b->h_counters);
    delete[] b->group;
    delete b;
    s->slab = nullptr;
}

extern "C" {

int pbf_slab_unique_id(void *out128) {
    if (!out128) { pbf_set_error("pbf_slab_unique_id: null"); return PBF_ERR_INVALID; }
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    PBF_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return PBF_OK;
}

int pbf_slab_init(pbf_handle s, const void *id128, int rank, int nranks, int z_lo, int z_hi, int gz_global,
                  uint32_t halo_capacity) {
    if (!s || !id128) { pbf_set_error("pbf_slab_init: null"); return PBF_ERR_INVALID; }
    int rc = nccl_load();
    if (rc) return rc;
    int prev;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    rc = slab_alloc(s, rank, nranks, z_lo, z_hi, gz_global, halo_capacity);
    if (rc == PBF_OK && nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        ncclResult_t r = g_nccl.CommInitRank(&s->slab->comm, nranks, id, rank);
        if (r != ncclSuccess) { pbf_set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); rc = PBF_ERR_NCCL; }
    }
    if (rc == PBF_OK && nranks == 1) {      // a lone rank has nobody to exchange with: device-side counts need no connection
        s->slab->p2p = true;
        enable_devcount(s);
    }
    cudaSetDevice(prev);
    return rc;
}

// Peer-memory halo refresh between processes: every rank exports its mailbox as a CUDA IPC handle (64 bytes), the host
// runtime hands each rank its neighbours' handles (NULL where there is none).  Needs peer access between the devices
// (NVLink on the B200 box); the mailboxes of all ranks must have been created with the same halo capacity.
int pbf_slab_p2p_handle(pbf_handle s, void *out64) {
    if (!s || !s->slab || !out64) { pbf_set_error("pbf_slab_p2p_handle: slab not initialised"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    PBF_CUDA(cudaIpcGetMemHandle(&h, s->slab->mbox));
    memcpy(out64, &h, 64);
    return PBF_OK;
}

int pbf_slab_p2p_connect(pbf_handle s, const void *lo64, const void *hi64) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_p2p_connect: slab not initialised"); return PBF_ERR_STATE; }
    pbf_slab_state *b = s->slab;
    if (b->group) { pbf_set_error("pbf_slab_p2p_connect: virtual ranks are connected by pbf_slab_init_group"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    const void *handles[2] = {lo64, hi64};
    for (int side = 0; side < 2; side++) {
        if (!b->has[side]) continue;
        if (!handles[side]) { pbf_set_error("pbf_slab_p2p_connect: missing neighbour handle"); return PBF_ERR_INVALID; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles[side], 64);
        void *base = nullptr;
        PBF_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        b->ipc_base[side] = base;
        p2p_attach(b, side, static_cast<char *>(base));
    }
    b->p2p = true;
    const char *fz = getenv("PBF_SLAB_FUSED");     // measured slower than the separate push kernel: opt in
    b->fused = fz && fz[0] == '1';
    enable_devcount(s);
    return PBF_OK;
}

// virtual ranks: n handles of this process (same device), rank r owns layers [z_planes[r], z_planes[r+1])
int pbf_slab_init_group(pbf_handle *hs, int n, const int32_t *z_planes, int gz_global, uint32_t halo_capacity) {
    if (!hs || n < 1 || !z_planes) { pbf_set_error("pbf_slab_init_group: bad argument"); return PBF_ERR_INVALID; }
    for (int r = 0; r < n; r++) {
        int rc = slab_alloc(hs[r], r, n, z_planes[r], z_planes[r + 1], gz_global, halo_capacity);
        if (rc) return rc;
        hs[r]->slab->group = new pbf_sim *[n];
        hs[r]->slab->group_size = n;
        for (int k = 0; k < n; k++) hs[r]->slab->group[k] = hs[k];
        if (r > 0) {   // one stream for the whole group keeps the copies ordered without events
            cudaStreamSynchronize(hs[r]->stream);
            cudaStreamDestroy(hs[r]->stream);
            hs[r]->stream = hs[0]->stream;
        }
    }
    // the halo refreshes of a virtual group take the same peer-memory path as real ranks, with plain pointers
    const char *np = getenv("PBF_SLAB_P2P");
    if (!(np && np[0] == '0'))
        for (int r = 0; r < n; r++) {
            pbf_slab_state *b = hs[r]->slab;
            if (b->has[0]) p2p_attach(b, 0, hs[r - 1]->slab->mbox);
            if (b->has[1]) p2p_attach(b, 1, hs[r + 1]->slab->mbox);
            b->p2p = true;
            const char *fz = getenv("PBF_SLAB_FUSED");
            b->fused = fz && fz[0] == '1';
            enable_devcount(hs[r]);
        }
    return PBF_OK;
}

// local particles of a slab: HOST arrays by slot + their global ids
int pbf_slab_upload(pbf_handle s, const float *pos4, const float *vel4, const uint32_t *gid, uint32_t n) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_upload: slab not initialised"); return PBF_ERR_STATE; }
    if (n > s->cap) { pbf_set_error("pbf_slab_upload: n exceeds capacity"); return PBF_ERR_CAPACITY; }
    if (n && (!pos4 || !gid)) { pbf_set_error("pbf_slab_upload: null buffer"); return PBF_ERR_INVALID; }
    int prev;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    PBF_CUDA(cudaMemcpyAsync(s->pos, pos4, (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    if (vel4) PBF_CUDA(cudaMemcpyAsync(s->vel, vel4, (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
    else PBF_CUDA(cudaMemsetAsync(s->vel, 0, (size_t)n * 16, s->stream));
    PBF_CUDA(cudaMemsetAsync(s->hl, 0, (size_t)n * 4, s->stream));
    PBF_CUDA(cudaMemcpyAsync(s->slab->gid, gid, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream));
    if (s->slab->devcount) {
        const u32 nn = n;
        PBF_CUDA(cudaMemcpyAsync(s->dn + DN_LOCAL, &nn, 4, cudaMemcpyHostToDevice, s->stream));
    }
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    s->n = n;
    s->slab->n_local = n;
    if (s->slab->devcount) set_bounds_exact(s, n);
    cudaSetDevice(prev);
    return PBF_OK;
}

int pbf_slab_download(pbf_handle s, float *pos4, float *vel4, uint32_t *gid, uint32_t *n) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_download: slab not initialised"); return PBF_ERR_STATE; }
    int prev;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    { const int rc = sync_counts(s); if (rc) { cudaSetDevice(prev); return rc; } }
    const u32 m = s->slab->n_local;
    if (n) *n = m;
    if (pos4) PBF_CUDA(cudaMemcpyAsync(pos4, s->pos, (size_t)m * 16, cudaMemcpyDeviceToHost, s->stream));
    if (vel4) PBF_CUDA(cudaMemcpyAsync(vel4, s->vel, (size_t)m * 16, cudaMemcpyDeviceToHost, s->stream));
    if (gid) PBF_CUDA(cudaMemcpyAsync(gid, s->slab->gid, (size_t)m * 4, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    cudaSetDevice(prev);
    return PBF_OK;
}

// highlight words of the local particles, by slot (same order as pbf_slab_download's arrays)
int pbf_slab_download_highlight(pbf_handle s, uint32_t *highlight) {
    if (!s || !s->slab || !highlight) { pbf_set_error("pbf_slab_download_highlight: slab not initialised or null buffer"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    { const int rc = sync_counts(s); if (rc) return rc; }
    PBF_CUDA(cudaMemcpyAsync(highlight, s->hl, (size_t)s->slab->n_local * 4, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    return PBF_OK;
}

// SPH::Run on this rank's slab (NCCL transport), or on every virtual rank of the handle's group
int pbf_slab_step(pbf_handle s, int nsteps) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_step: slab not initialised"); return PBF_ERR_STATE; }
    int prev;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    int rc = PBF_OK;
    for (int i = 0; i < nsteps && rc == PBF_OK; i++) {
        pbf_sim *one[1] = {s};
        pbf_sim **grp = s->slab->group ? s->slab->group : one;
        const int ng = s->slab->group ? s->slab->group_size : 1;
        rc = grp[0]->slab->devcount ? slab_step_dev(grp, ng) : slab_step(grp, ng);
    }
    cudaSetDevice(prev);
    return rc;
}

// End-to-end call of a slab rank: HOST arrays in (n_in local particles by slot + global ids), nsteps, HOST arrays out (the
// particles this rank owns afterwards; *n_out of them, at most `capacity`).  The arrays should be pinned: every copy is
// asynchronous on the handle's stream and the call synchronises once, at the end.
int pbf_slab_step_host(pbf_handle s, float *pos4, float *vel4, uint32_t *gid, uint32_t n_in, uint32_t capacity,
                       uint32_t *n_out, int nsteps) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_step_host: slab not initialised"); return PBF_ERR_STATE; }
    if (!pos4 || !vel4 || !gid || !n_out) { pbf_set_error("pbf_slab_step_host: null buffer"); return PBF_ERR_INVALID; }
    if (n_in > s->cap || n_in > capacity) { pbf_set_error("pbf_slab_step_host: n exceeds capacity"); return PBF_ERR_CAPACITY; }
    if (s->slab->group) { pbf_set_error("pbf_slab_step_host: one rank per process only (virtual groups step together)"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    PBF_CUDA(cudaMemcpyAsync(s->pos, pos4, (size_t)n_in * 16, cudaMemcpyHostToDevice, s->stream));
    PBF_CUDA(cudaMemcpyAsync(s->vel, vel4, (size_t)n_in * 16, cudaMemcpyHostToDevice, s->stream));
    PBF_CUDA(cudaMemcpyAsync(s->slab->gid, gid, (size_t)n_in * 4, cudaMemcpyHostToDevice, s->stream));
    s->n = n_in;
    s->slab->n_local = n_in;
    if (s->slab->devcount) {
        PBF_CUDA(cudaMemcpyAsync(s->dn + DN_LOCAL, &n_in, 4, cudaMemcpyHostToDevice, s->stream));
        set_bounds_exact(s, n_in);
    }
    int rc = pbf_slab_step(s, nsteps);
    if (rc) return rc;
    if (s->slab->devcount) {       // the one wait of the call: the counts the copies below are sized by
        rc = sync_counts(s);
        if (rc) return rc;
    }
    const u32 m = s->slab->n_local;
    if (m > capacity) { pbf_set_error("pbf_slab_step_host: more particles than the caller's arrays hold"); return PBF_ERR_CAPACITY; }
    PBF_CUDA(cudaMemcpyAsync(pos4, s->pos, (size_t)m * 16, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaMemcpyAsync(vel4, s->vel, (size_t)m * 16, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaMemcpyAsync(gid, s->slab->gid, (size_t)m * 4, cudaMemcpyDeviceToHost, s->stream));
    PBF_CUDA(cudaStreamSynchronize(s->stream));
    *n_out = m;
    return PBF_OK;
}

// Load balancing.  pbf_slab_layer_counts: this rank's particles per global cell layer (gz_global counters, HOST array); the
// host runtime sums them over the ranks and picks new planes.  pbf_slab_set_planes: this rank owns [z_lo, z_hi) from the
// next step on -- every rank must be given consistent planes before any of them steps; the particles that now lie outside
// move to the neighbours through the next step's ordinary migration, so a plane may only move by what one step's record
// capacity can carry (a layer or two).  The window (z_hi - z_lo + 2 layers) must fit the grid depth the handle was created
// with.
int pbf_slab_layer_counts(pbf_handle s, uint32_t *counts) {
    if (!s || !s->slab || !counts) { pbf_set_error("pbf_slab_layer_counts: slab not initialised or null buffer"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    pbf_slab_state *b = s->slab;
    const size_t bytes = (size_t)s->grid.gz_global * 4;
    u32 *d = nullptr;
    PBF_CUDA(cudaMalloc(&d, bytes));
    cudaMemsetAsync(d, 0, bytes, s->stream);
    const NRef nloc = b->devcount ? NRef{b->bound_local, s->dn + DN_LOCAL} : NRef{b->n_local, nullptr};
    k_layer_counts<<<nb(nloc.n ? nloc.n : 1), 256, 0, s->stream>>>(nloc, s->pos, s->grid, d);
    s->launches++;
    cudaMemcpyAsync(counts, d, bytes, cudaMemcpyDeviceToHost, s->stream);
    const cudaError_t e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (e != cudaSuccess) { pbf_set_error(cudaGetErrorString(e)); return PBF_ERR_CUDA; }
    return PBF_OK;
}

int pbf_slab_set_planes(pbf_handle s, int z_lo, int z_hi) {
    if (!s || !s->slab) { pbf_set_error("pbf_slab_set_planes: slab not initialised"); return PBF_ERR_STATE; }
    pbf_slab_state *b = s->slab;
    if (z_hi - z_lo < 2 || z_lo < 0 || z_hi > s->grid.gz_global || z_hi - z_lo + 2 > s->cfg.grid[2]) {
        pbf_set_error("pbf_slab_set_planes: a slab owns at least 2 layers and its window must fit the handle's grid depth");
        return PBF_ERR_INVALID;
    }
    if ((!b->has[0] && z_lo != b->z_lo) || (!b->has[1] && z_hi != b->z_hi)) {
        pbf_set_error("pbf_slab_set_planes: the outer planes of the domain cannot move");
        return PBF_ERR_INVALID;
    }
    if (z_lo == b->z_lo && z_hi == b->z_hi) return PBF_OK;
    DeviceGuard guard(s->device);
    // the cell tables are addressed relative to the window: undo the last step's entries under the OLD window first
    s->launches += launch_unclear_cells(s);
    s->n_prev_sorted = 0;
    if (b->devcount) PBF_CUDA(cudaMemsetAsync(s->dn + DN_PREV, 0, 4, s->stream));
    b->z_lo = z_lo; b->z_hi = z_hi;
    s->grid.zoff = z_lo - 1;
    s->grid.gz = z_hi - z_lo + 2;
    PBF_CUDA(cudaGetLastError());
    return PBF_OK;
}

// PBF_SLAB_PHASES=1: device time of the last step's phases on this rank -- [0] predict + migration out + compaction,
// [1] arrivals + boundary + ghosts out, [2] ghosts in + sort + cells + halo index, [3] solver with its halo refreshes,
// [4] ghost velocities + vorticity with its halo refresh
int pbf_slab_phase_times(pbf_handle s, float ms[5]) {
    if (!s || !s->slab || !ms) { pbf_set_error("pbf_slab_phase_times: slab not initialised"); return PBF_ERR_STATE; }
    pbf_slab_state *b = s->slab->group ? s->slab->group[0]->slab : s->slab;
    if (!b->phases || !b->ph_valid) { pbf_set_error("pbf_slab_phase_times: set PBF_SLAB_PHASES=1 before pbf_slab_init and step once"); return PBF_ERR_STATE; }
    DeviceGuard guard(s->device);
    PBF_CUDA(cudaEventSynchronize(b->ph_ev[5]));
    for (int k = 0; k < 5; k++) PBF_CUDA(cudaEventElapsedTime(&ms[k], b->ph_ev[k], b->ph_ev[k + 1]));
    if (b->trace && b->tr_n > 1) {
        std::string line = "PBF_SLAB_TRACE rank " + std::to_string(b->rank) + ":";
        for (int k = 1; k < b->tr_n; k++) {
            float t = 0.0f;
            cudaEventElapsedTime(&t, b->tr_ev[k - 1], b->tr_ev[k]);
            char buf[64];
            snprintf(buf, sizeof(buf), " %s=%.3f", b->tr_label[k], t);
            line += buf;
        }
        fprintf(stderr, "%s\n", line.c_str());
    }
    return PBF_OK;
}

// out[0] local particles, [1] ghosts from z-, [2] ghosts from z+, [3] boundary sent to z-, [4] sent to z+,
// [5] particles migrated away so far, [6] exchanges so far, [7] bytes sent so far (low 32 bits)
int pbf_slab_stats(pbf_handle s, uint64_t out[8]) {
    if (!s || !s->slab || !out) { pbf_set_error("pbf_slab_stats: slab not initialised"); return PBF_ERR_STATE; }
    pbf_slab_state *b = s->slab;
    { DeviceGuard guard(s->device); const int rc = sync_counts(s); if (rc) return rc; }
    out[0] = b->n_local; out[1] = b->n_ghost[0]; out[2] = b->n_ghost[1]; out[3] = b->n_bnd[0]; out[4] = b->n_bnd[1];
    out[5] = b->migrated; out[6] = b->exchanges; out[7] = b->bytes_sent;
    return PBF_OK;
}

}  // extern "C"
