// sweeps.cu -- the four neighbour sweeps of SPH::Run (K8 calclambda.glsl, K9 updatepos.glsl, K11 vorticity.glsl as two
// kernels) and the per-step plan they run on.
//
// Candidate set = FOR_EACH_NEIGHBOUR (shaders/sph/foreachneighbour.glsl:1-10): for a particle in cell (x,y,z) the nine
// rows (y+dy, z+dz), each the cells x-1..x+1 merged into one run (neighbourcells.glsl:37-47, :62-84).  Cells are ordered
// x-fastest (key = x + z*gx + y*gx*gz), so for a TILE of TL = 128 consecutive sorted particles the runs of row offset
// o = (dy,dz) of all its particles lie inside ONE contiguous range of the sorted array (64 to 400 records at the
// headline scene, depending on how dense that row is; ~1140 records over the nine ranges, 8.9 per particle).
//   * k_plan (once per step, after the cell tables): per particle its nine runs (this is K7, evaluated once per step
//     as in the reference), per tile the nine ranges [S_o, S_o + n_o) that cover them; the runs are stored relative to
//     the tile's shared-memory image (11-bit start, 5-bit count: 20 B per particle), longest first: a warp walks run
//     slot k of all its lanes in lock step and pays for the longest, so runs of similar length share a slot.
//   * every sweep: nine lanes issue one 1-D bulk copy each (cp.async.bulk -> UBLKCP, completion on an mbarrier) that
//     bring the tile's nine ranges into shared memory verbatim -- no per-record instructions, no LSU wavefronts for
//     staging; meanwhile the other threads fetch their runs.  A thread then walks its nine runs in the image, two
//     candidates per iteration (two LDS.128, packed f32x2 arithmetic).  Neighbouring lanes read neighbouring records
//     of the same row, so the loads are free of bank conflicts.
//   * warp 2 of every block meanwhile pulls what the block one wave later (blockIdx.x + PBF_PREFETCH_DIST) will need --
//     descriptor, packed runs, the ranges of its own row and of the row above; the other seven ranges are those two of
//     its neighbours -- into L2 (cp.async.bulk.prefetch.L2).
// What bounds the sweeps is instruction issue (FP32 + shared-memory loads + run bookkeeping) at the residency the
// registers allow (32 warps per SM), not HBM: see DESIGN.md section 4 and profiles/ (with the arithmetic removed a
// sweep still takes 56 % of its time, with the shared-memory loads removed too 45 %).
//
// A tile whose ranges do not fit the image in up to four staging phases, or with a run of 32+ candidates (very sparse or
// over-dense scenes), takes the general path: every thread walks its nine runs straight from global memory through L1
// (the first version of these kernels).  Both paths visit exactly the same candidate set.
#include <limits.h>

#include "neighbour.cuh"

namespace {

#ifndef PBF_TL
#define PBF_TL 128
#endif
constexpr int TL = PBF_TL;       // particles per tile = threads per block
static_assert(TL % 32 == 0 && TL >= 32 && TL <= 1024, "a tile is a whole number of warps");
#ifndef PBF_TL_CAP
#define PBF_TL_CAP 1664
#endif
#ifndef PBF_TL_CTAS
#define PBF_TL_CTAS 8
#endif
#ifndef PBF_PREFETCH_DIST
#define PBF_PREFETCH_DIST 1184   // one wave of blocks ahead on a B200 (148 SMs x 8 blocks)
#endif
constexpr int TL_CAP = PBF_TL_CAP;   // records of one shared-memory image, 16 B each (26 KB: eight blocks of four warps per
                                     // SM; 2.5 % of the headline scene's tiles need a second phase).  Measured on B200 at 8M
                                     // particles, whole step: TL 256 / cap 3328 / 4 blocks 4.69 ms, TL 128 / 1664 / 8 4.63 ms,
                                     // TL 128 / 1920 / 7 4.83 ms, TL 64 / 832 / 16 4.78 ms
#ifndef PBF_WALK_UNROLL
#define PBF_WALK_UNROLL 1
#endif
constexpr int WALK_UNROLL = PBF_WALK_UNROLL;   // pair iterations unrolled in walk_run (measured: DESIGN.md section 4)
constexpr int TL_PHASES = 4;     // a tile whose nine ranges exceed one image stages them in up to four phases
constexpr int TL_DESC = 32;      // ints per tile descriptor: phases, records, nine range starts, nine range lengths, cuts
constexpr int D_MODE = 0, D_TOTAL = 1, D_S = 2, D_N = 11, D_CUT = 20, D_AT = 21;   // D_AT: image offset of every range (one-image tiles)
constexpr int RUN_WORDS = 5;     // packed runs of one particle: nine 16-bit fields {image index:11 | count:5}; then
                                 // bit 16 of word 4: the particle meets itself in one of its runs
static_assert(PBF_TL_CAP + 4 <= 2048, "run fields hold an 11-bit image index");
constexpr int TL_PAD = 4;        // zeroed records after the last range: a walk reads up to 3 records past its run
constexpr size_t TL_IMG = (size_t)(TL_CAP + TL_PAD) * sizeof(float4);   // one image
constexpr size_t TL_SMEM1 = TL_IMG;                                // PBF_TL_CTAS blocks per SM
constexpr size_t TL_SMEM2 = 2 * TL_IMG;                            // two arrays: half as many

// ---- the plan: one block per tile ---------------------------------------------------------------------------------------
template <bool LOOP>
__global__ void __launch_bounds__(TL)
k_plan(NRef nr, const u32 *__restrict__ home, const int2 *__restrict__ runs3, const int2 *__restrict__ cells,
       int *__restrict__ desc, u32 *__restrict__ runs, GridInfo g, int allow) {
    constexpr int NW = TL / 32;
    __shared__ int wS[NW][9], wE[NW][9];          // per warp: first / one-past-last sorted slot its runs of row o touch
    __shared__ int sS[9], sN[9], sBase[9], sAt[9];  // per tile: range start, length, (image offset of the range) - start, image offset
    __shared__ int sMeta[4];                       // phases, records, cuts, "every range fits one image"
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 n = LOOP ? nref(nr) : nr.n;
    // one tile per block; the loop only turns when the grid was sized for fewer particles than there are (NRef, LOOP)
    for (u32 tile = blockIdx.x; tile * (u32)TL < n; tile += LOOP ? gridDim.x : 0x1000000u) {
    const u32 i = tile * (u32)TL + tid;
    int2 r[9];
#pragma unroll
    for (int o = 0; o < 9; o++) r[o] = make_int2(-1, 0);
    if (i < n) fetch_runs(home[i], g, runs3, cells, r);
#pragma unroll
    for (int o = 0; o < 9; o++) {
        const bool has = r[o].y > 0;
        const int a = __reduce_min_sync(0xffffffffu, has ? r[o].x : INT_MAX);
        const int b = __reduce_max_sync(0xffffffffu, has ? r[o].x + r[o].y : -1);
        if (lane == 0) { wS[warp][o] = a; wE[warp][o] = b; }
    }
    __syncthreads();
    // The per-tile arithmetic once, not in every thread: lane o of warp 0 merges the warps' bounds of range o, lane 0
    // lays the ranges out.  Greedy phases: ranges are staged in order o = 0..8; a phase ends before the range that would
    // overflow the image.  cut = 4-bit first-range index of every phase, closed by 9 (one phase: 0x90).
    if (warp == 0) {
        if (lane < 9) {
            int a = INT_MAX, b = -1;
#pragma unroll
            for (int w = 0; w < NW; w++) { a = min(a, wS[w][lane]); b = max(b, wE[w][lane]); }
            sS[lane] = b >= 0 ? a : 0;
            sN[lane] = b >= 0 ? b - a : 0;
        }
        __syncwarp();
        if (lane == 0) {
            int total = 0, fill = 0, nph = 1, ok = 1;
            u32 cut = 0;
            for (int o = 0; o < 9; o++) {
                const int no = sN[o];
                if (fill + no > TL_CAP) { if (nph < 7) cut |= (u32)o << (4 * nph); nph++; fill = 0; }
                sBase[o] = fill - sS[o];
                sAt[o] = fill;
                ok = ok && no <= TL_CAP;
                fill += no;
                total += no;
            }
            if (nph < 8) cut |= 9u << (4 * nph);
            sMeta[0] = nph; sMeta[1] = total; sMeta[2] = (int)cut; sMeta[3] = ok;
        }
    }
    __syncthreads();
    const int nph = sMeta[0];
    bool fits = true;
    // nine 16-bit fields {image index of the run's first record:11 | count:5}, two per word, and in bit 16 of the last
    // word "one of the runs holds the particle itself": RUN_WORDS words per particle
    u32 v[9];
#pragma unroll
    for (int o = 0; o < 9; o++) {
        // run descriptor {image index:11 | count:5}; above it, for the sort below: {8 - row:4 | pair iterations:5}
        const u32 iters = (u32)(((r[o].y & 31) + 1) >> 1);
        v[o] = r[o].y > 0 ? (u32)((r[o].x + sBase[o]) & 0x7ff) | ((u32)(r[o].y & 31) << 11) | ((u32)(8 - o) << 16) | (iters << 20)
                          : 0u;
        fits = fits && r[o].y < 32;
    }
#ifndef PBF_PLAN_UNSORTED
    // A warp walks run slot k of all its lanes in lock step, so it pays max_lanes(iterations) per slot.  In a one-image
    // tile a run descriptor is self-contained (image index + count), so every lane may walk its nine runs in ANY order:
    // longest first puts runs of similar length into the same slot (DESIGN.md section 4 has the numbers: -17 % slots on
    // a disordered scene, nothing on a regular lattice).  Only the order of a particle's floating-point sums changes.
    // Tiles staged in phases keep the row order: there a run belongs to the phase of its range.
    if (nph == 1 && !(allow & 2)) {                   // allow bit 1: canonical order, the runs stay in row order
        // sort key {pair iterations = ceil(count / 2):5 | 8 - row:4 | descriptor:16}: what a slot costs is its number of
        // iterations, and runs that cost the same keep the row order -- on a regular lattice hardly anything moves, so
        // neighbouring lanes keep reading neighbouring records of the same row (no extra bank conflicts), and the order of
        // the sums does not depend on where the tile's ranges happen to lie
        // 25-comparator sorting network for nine keys, descending
#define PBF_CSWAP(a, b) { const u32 hi_ = max(v[a], v[b]), lo_ = min(v[a], v[b]); v[a] = hi_; v[b] = lo_; }
        PBF_CSWAP(0, 1) PBF_CSWAP(3, 4) PBF_CSWAP(6, 7)
        PBF_CSWAP(1, 2) PBF_CSWAP(4, 5) PBF_CSWAP(7, 8)
        PBF_CSWAP(0, 1) PBF_CSWAP(3, 4) PBF_CSWAP(6, 7)
        PBF_CSWAP(0, 3) PBF_CSWAP(3, 6) PBF_CSWAP(0, 3)
        PBF_CSWAP(1, 4) PBF_CSWAP(4, 7) PBF_CSWAP(1, 4)
        PBF_CSWAP(2, 5) PBF_CSWAP(5, 8) PBF_CSWAP(2, 5)
        PBF_CSWAP(1, 3) PBF_CSWAP(5, 7)
        PBF_CSWAP(2, 6) PBF_CSWAP(4, 6) PBF_CSWAP(2, 4)
        PBF_CSWAP(2, 3) PBF_CSWAP(5, 6)
#undef PBF_CSWAP
    }
#endif
    u32 w[RUN_WORDS];
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = (v[2 * k] & 0xffffu) | (v[2 * k + 1] << 16);
    w[4] = v[8] & 0xffffu;
    // Is the particle itself among its candidates (FOR_EACH_NEIGHBOUR skips it by index, foreachneighbour.glsl:9)?  Normally
    // it sits in run 4, its own row -- but a particle outside the grid is filed under its CLAMPED cell (findcells.glsl)
    // while its runs are built around the UNCLAMPED one (neighbourcells.glsl:57), so any of the nine runs may hold it.
    bool self_in = (int)i >= r[4].x && (int)i < r[4].x + r[4].y;
    if (!self_in) {                                          // rare: outside the grid, or the quirk's invisible cell
#pragma unroll
        for (int o = 0; o < 9; o++) self_in = self_in || ((int)i >= r[o].x && (int)i < r[o].x + r[o].y);
    }
    if (self_in) w[4] |= 1u << 16;
    u32 *out = runs + (size_t)tile * RUN_WORDS * TL + tid;
#pragma unroll
    for (int k = 0; k < RUN_WORDS; k++) out[k * TL] = w[k];
    fits = __syncthreads_and(fits) && sMeta[3] && nph <= TL_PHASES && (allow & 1);
    int *d = desc + (size_t)tile * TL_DESC;
    if (tid == 0) { d[D_MODE] = fits ? nph : 0; d[D_TOTAL] = sMeta[1]; d[D_CUT] = sMeta[2]; }
    if (tid < 9) { d[D_S + tid] = sS[tid]; d[D_N + tid] = sN[tid]; d[D_AT + tid] = sAt[tid]; }
    if (LOOP) __syncthreads();                     // the shared tables are free for the next tile
    }
}

// Fused halo pull (HaloPull, pbf_internal.cuh): halo_pull_run by the first blocks of the grid before their first tile,
// halo_pull_wait by every tile whose descriptor says that it touches the halo layers.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p, bool sys) {
    unsigned long long v;
    if (sys) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Wait until *p >= seq.  A neighbour that never delivers (a bug, or a rank that died) must not hang the GPU: after 20 s the
// wait gives up for good and sets bit 4 of dn[DN_OVERFLOW], which the host reports with the next call.
__device__ __forceinline__ void halo_spin(const unsigned long long *p, unsigned long long seq, bool sys, u32 *dn, unsigned ns) {
    unsigned long long t0 = 0;
    u32 it = 0;
    while (ld_acquire_u64(p, sys) < seq) {
        __nanosleep(ns);
        if ((++it & 255u) == 0u) {
            if (*reinterpret_cast<volatile u32 *>(dn + DN_OVERFLOW) & 16u) break;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 20000000000ull) { atomicOr(dn + DN_OVERFLOW, 16u); break; }
        }
    }
}

// the first blocks of the grid, before their first tile (they are resident before any block that waits for them)
__device__ __forceinline__ void halo_pull_run(const HaloPull &pl, int tid, int nthreads) {
    if (pl.e == 0u || blockIdx.x >= (unsigned)PBF_PULL_BLOCKS) return;      // uniform
    const u32 nlo = pl.dn[DN_HALO_N + 2], nhi = pl.dn[DN_HALO_N + 3];
    const unsigned long long seq = 256ull * pl.dn[DN_STEP] + pl.e;
    if (tid == 0) {
        if (nlo) halo_spin(pl.flag[0], seq, true, pl.dn, 64);
        if (nhi) halo_spin(pl.flag[1], seq, true, pl.dn, 64);
    }
    __syncthreads();
    const u32 nblk = min(gridDim.x, (unsigned)PBF_PULL_BLOCKS);
    // four records per thread and trip: the chains (slot index -> mailbox value -> store) are latency bound
    const u32 stride = nblk * (u32)nthreads, ntot = nlo + nhi;
    for (u32 k0 = blockIdx.x * (u32)nthreads + (u32)tid; k0 < ntot; k0 += 4u * stride) {
        u32 idx[4];
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const u32 k = k0 + (u32)u * stride;
            idx[u] = k < ntot ? pl.ghost_sorted[k] : 0u;
            const char *src = k < nlo ? pl.data[0] : pl.data[1];
            const u32 j = k < nlo ? k : k - nlo;
            q[u] = make_uint4(0u, 0u, 0u, 0u);
            if (k < ntot) {
                if (pl.wide) q[u] = __ldcv(reinterpret_cast<const uint4 *>(src) + j);
                else q[u].w = __ldcv(reinterpret_cast<const u32 *>(src) + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (k0 + (u32)u * stride >= ntot) continue;
            if (pl.wide) pl.buf[idx[u]] = make_float4(__uint_as_float(q[u].x), __uint_as_float(q[u].y), __uint_as_float(q[u].z), __uint_as_float(q[u].w));
            else pl.buf[idx[u]].w = __uint_as_float(q[u].w);
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(pl.ready + blockIdx.x), "l"(seq) : "memory");
}

__device__ __forceinline__ void halo_pull_wait(const HaloPull &pl, int tid) {
    const u32 nblk = min(gridDim.x, (unsigned)PBF_PULL_BLOCKS);
    const unsigned long long seq = 256ull * pl.dn[DN_STEP] + pl.e;
    if ((u32)tid < nblk) halo_spin(pl.ready + tid, seq, false, pl.dn, 32);
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");         // the bulk copies that follow read what generic stores just wrote
}

// ---- tile frame of the sweeps -------------------------------------------------------------------------------------------
struct TileCtx {
    u32 tile;            // this block's tile (normally blockIdx.x) and the number of tiles of the launch's particle count
    u32 ntiles;
    int mode;            // staging phases of the tiled path (1 for almost every tile), 0 = general path
    u32 flags;           // slab rank: PBF_TILE_PUSH / PBF_TILE_WAIT >> 8 (bits 8.. of the descriptor's mode word, set by k_halo_index
                         // after the plan)
    u32 cut;             // first range of every phase, 4 bits each, closed by 9
    bool self_in;        // FOR_EACH_NEIGHBOUR would have met (and skipped) the particle itself
    unsigned img;        // shared-window address of image 0 (image 1 follows at + TL_IMG)
    u32 run[9];          // this thread's nine runs {image index:11 | count:5}
};

__device__ __forceinline__ Pair make_pair(const float4 &a, const float4 &b) {
    Pair p;
    p.x = make_float2(a.x, b.x); p.y = make_float2(a.y, b.y); p.z = make_float2(a.z, b.z); p.w = make_float2(a.w, b.w);
    return p;
}

__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {   // bytes: multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Thread 0: stage ranges [lo, hi) of NSRC arrays into the image(s): zeroed pad records after the last range (a walk
// reads up to TL_PAD - 1 records past its run; they are masked but must be finite), then one arrive.expect_tx -- its
// release makes the pad visible to everybody who passes the wait -- and one bulk copy per non-empty range.
template <int NSRC>
__device__ __forceinline__ void tile_stage(unsigned char *dsm, unsigned mb, const float4 *__restrict__ src0,
                                           const float4 *__restrict__ src1, const int *__restrict__ dg, int lo, int hi) {
    float4 *sm0 = reinterpret_cast<float4 *>(dsm);
    float4 *sm1 = sm0 + (TL_CAP + TL_PAD);
    // the whole descriptor in five independent 16-byte loads, before anything waits on it
    const int4 *d4 = reinterpret_cast<const int4 *>(dg);
    const int4 q0 = __ldg(d4), q1 = __ldg(d4 + 1), q2 = __ldg(d4 + 2), q3 = __ldg(d4 + 3), q4 = __ldg(d4 + 4);
    const int dv[20] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w,
                        q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w};
    int total = 0;
#pragma unroll
    for (int o = 0; o < 9; o++) total += (o >= lo && o < hi) ? dv[D_N + o] : 0;
#pragma unroll
    for (int k = 0; k < TL_PAD; k++) {
        sm0[total + k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NSRC == 2) sm1[total + k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(total * 16 * NSRC)) : "memory");
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(sm0), a1 = (unsigned)__cvta_generic_to_shared(sm1);
    int at = 0;
#pragma unroll
    for (int o = 0; o < 9; o++) {
        const int so = dv[D_S + o], no = (o >= lo && o < hi) ? dv[D_N + o] : 0;
        if (no > 0) {
            bulk_g2s(a0 + 16u * at, src0 + so, 16u * no, mb);
            if (NSRC == 2) bulk_g2s(a1 + 16u * at, src1 + so, 16u * no, mb);
        }
        at += no;
    }
}

// All threads of the block call this.  Thread 0 initialises the mbarrier and stages the first phase; meanwhile the
// others fetch their runs (the caller overlaps its own loads before tile_sweep).
template <int NSRC, bool LOOP>
__device__ __forceinline__ TileCtx tile_begin(unsigned char *dsm, unsigned long long *mbar, const float4 *__restrict__ src0,
                                              const float4 *__restrict__ src1, const int *__restrict__ desc,
                                              const u32 *__restrict__ runs, int tid, u32 tile, u32 ntiles_, const HaloPull &pl,
                                              int ptid = -1) {
    // ptid: the thread's particle within the tile when it differs from tid (two threads per particle, k_vorticity_a)
    if (ptid < 0) ptid = tid;
    const int *dg = desc + (size_t)tile * TL_DESC;
    TileCtx c;
    c.tile = tile;
    c.ntiles = ntiles_;
    // The nine range lengths and starts of the one-image path are fetched TOGETHER with the mode word, not after it is known:
    // the descriptor spans three 32-byte sectors, and a second dependent round trip to L2 would sit on the critical path of
    // every tile (descriptor -> bulk copies -> image), during which the whole block idles.
    const int no_spec = tid < 9 ? __ldg(dg + D_N + tid) : 0;
    const int so_spec = tid < 9 ? __ldg(dg + D_S + tid) : 0;
    const int at_spec = tid < 9 ? __ldg(dg + D_AT + tid) : 0;      // image offsets laid out by k_plan: no scan here
    const int total_spec = tid == 0 ? __ldg(dg + D_TOTAL) : 0;
    c.mode = __ldg(dg + D_MODE);
    c.flags = 0u;
    if (LOOP) {                      // only a slab rank's descriptors carry flags (and only its kernels are LOOP kernels)
        c.flags = (u32)c.mode >> 8;
        c.mode &= 0xff;
        if (pl.e && (c.flags & 2u)) halo_pull_wait(pl, tid);     // block-uniform
    }
    c.cut = (u32)__ldg(dg + D_CUT);
    c.img = (unsigned)__cvta_generic_to_shared(dsm);
    const u32 *rp = runs + (size_t)tile * RUN_WORDS * TL + ptid;
    u32 w[RUN_WORDS];
#pragma unroll
    for (int k = 0; k < RUN_WORDS; k++) w[k] = __ldg(rp + k * TL);   // in flight while the bulk copies land
#pragma unroll
    for (int k = 0; k < 4; k++) { c.run[2 * k] = w[k] & 0xffffu; c.run[2 * k + 1] = w[k] >> 16; }
    c.run[8] = w[4] & 0xffffu;
    c.self_in = ((w[4] >> 16) & 1u) != 0u;
#if PBF_PREFETCH_DIST > 0
    // One wave ahead: pull what tile blockIdx.x + DIST will need into L2 -- its descriptor (by loading it), its packed
    // runs, the range of its OWN row (dy = dz = 0) and the range of the row above it (dy = +1, dz = 0) -- so that its prologue
    // (descriptor -> bulk copies -> data) runs on L2 hits instead of DRAM round trips.  The other seven ranges are covered by
    // the same two prefetches of the tiles around it: rows (y, z +- 1) and (y + 1, z +- 1) by the tiles two further on or
    // back, which prefetch at almost the same time, and the rows of layer y - 1 were staged one layer of tiles ago and are
    // still in L2.  Every record is prefetched twice, not nine times (PBF_PREFETCH_ALL: all nine ranges, 10 instead of 3
    // bulk prefetches per tile).
    constexpr int PF0 = TL >= 128 ? 64 : TL / 2;       // first of the prefetching threads (warp 2: idle until the image lands)
#if defined(PBF_PREFETCH_ALL)
    constexpr int PFN = 10;
#elif defined(PBF_PREFETCH_OWN_ROW)                   // own row only: enough while a layer of tiles is shorter than DIST
    constexpr int PFN = 2;
#else
    constexpr int PFN = 3;
#endif
    if (tid >= PF0 && tid < PF0 + PFN) {
        const u32 ft = tile + (u32)PBF_PREFETCH_DIST;
        if (ft < ntiles_) {
            const int *fd = desc + (size_t)ft * TL_DESC;
            const int k = tid - PF0;
            if (k < PFN - 1) {
                const int o = PFN <= 3 ? 4 + 3 * k : k;                // 4 = (dy 0, dz 0), 7 = (dy +1, dz 0)
                const int so = __ldg(fd + D_S + o), no = __ldg(fd + D_N + o);
                // the descriptor's other sectors (lengths, image offsets): loaded only to have them in L2 as well
                const int a0 = __ldg(fd + D_N + 8), a1 = __ldg(fd + D_AT + 8);
                if (__ldg(fd + D_MODE) && (no > 0) && (a0 | a1) >= 0) {
                    bulk_prefetch_l2(src0 + so, 16u * (unsigned)no);
                    if (NSRC == 2) bulk_prefetch_l2(src1 + so, 16u * (unsigned)no);
                }
            } else {
                bulk_prefetch_l2(runs + (size_t)ft * RUN_WORDS * TL, (unsigned)(RUN_WORDS * TL * sizeof(u32)));
            }
        }
    }
#endif
    if (c.mode == 1 && tid < 32) {
        // One image (all but a few per cent of the tiles): lanes 0..8 of warp 0 issue one bulk copy each -- range start,
        // length and image offset come straight from the descriptor -- instead of one thread issuing nine in a row
        // (~300 instructions on the block's critical path).
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        const int no = no_spec, so = so_spec;
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        float4 *sm0 = reinterpret_cast<float4 *>(dsm);
        float4 *sm1 = sm0 + (TL_CAP + TL_PAD);
        if (tid == 0) {
            const int total = total_spec;
#pragma unroll
            for (int k = 0; k < TL_PAD; k++) {              // see tile_stage
                sm0[total + k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (NSRC == 2) sm1[total + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(total * 16 * NSRC)) : "memory");
        }
        __syncwarp();
        if (no > 0) {
            const int at = at_spec;
            bulk_g2s((unsigned)__cvta_generic_to_shared(sm0) + 16u * at, src0 + so, 16u * no, mb);
            if (NSRC == 2) bulk_g2s((unsigned)__cvta_generic_to_shared(sm1) + 16u * at, src1 + so, 16u * no, mb);
        }
    } else if (c.mode > 1 && tid == 0) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tile_stage<NSRC>(dsm, mb, src0, src1, dg, 0, (int)((c.cut >> 4) & 15u));
    }
    return c;
}

__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(mb), "r"(parity) : "memory");
}

__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Tiled sweep of one particle, all threads of the block (threads without a particle hold nine empty runs): per
// phase wait for the image, then walk the runs staged in it, two candidates per iteration.
// body(candidate pair of array 0, same pair of array 1, valid0, valid1)
template <int NSRC, class F>
__device__ __forceinline__ void walk_run(unsigned a, const unsigned end, F &body) {
    // second candidate of a pair valid <=> a + 16 < end; inside the loop end >= a + 16 >= 16, so the comparison against
    // end - 16 (hoisted) is the same and saves an add per iteration
    const unsigned last = end - 16u;
#pragma unroll WALK_UNROLL
    for (; a < end; a += 32u) {
        const Pair p = make_pair(lds128(a), lds128(a + 16u));
        if (NSRC == 2) body(p, make_pair(lds128(a + (unsigned)TL_IMG), lds128(a + (unsigned)TL_IMG + 16u)), true, a < last);
        else body(p, p, true, a < last);
    }
}

// split (two threads per particle): 0 = this thread walks all nine runs; 1 / 2 = in a one-image tile the even / odd run
// slots (the runs are sorted by length, so the halves weigh about the same), in a tile staged in phases all / none; 3 = none
// (canonical order: the first thread walks everything in row order).
template <int NSRC, class F>
__device__ __forceinline__ void tile_sweep(TileCtx &c, unsigned char *dsm, unsigned long long *mbar,
                                           const float4 *__restrict__ src0, const float4 *__restrict__ src1,
                                           const int *__restrict__ desc, int tid, F body, int split = 0) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
    __syncthreads();                                       // the barrier is initialised
    if (c.mode == 1) {                                     // all nine ranges in one image (all but a handful of tiles):
        mbar_wait(mb, 0u);                                 // no per-run phase test
#pragma unroll
        for (int o = 0; o < 9; o++) {
            if (split && ((o & 1) + 1) != split) continue;
            const unsigned a = c.img + 16u * (c.run[o] & 0x7ffu);
            walk_run<NSRC>(a, a + 16u * (c.run[o] >> 11), body);
        }
        return;
    }
    int lo = 0;
#pragma unroll 1
    for (int ph = 0; ph < c.mode; ph++) {
        const int hi = (int)((c.cut >> (4 * ph + 4)) & 15u);
        if (ph > 0) {
            __syncthreads();                               // everybody is done with the previous phase's image
            if (tid == 0) tile_stage<NSRC>(dsm, mb, src0, src1, desc + (size_t)c.tile * TL_DESC, lo, hi);
        }
        mbar_wait(mb, (unsigned)(ph & 1));
#pragma unroll 1
        for (int o = lo; o < (split >= 2 ? lo : hi); o++) {   // rare path: rolled, the runs come from a switch
            u32 r = 0;
#pragma unroll
            for (int k = 0; k < 9; k++) r = (o == k) ? c.run[k] : r;
            const unsigned a = c.img + 16u * (r & 0x7ffu);
            walk_run<NSRC>(a, a + 16u * (r >> 11), body);
        }
        lo = hi;
    }
}

// General walk of one particle: its nine merged runs from global memory (aligned pairs, masked edges).
// CG: candidate loads through L2 only (kernels whose first blocks overwrite ghost slots of the array they sweep)
template <int NSRC, int RAD, bool CG, class F>
__device__ __forceinline__ void general_walk(TileCtx &c, unsigned char *dsm, const float4 *__restrict__ src0,
                                             const float4 *__restrict__ src1, const u32 *__restrict__ home,
                                             const int2 *__restrict__ runs3, const int2 *__restrict__ cells,
                                             const GridInfo &g, u32 i, int tid, bool rel, F body) {
    int2 *srun = reinterpret_cast<int2 *>(dsm);            // 9 (25 with full support) x TL run descriptors
    static_assert((size_t)(2 * RAD + 1) * (2 * RAD + 1) * TL * sizeof(int2) <= TL_SMEM1, "run descriptors fit the image area");
    int slots_;
    load_runs<TL, RAD>(home[i], i, g, runs3, cells, srun, tid, &slots_, &c.self_in, rel);
    for_each_pair<TL>(srun, tid, slots_, rel, [&](int c0, bool v0, bool v1) {
        if (rel) {
            // canonical order: the pair starts at the run's first candidate (any parity): two 16-byte loads, the second only
            // if it belongs to the run (it may lie past the end of the array)
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const Pair p = make_pair(__ldcg(src0 + c0), v1 ? __ldcg(src0 + c0 + 1) : z4);
            if (NSRC == 2) body(p, make_pair(__ldcg(src1 + c0), v1 ? __ldcg(src1 + c0 + 1) : z4), v0, v1);
            else body(p, p, v0, v1);
            return;
        }
        const Pair p = CG ? ldcg_pair(src0 + c0) : ldg_pair(src0 + c0);
        if (NSRC == 2) body(p, CG ? ldcg_pair(src1 + c0) : ldg_pair(src1 + c0), v0, v1);
        else body(p, p, v0, v1);
    });
}

// both paths; every thread of the block calls this (the tiled path synchronises the block), `live` = has a particle
template <int NSRC, bool FULL, bool CG, class F>
__device__ __forceinline__ void walk(TileCtx &c, unsigned char *dsm, unsigned long long *mbar, const float4 *__restrict__ src0,
                                     const float4 *__restrict__ src1, const u32 *__restrict__ home,
                                     const int2 *__restrict__ runs3, const int2 *__restrict__ cells,
                                     const int *__restrict__ desc, const GridInfo &g, u32 i, bool live, int tid, bool rel, F body,
                                     int split = 0) {
    // split >= 2: a particle's second thread (3: it has no share at all, canonical order) -- none in the general walk
    if (split >= 2 && (FULL || c.mode == 0)) return;
    if (FULL) {                                            // full-support search: 25 rows of five cells from global memory
        if (live) general_walk<NSRC, 2, CG>(c, dsm, src0, src1, home, runs3, cells, g, i, tid, rel, body);
    } else if (c.mode) {
        tile_sweep<NSRC>(c, dsm, mbar, src0, src1, desc, tid, body, split);
    } else if (live) {
        general_walk<NSRC, 1, CG>(c, dsm, src0, src1, home, runs3, cells, g, i, tid, rel, body);
    }
}

// tile frame of a FULL kernel: no plan, no image
__device__ __forceinline__ TileCtx tile_none(u32 tile, u32 ntiles_) {
    TileCtx c;
    c.tile = tile; c.ntiles = ntiles_; c.mode = 0; c.flags = 0u; c.cut = 0; c.self_in = false; c.img = 0;
    return c;
}

// Fused halo push (slab runtime): `wide` sends the 16-byte record, otherwise its .w; every thread of the block calls this
// after its output is written.
__device__ __forceinline__ void halo_push(const HaloPush &hp, u32 i, bool live, const float4 &out, bool wide, int tid) {
    if (hp.map == nullptr) return;                           // uniform: single-GPU handles stop here
    bool wrote = false;
    if (live) {
        const u32 t = hp.map[i];
        if (t) {
            char *dst = (t >> 31) ? hp.data[1] : hp.data[0];         // selects, not indexing: the block stays in constant memory
            const u32 k = (t & 0x7fffffffu) - 1u;
            if (wide) reinterpret_cast<float4 *>(dst)[k] = out;
            else reinterpret_cast<float *>(dst)[k] = out.w;
            wrote = true;
        }
    }
    // Only tiles with boundary particles (about 2 % of them) pay for the system-scope fence and the counter; their number
    // is known from the halo index of this step.  ONE fence per tile: the block barrier orders every thread's stores before
    // thread 0's fence, and fences are cumulative (with a fence per storing thread a sweep took 4 us longer).
    if (!__syncthreads_or(wrote)) return;
    if (tid == 0) __threadfence_system();
    if (tid == 0 && atomicAdd(hp.done, 1u) + 1u == *hp.expect) {
        *hp.done = 0u;
        __threadfence_system();
        // counts on the device (one graph per step): the refresh number comes from the step counter
        const unsigned long long seq = hp.dn ? 256ull * hp.dn[DN_STEP] + hp.e : hp.seq;
        const u32 c0 = hp.dn ? hp.dn[DN_HALO_N] : hp.count[0], c1 = hp.dn ? hp.dn[DN_HALO_N + 1] : hp.count[1];
        if (c0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hp.flag[0]), "l"(seq) : "memory");
        if (c1) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hp.flag[1]), "l"(seq) : "memory");
    }
}

// Every sweep kernel: one tile per block; the loop only turns when the grid was sized for fewer particles than there are
// (NRef: a slab rank's count lives on the device).  Before a block reuses its image and its mbarrier for another tile
// everybody must be done with them and the barrier object must be invalidated.
// LOOP = false (count known on the host, grid = tiles): the loop folds into `if (blockIdx.x < tiles)`, the original kernel.
#define TILE_LOOP_BEGIN                                                                  \
    const u32 n = LOOP ? nref(nr) : nr.n, ntl = (n + (u32)TL - 1u) / (u32)TL;            \
    for (u32 tile = blockIdx.x; tile < ntl; tile += LOOP ? gridDim.x : 0x40000000u) {
#define TILE_LOOP_END(tc)                                                                \
        if (LOOP && tile + gridDim.x < ntl) {                                            \
            __syncthreads();                                                             \
            if ((tc).mode && tid == 0)                                                   \
                asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(&mbar)) : "memory"); \
            __syncthreads();                                                             \
        }                                                                                \
    }

#define TILE_ARGS const u32 *__restrict__ home, const int2 *__restrict__ runs3, const int2 *__restrict__ cells,   \
                  const int *__restrict__ desc, const u32 *__restrict__ runs

// ---- K8 calclambda.glsl:66-103 ------------------------------------------------------------------------------------
// out {x,y,z,lambda}.  rho (self excluded), S = sum |g_j|^2 + |sum g_j|^2, lambda = -C/(S+eps).
template <bool DIAG, bool FULL, bool LOOP>
__global__ void __launch_bounds__(TL, PBF_TL_CTAS)
k_lambda(NRef nr, const float4 *__restrict__ A, TILE_ARGS, float4 *__restrict__ B, GridInfo g, SimParams P, double *diag,
         const HaloPush hp, const HaloPull pl) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (LOOP) halo_pull_run(pl, tid, TL);
    TILE_LOOP_BEGIN
    TileCtx tc = FULL ? tile_none(tile, ntl) : tile_begin<1, LOOP>(dsm, &mbar, A, A, desc, runs, tid, tile, ntl, pl);
    const u32 tflag = LOOP ? tc.flags : 1u;
    const u32 i = tile * TL + tid;
    const bool live = i < n;
    float err = 0.0f;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 pi = live ? (LOOP ? __ldcg(A + i) : A[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 rho = make_float2(0.f, 0.f), S = rho, gx = rho, gy = rho, gz = rho;
    walk<1, FULL, LOOP>(tc, dsm, &mbar, A, A, home, runs3, cells, desc, g, i, live, tid, P.canonical != 0, [&](const Pair &c, const Pair &, bool v0, bool v1) {
            const PairGeom q = pair_geom(pi, c, v0, v1);
            rho = __ffma2_rn(__fmul2_rn(q.t, q.t), q.t, rho);                 // -sum (h^2-r^2)^3 (q.t is negated)
            const float2 tt = __fmul2_rn(q.t2, q.t2);                         // (h-l)^2
            const float2 cc = __fmul2_rn(tt, q.il);                           // (h-l)^2 / l
            // sum |grad|^2 (up to a constant).  Not (h-l)^4 = tt * tt, one op less: the exact r2 is what makes a coincident
            // particle (and the particle itself) contribute nothing, as gradWspiky's l == 0 branch demands
            S = __ffma2_rn(__fmul2_rn(cc, cc), q.r2, S);
            gx = __ffma2_rn(cc, q.dx, gx);
            gy = __ffma2_rn(cc, q.dy, gy);
            gz = __ffma2_rn(cc, q.dz, gz);
    });
    if (live) {
        // FOR_EACH_NEIGHBOUR skips j == i (foreachneighbour.glsl:9): self only ever adds (h^2)^3 = 64 to the poly6 sum
        float rs = -(rho.x + rho.y);
        if (tc.self_in != (P.self_term != 0)) rs += P.self_term ? 64.0f : -64.0f;   // pbf_options::density_self_term keeps it
        const float r = POLY6 * rs;
        const float cg = SPIKY_GRAD * P.one_over_rho_0;
        const float sx = cg * (gx.x + gx.y), sy = cg * (gy.x + gy.y), sz = cg * (gz.x + gz.y);
        const float Ssum = cg * cg * (S.x + S.y) + (sx * sx + sy * sy + sz * sz);
        const float C = r * P.one_over_rho_0 - 1.0f;
        if (DIAG) err = fabsf(C);
        else B[i] = out = make_float4(pi.x, pi.y, pi.z, -C / (Ssum + P.epsilon));
    }
    if (!DIAG && (tflag & 1u)) halo_push(hp, i, live, out, false, tid);
    if (DIAG) {
        __shared__ float red[TL / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if ((tid & 31) == 0) red[tid >> 5] = err;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < TL / 32; w++) s += (double)red[w];
            atomicAdd(diag, s);
        }
        __syncthreads();
    }
    TILE_LOOP_END(tc)
}

// ---- K9 updatepos.glsl:43-105, Jacobi: reads B {p, lambda}, writes A -----------------------------------------------
// FINAL (the last solver iteration of a step): K10 update.glsl:16-28 runs in the epilogue -- the thread that has just
// computed a particle's final position also derives its velocity and writes both back by particle id.  The gather of the
// old position and the two scatters are uncoalesced 16-byte accesses (a separate kernel ran them at 2.7 TB/s, 0.19 ms);
// here they hide under a sweep that leaves DRAM idle.  FINAL = 1: velocity by id; FINAL = 2: into the sorted array the
// vorticity kernels read.
struct UpdateArgs {
    const u32 *perm;
    float4 *pos, *vel, *svel;
    const u32 *n_local;      // slab rank: ids at and beyond *n_local are ghosts, whose by-slot position (the owner's OLD
                             // position) must survive for k_ghost_velocity; null on a single domain
};

template <int FINAL, bool FULL, bool LOOP>
__global__ void __launch_bounds__(TL, PBF_TL_CTAS)
k_delta_p(NRef nr, const float4 *__restrict__ B, TILE_ARGS, float4 *__restrict__ A, GridInfo g, SimParams P, const HaloPush hp,
          const UpdateArgs up, const HaloPull pl) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (LOOP) halo_pull_run(pl, tid, TL);
    TILE_LOOP_BEGIN
    TileCtx tc = FULL ? tile_none(tile, ntl) : tile_begin<1, LOOP>(dsm, &mbar, B, B, desc, runs, tid, tile, ntl, pl);
    const u32 tflag = LOOP ? tc.flags : 1u;
    const u32 i = tile * TL + tid;
    const bool live = i < n;
    const float4 pi = live ? (LOOP ? __ldcg(B + i) : B[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    u32 id = 0;
    float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FINAL && live) {                                    // in flight while the sweep runs
        id = __ldg(up.perm + i);
        old = up.pos[id];
    }
    float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
    // scorr = -k (scale W)^4 = -(k scale^4 POLY6^4) t^12 with t = max(h^2 - r^2, 0)        (updatepos.glsl:57-60)
    float sc4 = P.tensile_scale * POLY6;
    sc4 *= sc4;
    sc4 *= sc4;
    const float nk = -P.tensile_k * sc4;
    const float2 nk2 = make_float2(nk, nk), li2 = make_float2(pi.w, pi.w);
    walk<1, FULL, LOOP>(tc, dsm, &mbar, B, B, home, runs3, cells, desc, g, i, live, tid, P.canonical != 0, [&](const Pair &c, const Pair &, bool v0, bool v1) {
        const PairGeom q = pair_geom(pi, c, v0, v1);
        float2 t3 = __fmul2_rn(__fmul2_rn(q.t, q.t), q.t);
        t3 = __fmul2_rn(t3, t3);
        t3 = __fmul2_rn(t3, t3);                                              // t^12
        const float2 f = __ffma2_rn(nk2, t3, __fadd2_rn(li2, c.w));           // lambda_i + lambda_j + scorr
        const float2 cc = __fmul2_rn(f, __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il));
        ax = __ffma2_rn(cc, q.dx, ax);
        ay = __ffma2_rn(cc, q.dy, ay);
        az = __ffma2_rn(cc, q.dz, az);
    });
    const float s = SPIKY_GRAD * P.one_over_rho_0;
    float x = pi.x + s * (ax.x + ax.y), y = pi.y + s * (ay.x + ay.y), z = pi.z + s * (az.x + az.y);
    x = fminf(fmaxf(x, g.wlo[0]), g.whi[0]);                               // updatepos.glsl:98-100
    y = fminf(fmaxf(y, g.wlo[1]), g.whi[1]);
    z = fminf(fmaxf(z, g.wlo[2]), g.whi[2]);
    const float4 out = make_float4(x, y, z, 0.0f);
    if (live) A[i] = out;
    if (FINAL && live) {                                    // update.glsl:16-28, same arithmetic as k_update
        float4 v;
        v.x = __fdiv_rn(__fsub_rn(x, old.x), P.timestep);
        v.y = __fdiv_rn(__fsub_rn(y, old.y), P.timestep);
        v.z = __fdiv_rn(__fsub_rn(z, old.z), P.timestep);
        v.w = 0.0f;
        if (P.restitution >= 0.0f) {
            if ((x <= g.wlo[0] && v.x < 0.0f) || (x >= g.whi[0] && v.x > 0.0f)) v.x *= -P.restitution;
            if ((y <= g.wlo[1] && v.y < 0.0f) || (y >= g.whi[1] && v.y > 0.0f)) v.y *= -P.restitution;
            if ((z <= g.wlo[2] && v.z < 0.0f) || (z >= g.whi[2] && v.z > 0.0f)) v.z *= -P.restitution;
        }
        if (up.n_local == nullptr || id < *up.n_local) {
            up.pos[id] = out;
            if (FINAL == 1) up.vel[id] = v;
        }
        if (FINAL == 2) up.svel[i] = v;
    }
    if (tflag & 1u) halo_push(hp, i, live, out, true, tid);
    TILE_LOOP_END(tc)
}

// ---- K11 vorticity.glsl:34-60 (first sweep): XSPH + vorticity ------------------------------------------------------------
// out: vprime = v + c*sum v_ij W, omega = sum v_ij x gradW, B = {p, |omega|}
// TWO THREADS PER PARTICLE: this kernel stages two images (positions and velocities, 52 KB), so only four blocks fit an SM;
// with one thread per particle that is 16 warps per SM, and the kernel spent 45 % of its issue slots waiting (ncu r02h).
// Blocks of 2 x TL threads give it the 32 warps the other sweeps have: thread t and thread t + TL share particle t, the
// first walks the even run slots, the second the odd ones (the plan sorts a particle's runs by length, so the halves weigh
// about the same), and the partial sums meet in shared memory.
template <bool FULL, bool LOOP>
__global__ void __launch_bounds__(2 * TL, 4)
k_vorticity_a(NRef nr, const float4 *__restrict__ A, const float4 *__restrict__ svel, TILE_ARGS, float4 *__restrict__ B,
              float4 *__restrict__ vprime, float4 *__restrict__ omega, GridInfo g, SimParams P, const HaloPush hp) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar;
    float (*part)[TL] = reinterpret_cast<float (*)[TL]>(dsm);   // 3 KB of the image area, once everybody is done walking it
    const int tid = threadIdx.x, ptid = tid & (TL - 1), second = tid >= TL ? 1 : 0;
    TILE_LOOP_BEGIN
    TileCtx tc = FULL ? tile_none(tile, ntl) : tile_begin<2, LOOP>(dsm, &mbar, A, svel, desc, runs, tid, tile, ntl, HaloPull{}, ptid);
    const u32 tflag = LOOP ? tc.flags : 1u;
    const u32 i = tile * TL + ptid;
    const bool live = i < n;
    const float4 pi = live ? A[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 vi = live ? svel[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 vx = make_float2(0.f, 0.f), vy = vx, vz = vx, wx = vx, wy = vx, wz = vx;
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    walk<2, FULL, LOOP>(tc, dsm, &mbar, A, svel, home, runs3, cells, desc, g, i, live, tid, P.canonical != 0, [&](const Pair &c, const Pair &u, bool v0, bool v1) {
        const PairGeom q = pair_geom(pi, c, v0, v1);
        const float2 ux = make_float2(u.x.x - vi.x, u.x.y - vi.x);           // v_ij = v_j - v_i
        const float2 uy = make_float2(u.y.x - vi.y, u.y.y - vi.y);
        const float2 uz = make_float2(u.z.x - vi.z, u.z.y - vi.z);
        const float2 w = __fmul2_rn(__fmul2_rn(q.t, q.t), q.t);              // -Wpoly6 / POLY6 (q.t is negated)
        vx = __ffma2_rn(ux, w, vx);
        vy = __ffma2_rn(uy, w, vy);
        vz = __ffma2_rn(uz, w, vz);
        const float2 cc = __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il);          // grad = SPIKY_GRAD * cc * d
        const float2 gx = __fmul2_rn(cc, q.dx), gy = __fmul2_rn(cc, q.dy), gz = __fmul2_rn(cc, q.dz);
        // cross(v_ij, grad)
        wx = __ffma2_rn(uy, gz, __ffma2_rn(__fmul2_rn(gy, uz), neg1, wx));
        wy = __ffma2_rn(uz, gx, __ffma2_rn(__fmul2_rn(gz, ux), neg1, wy));
        wz = __ffma2_rn(ux, gy, __ffma2_rn(__fmul2_rn(gx, uy), neg1, wz));
    }, P.canonical ? (second ? 3 : 0) : 1 + second);
    // the second thread hands its share over (zeros where it had none: general path, tiles staged in phases)
    __syncthreads();
    if (second) {
        part[0][ptid] = vx.x + vx.y; part[1][ptid] = vy.x + vy.y; part[2][ptid] = vz.x + vz.y;
        part[3][ptid] = wx.x + wx.y; part[4][ptid] = wy.x + wy.y; part[5][ptid] = wz.x + wz.y;
    }
    __syncthreads();
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool mine = live && !second;
    if (mine) {
        const float sx = (vx.x + vx.y) + part[0][ptid], sy = (vy.x + vy.y) + part[1][ptid], sz = (vz.x + vz.y) + part[2][ptid];
        const float cw = -P.xsph_c * POLY6;
        const float ox = SPIKY_GRAD * ((wx.x + wx.y) + part[3][ptid]), oy = SPIKY_GRAD * ((wy.x + wy.y) + part[4][ptid]),
                    oz = SPIKY_GRAD * ((wz.x + wz.y) + part[5][ptid]);
        out = make_float4(pi.x, pi.y, pi.z, sqrtf(ox * ox + oy * oy + oz * oz));   // vorticity.glsl:60
        vprime[i] = make_float4(vi.x + cw * sx, vi.y + cw * sy, vi.z + cw * sz, 0.0f);
        omega[i] = make_float4(ox, oy, oz, 0.0f);
        B[i] = out;
    }
    if (tflag & 1u) halo_push(hp, i, mine, out, false, tid);
    if (LOOP) __syncthreads();                             // part[] is free for the next tile
    TILE_LOOP_END(tc)
}

// ---- K11 vorticity.glsl:65-85 (second sweep): confinement force, velocity[id] written once -------------------------------
template <bool FULL, bool LOOP>
__global__ void __launch_bounds__(TL, PBF_TL_CTAS)
k_vorticity_b(NRef nr, const float4 *__restrict__ B, const float4 *__restrict__ vprime, const float4 *__restrict__ omega,
              const u32 *__restrict__ perm, TILE_ARGS, float4 *__restrict__ vel, GridInfo g, SimParams P, const HaloPull pl) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (LOOP) halo_pull_run(pl, tid, TL);
    TILE_LOOP_BEGIN
    TileCtx tc = FULL ? tile_none(tile, ntl) : tile_begin<1, LOOP>(dsm, &mbar, B, B, desc, runs, tid, tile, ntl, pl);
    const u32 i = tile * TL + tid;
    const bool live = i < n;
    const float4 pi = live ? (LOOP ? __ldcg(B + i) : B[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 ex = make_float2(0.f, 0.f), ey = ex, ez = ex;
    walk<1, FULL, LOOP>(tc, dsm, &mbar, B, B, home, runs3, cells, desc, g, i, live, tid, P.canonical != 0, [&](const Pair &c, const Pair &, bool v0, bool v1) {
        const PairGeom q = pair_geom(pi, c, v0, v1);
        const float2 cc = __fmul2_rn(c.w, __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il));   // |omega_j| * grad factor
        ex = __ffma2_rn(cc, q.dx, ex);
        ey = __ffma2_rn(cc, q.dy, ey);
        ez = __ffma2_rn(cc, q.dz, ez);
    });
    if (live) {
        float nx = SPIKY_GRAD * (ex.x + ex.y), ny = SPIKY_GRAD * (ey.x + ey.y), nz = SPIKY_GRAD * (ez.x + ez.y);
        const float l = sqrtf(nx * nx + ny * ny + nz * nz);
        if (l > 0.0f) { nx /= l; ny /= l; nz /= l; }
        const float4 w = omega[i];
        const float4 v = vprime[i];
        const float s = P.timestep * P.vort_eps;
        vel[perm[i]] = make_float4(v.x + s * (ny * w.z - w.y * nz), v.y + s * (nz * w.x - w.z * nx),
                                   v.z + s * (nx * w.y - w.x * ny), 0.0f);   // cross(N, omega)
    }
    TILE_LOOP_END(tc)
}

inline int ntiles(u32 n) { return (int)((n + TL - 1) / TL); }

}  // namespace

#define TILE_PASS s->home, s->runs3, s->cells, s->tile_desc, s->tile_runs

u32 plan_tile_size(void) { return (u32)TL; }
u32 plan_desc_stride(void) { return (u32)TL_DESC; }
size_t plan_desc_ints(u32 cap) { return (size_t)ntiles(cap) * TL_DESC; }
size_t plan_run_words(u32 cap) { return (size_t)ntiles(cap) * RUN_WORDS * TL; }

int sweeps_init(void) {
    cudaError_t e = cudaSuccess;
    auto opt = [&](const void *f, size_t bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
#define OPT_ALL(kern, smem, ...)                                                                      \
    opt((const void *)kern<__VA_ARGS__, false, false>, smem); opt((const void *)kern<__VA_ARGS__, false, true>, smem); \
    opt((const void *)kern<__VA_ARGS__, true, false>, smem); opt((const void *)kern<__VA_ARGS__, true, true>, smem);
    OPT_ALL(k_lambda, TL_SMEM1, false) OPT_ALL(k_lambda, TL_SMEM1, true)
    OPT_ALL(k_delta_p, TL_SMEM1, 0) OPT_ALL(k_delta_p, TL_SMEM1, 1) OPT_ALL(k_delta_p, TL_SMEM1, 2)
#undef OPT_ALL
    opt((const void *)k_vorticity_b<false, false>, TL_SMEM1); opt((const void *)k_vorticity_b<false, true>, TL_SMEM1);
    opt((const void *)k_vorticity_b<true, false>, TL_SMEM1); opt((const void *)k_vorticity_b<true, true>, TL_SMEM1);
    opt((const void *)k_vorticity_a<false, false>, TL_SMEM2); opt((const void *)k_vorticity_a<false, true>, TL_SMEM2);
    opt((const void *)k_vorticity_a<true, false>, TL_SMEM2); opt((const void *)k_vorticity_a<true, true>, TL_SMEM2);
    return e == cudaSuccess ? 0 : -1;
}

// pbf_options::full_support: the 5 x 5 x 5 search needs no plan (every particle walks its 25 rows from global memory)
static inline bool full(const pbf_sim *s) { return s->options.full_support != 0; }

int launch_plan(pbf_sim *s) {
    if (full(s)) return 0;
    if (s->n_dev)
        k_plan<true><<<ntiles(s->n), TL, 0, s->stream>>>(nref_total(s), s->home, s->runs3, s->cells, s->tile_desc, s->tile_runs, s->grid,
                                                         (s->tiled_sweeps ? 1 : 0) | (s->canonical ? 2 : 0));
    else
        k_plan<false><<<ntiles(s->n), TL, 0, s->stream>>>(nref_total(s), s->home, s->runs3, s->cells, s->tile_desc, s->tile_runs, s->grid,
                                                          (s->tiled_sweeps ? 1 : 0) | (s->canonical ? 2 : 0));
    return 1;
}

static const HaloPush NO_PUSH = {};
static const HaloPull NO_PULL = {};
// kernel<..., FULL, LOOP>: FULL = 5 x 5 x 5 search, LOOP = the count lives on the device (slab rank)
#define SWEEP_LAUNCH_T(K, threads, smem, ...)                                                                     \
    do {                                                                                                          \
        const bool loop_ = s->n_dev != nullptr;                                                                   \
        if (full(s)) { if (loop_) K(true, true)<<<ntiles(s->n), threads, smem, s->stream>>>(__VA_ARGS__);         \
                       else K(true, false)<<<ntiles(s->n), threads, smem, s->stream>>>(__VA_ARGS__); }            \
        else { if (loop_) K(false, true)<<<ntiles(s->n), threads, smem, s->stream>>>(__VA_ARGS__);                \
               else K(false, false)<<<ntiles(s->n), threads, smem, s->stream>>>(__VA_ARGS__); }                   \
    } while (0)
#define SWEEP_LAUNCH(K, smem, ...) SWEEP_LAUNCH_T(K, TL, smem, __VA_ARGS__)
#define KL_LAMBDA(F, L) k_lambda<false, F, L>
#define KL_DIAG(F, L) k_lambda<true, F, L>
#define KL_DP0(F, L) k_delta_p<0, F, L>
#define KL_DP1(F, L) k_delta_p<1, F, L>
#define KL_DP2(F, L) k_delta_p<2, F, L>
#define KL_VA(F, L) k_vorticity_a<F, L>
#define KL_VB(F, L) k_vorticity_b<F, L>

// push: the sweep stores its boundary particles' results into the neighbours' mailboxes; pull: its first blocks fetch the
// ghost particles' inputs from this rank's mailbox (slab ranks only, slab.cu)
int launch_lambda(pbf_sim *s, const HaloPush *push, const HaloPull *pull) {
    SWEEP_LAUNCH(KL_LAMBDA, TL_SMEM1, nref_total(s), s->bufA, TILE_PASS, s->bufB, s->grid,
                 sim_params(s), nullptr, push ? *push : NO_PUSH, pull ? *pull : NO_PULL);
    return 1;
}

int launch_delta_p(pbf_sim *s, const HaloPush *push, const HaloPull *pull) {
    SWEEP_LAUNCH(KL_DP0, TL_SMEM1, nref_total(s), s->bufB, TILE_PASS, s->bufA, s->grid,
                 sim_params(s), push ? *push : NO_PUSH, UpdateArgs{}, pull ? *pull : NO_PULL);
    return 1;
}

// the last solver iteration of a step: delta-p with update.glsl in its epilogue (replaces launch_delta_p + launch_update)
int launch_delta_p_update(pbf_sim *s, const HaloPush *push, const HaloPull *pull) {
    const UpdateArgs up = {s->perm, s->pos, s->vel, s->svel, s->n_dev ? s->dn + DN_LOCAL : nullptr};
    if (s->params.vorticity_confinement)
        SWEEP_LAUNCH(KL_DP2, TL_SMEM1, nref_total(s), s->bufB, TILE_PASS, s->bufA, s->grid,
                     sim_params(s), push ? *push : NO_PUSH, up, pull ? *pull : NO_PULL);
    else
        SWEEP_LAUNCH(KL_DP1, TL_SMEM1, nref_total(s), s->bufB, TILE_PASS, s->bufA, s->grid,
                     sim_params(s), push ? *push : NO_PUSH, up, pull ? *pull : NO_PULL);
    return 1;
}

int launch_vorticity_a(pbf_sim *s, const HaloPush *push) {
    SWEEP_LAUNCH_T(KL_VA, 2 * TL, TL_SMEM2, nref_total(s), s->bufA, s->svel, TILE_PASS, s->bufB, s->vprime,
                 s->omega, s->grid, sim_params(s), push ? *push : NO_PUSH);
    return 1;
}

int launch_vorticity_b(pbf_sim *s, const HaloPull *pull) {
    SWEEP_LAUNCH(KL_VB, TL_SMEM1, nref_total(s), s->bufB, s->vprime, s->omega, s->perm, TILE_PASS,
                 s->vel, s->grid, sim_params(s), pull ? *pull : NO_PULL);
    return 1;
}

int launch_vorticity(pbf_sim *s) { return launch_vorticity_a(s, nullptr) + launch_vorticity_b(s); }

int launch_density_diag(pbf_sim *s) {
    SWEEP_LAUNCH(KL_DIAG, TL_SMEM1, nref_total(s), s->bufA, TILE_PASS, s->bufB, s->grid,
                 sim_params(s), s->diag, NO_PUSH, NO_PULL);
    return 1;
}
