// sort.cu -- onesweep LSD radix sort of (cell key, particle index) pairs.
//
// Replaces RadixSort::Run (reference src/RadixSort.cpp:124-200 and shaders/radixsort/*.glsl): the reference
// sorts the 16-byte records themselves with 2 bits per pass (10-13 passes x 4 dispatches, ~60 B/particle/pass).
// Here up to 9 bits per pass on 8-byte (key, index) pairs, one kernel per pass: digit ranking inside a warp, per-tile
// digit counts chained between tiles by decoupled look-back, and a shared-memory staged scatter so that global stores are
// runs of consecutive addresses.  The result is the same permutation: a STABLE sort on the low plan.bits bits of the key
// (globalsort.glsl:62-64 proves the reference is stable; bits above plan.bits, e.g. the ceiling wrap of SURVEY.md a3,
// ride along unsorted exactly as in the reference).
//
// Digit widths: the sorted bits are split evenly over ceil(bits / 9) passes -- 26 bits (512 x 256 x 512 cells, the
// headline grid) = 9 + 9 + 8, i.e. THREE passes whose digits are exactly the cell's x, z and y; 20 and 23 bits take three
// passes of 7 and 8 bits; only grids beyond 2^27 cells need a fourth.
//
// Ranking inside a warp.  Lanes hold consecutive elements, and cell keys of consecutive elements are either equal in a
// digit (the x and z passes of a lattice: whole warps share one digit) or increasing (the y pass: 32 lanes, ~30 distinct
// digits -- the case in which match.any, which loops over the distinct values, cost 2.4x a normal pass).  Equal digits
// therefore sit in RUNS of neighbouring lanes: one ballot of the run heads gives every lane its run and its place in it,
// the head lane adds the run length to the warp's digit counter.  That is only valid if no two runs of the warp share a
// digit, which is certain when the digits do not decrease from lane to lane (one vote); a warp whose digits do ranks that
// one item with match.any instead (random keys, e.g. pbf_sort_pairs on arbitrary input, almost always do).  A warp that
// holds a single digit keeps that digit's count in a register from item to item.
//
// Algorithmic traffic: 4 B/particle for the digit histograms (read back from L2 right after k_predict wrote the keys)
// and 16 B per pass.
#include "pbf_internal.cuh"

namespace {

constexpr int SORT_BLOCK = 256;
#ifndef PBF_SORT_ITEMS
#define PBF_SORT_ITEMS 16
#endif
#ifndef PBF_SORT_CTAS
#define PBF_SORT_CTAS 3      // measured at 8M particles, three 9-bit passes: 3 blocks per SM 0.241 ms, 4 blocks (64 registers) 0.261 ms,
                             // 8 items per thread 0.290 ms; four 8-bit passes 0.288 ms
#endif
constexpr int SORT_ITEMS = PBF_SORT_ITEMS;
constexpr int SORT_TILE = SORT_BLOCK * SORT_ITEMS;  // 4096 pairs per tile
constexpr int SORT_WARPS = SORT_BLOCK / 32;
constexpr int R = PBF_RADIX;                       // 512: digit slots per pass (a pass may use fewer)
static_assert(R == 2 * SORT_BLOCK, "the look-back handles two digits per thread");

constexpr u32 ST_AGG = 1u << 30;                  // tile aggregate published
constexpr u32 ST_PREFIX = 2u << 30;               // inclusive prefix published
constexpr u32 ST_VALUE = (1u << 30) - 1;

// shared memory of k_onesweep (dynamic: 52 KB, above the 48 KB a static allocation may have; four blocks per SM fit)
struct SortSmem {
    u32 wh[SORT_WARPS][R];        // per-warp digit counts, then exclusive offsets over warps
    u32 dstart[R];                // first tile-local slot of each digit
    u32 goff[R];                  // global slot of tile-local slot 0 of each digit
    u32 keys[SORT_TILE];
    u32 vals[SORT_TILE];
    u32 wsum[2][SORT_WARPS];
    u32 tile;
};
static_assert(PBF_SORT_CTAS * (sizeof(SortSmem) + 1024) <= 233472, "blocks per SM");

// exclusive scan of the per-pass digit histograms -> global digit bases; clears the histograms and the tile
// counters for the next step.
__global__ void __launch_bounds__(R) k_sort_scan(u32 *__restrict__ hist, u32 *__restrict__ gbase,
                                                  u32 *__restrict__ tile_counter) {
    __shared__ u32 wsum[R / 32];
    const int pass = blockIdx.x, d = threadIdx.x, lane = d & 31, warp = d >> 5;
    u32 c = hist[pass * R + d];
    hist[pass * R + d] = 0;
    u32 inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    u32 base = 0;
#pragma unroll
    for (int w = 0; w < R / 32; w++)
        if (w < warp) base += wsum[w];
    gbase[pass * R + d] = base + inc - c;
    if (d == 0) tile_counter[pass] = 0;
}

// digit histograms of every pass.  Cell keys of neighbouring ids share most digits (plain shared atomics would
// serialise 32 lanes on one bin, match.any costs a pass over the distinct values): every thread takes 16 consecutive
// keys and run-length encodes each digit stream in registers, one shared atomic per run.
__global__ void __launch_bounds__(256) k_sort_hist(const u32 *__restrict__ keys, NRef nr, SortPlan plan,
                                                    u32 *__restrict__ hist) {
    __shared__ u32 sh[4 * R];
    const u32 n = nref(nr);
    for (int i = threadIdx.x; i < 4 * R; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u32 chunks = (n + 15u) >> 4;
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x) {
        u32 k[16];
        const u32 base = c << 4;
        if (base + 16u <= n && (reinterpret_cast<uintptr_t>(keys) & 15u) == 0) {
            const uint4 *q = reinterpret_cast<const uint4 *>(keys + base);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 v = __ldg(q + j);
                k[4 * j] = v.x; k[4 * j + 1] = v.y; k[4 * j + 2] = v.z; k[4 * j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) k[j] = base + j < n ? __ldg(keys + base + j) : 0u;
        }
        const int cnt = (int)min(16u, n - base);
#pragma unroll
        for (int p = 0; p < 4; p++) {
            if (p < plan.passes) {
                const int sft = plan.shift[p];
                const u32 msk = plan.mask[p];
                u32 cur = (k[0] >> sft) & msk, run = 1;
#pragma unroll
                for (int j = 1; j < 16; j++) {
                    const u32 d = (k[j] >> sft) & msk;
                    if (j < cnt) {
                        if (d != cur) { atomicAdd(&sh[p * R + cur], run); cur = d; run = 0; }
                        run++;
                    }
                }
                atomicAdd(&sh[p * R + cur], run);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * R; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

template <bool IOTA>
__global__ void __launch_bounds__(SORT_BLOCK, PBF_SORT_CTAS)
k_onesweep(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u32 *__restrict__ keys_out,
           u32 *__restrict__ vals_out, NRef nr, int shift, u32 dmask, const u32 *__restrict__ gbase,
           u32 *status, u32 *tile_counter) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem &S = *reinterpret_cast<SortSmem *>(sort_smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 n = nref(nr);
    // Tiles are handed out in launch order so that every tile a look-back waits on is already resident (or done).  A block
    // takes tiles until none is left: normally exactly one (the grid is sized for n), more when the grid was sized for a
    // stale bound of a count that lives on the device (NRef).
    for (;;) {
    __syncthreads();                                   // the previous tile's staging area is free
    if (tid == 0) S.tile = atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) { S.wh[w][tid] = 0; S.wh[w][tid + SORT_BLOCK] = 0; }
    __syncthreads();
    const u32 tile = S.tile;
    const u32 base = tile * SORT_TILE;
    if (base >= n) return;

    // warp-striped load: warp w owns SORT_ITEMS*32 consecutive pairs, item i of lane l is element i*32+l
    const u32 wbase = base + warp * (SORT_ITEMS * 32) + lane;
    u32 key[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 idx = wbase + i * 32;
        key[i] = idx < n ? __ldg(keys_in + idx) : 0xffffffffu;
    }

    // rank inside the warp, in element order (see the header: runs of equal digits in neighbouring lanes)
    const u32 le = 0xffffffffu >> (31 - lane);          // lanes 0..lane
    u32 rank2[SORT_ITEMS / 2];                        // two 16-bit ranks per register
    u32 *wh = S.wh[warp];
    // A whole warp often holds ONE digit for several items in a row (the x and z passes of a lattice): that digit's count
    // is then kept in a register (cd = the digit, cc = its count so far; both warp-uniform) and shared memory is only
    // touched when the digit changes.
    u32 cd = 0xffffffffu, cc = 0;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const u32 idx = wbase + i * 32;
        const u32 d = idx < n ? ((key[i] >> shift) & dmask) : dmask;   // padding ranks after all real keys
        const u32 dprev = __shfl_up_sync(0xffffffffu, d, 1);
        const bool head = lane == 0 || d != dprev;
        const u32 heads = __ballot_sync(0xffffffffu, head);
        u32 rk;
        if (heads == 1u) {                                               // one digit in the whole warp
            if (d != cd) {
                if (cd != 0xffffffffu && lane == 0) wh[cd] = cc;
                __syncwarp();
                cd = d;
                cc = wh[d];
                __syncwarp();          // every lane has read the count before lane 0 may overwrite it at the next digit change
            }
            rk = cc + (u32)lane;
            cc += 32u;
        } else {
            if (cd != 0xffffffffu) {
                if (lane == 0) wh[cd] = cc;
                cd = 0xffffffffu;
                __syncwarp();
            }
            u32 prev = 0;
            // digits that do not decrease from lane to lane: equal digits are neighbours, so the runs are the whole story
            if (__all_sync(0xffffffffu, lane == 0 || d >= dprev)) {
                const int start = 31 - __clz(heads & le);               // my run's first lane
                const u32 above = heads & ~le;
                const int end = above ? __ffs(above) - 1 : 32;          // one past my run's last lane
                if (head) {
                    prev = wh[d];
                    wh[d] = prev + (u32)(end - start);
                }
                prev = __shfl_sync(0xffffffffu, prev, start);
                rk = prev + (u32)(lane - start);
            } else {
                const u32 m = __match_any_sync(0xffffffffu, d);
                const int leader = __ffs(m) - 1;
                if (lane == leader) {
                    prev = wh[d];
                    wh[d] = prev + __popc(m);
                }
                prev = __shfl_sync(0xffffffffu, prev, leader);
                rk = prev + __popc(m & (le >> 1));
            }
            __syncwarp();
        }
        if (i & 1) rank2[i >> 1] |= rk << 16;
        else rank2[i >> 1] = rk;
    }
    if (cd != 0xffffffffu && lane == 0) wh[cd] = cc;
    __syncthreads();

    // one thread per digit pair (d0 = tid, d1 = tid + 256): offsets over warps, tile count, decoupled look-back
    {
        const int d0 = tid, d1 = tid + SORT_BLOCK;
        u32 run0 = 0, run1 = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const u32 t0 = S.wh[w][d0], t1 = S.wh[w][d1];
            S.wh[w][d0] = run0; S.wh[w][d1] = run1;
            run0 += t0; run1 += t1;
        }
        const u32 count0 = run0, count1 = run1;
        const bool two = dmask >= (u32)SORT_BLOCK;       // digits above 255 exist in this pass
        volatile u32 *st = status;
        u32 excl0 = 0, excl1 = 0;
        if (tile == 0) {
            st[d0] = ST_PREFIX | count0;
            if (two) st[d1] = ST_PREFIX | count1;
        } else {
            st[(size_t)tile * R + d0] = ST_AGG | count0;
            if (two) st[(size_t)tile * R + d1] = ST_AGG | count1;
            int t0 = (int)tile - 1, t1 = two ? t0 : -1;
            while (t0 >= 0 || t1 >= 0) {                 // both chains advance in the same loop: their loads overlap
                u32 s0 = 0, s1 = 0;
                if (t0 >= 0) s0 = st[(size_t)t0 * R + d0];
                if (t1 >= 0) s1 = st[(size_t)t1 * R + d1];
                if (t0 >= 0 && (s0 >> 30) != 0) {
                    excl0 += s0 & ST_VALUE;
                    t0 = (s0 >> 30) == 2 ? -1 : t0 - 1;
                }
                if (t1 >= 0 && (s1 >> 30) != 0) {
                    excl1 += s1 & ST_VALUE;
                    t1 = (s1 >> 30) == 2 ? -1 : t1 - 1;
                }
            }
            st[(size_t)tile * R + d0] = ST_PREFIX | (excl0 + count0);
            if (two) st[(size_t)tile * R + d1] = ST_PREFIX | (excl1 + count1);
        }
        // exclusive scan of the 512 digit counts inside the tile: digits 0..255 over the threads, then 256..511
        u32 inc0 = count0, inc1 = count1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, inc0, o), b = __shfl_up_sync(0xffffffffu, inc1, o);
            if (lane >= o) { inc0 += a; inc1 += b; }
        }
        if (lane == 31) { S.wsum[0][warp] = inc0; S.wsum[1][warp] = inc1; }
        __syncthreads();
        u32 wb0 = 0, wb1 = 0, total0 = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            if (w < warp) { wb0 += S.wsum[0][w]; wb1 += S.wsum[1][w]; }
            total0 += S.wsum[0][w];
        }
        const u32 ds0 = wb0 + inc0 - count0, ds1 = total0 + wb1 + inc1 - count1;
        S.dstart[d0] = ds0;
        S.dstart[d1] = ds1;
        S.goff[d0] = __ldg(gbase + d0) + excl0 - ds0;
        S.goff[d1] = two ? __ldg(gbase + d1) + excl1 - ds1 : 0u;
    }
    __syncthreads();

    // stage the tile in digit order
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 idx = wbase + i * 32;
        u32 d = idx < n ? ((key[i] >> shift) & dmask) : dmask;
        u32 slot = S.dstart[d] + S.wh[warp][d] + ((rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
        S.keys[slot] = key[i];
        S.vals[slot] = IOTA ? idx : (idx < n ? __ldg(vals_in + idx) : 0u);
    }
    __syncthreads();

    const u32 tile_n = min((u32)SORT_TILE, n - base);
#pragma unroll 4
    for (u32 slot = tid; slot < tile_n; slot += SORT_BLOCK) {
        u32 k = S.keys[slot];
        u32 dst = S.goff[(k >> shift) & dmask] + slot;
        keys_out[dst] = k;
        vals_out[dst] = S.vals[slot];
    }
    if (nr.p == nullptr) return;                       // count known on the host: the grid has one block per tile
    }
}

int run_passes(pbf_sim *s, const SortPlan &plan, const u32 *kin, const u32 *vin, u32 *kout, u32 *vout, NRef n, const u32 *gbase) {
    // status words for every tile the count could ever need (capacity), blocks for the bound
    const u32 tiles = (n.n + SORT_TILE - 1) / SORT_TILE, st_tiles = n.p ? s->max_tiles : tiles;
    if (tiles == 0) return 0;
    cudaMemsetAsync(s->status, 0, (size_t)plan.passes * st_tiles * R * sizeof(u32), s->stream);
    int launched = 0;
    for (int p = 0; p < plan.passes; p++) {
        const u32 *ki = p == 0 ? kin : s->ktmp[(p - 1) & 1];
        const u32 *vi = p == 0 ? vin : s->vtmp[(p - 1) & 1];
        u32 *ko = p == plan.passes - 1 ? kout : s->ktmp[p & 1];
        u32 *vo = p == plan.passes - 1 ? vout : s->vtmp[p & 1];
        u32 *st = s->status + (size_t)p * st_tiles * R;
        if (p == 0 && vin == nullptr)
            k_onesweep<true><<<tiles, SORT_BLOCK, sizeof(SortSmem), s->stream>>>(ki, nullptr, ko, vo, n, plan.shift[p], plan.mask[p],
                                                                                 gbase + p * R, st, s->tile_counter + p);
        else
            k_onesweep<false><<<tiles, SORT_BLOCK, sizeof(SortSmem), s->stream>>>(ki, vi, ko, vo, n, plan.shift[p], plan.mask[p],
                                                                                  gbase + p * R, st, s->tile_counter + p);
        launched++;
    }
    return launched;
}

}  // namespace

int sort_init(void) {
    cudaError_t e = cudaFuncSetAttribute(k_onesweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_onesweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
    return e == cudaSuccess ? 0 : -1;
}

SortPlan make_sort_plan(int bits) {
    SortPlan plan{};
    plan.bits = bits;
#ifndef PBF_SORT_MAX_BITS
#define PBF_SORT_MAX_BITS 9
#endif
    plan.passes = (bits + PBF_SORT_MAX_BITS - 1) / PBF_SORT_MAX_BITS;   // up to 9 bits per pass ...
    const int width = (bits + plan.passes - 1) / plan.passes;   // ... spread evenly: 26 -> 9 + 9 + 8
    int at = 0;
    for (int p = 0; p < plan.passes; p++) {
        const int nb = bits - at < width ? bits - at : width;
        plan.shift[p] = at;
        plan.mask[p] = (1u << nb) - 1u;
        at += nb;
    }
    return plan;
}

u32 sort_max_tiles(u32 cap) { return (cap + SORT_TILE - 1) / SORT_TILE; }

int launch_sort_scan(pbf_sim *s) {
    k_sort_scan<<<s->plan.passes, R, 0, s->stream>>>(s->hist, s->gbase, s->tile_counter);
    return 1;
}

// the simulation's sort: keys by id (histograms already accumulated by k_predict), values = iota
int launch_sort_passes(pbf_sim *s) {
    return run_passes(s, s->plan, s->keys, nullptr, s->skey, s->perm, nref_total(s), s->gbase);
}

// digit histograms of every pass of the handle's plan over keys[0..n) (slab mode: keys change after k_predict)
int launch_sort_hist(pbf_sim *s, const u32 *keys, NRef n) {
    int blocks = (int)(((n.n + 15) / 16 + 255) / 256);
    int maxb = s->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    // the histograms are accumulated with atomics: start from zero whatever ran before (pbf_predict twice, pbf_predict
    // followed by pbf_step, ... -- k_sort_scan also clears them, but only when a sort follows)
    cudaMemsetAsync(s->hist, 0, 4 * R * sizeof(u32), s->stream);
    k_sort_hist<<<blocks, 256, 0, s->stream>>>(keys, n, s->plan, s->hist);
    return 1;
}

int launch_sort_pairs(pbf_sim *s, const u32 *kin, const u32 *vin, u32 *kout, u32 *vout, u32 n, int bits) {
    SortPlan plan = make_sort_plan(bits);
    int blocks = (int)(((n + 15) / 16 + 255) / 256);
    int maxb = s->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    // own histogram / digit-base scratch (second half of the allocations): a standalone sort between pbf_predict and
    // pbf_sort must not disturb the histograms the simulation's sort is about to scan
    u32 *hist = s->hist + 4 * R, *gbase = s->gbase + 4 * R;
    cudaMemsetAsync(hist, 0, 4 * R * sizeof(u32), s->stream);
    k_sort_hist<<<blocks, 256, 0, s->stream>>>(kin, NRef{n, nullptr}, plan, hist);
    k_sort_scan<<<plan.passes, R, 0, s->stream>>>(hist, gbase, s->tile_counter);
    return 2 + run_passes(s, plan, kin, vin, kout, vout, NRef{n, nullptr}, gbase);
}
