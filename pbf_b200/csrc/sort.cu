// sort.cu -- onesweep LSD radix sort of (cell key, particle index) pairs.
//
// Replaces RadixSort::Run (reference src/RadixSort.cpp:124-200 and shaders/radixsort/*.glsl): the reference
// sorts the 16-byte records themselves with 2 bits per pass (10-13 passes x 4 dispatches, ~60 B/particle/pass).
// Here 8 bits per pass on 8-byte (key, index) pairs, one kernel per pass: warp-level digit ranking with
// match.any, per-tile digit counts chained between tiles by decoupled look-back, and a shared-memory staged
// scatter so that global stores are runs of consecutive addresses.  The result is the same permutation: a
// STABLE sort on the low plan.bits bits of the key (globalsort.glsl:62-64 proves the reference is stable; bits
// above plan.bits, e.g. the ceiling wrap of SURVEY.md a3, ride along unsorted exactly as in the reference).
//
// Algorithmic traffic: 4 B/particle for the digit histograms (read back from L2 right after k_predict wrote the keys)
// and 16 B per pass.
#include "pbf_internal.cuh"

namespace {

constexpr int SORT_BLOCK = 256;                   // = PBF_RADIX, one thread per digit in the look-back
#ifndef PBF_SORT_ITEMS
#define PBF_SORT_ITEMS 16
#endif
#ifndef PBF_SORT_CTAS
#define PBF_SORT_CTAS 4      // 64 registers, no spills: 4 blocks per SM instead of 3 (0.288 vs 0.300 ms for the four passes)
#endif
constexpr int SORT_ITEMS = PBF_SORT_ITEMS;
constexpr int SORT_TILE = SORT_BLOCK * SORT_ITEMS;  // 4096 pairs per tile
constexpr int SORT_WARPS = SORT_BLOCK / 32;

constexpr u32 ST_AGG = 1u << 30;                  // tile aggregate published
constexpr u32 ST_PREFIX = 2u << 30;               // inclusive prefix published
constexpr u32 ST_VALUE = (1u << 30) - 1;

// exclusive scan of the per-pass digit histograms -> global digit bases; clears the histograms and the tile
// counters for the next step.
__global__ void __launch_bounds__(PBF_RADIX) k_sort_scan(u32 *__restrict__ hist, u32 *__restrict__ gbase,
                                                          u32 *__restrict__ tile_counter) {
    __shared__ u32 wsum[PBF_RADIX / 32];
    const int pass = blockIdx.x, d = threadIdx.x, lane = d & 31, warp = d >> 5;
    u32 c = hist[pass * PBF_RADIX + d];
    hist[pass * PBF_RADIX + d] = 0;
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
    for (int w = 0; w < PBF_RADIX / 32; w++)
        if (w < warp) base += wsum[w];
    gbase[pass * PBF_RADIX + d] = base + inc - c;
    if (d == 0) tile_counter[pass] = 0;
}

// digit histograms of every pass.  Cell keys of neighbouring ids share most digits (plain shared atomics would
// serialise 32 lanes on one bin, match.any costs a pass over the distinct values): every thread takes 16 consecutive
// keys and run-length encodes each digit stream in registers, one shared atomic per run.
__global__ void __launch_bounds__(256) k_sort_hist(const u32 *__restrict__ keys, u32 n, SortPlan plan,
                                                    u32 *__restrict__ hist) {
    __shared__ u32 sh[4 * PBF_RADIX];
    for (int i = threadIdx.x; i < 4 * PBF_RADIX; i += blockDim.x) sh[i] = 0;
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
                        if (d != cur) { atomicAdd(&sh[p * PBF_RADIX + cur], run); cur = d; run = 0; }
                        run++;
                    }
                }
                atomicAdd(&sh[p * PBF_RADIX + cur], run);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * PBF_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

template <bool IOTA>
__global__ void __launch_bounds__(SORT_BLOCK, PBF_SORT_CTAS)
k_onesweep(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u32 *__restrict__ keys_out,
           u32 *__restrict__ vals_out, u32 n, int shift, u32 dmask, const u32 *__restrict__ gbase,
           u32 *status, u32 *tile_counter) {
    __shared__ u32 s_wh[SORT_WARPS][PBF_RADIX];   // per-warp digit counts, then exclusive offsets over warps
    __shared__ u32 s_dstart[PBF_RADIX];           // first tile-local slot of each digit
    __shared__ u32 s_goff[PBF_RADIX];             // global slot of tile-local slot 0 of each digit
    __shared__ u32 s_keys[SORT_TILE];
    __shared__ u32 s_vals[SORT_TILE];
    __shared__ u32 s_wsum[SORT_WARPS];
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tiles are handed out in launch order so that every tile a look-back waits on is already resident
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) s_wh[w][tid] = 0;
    __syncthreads();
    const u32 tile = s_tile;
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

    // rank inside the warp, in element order: lanes holding the same digit find each other with match.any
    const u32 lt = (1u << lane) - 1u;
    unsigned short rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 idx = wbase + i * 32;
        u32 d = idx < n ? ((key[i] >> shift) & dmask) : dmask;   // padding ranks after all real keys
        u32 m = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(m) - 1;
        u32 prev = 0;
        if (lane == leader) {
            prev = s_wh[warp][d];
            s_wh[warp][d] = prev + __popc(m);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = (unsigned short)(prev + __popc(m & lt));
        __syncwarp();
    }
    __syncthreads();

    // one thread per digit: offsets over warps, tile count, decoupled look-back over earlier tiles
    {
        const int d = tid;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            u32 t = s_wh[w][d];
            s_wh[w][d] = run;
            run += t;
        }
        const u32 count = run;
        volatile u32 *st = status;
        u32 excl = 0;
        if (tile == 0) {
            st[d] = ST_PREFIX | count;
        } else {
            st[(size_t)tile * PBF_RADIX + d] = ST_AGG | count;
            int t = (int)tile - 1;
            while (true) {
                u32 sv;
                do { sv = st[(size_t)t * PBF_RADIX + d]; } while ((sv >> 30) == 0);
                excl += sv & ST_VALUE;
                if ((sv >> 30) == 2) break;
                t--;
            }
            st[(size_t)tile * PBF_RADIX + d] = ST_PREFIX | (excl + count);
        }
        // exclusive scan of the digit counts inside the tile
        u32 inc = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        u32 wb = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++)
            if (w < warp) wb += s_wsum[w];
        const u32 dstart = wb + inc - count;
        s_dstart[d] = dstart;
        s_goff[d] = __ldg(gbase + d) + excl - dstart;
    }
    __syncthreads();

    // stage the tile in digit order
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        u32 idx = wbase + i * 32;
        u32 d = idx < n ? ((key[i] >> shift) & dmask) : dmask;
        u32 slot = s_dstart[d] + s_wh[warp][d] + rank[i];
        s_keys[slot] = key[i];
        s_vals[slot] = IOTA ? idx : (idx < n ? __ldg(vals_in + idx) : 0u);
    }
    __syncthreads();

    const u32 tile_n = min((u32)SORT_TILE, n - base);
#pragma unroll 4
    for (u32 slot = tid; slot < tile_n; slot += SORT_BLOCK) {
        u32 k = s_keys[slot];
        u32 dst = s_goff[(k >> shift) & dmask] + slot;
        keys_out[dst] = k;
        vals_out[dst] = s_vals[slot];
    }
}

int run_passes(pbf_sim *s, const SortPlan &plan, const u32 *kin, const u32 *vin, u32 *kout, u32 *vout, u32 n, const u32 *gbase) {
    const u32 tiles = (n + SORT_TILE - 1) / SORT_TILE;
    if (tiles == 0) return 0;
    cudaMemsetAsync(s->status, 0, (size_t)plan.passes * tiles * PBF_RADIX * sizeof(u32), s->stream);
    int launched = 0;
    for (int p = 0; p < plan.passes; p++) {
        const u32 *ki = p == 0 ? kin : s->ktmp[(p - 1) & 1];
        const u32 *vi = p == 0 ? vin : s->vtmp[(p - 1) & 1];
        u32 *ko = p == plan.passes - 1 ? kout : s->ktmp[p & 1];
        u32 *vo = p == plan.passes - 1 ? vout : s->vtmp[p & 1];
        u32 *st = s->status + (size_t)p * tiles * PBF_RADIX;
        if (p == 0 && vin == nullptr)
            k_onesweep<true><<<tiles, SORT_BLOCK, 0, s->stream>>>(ki, nullptr, ko, vo, n, plan.shift[p], plan.mask[p],
                                                                  gbase + p * PBF_RADIX, st, s->tile_counter + p);
        else
            k_onesweep<false><<<tiles, SORT_BLOCK, 0, s->stream>>>(ki, vi, ko, vo, n, plan.shift[p], plan.mask[p],
                                                                   gbase + p * PBF_RADIX, st, s->tile_counter + p);
        launched++;
    }
    return launched;
}

}  // namespace

SortPlan make_sort_plan(int bits) {
    SortPlan plan{};
    plan.bits = bits;
    plan.passes = (bits + 7) / 8;
    for (int p = 0; p < plan.passes; p++) {
        int nb = bits - 8 * p < 8 ? bits - 8 * p : 8;
        plan.shift[p] = 8 * p;
        plan.mask[p] = (1u << nb) - 1u;
    }
    return plan;
}

u32 sort_max_tiles(u32 cap) { return (cap + SORT_TILE - 1) / SORT_TILE; }

int launch_sort_scan(pbf_sim *s) {
    k_sort_scan<<<s->plan.passes, PBF_RADIX, 0, s->stream>>>(s->hist, s->gbase, s->tile_counter);
    return 1;
}

// the simulation's sort: keys by id (histograms already accumulated by k_predict), values = iota
int launch_sort_passes(pbf_sim *s) {
    return run_passes(s, s->plan, s->keys, nullptr, s->skey, s->perm, s->n, s->gbase);
}

// digit histograms of every pass of the handle's plan over keys[0..n) (slab mode: keys change after k_predict)
int launch_sort_hist(pbf_sim *s, const u32 *keys, u32 n) {
    int blocks = (int)(((n + 15) / 16 + 255) / 256);
    int maxb = s->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    // the histograms are accumulated with atomics: start from zero whatever ran before (pbf_predict twice, pbf_predict
    // followed by pbf_step, ... -- k_sort_scan also clears them, but only when a sort follows)
    cudaMemsetAsync(s->hist, 0, 4 * PBF_RADIX * sizeof(u32), s->stream);
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
    u32 *hist = s->hist + 4 * PBF_RADIX, *gbase = s->gbase + 4 * PBF_RADIX;
    cudaMemsetAsync(hist, 0, 4 * PBF_RADIX * sizeof(u32), s->stream);
    k_sort_hist<<<blocks, 256, 0, s->stream>>>(kin, n, plan, hist);
    k_sort_scan<<<plan.passes, PBF_RADIX, 0, s->stream>>>(hist, gbase, s->tile_counter);
    return 2 + run_passes(s, plan, kin, vin, kout, vout, n, gbase);
}
