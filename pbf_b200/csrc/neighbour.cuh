// neighbour.cuh -- device helpers shared by the neighbour sweeps (sweeps.cu) and the table/debug kernels
// (sim_kernels.cu): kernel constants, the merged-run lookup of neighbourcells.glsl:52-91, 256-bit candidate-pair loads
// and the packed f32x2 pair geometry.
#pragma once
#include "pbf_internal.cuh"

namespace {

constexpr int NB_BLOCK = 256;   // threads per block of the neighbour kernels (= BLOCKSIZE of src/SPH.cpp:60)

constexpr float H = 2.0f;                                // src/SPH.cpp:58
constexpr float H2 = 4.0f;
constexpr float POLY6 = 1.56668147106f / 512.0f;         // calclambda.glsl:46, /h^9
constexpr float SPIKY_GRAD = -3.0f * 4.774648292756860f / 64.0f;   // calclambda.glsl:63, /h^6
constexpr float FAR2 = 1.0e8f;      // r2 of a masked candidate: clamped to h^2 -> both kernels vanish, no inf/nan
constexpr float TINY = 1.0e-24f;    // r2 clamp: rsqrt stays finite and c*d = 0 for coincident particles (l == 0 branch)

// neighbourcells.glsl:62-84 for the window x-1..x+1 of one row: first existing start, summed sizes
__device__ __forceinline__ int2 merge3(const int2 *__restrict__ cells, int base, int x, int gx) {
    int cell = -1, entries = 0;
#pragma unroll
    for (int j = -1; j <= 1; j++) {
        const int xx = x + j;
        if (xx >= 0 && xx < gx) {
            int2 c = cells[base + xx];
            if (cell == -1) cell = c.x;
            if (c.x != -1) entries += c.y - c.x;
        }
    }
    return make_int2(cell, cell == -1 ? 0 : entries);
}

// ---- K7 neighbourcells.glsl:52-91: the nine merged runs {start, count} of a particle whose unclamped cell is `home` ---
__device__ __forceinline__ void fetch_runs(const u32 home, const GridInfo &g, const int2 *__restrict__ runs3,
                                           const int2 *__restrict__ cells, int2 r[9]) {
    const int cx = (int)(home & ((1u << g.bx) - 1u)) - 2;
    const int cz = (int)((home >> g.bx) & ((1u << g.bz) - 1u)) - 2;
    const int cy = (int)(home >> (g.bx + g.bz)) - 2;
    const bool fast = cx >= 0 && cx < g.gx;
#pragma unroll
    for (int o = 0; o < 9; o++) {
        const int yy = cy + (o / 3 - 1), zz = cz + (o % 3 - 1);   // gridoffsets[o] = (0, dy, dz)
        r[o] = make_int2(-1, 0);
        if (yy >= 0 && yy < g.gy && zz >= 0 && zz < g.gz) {
            const int base = yy * g.gxgz + zz * g.gx;
            if (fast) r[o] = __ldg(runs3 + base + cx);
            else r[o] = merge3(cells, base, cx, g.gx);             // particle outside the grid in x: rare
        }
    }
}

// Opt-in "full support" search (pbf_options::full_support): the kernels have support h = 2 cells but
// the reference only visits the 27 cells around a particle (neighbourcells.glsl:37-47), truncating every sum.  RAD = 2
// visits 5 x 5 x 5 cells: 25 rows (dy, dz in -2..2, dy outermost as in gridoffsets[]), each the cells x-2..x+2 merged
// like neighbourcells.glsl:62-84 merges three.
template <int RAD>
__device__ __forceinline__ int2 merge_row(const int2 *__restrict__ cells, int base, int x, int gx) {
    int cell = -1, entries = 0;
#pragma unroll
    for (int j = -RAD; j <= RAD; j++) {
        const int xx = x + j;
        if (xx >= 0 && xx < gx) {
            const int2 c = cells[base + xx];
            if (cell == -1) cell = c.x;
            if (c.x != -1) entries += c.y - c.x;
        }
    }
    return make_int2(cell, cell == -1 ? 0 : entries);
}

// The thread's non-empty runs {start,end} into shared memory.  Returns their number; *slots = number of aligned
// candidate pairs over all runs; *self_in = whether the particle's own slot lies in one of its runs, i.e. whether
// FOR_EACH_NEIGHBOUR would have skipped `self`.
// rel (canonical order, pbf_set_canonical_order): pairs are counted from the run's first candidate, not from even indices
template <int BLOCK, int RAD = 1>
__device__ __forceinline__ int load_runs(const u32 home, const u32 i, const GridInfo &g, const int2 *__restrict__ runs3,
                                         const int2 *__restrict__ cells, int2 *srun, int tid, int *slots, bool *self_in,
                                         bool rel = false) {
    int cnt = 0, tot = 0;
    bool self = false;
    if (RAD == 1) {
        int2 r[9];
        fetch_runs(home, g, runs3, cells, r);
#pragma unroll
        for (int o = 0; o < 9; o++) {
            self = self || ((int)i >= r[o].x && (int)i < r[o].x + r[o].y);   // see k_plan: any run may hold the particle itself
            if (r[o].y > 0) {
                const int s = r[o].x, e = r[o].x + r[o].y;
                srun[cnt * BLOCK + tid] = make_int2(s, e);
                cnt++;
                tot += rel ? (e - s + 1) >> 1 : ((e + 1) >> 1) - (s >> 1);
            }
        }
    } else {
        const int cx = (int)(home & ((1u << g.bx) - 1u)) - 2;
        const int cz = (int)((home >> g.bx) & ((1u << g.bz) - 1u)) - 2;
        const int cy = (int)(home >> (g.bx + g.bz)) - 2;
#pragma unroll 1
        for (int o = 0; o < (2 * RAD + 1) * (2 * RAD + 1); o++) {
            const int yy = cy + (o / (2 * RAD + 1) - RAD), zz = cz + (o % (2 * RAD + 1) - RAD);
            if (yy < 0 || yy >= g.gy || zz < 0 || zz >= g.gz) continue;
            const int2 r = merge_row<RAD>(cells, yy * g.gxgz + zz * g.gx, cx, g.gx);
            self = self || ((int)i >= r.x && (int)i < r.x + r.y);
            if (r.y > 0) {
                srun[cnt * BLOCK + tid] = make_int2(r.x, r.x + r.y);
                cnt++;
                tot += rel ? (r.y + 1) >> 1 : ((r.x + r.y + 1) >> 1) - (r.x >> 1);
            }
        }
    }
    *self_in = self;
    *slots = tot;
    return cnt;
}

struct Pair {   // candidates 2m (.x of each float2) and 2m+1 (.y)
    float2 x, y, z, w;
};

__device__ __forceinline__ Pair ldg_pair(const float4 *p) {   // one 256-bit read-only load, p is 32-byte aligned
    Pair r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.x.x), "=f"(r.y.x), "=f"(r.z.x), "=f"(r.w.x), "=f"(r.x.y), "=f"(r.y.y), "=f"(r.z.y), "=f"(r.w.y)
                 : "l"(p));
    return r;
}

// the same through L2 only: for arrays that other blocks of the SAME kernel write (ghost slots, fused halo pull)
__device__ __forceinline__ Pair ldcg_pair(const float4 *p) {
    Pair r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.x.x), "=f"(r.y.x), "=f"(r.z.x), "=f"(r.w.x), "=f"(r.x.y), "=f"(r.y.y), "=f"(r.z.y), "=f"(r.w.y)
                 : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// geometry of one candidate pair against particle p: d = p - c, r2, 1/l, and the two NEGATED factors
// tn = rc - h^2 = -(h^2 - r^2)+ and t2n = l - h = -(h - l)+ with rc = min(r2 + TINY, h^2), l = rc / sqrt(rc): beyond the
// support radius both factors vanish without a separate clamp to zero, at r = 0 rsqrt stays finite and c*d = 0 like
// the l == 0 branch of gradWspiky (callers multiply gradient magnitudes by the EXACT r2 or d, which are 0 there).  Odd powers carry a minus sign the
// callers fold into their final constants.  Members outside the run (v0/v1 false) get rc = h^2, i.e. they leave kernel
// support; their d is finite (a real record or a zeroed pad record), so 0 * d = 0.
struct PairGeom {
    float2 dx, dy, dz, r2, il, t, t2;
};

__device__ __forceinline__ PairGeom pair_geom(const float4 &p, const Pair &c, bool v0, bool v1) {
    PairGeom q;
    q.dx = make_float2(p.x - c.x.x, p.x - c.x.y);
    q.dy = make_float2(p.y - c.y.x, p.y - c.y.y);
    q.dz = make_float2(p.z - c.z.x, p.z - c.z.y);
    q.r2 = __ffma2_rn(q.dz, q.dz, __ffma2_rn(q.dy, q.dy, __fmul2_rn(q.dx, q.dx)));      // exact 0 for coincident particles
#ifndef PBF_SWEEP_CLAMP_TINY
    // Measured at 8M particles (lambda / delta-p sweep, ms): this form 0.346 / 0.356; TINY folded into the first
    // multiply-add of r2 and S from (h-l)^4 0.327 / 0.345 -- two instructions fewer, but a coincident particle then adds 16
    // to S (gradWspiky's l == 0 branch says 0; particles clamped into the same corner do coincide), so not adopted;
    // clamp(r2, TINY, h^2) with min/max on the ALU pipe instead of the add 0.350 / 0.360 -- one instruction MORE and slower:
    // the sweeps are bound by issue slots, not by the FP32 pipe alone.
    const float2 re = __fadd2_rn(q.r2, make_float2(TINY, TINY));
    const float2 rc = make_float2(fminf(v0 ? re.x : FAR2, H2), fminf(v1 ? re.y : FAR2, H2));
#else
    const float cx_ = fminf(fmaxf(q.r2.x, TINY), H2), cy_ = fminf(fmaxf(q.r2.y, TINY), H2);
    const float2 rc = make_float2(v0 ? cx_ : H2, v1 ? cy_ : H2);
#endif
    q.il = make_float2(rsqrt_ftz(rc.x), rsqrt_ftz(rc.y));
    q.t2 = __ffma2_rn(rc, q.il, make_float2(-H, -H));                // l - h: 0 at and beyond the support radius
    q.t = __fadd2_rn(rc, make_float2(-H2, -H2));                     // r^2 - h^2, exactly <= 0
    return q;
}

// flattened walk over the candidate pairs of all runs; body(index of the pair's first candidate, valid0, valid1).
// Default: ALIGNED pairs (candidates 2m, 2m+1: one 256-bit load each), members outside the run masked.  rel: pairs counted
// from the run's first candidate, like the tiled path of sweeps.cu walks its shared-memory image -- the order of a
// particle's floating-point sums is then the same on both paths (canonical order).
template <int BLOCK, class F>
__device__ __forceinline__ void for_each_pair(const int2 *srun, int tid, int slots, bool rel, F body) {
    const int2 *sp = srun + tid;
    int c = 0, cend = 0, s = 0, e = 0;
#pragma unroll 1
    for (int k = 0; k < slots; k++) {
        if (c >= cend) {
            const int2 r = *sp;
            sp += BLOCK;
            s = r.x; e = r.y;
            c = rel ? s : s & ~1;
            cend = e;
        }
        body(c, c >= s, c + 1 < e);
        c += 2;
    }
}

}  // namespace
