// sim_kernels.cu -- the per-timestep PBF kernels (everything of SPH::Run except the sort).
//
// Data layout in HBM (all SoA, 16-byte records):
//   pos/vel/hl   by persistent particle id (the renderer-visible buffers, src/SPH.cpp:106-133)
//   pred, keys   by id: predicted record {p*, id} (predictpos.glsl:37) and its cell key
//   skey, perm   by sorted slot: cell key and particle id (output of the onesweep sort)
//   home         by sorted slot: packed unclamped cell of the predicted position (what K7 evaluates once per step)
//   cells        dense table, int2 {start,end} per cell, index x + z*gx + y*gx*gz (findcells.glsl:45-52)
//   runs3        dense table, int2 {start,count} per cell: cells x-1,x,x+1 merged as neighbourcells.glsl:62-84 does
//   bufA         by sorted slot {x,y,z,-}: positions the solver iterates on
//   bufB         by sorted slot {x,y,z,lambda}: written by the lambda kernel so that the delta-p kernel needs ONE
//                load per neighbour (the reference does a record load + a lambda texel fetch)
//   svel, vprime, omega  by sorted slot, vorticity/XSPH only
//
// Neighbour iteration visits exactly the candidate set of FOR_EACH_NEIGHBOUR (shaders/sph/foreachneighbour.glsl:1-10):
// 9 runs in the order of neighbourcells.glsl:37-47, each the merged cells x-1..x+1.  The reference stores these runs
// per particle (48 B, K7, re-read by every sweep); here a thread fetches its 9 runs from runs3 (L1/L2 resident,
// shared by neighbouring particles), keeps the non-empty ones in shared memory and walks them as ONE flattened loop
// over 32-byte aligned candidate PAIRS: one 256-bit load (LDG.E.256) brings candidates 2m and 2m+1, the pair
// is evaluated with packed f32x2 arithmetic (FFMA2/FMUL2), members outside [start,end) are masked by moving them
// out of kernel support.  The kernels are FP32-issue bound, not HBM bound (profiles/): these measures cut the
// instructions per candidate from 45 to about 20.
#include "pbf_internal.cuh"

namespace {

constexpr int NB_BLOCK = 256;   // threads per block of the neighbour kernels (= BLOCKSIZE of src/SPH.cpp:60)

constexpr float H = 2.0f;                                // src/SPH.cpp:58
constexpr float H2 = 4.0f;
constexpr float POLY6 = 1.56668147106f / 512.0f;         // calclambda.glsl:46, /h^9
constexpr float SPIKY_GRAD = -3.0f * 4.774648292756860f / 64.0f;   // calclambda.glsl:63, /h^6
constexpr float FAR = 1.0e8f;       // a masked candidate is moved here: r2 = 1e16 -> both kernels vanish, no inf/nan
constexpr float TINY = 1.0e-24f;    // r2 clamp: rsqrt stays finite and c*d = 0 for coincident particles (l == 0 branch)

__device__ __forceinline__ u32 cell_key(float x, float y, float z, const GridInfo &g) {
    // ivec3(clamp(pos, 0, GRID_SIZE)) . (1, gx*gz, gx)   (counting.glsl:53-57); clamp is inclusive
    int cx = (int)fminf(fmaxf(x, 0.0f), (float)g.gx);
    int cy = (int)fminf(fmaxf(y, 0.0f), (float)g.gy);
    const int czg = (int)fminf(fmaxf(z, 0.0f), (float)g.gz_global);   // global cell layer
    const int cz = min(max(czg - g.zoff, 0), g.gz);                  // layer inside this handle's window
    u32 k = (u32)cx + (u32)cz * (u32)g.gx + (u32)cy * (u32)g.gxgz;
    if (cx >= g.gx || cy >= g.gy || cz >= g.gz || czg >= g.gz_global || czg < g.zoff) k |= PBF_KEY_NOCELL;
    return k;
}

// ---- K1 predictpos.glsl:18-38 + cell key + digit histograms of every sort pass + clearhighlight.glsl ----------
__global__ void __launch_bounds__(256)
k_predict(u32 first, u32 n, const float4 *__restrict__ pos, const float4 *__restrict__ vel, u32 *__restrict__ hl,
          float4 *__restrict__ pred, u32 *__restrict__ keys, u32 *__restrict__ hist, u32 *__restrict__ flags,
          GridInfo g, SimParams P, SortPlan plan) {
    __shared__ u32 sh[4 * PBF_RADIX];
    for (int i = threadIdx.x; i < 4 * PBF_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 stride = gridDim.x * blockDim.x;
    const u32 nround = (n + 31u) & ~31u;
    bool any_hl = false;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += stride) {
        const bool valid = j < n;
        const u32 i = first + j;
        u32 key = 0;
        if (valid) {
            float4 p = pos[i];
            float4 v = vel[i];
            // exact, uncontracted binary32 in the order of the shader so that keys are bit-identical to the oracle
            if (P.extforce && p.z > (float)g.gz / 2.0f)
                v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, P.gravity), -1.0f), P.timestep));
            v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(P.gravity, -1.0f), P.timestep));
            p.x = __fadd_rn(p.x, __fmul_rn(P.timestep, v.x));
            p.y = __fadd_rn(p.y, __fmul_rn(P.timestep, v.y));
            p.z = __fadd_rn(p.z, __fmul_rn(P.timestep, v.z));
            p.w = __int_as_float((int)i);
            pred[i] = p;
            key = cell_key(p.x, p.y, p.z, g);
            keys[i] = key;
            // clearhighlight.glsl: flag &= 1 (written back only when it changes anything)
            u32 h = hl[i];
            if (h & ~1u) hl[i] = h & 1u;
            any_hl |= (h & 1u) != 0;
        }
        for (int p = 0; p < plan.passes; p++) {    // plan.passes = 0: the caller histograms later (slab mode)
            u32 d = valid ? ((key >> plan.shift[p]) & plan.mask[p]) : 0xffffffffu;
            u32 m = __match_any_sync(0xffffffffu, d);
            if (valid && lane == (u32)(__ffs(m) - 1)) atomicAdd(&sh[p * PBF_RADIX + d], (u32)__popc(m));
        }
    }
    if (__any_sync(0xffffffffu, any_hl) && lane == 0) flags[0] = 1u;
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * PBF_RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// appended halo records already hold p*: cell keys only
__global__ void __launch_bounds__(256)
k_keys_only(u32 first, u32 count, const float4 *__restrict__ pred, u32 *__restrict__ keys, GridInfo g) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    float4 p = pred[first + j];
    keys[first + j] = cell_key(p.x, p.y, p.z, g);
}

// ---- cell tables ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_tables(size_t ncell, int2 *__restrict__ cells, int2 *__restrict__ runs3) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    cells[c] = make_int2(-1, 0);    // start = -1 (gridtexture clear, src/NeighbourCellFinder.cpp:116-126)
    runs3[c] = make_int2(-1, 0);
}

// Reset what the previous step wrote (replaces the per-step clear of the whole start image: 4 B per previously
// sorted particle instead of 4 B per cell).
__global__ void __launch_bounds__(256)
k_unclear_cells(u32 n, const u32 *__restrict__ skey, int2 *__restrict__ cells, int2 *__restrict__ runs3, GridInfo g) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 k = skey[i];
    if (k & PBF_KEY_NOCELL) return;
    if (i == 0 || skey[i - 1] != k) {
        cells[k].x = -1;
        const int x = (int)(k % (u32)g.gx);
        const int2 empty = make_int2(-1, 0);
        if (x > 0) runs3[k - 1] = empty;
        runs3[k] = empty;
        if (x + 1 < g.gx) runs3[k + 1] = empty;
    }
    if (i == 0 && g.ref_quirks) { runs3[0] = make_int2(-1, 0); if (g.gx > 1) runs3[1] = make_int2(-1, 0); }
}

// packed unclamped cell of a position (neighbourcells.glsl:57 `ivec3(pos)`), each coordinate saturated to [-2, g+1]
__device__ __forceinline__ u32 pack_home(float x, float y, float z, const GridInfo &g) {
    int cx = min(max((int)x, -2), g.gx + 1) + 2;
    int cy = min(max((int)y, -2), g.gy + 1) + 2;
    int cz = min(max((int)z - g.zoff, -2), g.gz + 1) + 2;
    return (u32)cx | ((u32)cz << g.bx) | ((u32)cy << (g.bx + g.bz));
}

// ---- reorder (gather the predicted record of each sorted slot) + K6 findcells.glsl:34-53 -----------------------
__global__ void __launch_bounds__(256)
k_reorder_cells(u32 n, const u32 *__restrict__ skey, const u32 *__restrict__ perm, const float4 *__restrict__ pred,
                float4 *__restrict__ bufA, u32 *__restrict__ home, int2 *__restrict__ cells, GridInfo g) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 id = perm[i];
    const float4 p = pred[id];
    bufA[i] = p;
    home[i] = pack_home(p.x, p.y, p.z, g);
    u32 k = skey[i];
    if (i == 0) {
        if (g.ref_quirks) cells[0].x = 0;                         // findcells.glsl:39-43
        else if (!(k & PBF_KEY_NOCELL)) cells[k].x = 0;
    } else {
        u32 kp = skey[i - 1];
        if (k != kp) {                                            // findcells.glsl:48-52
            if (!(k & PBF_KEY_NOCELL)) cells[k].x = (int)i;
            if (!(kp & PBF_KEY_NOCELL)) cells[kp].y = (int)i;
        }
    }
    if (i == n - 1 && !(k & PBF_KEY_NOCELL)) cells[k].y = (int)n;   // policy: end of the last occupied cell
}

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

// one thread per first-particle-of-a-cell: refresh the merged runs of the cells whose window contains that cell
__global__ void __launch_bounds__(256)
k_build_runs(u32 n, const u32 *__restrict__ skey, const int2 *__restrict__ cells, int2 *__restrict__ runs3, GridInfo g) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 && g.ref_quirks) {                                // start[(0,0,0)] = 0 is visible to cells 0 and 1
        runs3[0] = merge3(cells, 0, 0, g.gx);
        if (g.gx > 1) runs3[1] = merge3(cells, 0, 1, g.gx);
    }
    u32 k = skey[i];
    if (k & PBF_KEY_NOCELL) return;
    if (i != 0 && skey[i - 1] == k) return;
    const int x = (int)(k % (u32)g.gx);
    const int base = (int)k - x;
#pragma unroll
    for (int t = x - 1; t <= x + 1; t++)
        if (t >= 0 && t < g.gx) runs3[base + t] = merge3(cells, base, t, g.gx);
}

// ---- K7 neighbourcells.glsl:52-91: the thread's non-empty runs {start,end} into shared memory ----------------------
// returns the number of non-empty runs; *slots = number of aligned candidate pairs over all runs; *self_in = whether
// the particle's own slot lies in run 4 (its own row), i.e. whether FOR_EACH_NEIGHBOUR would have skipped `self`.
template <int BLOCK>
__device__ __forceinline__ int load_runs(const u32 home, const u32 i, const GridInfo &g, const int2 *__restrict__ runs3,
                                         const int2 *__restrict__ cells, int2 *srun, int tid, int *slots, bool *self_in) {
    const int cx = (int)(home & ((1u << g.bx) - 1u)) - 2;
    const int cz = (int)((home >> g.bx) & ((1u << g.bz) - 1u)) - 2;
    const int cy = (int)(home >> (g.bx + g.bz)) - 2;
    const bool fast = cx >= 0 && cx < g.gx;
    int2 r[9];
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
    int cnt = 0, tot = 0;
    *self_in = (int)i >= r[4].x && (int)i < r[4].x + r[4].y;
#pragma unroll
    for (int o = 0; o < 9; o++) {
        if (r[o].y > 0) {
            const int s = r[o].x, e = r[o].x + r[o].y;
            srun[cnt * BLOCK + tid] = make_int2(s, e);
            cnt++;
            tot += ((e + 1) >> 1) - (s >> 1);
        }
    }
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

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// geometry of one candidate pair against particle p: d = p - c, r2, 1/l, max(h^2 - r2, 0), max(h - l, 0)
struct PairGeom {
    float2 dx, dy, dz, r2, il, t, t2;
};

__device__ __forceinline__ PairGeom pair_geom(const float4 &p, const Pair &c, bool v0, bool v1) {
    PairGeom q;
    const float x0 = v0 ? c.x.x : FAR, x1 = v1 ? c.x.y : FAR;     // out-of-run members leave kernel support
    q.dx = make_float2(p.x - x0, p.x - x1);
    q.dy = make_float2(p.y - c.y.x, p.y - c.y.y);
    q.dz = make_float2(p.z - c.z.x, p.z - c.z.y);
    q.r2 = __ffma2_rn(q.dz, q.dz, __ffma2_rn(q.dy, q.dy, __fmul2_rn(q.dx, q.dx)));
    const float2 rc = __fadd2_rn(q.r2, make_float2(TINY, TINY));   // keeps rsqrt finite at r = 0 (self, coincident)
    q.il = make_float2(rsqrt_ftz(rc.x), rsqrt_ftz(rc.y));
    const float2 l = __fmul2_rn(q.r2, q.il);
    q.t2 = __ffma2_rn(l, make_float2(-1.0f, -1.0f), make_float2(H, H));
    q.t2 = make_float2(fmaxf(q.t2.x, 0.0f), fmaxf(q.t2.y, 0.0f));   // gradWspiky = 0 for l > h
    q.t = __fmul2_rn(q.t2, __fadd2_rn(l, make_float2(H, H)));       // h^2 - r^2 = (h-l)(h+l), 0 for r > h (Wpoly6)
    return q;
}

// flattened walk over the aligned candidate pairs of all runs; body(pair index m, valid0, valid1)
template <int BLOCK, class F>
__device__ __forceinline__ void for_each_pair(const int2 *srun, int tid, int slots, F body) {
    const int2 *sp = srun + tid;
    int m = 0, mend = 0, s = 0, e = 0;
#pragma unroll 1
    for (int k = 0; k < slots; k++) {
        if (m >= mend) {
            const int2 r = *sp;
            sp += BLOCK;
            s = r.x; e = r.y;
            m = s >> 1; mend = (e + 1) >> 1;
        }
        body(m, 2 * m >= s, 2 * m + 1 < e);
        m++;
    }
}

// ---- K8 calclambda.glsl:66-103 ------------------------------------------------------------------------------------
// out {x,y,z,lambda}.  rho (self excluded), S = sum |g_j|^2 + |sum g_j|^2, lambda = -C/(S+eps).
template <bool DIAG>
__global__ void __launch_bounds__(NB_BLOCK)
k_lambda(u32 n, const float4 *__restrict__ A, const u32 *__restrict__ home, const int2 *__restrict__ runs3,
         const int2 *__restrict__ cells, float4 *__restrict__ B, GridInfo g, SimParams P, double *diag) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    float err = 0.0f;
    if (i < n) {
        const float4 pi = A[i];
        int slots;
        bool self_in;
        load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
        float2 rho = make_float2(0.f, 0.f), S = rho, gx = rho, gy = rho, gz = rho;
        for_each_pair<NB_BLOCK>(srun, tid, slots, [&](int m, bool v0, bool v1) {
            const Pair c = ldg_pair(A + 2 * (size_t)m);
            const PairGeom q = pair_geom(pi, c, v0, v1);
            rho = __ffma2_rn(__fmul2_rn(q.t, q.t), q.t, rho);                 // sum (h^2-r^2)^3
            const float2 cc = __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il);       // (h-l)^2 / l
            S = __ffma2_rn(__fmul2_rn(cc, cc), q.r2, S);                      // sum |grad|^2 (up to a constant)
            gx = __ffma2_rn(cc, q.dx, gx);
            gy = __ffma2_rn(cc, q.dy, gy);
            gz = __ffma2_rn(cc, q.dz, gz);
        });
        // FOR_EACH_NEIGHBOUR skips j == i (foreachneighbour.glsl:9): self only ever adds (h^2)^3 = 64 to the poly6 sum
        float rs = rho.x + rho.y;
        if (self_in) rs -= 64.0f;
        const float r = POLY6 * rs;
        const float cg = SPIKY_GRAD * P.one_over_rho_0;
        const float sx = cg * (gx.x + gx.y), sy = cg * (gy.x + gy.y), sz = cg * (gz.x + gz.y);
        const float Ssum = cg * cg * (S.x + S.y) + (sx * sx + sy * sy + sz * sz);
        const float C = r * P.one_over_rho_0 - 1.0f;
        if (DIAG) err = fabsf(C);
        else B[i] = make_float4(pi.x, pi.y, pi.z, -C / (Ssum + P.epsilon));
    }
    if (DIAG) {
        __shared__ float red[NB_BLOCK / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
        if ((tid & 31) == 0) red[tid >> 5] = err;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NB_BLOCK / 32; w++) s += (double)red[w];
            atomicAdd(diag, s);
        }
    }
}

// ---- K9 updatepos.glsl:43-105, Jacobi: reads B {p, lambda}, writes A -----------------------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_delta_p(u32 n, const float4 *__restrict__ B, const u32 *__restrict__ home, const int2 *__restrict__ runs3,
          const int2 *__restrict__ cells, float4 *__restrict__ A, GridInfo g, SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = B[i];
    int slots;
    bool self_in;
    load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
    float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
    // scorr = -k (scale W)^4 = -(k scale^4 POLY6^4) t^12 with t = max(h^2 - r^2, 0)        (updatepos.glsl:57-60)
    float sc4 = P.tensile_scale * POLY6;
    sc4 *= sc4;
    sc4 *= sc4;
    const float nk = -P.tensile_k * sc4;
    const float2 nk2 = make_float2(nk, nk), li2 = make_float2(pi.w, pi.w);
    for_each_pair<NB_BLOCK>(srun, tid, slots, [&](int m, bool v0, bool v1) {
        const Pair c = ldg_pair(B + 2 * (size_t)m);
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
    A[i] = make_float4(x, y, z, 0.0f);
}

// ---- K10 update.glsl:16-28 -------------------------------------------------------------------------------------------
// position[id] <- p, velocity = (p - old)/dt.  With vorticity on, the velocity goes to the sorted array the
// vorticity kernels read (the by-id velocity is then written once, by k_vorticity_b).
template <bool VORT>
__global__ void __launch_bounds__(256)
k_update(u32 n, const float4 *__restrict__ A, const u32 *__restrict__ perm, float4 *__restrict__ pos,
         float4 *__restrict__ vel, float4 *__restrict__ svel, SimParams P) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = A[i];
    const u32 id = perm[i];
    const float4 o = pos[id];
    float4 v;
    v.x = __fdiv_rn(__fsub_rn(p.x, o.x), P.timestep);
    v.y = __fdiv_rn(__fsub_rn(p.y, o.y), P.timestep);
    v.z = __fdiv_rn(__fsub_rn(p.z, o.z), P.timestep);
    v.w = 0.0f;
    pos[id] = make_float4(p.x, p.y, p.z, 0.0f);
    if (VORT) svel[i] = v;
    else vel[id] = v;
}

// ---- K11 vorticity.glsl:34-60 (first sweep): XSPH + vorticity ------------------------------------------------------------
// out: vprime = v + c*sum v_ij W, omega = sum v_ij x gradW, B = {p, |omega|}
__global__ void __launch_bounds__(NB_BLOCK)
k_vorticity_a(u32 n, const float4 *__restrict__ A, const float4 *__restrict__ svel, const u32 *__restrict__ home,
              const int2 *__restrict__ runs3, const int2 *__restrict__ cells, float4 *__restrict__ B,
              float4 *__restrict__ vprime, float4 *__restrict__ omega, GridInfo g, SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = A[i];
    const float4 vi = svel[i];
    int slots;
    bool self_in;
    load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
    float2 vx = make_float2(0.f, 0.f), vy = vx, vz = vx, wx = vx, wy = vx, wz = vx;
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    for_each_pair<NB_BLOCK>(srun, tid, slots, [&](int m, bool v0, bool v1) {
        const Pair c = ldg_pair(A + 2 * (size_t)m);
        const Pair u = ldg_pair(svel + 2 * (size_t)m);
        const PairGeom q = pair_geom(pi, c, v0, v1);
        const float2 ux = make_float2(u.x.x - vi.x, u.x.y - vi.x);           // v_ij = v_j - v_i
        const float2 uy = make_float2(u.y.x - vi.y, u.y.y - vi.y);
        const float2 uz = make_float2(u.z.x - vi.z, u.z.y - vi.z);
        const float2 w = __fmul2_rn(__fmul2_rn(q.t, q.t), q.t);              // Wpoly6 / POLY6
        vx = __ffma2_rn(ux, w, vx);
        vy = __ffma2_rn(uy, w, vy);
        vz = __ffma2_rn(uz, w, vz);
        const float2 cc = __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il);          // grad = SPIKY_GRAD * cc * d
        const float2 gx = __fmul2_rn(cc, q.dx), gy = __fmul2_rn(cc, q.dy), gz = __fmul2_rn(cc, q.dz);
        // cross(v_ij, grad)
        wx = __ffma2_rn(uy, gz, __ffma2_rn(__fmul2_rn(gy, uz), neg1, wx));
        wy = __ffma2_rn(uz, gx, __ffma2_rn(__fmul2_rn(gz, ux), neg1, wy));
        wz = __ffma2_rn(ux, gy, __ffma2_rn(__fmul2_rn(gx, uy), neg1, wz));
    });
    const float cw = P.xsph_c * POLY6;
    vprime[i] = make_float4(vi.x + cw * (vx.x + vx.y), vi.y + cw * (vy.x + vy.y), vi.z + cw * (vz.x + vz.y), 0.0f);
    const float ox = SPIKY_GRAD * (wx.x + wx.y), oy = SPIKY_GRAD * (wy.x + wy.y), oz = SPIKY_GRAD * (wz.x + wz.y);
    omega[i] = make_float4(ox, oy, oz, 0.0f);
    B[i] = make_float4(pi.x, pi.y, pi.z, sqrtf(ox * ox + oy * oy + oz * oz));   // vorticity.glsl:60
}

// ---- K11 vorticity.glsl:65-85 (second sweep): confinement force, velocity[id] written once -------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_vorticity_b(u32 n, const float4 *__restrict__ B, const float4 *__restrict__ vprime, const float4 *__restrict__ omega,
              const u32 *__restrict__ perm, const u32 *__restrict__ home, const int2 *__restrict__ runs3,
              const int2 *__restrict__ cells, float4 *__restrict__ vel, GridInfo g, SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = B[i];
    int slots;
    bool self_in;
    load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
    float2 ex = make_float2(0.f, 0.f), ey = ex, ez = ex;
    for_each_pair<NB_BLOCK>(srun, tid, slots, [&](int m, bool v0, bool v1) {
        const Pair c = ldg_pair(B + 2 * (size_t)m);
        const PairGeom q = pair_geom(pi, c, v0, v1);
        const float2 cc = __fmul2_rn(c.w, __fmul2_rn(__fmul2_rn(q.t2, q.t2), q.il));   // |omega_j| * grad factor
        ex = __ffma2_rn(cc, q.dx, ex);
        ey = __ffma2_rn(cc, q.dy, ey);
        ez = __ffma2_rn(cc, q.dz, ez);
    });
    float nx = SPIKY_GRAD * (ex.x + ex.y), ny = SPIKY_GRAD * (ey.x + ey.y), nz = SPIKY_GRAD * (ez.x + ez.y);
    const float l = sqrtf(nx * nx + ny * ny + nz * nz);
    if (l > 0.0f) { nx /= l; ny /= l; nz /= l; }
    const float4 w = omega[i];
    const float4 v = vprime[i];
    const float s = P.timestep * P.vort_eps;
    vel[perm[i]] = make_float4(v.x + s * (ny * w.z - w.y * nz), v.y + s * (nz * w.x - w.z * nx),
                               v.z + s * (nx * w.y - w.x * ny), 0.0f);       // cross(N, omega)
}

// ---- K12 highlight.glsl:17-30 (clearhighlight is fused into k_predict) ---------------------------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_highlight(u32 n, const u32 *__restrict__ home, const u32 *__restrict__ perm, const int2 *__restrict__ runs3,
            const int2 *__restrict__ cells, u32 *__restrict__ hl, const u32 *__restrict__ flags, GridInfo g) {
    __shared__ int2 srun[9 * NB_BLOCK];
    if (flags[0] == 0u) return;                      // nobody carries bit 0: the kernel is a no-op
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    if ((hl[perm[i]] & 1u) == 0u) return;
    int slots;
    bool self_in;
    const int cnt = load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
    for (int o = 0; o < cnt; o++) {
        const int2 r = srun[o * NB_BLOCK + tid];
        for (int j = r.x; j < r.y; j++)
            if (j != (int)i) atomicOr(&hl[perm[j]], 2u);
    }
}

// ---- debug / diagnostics ---------------------------------------------------------------------------------------------------
__global__ void k_compose_records(u32 n, const float4 *__restrict__ A, const u32 *__restrict__ perm, float4 *__restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = A[i];
    p.w = __int_as_float((int)perm[i]);
    out[i] = p;
}

// the 9 runs in reference order, unpacked (start, count); goes through the same runs3 / slow-path lookups as the sweeps
__global__ void __launch_bounds__(256)
k_neighbour_runs(u32 n, const u32 *__restrict__ home_, const int2 *__restrict__ runs3, const int2 *__restrict__ cells,
                 int *__restrict__ rs, int *__restrict__ rc, GridInfo g) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 home = home_[i];
    const int cx = (int)(home & ((1u << g.bx) - 1u)) - 2;
    const int cz = (int)((home >> g.bx) & ((1u << g.bz) - 1u)) - 2;
    const int cy = (int)(home >> (g.bx + g.bz)) - 2;
    for (int o = 0; o < 9; o++) {
        const int yy = cy + (o / 3 - 1), zz = cz + (o % 3 - 1);
        int2 r = make_int2(-1, 0);
        if (yy >= 0 && yy < g.gy && zz >= 0 && zz < g.gz) {
            const int base = yy * g.gxgz + zz * g.gx;
            r = (cx >= 0 && cx < g.gx) ? runs3[base + cx] : merge3(cells, base, cx, g.gx);
        }
        rs[9 * (size_t)i + o] = r.x;
        rc[9 * (size_t)i + o] = r.y;
    }
}

__global__ void __launch_bounds__(256) k_kinetic(u32 n, const float4 *__restrict__ vel, double *diag) {
    __shared__ double red[8];
    double e = 0.0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = vel[i];
        e += 0.5 * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w];
        atomicAdd(diag + 1, s);
    }
}

SimParams sim_params(const pbf_sim *s) {
    SimParams P;
    P.one_over_rho_0 = s->params.one_over_rho_0;
    P.epsilon = s->params.epsilon;
    P.gravity = s->params.gravity;
    P.timestep = s->params.timestep;
    P.tensile_k = s->params.tensile_instability_k;
    P.tensile_scale = s->params.tensile_instability_scale;
    P.xsph_c = s->params.xsph_viscosity_c;
    P.vort_eps = s->params.vorticity_epsilon;
    P.extforce = s->params.external_force;
    return P;
}

inline int nblocks(size_t n, int b) { return (int)((n + b - 1) / b); }

}  // namespace

int launch_fill_tables(pbf_sim *s) {
    k_fill_tables<<<nblocks(s->ncell, 256), 256, 0, s->stream>>>(s->ncell, s->cells, s->runs3);
    return 1;
}

int launch_unclear_cells(pbf_sim *s) {
    if (s->n_prev_sorted == 0) return 0;
    k_unclear_cells<<<nblocks(s->n_prev_sorted, 256), 256, 0, s->stream>>>(s->n_prev_sorted, s->skey, s->cells, s->runs3,
                                                                           s->grid);
    return 1;
}

int launch_predict_range(pbf_sim *s, u32 first, u32 count, bool with_hist) {
    if (count == 0) return 0;
    int blocks = nblocks(count, 256);
    int maxb = s->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    SortPlan plan = s->plan;
    if (!with_hist) plan.passes = 0;
    k_predict<<<blocks, 256, 0, s->stream>>>(first, count, s->pos, s->vel, s->hl, s->pred, s->keys, s->hist, s->flags,
                                             s->grid, sim_params(s), plan);
    return 1;
}

int launch_predict(pbf_sim *s) { return launch_predict_range(s, 0, s->n, true); }

int launch_keys_only(pbf_sim *s, u32 first, u32 count) {
    if (count == 0) return 0;
    k_keys_only<<<nblocks(count, 256), 256, 0, s->stream>>>(first, count, s->pred, s->keys, s->grid);
    return 1;
}

int launch_reorder_cells(pbf_sim *s) {
    k_reorder_cells<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->skey, s->perm, s->pred, s->bufA, s->home,
                                                               s->cells, s->grid);
    k_build_runs<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->skey, s->cells, s->runs3, s->grid);
    s->n_prev_sorted = s->n;
    return 2;
}

int launch_highlight(pbf_sim *s) {
    k_highlight<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->home, s->perm, s->runs3, s->cells, s->hl,
                                                                     s->flags, s->grid);
    return 1;
}

int launch_lambda(pbf_sim *s) {
    k_lambda<false><<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->home, s->runs3, s->cells,
                                                                         s->bufB, s->grid, sim_params(s), nullptr);
    return 1;
}

int launch_delta_p(pbf_sim *s) {
    k_delta_p<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufB, s->home, s->runs3, s->cells, s->bufA,
                                                                   s->grid, sim_params(s));
    return 1;
}

int launch_update(pbf_sim *s) {
    if (s->params.vorticity_confinement)
        k_update<true><<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->bufA, s->perm, s->pos, s->vel, s->svel,
                                                                  sim_params(s));
    else
        k_update<false><<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->bufA, s->perm, s->pos, s->vel, s->svel,
                                                                   sim_params(s));
    return 1;
}

int launch_vorticity_a(pbf_sim *s) {
    k_vorticity_a<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->svel, s->home, s->runs3,
                                                                       s->cells, s->bufB, s->vprime, s->omega, s->grid,
                                                                       sim_params(s));
    return 1;
}

int launch_vorticity_b(pbf_sim *s) {
    k_vorticity_b<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufB, s->vprime, s->omega, s->perm,
                                                                       s->home, s->runs3, s->cells, s->vel, s->grid,
                                                                       sim_params(s));
    return 1;
}

int launch_vorticity(pbf_sim *s) { return launch_vorticity_a(s) + launch_vorticity_b(s); }

int launch_density_diag(pbf_sim *s) {
    k_lambda<true><<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->home, s->runs3, s->cells,
                                                                        s->bufB, s->grid, sim_params(s), s->diag);
    return 1;
}

int launch_kinetic_diag(pbf_sim *s) {
    int blocks = nblocks(s->n, 256);
    if (blocks > s->sm_count * 8) blocks = s->sm_count * 8;
    k_kinetic<<<blocks, 256, 0, s->stream>>>(s->n, s->vel, s->diag);
    return 1;
}

int launch_compose_records(pbf_sim *s, float4 *out) {
    k_compose_records<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->bufA, s->perm, out);
    return 1;
}

int launch_neighbour_runs(pbf_sim *s, int *run_start, int *run_count) {
    k_neighbour_runs<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->home, s->runs3, s->cells, run_start, run_count,
                                                                s->grid);
    return 1;
}
