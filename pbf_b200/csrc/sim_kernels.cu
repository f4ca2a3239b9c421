// sim_kernels.cu -- the per-timestep PBF kernels (everything of SPH::Run except the sort).
//
// Data layout in HBM (all SoA, 16-byte records):
//   pos/vel/hl   by persistent particle id (the renderer-visible buffers, src/SPH.cpp:106-133)
//   pred, keys   by id: predicted record {p*, id} (predictpos.glsl:37) and its cell key
//   skey, perm   by sorted slot: cell key and particle id (output of the onesweep sort)
//   cells        dense table, int2 {start,end} per cell, index x + z*gx + y*gx*gz (findcells.glsl:45-52)
//   bufA         by sorted slot {x,y,z,-}: positions the solver iterates on
//   bufB         by sorted slot {x,y,z,lambda}: written by the lambda kernel so that the delta-p kernel needs ONE
//                16-byte load per neighbour (the reference does a record load + a lambda texel fetch)
//   svel, vprime, omega  by sorted slot, vorticity/XSPH only
//
// Neighbour iteration reproduces FOR_EACH_NEIGHBOUR (shaders/sph/foreachneighbour.glsl:1-10) exactly: 9 runs in
// the order of neighbourcells.glsl:37-47, each the merged cells x-1..x+1, ascending index.  The reference stores
// these runs per particle (48 B, K7); here each thread rebuilds them from the cell table (L1/L2 resident) into
// shared memory and walks them as ONE flattened loop, so a warp runs max(total) iterations instead of the sum
// over runs of max(run length).
#include "pbf_internal.cuh"

namespace {

constexpr int NB_BLOCK = 256;   // threads per block of the neighbour kernels (= BLOCKSIZE of src/SPH.cpp:60)

constexpr float H = 2.0f;                                // src/SPH.cpp:58
constexpr float H2 = 4.0f;
constexpr float POLY6 = 1.56668147106f / 512.0f;         // calclambda.glsl:46, /h^9
constexpr float SPIKY_GRAD = -3.0f * 4.774648292756860f / 64.0f;   // calclambda.glsl:63, /h^6

__device__ __forceinline__ u32 cell_key(float x, float y, float z, const GridInfo &g) {
    // ivec3(clamp(pos, 0, GRID_SIZE)) . (1, gx*gz, gx)   (counting.glsl:53-57); clamp is inclusive
    int cx = (int)fminf(fmaxf(x, 0.0f), (float)g.gx);
    int cy = (int)fminf(fmaxf(y, 0.0f), (float)g.gy);
    int cz = (int)fminf(fmaxf(z, 0.0f), (float)g.gz);
    u32 k = (u32)cx + (u32)cz * (u32)g.gx + (u32)cy * (u32)g.gxgz;
    if (cx >= g.gx || cy >= g.gy || cz >= g.gz) k |= PBF_KEY_NOCELL;
    return k;
}

// ---- K1 predictpos.glsl:18-38 + cell key + digit histograms of every sort pass + clearhighlight.glsl ----------
__global__ void __launch_bounds__(256)
k_predict(u32 n, const float4 *__restrict__ pos, const float4 *__restrict__ vel, u32 *__restrict__ hl,
          float4 *__restrict__ pred, u32 *__restrict__ keys, u32 *__restrict__ hist, u32 *__restrict__ flags,
          GridInfo g, SimParams P, SortPlan plan) {
    __shared__ u32 sh[4 * PBF_RADIX];
    for (int i = threadIdx.x; i < 4 * PBF_RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 stride = gridDim.x * blockDim.x;
    const u32 nround = (n + 31u) & ~31u;
    bool any_hl = false;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        const bool valid = i < n;
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
        for (int p = 0; p < plan.passes; p++) {
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

// appended (halo) records already hold p*: only keys + histograms
__global__ void __launch_bounds__(256)
k_keys_only(u32 first, u32 n, const float4 *__restrict__ pred, u32 *__restrict__ keys, u32 *__restrict__ hist,
            GridInfo g, SortPlan plan) {
    u32 i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pred[i];
    u32 key = cell_key(p.x, p.y, p.z, g);
    keys[i] = key;
    for (int q = 0; q < plan.passes; q++) atomicAdd(&hist[q * PBF_RADIX + ((key >> plan.shift[q]) & plan.mask[q])], 1u);
}

// ---- reset the start entries the previous step wrote (replaces the per-step clear of the whole start image,
//      src/NeighbourCellFinder.cpp:116-126: 4 B per previously sorted particle instead of 4 B per cell) ----------
__global__ void __launch_bounds__(256) k_unclear_cells(u32 n, const u32 *__restrict__ skey, int2 *__restrict__ cells) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 k = skey[i];
    if (k & PBF_KEY_NOCELL) return;
    if (i == 0 || skey[i - 1] != k) cells[k].x = -1;
}

__device__ __forceinline__ u32 pack_home(float x, float y, float z, const GridInfo &g);

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

// ---- K7 neighbourcells.glsl:52-91, per thread into shared memory ------------------------------------------------
// The reference evaluates K7 once per step on the PREDICTED position (ivec3(pos), not clamped, :57) and keeps the
// 9 runs for every later sweep.  Here k_reorder_cells stores that home cell packed in 4 bytes (each coordinate
// saturated to [-2, g+1], biased by 2) and every sweep rebuilds the runs from it.
__device__ __forceinline__ u32 pack_home(float x, float y, float z, const GridInfo &g) {
    int cx = min(max((int)x, -2), g.gx + 1) + 2;
    int cy = min(max((int)y, -2), g.gy + 1) + 2;
    int cz = min(max((int)z, -2), g.gz + 1) + 2;
    return (u32)cx | ((u32)cz << g.bx) | ((u32)cy << (g.bx + g.bz));
}

template <int BLOCK>
__device__ __forceinline__ void build_runs(const u32 home, const GridInfo &g, const int2 *__restrict__ cells,
                                           int2 *srun, int tid) {
    const int cx = (int)(home & ((1u << g.bx) - 1u)) - 2;
    const int cz = (int)((home >> g.bx) & ((1u << g.bz) - 1u)) - 2;
    const int cy = (int)(home >> (g.bx + g.bz)) - 2;
#pragma unroll
    for (int o = 0; o < 9; o++) {
        const int yy = cy + (o / 3 - 1), zz = cz + (o % 3 - 1);   // gridoffsets[o] = (0, dy, dz)
        int cell = -1, entries = 0;
        if (yy >= 0 && yy < g.gy && zz >= 0 && zz < g.gz) {
            const int base = yy * g.gxgz + zz * g.gx;
#pragma unroll
            for (int j = -1; j <= 1; j++) {
                const int xx = cx + j;
                if (xx >= 0 && xx < g.gx) {
                    int2 c = __ldg(cells + base + xx);
                    if (cell == -1) cell = c.x;
                    if (c.x != -1) entries += c.y - c.x;
                }
            }
        }
        srun[o * BLOCK + tid] = make_int2(cell, cell == -1 ? 0 : entries);
    }
}

// flattened FOR_EACH_NEIGHBOUR; body(j) is also called for j == i (every body is a no-op or masks it)
template <int BLOCK, class F>
__device__ __forceinline__ void for_each_neighbour(const int2 *srun, int tid, F body) {
    int o = 0;
    int2 r = srun[tid];
    int j = r.x, e = r.x + r.y;
    while (true) {
        while (j >= e) {
            if (++o >= 9) return;
            r = srun[o * BLOCK + tid];
            j = r.x;
            e = r.x + r.y;
        }
        body(j);
        j++;
    }
}

// ---- K8 calclambda.glsl:66-103 ------------------------------------------------------------------------------------
// out {x,y,z,lambda}.  rho (self excluded), S = sum |g_j|^2 + |sum g_j|^2, lambda = -C/(S+eps).
template <bool DIAG>
__global__ void __launch_bounds__(NB_BLOCK)
k_lambda(u32 n, const float4 *__restrict__ A, const u32 *__restrict__ home, const int2 *__restrict__ cells, float4 *__restrict__ B, GridInfo g,
         SimParams P, double *diag) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    float err = 0.0f;
    if (i < n) {
        const float4 pi = A[i];
        build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
        float rho = 0.0f, S = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
        const float cg = SPIKY_GRAD * P.one_over_rho_0;
        for_each_neighbour<NB_BLOCK>(srun, tid, [&](int j) {
            const float4 pj = __ldg(A + j);
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float r2 = dx * dx + dy * dy + dz * dz;
            const float t = (j != (int)i) ? fmaxf(H2 - r2, 0.0f) : 0.0f;   // Wpoly6 = 0 for r > h; self skipped
            rho += POLY6 * t * t * t;
            const float il = r2 > 0.0f ? rsqrtf(r2) : 0.0f;                // gradWspiky = 0 for l == 0
            const float t2 = fmaxf(H - r2 * il, 0.0f);                     //            = 0 for l > h
            const float c = cg * t2 * t2 * il;
            S += c * c * r2;
            gx += c * dx; gy += c * dy; gz += c * dz;
        });
        S += gx * gx + gy * gy + gz * gz;
        const float C = rho * P.one_over_rho_0 - 1.0f;
        if (DIAG) err = fabsf(C);
        else B[i] = make_float4(pi.x, pi.y, pi.z, -C / (S + P.epsilon));
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
k_delta_p(u32 n, const float4 *__restrict__ B, const u32 *__restrict__ home, const int2 *__restrict__ cells, float4 *__restrict__ A, GridInfo g,
          SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = B[i];
    build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    const float sc = P.tensile_scale * POLY6;
    for_each_neighbour<NB_BLOCK>(srun, tid, [&](int j) {
        const float4 pj = __ldg(B + j);
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float t = fmaxf(H2 - r2, 0.0f);
        float s = sc * t * t * t;                                          // scale * Wpoly6
        s *= s;
        s *= s;
        const float il = r2 > 0.0f ? rsqrtf(r2) : 0.0f;                    // self / coincident: gradient 0
        const float t2 = fmaxf(H - r2 * il, 0.0f);
        const float c = (pi.w + pj.w - P.tensile_k * s) * (SPIKY_GRAD * t2 * t2 * il);
        ax += c * dx; ay += c * dy; az += c * dz;
    });
    float x = pi.x + P.one_over_rho_0 * ax, y = pi.y + P.one_over_rho_0 * ay, z = pi.z + P.one_over_rho_0 * az;
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
              const int2 *__restrict__ cells,
              float4 *__restrict__ B, float4 *__restrict__ vprime, float4 *__restrict__ omega, GridInfo g, SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = A[i];
    const float4 vi = svel[i];
    build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
    float vx = 0.0f, vy = 0.0f, vz = 0.0f, wx = 0.0f, wy = 0.0f, wz = 0.0f;
    for_each_neighbour<NB_BLOCK>(srun, tid, [&](int j) {
        const float4 pj = __ldg(A + j);
        const float4 vj = __ldg(svel + j);
        const float ux = vj.x - vi.x, uy = vj.y - vi.y, uz = vj.z - vi.z;    // v_ij
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;    // p_ij
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float t = fmaxf(H2 - r2, 0.0f);
        const float w = POLY6 * t * t * t;
        vx += ux * w; vy += uy * w; vz += uz * w;
        const float il = r2 > 0.0f ? rsqrtf(r2) : 0.0f;
        const float t2 = fmaxf(H - r2 * il, 0.0f);
        const float c = SPIKY_GRAD * t2 * t2 * il;
        const float gx = c * dx, gy = c * dy, gz = c * dz;
        wx += uy * gz - gy * uz;                                              // cross(v_ij, grad)
        wy += uz * gx - gz * ux;
        wz += ux * gy - gx * uy;
    });
    vprime[i] = make_float4(vi.x + P.xsph_c * vx, vi.y + P.xsph_c * vy, vi.z + P.xsph_c * vz, 0.0f);
    omega[i] = make_float4(wx, wy, wz, 0.0f);
    B[i] = make_float4(pi.x, pi.y, pi.z, sqrtf(wx * wx + wy * wy + wz * wz));   // vorticity.glsl:60
}

// ---- K11 vorticity.glsl:65-85 (second sweep): confinement force, velocity[id] written once -------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_vorticity_b(u32 n, const float4 *__restrict__ B, const float4 *__restrict__ vprime, const float4 *__restrict__ omega,
              const u32 *__restrict__ perm, const u32 *__restrict__ home, const int2 *__restrict__ cells,
              float4 *__restrict__ vel, GridInfo g,
              SimParams P) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    const float4 pi = B[i];
    build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
    float ex = 0.0f, ey = 0.0f, ez = 0.0f;
    for_each_neighbour<NB_BLOCK>(srun, tid, [&](int j) {
        const float4 pj = __ldg(B + j);
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        const float il = r2 > 0.0f ? rsqrtf(r2) : 0.0f;
        const float t2 = fmaxf(H - r2 * il, 0.0f);
        const float c = pj.w * (SPIKY_GRAD * t2 * t2 * il);
        ex += c * dx; ey += c * dy; ez += c * dz;
    });
    const float l = sqrtf(ex * ex + ey * ey + ez * ez);
    if (l > 0.0f) { ex /= l; ey /= l; ez /= l; }
    const float4 w = omega[i];
    const float4 v = vprime[i];
    const float s = P.timestep * P.vort_eps;
    vel[perm[i]] = make_float4(v.x + s * (ey * w.z - w.y * ez), v.y + s * (ez * w.x - w.z * ex),
                               v.z + s * (ex * w.y - w.x * ey), 0.0f);       // cross(N, omega)
}

// ---- K12 highlight.glsl:17-30 (clearhighlight is fused into k_predict) ---------------------------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_highlight(u32 n, const u32 *__restrict__ home, const u32 *__restrict__ perm, const int2 *__restrict__ cells,
            u32 *__restrict__ hl, const u32 *__restrict__ flags, GridInfo g) {
    __shared__ int2 srun[9 * NB_BLOCK];
    if (flags[0] == 0u) return;                      // nobody carries bit 0: the kernel is a no-op
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    if ((hl[perm[i]] & 1u) == 0u) return;
    build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
    for_each_neighbour<NB_BLOCK>(srun, tid, [&](int j) {
        if (j != (int)i) atomicOr(&hl[perm[j]], 2u);
    });
}

// ---- debug / diagnostics ---------------------------------------------------------------------------------------------------
__global__ void k_compose_records(u32 n, const float4 *__restrict__ A, const u32 *__restrict__ perm, float4 *__restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = A[i];
    p.w = __int_as_float((int)perm[i]);
    out[i] = p;
}

__global__ void __launch_bounds__(NB_BLOCK)
k_neighbour_runs(u32 n, const u32 *__restrict__ home, const int2 *__restrict__ cells, int *__restrict__ rs,
                 int *__restrict__ rc, GridInfo g) {
    __shared__ int2 srun[9 * NB_BLOCK];
    const int tid = threadIdx.x;
    const u32 i = blockIdx.x * NB_BLOCK + tid;
    if (i >= n) return;
    build_runs<NB_BLOCK>(home[i], g, cells, srun, tid);
    for (int o = 0; o < 9; o++) {
        int2 r = srun[o * NB_BLOCK + tid];
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

inline int nblocks(u32 n, int b) { return (int)((n + b - 1) / b); }

}  // namespace

int launch_unclear_cells(pbf_sim *s) {
    if (s->n_prev_sorted == 0) return 0;
    k_unclear_cells<<<nblocks(s->n_prev_sorted, 256), 256, 0, s->stream>>>(s->n_prev_sorted, s->skey, s->cells);
    return 1;
}

int launch_predict(pbf_sim *s) {
    int blocks = nblocks(s->n, 256);
    int maxb = s->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    k_predict<<<blocks, 256, 0, s->stream>>>(s->n, s->pos, s->vel, s->hl, s->pred, s->keys, s->hist, s->flags, s->grid,
                                             sim_params(s), s->plan);
    return 1;
}

int launch_reorder_cells(pbf_sim *s) {
    k_reorder_cells<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->skey, s->perm, s->pred, s->bufA, s->home,
                                                               s->cells, s->grid);
    s->n_prev_sorted = s->n;
    return 1;
}

int launch_highlight(pbf_sim *s) {
    k_highlight<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->home, s->perm, s->cells, s->hl, s->flags,
                                                                     s->grid);
    return 1;
}

int launch_lambda(pbf_sim *s) {
    k_lambda<false><<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->home, s->cells, s->bufB, s->grid,
                                                                         sim_params(s), nullptr);
    return 1;
}

int launch_delta_p(pbf_sim *s) {
    k_delta_p<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufB, s->home, s->cells, s->bufA, s->grid,
                                                                   sim_params(s));
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

int launch_vorticity(pbf_sim *s) {
    k_vorticity_a<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->svel, s->home, s->cells, s->bufB,
                                                                       s->vprime, s->omega, s->grid, sim_params(s));
    k_vorticity_b<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufB, s->vprime, s->omega, s->perm,
                                                                       s->home, s->cells, s->vel, s->grid, sim_params(s));
    return 2;
}

int launch_density_diag(pbf_sim *s) {
    k_lambda<true><<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->bufA, s->home, s->cells, s->bufB, s->grid,
                                                                        sim_params(s), s->diag);
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
    k_neighbour_runs<<<nblocks(s->n, NB_BLOCK), NB_BLOCK, 0, s->stream>>>(s->n, s->home, s->cells, run_start, run_count,
                                                                          s->grid);
    return 1;
}
