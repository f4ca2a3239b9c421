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
// The neighbour sweeps (lambda, delta-p, vorticity) live in sweeps.cu; the helpers both share are in neighbour.cuh.
#include "neighbour.cuh"

namespace {

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

// Reset what the previous step wrote into the cell tables (replaces the per-step clear of the whole start image,
// src/NeighbourCellFinder.cpp:116-126: 4 B per previously sorted particle instead of 4 B per cell).  i = a sorted slot of
// the previous step.
__device__ __forceinline__ void unclear_slot(u32 i, const u32 *__restrict__ skey, int2 *__restrict__ cells,
                                             int2 *__restrict__ runs3, const GridInfo &g) {
    const u32 k = skey[i];
    if (!(k & PBF_KEY_NOCELL) && (i == 0 || skey[i - 1] != k)) {
        cells[k].x = -1;
        const int x = (int)(k % (u32)g.gx);
        const int2 empty = make_int2(-1, 0);
        if (x > 0) runs3[k - 1] = empty;
        runs3[k] = empty;
        if (x + 1 < g.gx) runs3[k + 1] = empty;
    }
    if (i == 0 && g.ref_quirks) { runs3[0] = make_int2(-1, 0); if (g.gx > 1) runs3[1] = make_int2(-1, 0); }
}

// ---- K1 predictpos.glsl:18-38 + cell key + clearhighlight.glsl: one particle per thread, pure streaming --------------
// UNCLEAR: thread j also resets the table entries of the previous step's sorted slot j (the two jobs share nothing but the
// launch: the scattered resets hide under the streaming loads).
// SLAB: the thread also files its particle under "left through z-" / "left through z+" (slab.cu, migration).
template <bool UNCLEAR, bool SLAB>
__global__ void __launch_bounds__(256)
k_predict(u32 first, NRef nr, const float4 *__restrict__ pos, const float4 *__restrict__ vel, u32 *__restrict__ hl,
          float4 *__restrict__ pred, u32 *__restrict__ keys, u32 *__restrict__ flags, GridInfo g, SimParams P,
          NRef nprev, const u32 *__restrict__ skey, int2 *__restrict__ cells, int2 *__restrict__ runs3, LeaveArgs la) {
    const u32 n = nref(nr), n_prev = UNCLEAR ? nref(nprev) : 0u, m = n > n_prev ? n : n_prev;
    // one element per thread; the loop only turns when the grid was sized for fewer particles than there are (NRef)
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < ((m + 31u) & ~31u); j += gridDim.x * blockDim.x) {
    if (UNCLEAR && j < n_prev) unclear_slot(j, skey, cells, runs3, g);
    bool any_hl = false;
    if (j < n) {
        const u32 i = first + j;
        float4 p = pos[i];
        float4 v = vel[i];
        const u32 h = hl[i];
        // exact, uncontracted binary32 in the order of the shader so that keys are bit-identical to the oracle
        // GRID_SIZE.z/2 of the whole domain (predictpos.glsl:27): on a slab handle g.gz is only the window's depth
        if (P.extforce && p.z > (float)g.gz_global / 2.0f)
            v.z = __fadd_rn(v.z, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, P.gravity), -1.0f), P.timestep));
        v.y = __fadd_rn(v.y, __fmul_rn(__fmul_rn(P.gravity, -1.0f), P.timestep));
        p.x = __fadd_rn(p.x, __fmul_rn(P.timestep, v.x));
        p.y = __fadd_rn(p.y, __fmul_rn(P.timestep, v.y));
        p.z = __fadd_rn(p.z, __fmul_rn(P.timestep, v.z));
        p.w = __int_as_float((int)i);
        pred[i] = p;
        keys[i] = cell_key(p.x, p.y, p.z, g);
        // clearhighlight.glsl: flag &= 1 (written back only when it changes anything)
        if (h & ~1u) hl[i] = h & 1u;
        any_hl = (h & 1u) != 0;
        if (SLAB) {
            const int cz = (int)fminf(fmaxf(p.z, 0.0f), (float)g.gz_global);     // global cell layer of p*
            u32 tag = 0;
            if (la.has_lo && cz < la.z_lo) {
                const u32 k = atomicAdd(&la.cnt[0], 1u);
                if (k < la.cap) la.leave_lo[k] = i;
                tag = 0xffffffffu;
            } else if (la.has_hi && cz >= la.z_hi) {
                const u32 k = atomicAdd(&la.cnt[1], 1u);
                if (k < la.cap) la.leave_hi[k] = i;
                tag = 0xffffffffu;
            } else if (la.bnd_lo && la.has_lo && cz == la.z_lo) {          // the neighbours' ghosts-to-be (k_mark_boundary)
                const u32 k = atomicAdd(&la.cnt[2], 1u);
                if (k < la.cap) { la.bnd_lo[k] = i; tag = k + 1u; }
            } else if (la.bnd_hi && la.has_hi && cz == la.z_hi - 1) {
                const u32 k = atomicAdd(&la.cnt[3], 1u);
                if (k < la.cap) { la.bnd_hi[k] = i; tag = 0x80000000u | (k + 1u); }
            }
            la.btag[i] = tag;
        }
    }
    if (__any_sync(0xffffffffu, any_hl) && (threadIdx.x & 31) == 0) flags[0] = 1u;
    }
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

__global__ void __launch_bounds__(256)
k_unclear_cells(NRef nr, const u32 *__restrict__ skey, int2 *__restrict__ cells, int2 *__restrict__ runs3, GridInfo g) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) unclear_slot(i, skey, cells, runs3, g);
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
k_reorder_cells(NRef nr, const u32 *__restrict__ skey, const u32 *__restrict__ perm, const float4 *__restrict__ pred,
                float4 *__restrict__ bufA, u32 *__restrict__ home, int2 *__restrict__ cells, GridInfo g) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
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
}

// one thread per first-particle-of-a-cell: refresh the merged runs of the cells whose window contains that cell
__global__ void __launch_bounds__(256)
k_build_runs(NRef nr, const u32 *__restrict__ skey, const int2 *__restrict__ cells, int2 *__restrict__ runs3, GridInfo g) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i == 0 && g.ref_quirks) {                                // start[(0,0,0)] = 0 is visible to cells 0 and 1
            runs3[0] = merge3(cells, 0, 0, g.gx);
            if (g.gx > 1) runs3[1] = merge3(cells, 0, 1, g.gx);
        }
        u32 k = skey[i];
        if (k & PBF_KEY_NOCELL) continue;
        if (i != 0 && skey[i - 1] == k) continue;
        const int x = (int)(k % (u32)g.gx);
        const int base = (int)k - x;
#pragma unroll
        for (int t = x - 1; t <= x + 1; t++)
            if (t >= 0 && t < g.gx) runs3[base + t] = merge3(cells, base, t, g.gx);
    }
}

// ---- K10 update.glsl:16-28 -------------------------------------------------------------------------------------------
// position[id] <- p, velocity = (p - old)/dt.  With vorticity on, the velocity goes to the sorted array the
// vorticity kernels read (the by-id velocity is then written once, by k_vorticity_b).
template <bool VORT>
__global__ void __launch_bounds__(256)
k_update(NRef nr, const float4 *__restrict__ A, const u32 *__restrict__ perm, float4 *__restrict__ pos,
         float4 *__restrict__ vel, float4 *__restrict__ svel, SimParams P, GridInfo g) {
    const u32 n = nref(nr);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = A[i];
    const u32 id = perm[i];
    const float4 o = pos[id];
    float4 v;
    v.x = __fdiv_rn(__fsub_rn(p.x, o.x), P.timestep);
    v.y = __fdiv_rn(__fsub_rn(p.y, o.y), P.timestep);
    v.z = __fdiv_rn(__fsub_rn(p.z, o.z), P.timestep);
    v.w = 0.0f;
    if (P.restitution >= 0.0f) {             // pbf_options::wall_restitution (not in the reference): reflect off the walls
        if ((p.x <= g.wlo[0] && v.x < 0.0f) || (p.x >= g.whi[0] && v.x > 0.0f)) v.x *= -P.restitution;
        if ((p.y <= g.wlo[1] && v.y < 0.0f) || (p.y >= g.whi[1] && v.y > 0.0f)) v.y *= -P.restitution;
        if ((p.z <= g.wlo[2] && v.z < 0.0f) || (p.z >= g.whi[2] && v.z > 0.0f)) v.z *= -P.restitution;
    }
    pos[id] = make_float4(p.x, p.y, p.z, 0.0f);
    if (VORT) svel[i] = v;
    else vel[id] = v;
    }
}

// ---- K12 highlight.glsl:17-30 (clearhighlight is fused into k_predict) ---------------------------------------------------
__global__ void __launch_bounds__(NB_BLOCK)
k_highlight(NRef nr, const u32 *__restrict__ home, const u32 *__restrict__ perm, const int2 *__restrict__ runs3,
            const int2 *__restrict__ cells, u32 *__restrict__ hl, const u32 *__restrict__ flags, GridInfo g) {
    __shared__ int2 srun[9 * NB_BLOCK];
    if (flags[0] == 0u) return;                      // nobody carries bit 0: the kernel is a no-op
    const u32 n = nref(nr);
    const int tid = threadIdx.x;
    for (u32 i = blockIdx.x * NB_BLOCK + tid; i < n; i += gridDim.x * NB_BLOCK) {
        if ((hl[perm[i]] & 1u) == 0u) continue;
        int slots;
        bool self_in;
        const int cnt = load_runs<NB_BLOCK>(home[i], i, g, runs3, cells, srun, tid, &slots, &self_in);
        for (int o = 0; o < cnt; o++) {
            const int2 r = srun[o * NB_BLOCK + tid];
            for (int j = r.x; j < r.y; j++)
                if (j != (int)i) atomicOr(&hl[perm[j]], 2u);
        }
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

// ---- picking (Selection::GetParticle, src/Selection.cpp:55-85, done as a ray cast) ---------------------------------------
// The reference renders every particle as a sphere of radius 0.1 render units = 0.5 grid units (shaders/selection/
// fragment.glsl, world -> render = 0.2 * pos, shaders/particles/vertex.glsl:44) into an id buffer and reads the pixel under
// the cursor: the id of the nearest sphere on that pixel's ray, -1 for none.  Here: nearest ray/sphere hit over all
// particles, (t, id) packed so that one 64-bit atomicMin keeps the nearest (ties: lowest id).
__global__ void __launch_bounds__(256)
k_pick(u32 n, const float4 *__restrict__ pos, float ox, float oy, float oz, float dx, float dy, float dz, float radius,
       unsigned long long *best) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = ~0ull;
    if (i < n) {
        const float4 p = pos[i];
        const float cx = p.x - ox, cy = p.y - oy, cz = p.z - oz;
        const float b = cx * dx + cy * dy + cz * dz;                       // d is a unit vector
        const float disc = b * b - (cx * cx + cy * cy + cz * cz) + radius * radius;
        if (disc >= 0.0f) {
            const float t = b - sqrtf(disc);                               // front hit
            if (t >= 0.0f) mine = ((unsigned long long)__float_as_uint(t) << 32) | i;
        }
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other < mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0 && mine != ~0ull) atomicMin(best, mine);
}

__global__ void k_toggle_highlight(u32 *hl, u32 id) { hl[id] = hl[id] > 0u ? 0u : 1u; }   // src/Simulation.cpp:182-186

inline int nblocks(size_t n, int b) { return (int)((n + b - 1) / b); }

}  // namespace

int launch_pick(pbf_sim *s, const float o[3], const float d[3], float radius, unsigned long long *best) {
    k_pick<<<nblocks(s->n, 256), 256, 0, s->stream>>>(s->n, s->pos, o[0], o[1], o[2], d[0], d[1], d[2], radius, best);
    return 1;
}

int launch_toggle_highlight(pbf_sim *s, u32 id) {
    k_toggle_highlight<<<1, 1, 0, s->stream>>>(s->hl, id);
    return 1;
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
    P.self_term = s->options.density_self_term;
    P.restitution = s->options.wall_restitution;
    P.canonical = s->canonical ? 1 : 0;
    return P;
}

int launch_fill_tables(pbf_sim *s) {
    k_fill_tables<<<nblocks(s->ncell, 256), 256, 0, s->stream>>>(s->ncell, s->cells, s->runs3);
    return 1;
}

int launch_unclear_cells(pbf_sim *s) {
    if (s->n_prev_dev) {     // slab rank with device-side counts: always the same launch (capturable), the device knows how many
        k_unclear_cells<<<nblocks(s->n ? s->n : 1, 256), 256, 0, s->stream>>>(NRef{s->n, s->n_prev_dev}, s->skey, s->cells, s->runs3, s->grid);
        return 1;
    }
    if (s->n_prev_sorted == 0) return 0;
    k_unclear_cells<<<nblocks(s->n_prev_sorted, 256), 256, 0, s->stream>>>(nref_prev(s), s->skey, s->cells, s->runs3, s->grid);
    return 1;
}

int launch_predict_range(pbf_sim *s, u32 first, NRef count, bool with_hist) {
    if (count.n == 0) return 0;
    k_predict<false, false><<<nblocks(count.n, 256), 256, 0, s->stream>>>(first, count, s->pos, s->vel, s->hl, s->pred, s->keys, s->flags,
                                                                          s->grid, sim_params(s), NRef{0u, nullptr}, nullptr, nullptr, nullptr,
                                                                          LeaveArgs{});
    // digit histograms of all sort passes: the keys just written are still in L2
    return 1 + (with_hist ? launch_sort_hist(s, s->keys + first, count) : 0);
}

// whole-handle predict of the single-domain step: also undoes the previous step's cell-table writes (launch_unclear_cells)
int launch_predict(pbf_sim *s) {
    const u32 np = s->n_prev_sorted, m = s->n > np ? s->n : np;
    k_predict<true, false><<<nblocks(m, 256), 256, 0, s->stream>>>(0u, nref_total(s), s->pos, s->vel, s->hl, s->pred, s->keys, s->flags,
                                                                   s->grid, sim_params(s), nref_prev(s), s->skey, s->cells, s->runs3,
                                                                   LeaveArgs{});
    return 1 + launch_sort_hist(s, s->keys, nref_total(s));
}

// slab rank (device-side counts): predict of the local particles + reset of the previous step's table entries + leavers
int launch_predict_slab(pbf_sim *s, NRef n_local, const LeaveArgs &la) {
    const u32 m = s->n > n_local.n ? s->n : n_local.n;
    k_predict<true, true><<<nblocks(m ? m : 1, 256), 256, 0, s->stream>>>(0u, n_local, s->pos, s->vel, s->hl, s->pred, s->keys, s->flags,
                                                                         s->grid, sim_params(s), NRef{s->n, s->n_prev_dev}, s->skey,
                                                                         s->cells, s->runs3, la);
    return 1;
}

int launch_keys_only(pbf_sim *s, u32 first, u32 count) {
    if (count == 0) return 0;
    k_keys_only<<<nblocks(count, 256), 256, 0, s->stream>>>(first, count, s->pred, s->keys, s->grid);
    return 1;
}

int launch_reorder_cells(pbf_sim *s) {
    k_reorder_cells<<<nblocks(s->n, 256), 256, 0, s->stream>>>(nref_total(s), s->skey, s->perm, s->pred, s->bufA, s->home,
                                                               s->cells, s->grid);
    k_build_runs<<<nblocks(s->n, 256), 256, 0, s->stream>>>(nref_total(s), s->skey, s->cells, s->runs3, s->grid);
    s->n_prev_sorted = s->n;
    return 2 + launch_plan(s);   // staging plan of the tiled sweeps (sweeps.cu), valid until the next sort
}

int launch_highlight(pbf_sim *s) {
    // a few blocks per SM striding over the particles: the usual case is "nothing selected", a no-op worth ~2 us
    int blocks = nblocks(s->n, NB_BLOCK);
    if (blocks > s->sm_count * 4) blocks = s->sm_count * 4;
    k_highlight<<<blocks, NB_BLOCK, 0, s->stream>>>(nref_total(s), s->home, s->perm, s->runs3, s->cells, s->hl, s->flags, s->grid);
    return 1;
}

int launch_update(pbf_sim *s) {
    if (s->params.vorticity_confinement)
        k_update<true><<<nblocks(s->n, 256), 256, 0, s->stream>>>(nref_total(s), s->bufA, s->perm, s->pos, s->vel, s->svel,
                                                                  sim_params(s), s->grid);
    else
        k_update<false><<<nblocks(s->n, 256), 256, 0, s->stream>>>(nref_total(s), s->bufA, s->perm, s->pos, s->vel, s->svel,
                                                                   sim_params(s), s->grid);
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
