"""pbf_b200 -- B200-native per-timestep Position Based Fluids (the SPH::Run path of ekpyron/pbf).

This package is a thin ctypes binding of the C ABI in include/pbf_c.h (libpbf_b200.so, hand-written sm_100a
CUDA).  `SPH` mirrors the reference class of the same name (reference src/SPH.h:33-447): same method names,
argument meaning and error behaviour (errors raise RuntimeError, as the reference throws std::runtime_error).
There is no CPU fallback: without the compiled extension or without a B200 every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBF_B200_LIB: another build of the same library (kernel experiments); there is no other implementation to fall back to
LIB_PATH = os.environ.get("PBF_B200_LIB") or os.path.join(_HERE, "libpbf_b200.so")

PBF_KEY_NOCELL = 0x80000000


class Config(C.Structure):
    _fields_ = [("num_particles", C.c_uint32), ("capacity", C.c_uint32), ("grid", C.c_int32 * 3),
                ("wall", C.c_float * 3), ("ref_quirks", C.c_int32), ("device", C.c_int32), ("use_graph", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("one_over_rho_0", C.c_float), ("epsilon", C.c_float), ("gravity", C.c_float),
                ("timestep", C.c_float), ("tensile_instability_k", C.c_float),
                ("tensile_instability_scale", C.c_float), ("xsph_viscosity_c", C.c_float),
                ("vorticity_epsilon", C.c_float), ("num_solver_iterations", C.c_int32),
                ("vorticity_confinement", C.c_int32), ("external_force", C.c_int32)]


class Options(C.Structure):
    """pbf_options (include/pbf_c.h): opt-in corrections of reference defects, all off by default."""
    _fields_ = [("density_self_term", C.c_int32), ("wall_restitution", C.c_float), ("full_support", C.c_int32)]


class StateInfo(C.Structure):
    """pbf_state_info (include/pbf_c.h): what a state file's header holds."""
    _fields_ = [("num_particles", C.c_uint32), ("grid", C.c_int32 * 3), ("wall", C.c_float * 3),
                ("ref_quirks", C.c_int32), ("params", Params), ("steps", C.c_uint64), ("options", Options)]


# pbf_map_fn / pbf_unmap_fn (include/pbf_c.h): the owner of the by-id buffers hands them out between map and unmap
MAP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))
UNMAP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)

_lib = None


def lib():
    """Loads libpbf_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("pbf_b200: %s is missing -- run `python -m pbf_b200.build` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pbf_last_error.restype = C.c_char_p
        L.pbf_wpoly6.restype = C.c_float
        L.pbf_wpoly6.argtypes = [C.c_float, C.c_float]
        L.pbf_num_particles.restype = C.c_uint32
        L.pbf_num_particles.argtypes = [C.c_void_p]
        L.pbf_get_tile_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.pbf_kernel_launches.restype = C.c_uint64
        L.pbf_kernel_launches.argtypes = [C.c_void_p]
        L.pbf_stream.restype = C.c_void_p
        L.pbf_stream.argtypes = [C.c_void_p]
        L.pbf_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        for name in ("pbf_destroy", "pbf_sync", "pbf_predict", "pbf_sort", "pbf_build_cells", "pbf_highlight",
                     "pbf_calc_lambda", "pbf_update_positions", "pbf_finalize", "pbf_vorticity"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.pbf_step.argtypes = [C.c_void_p, C.c_int]
        L.pbf_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.pbf_set_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        L.pbf_set_canonical_order.argtypes = [C.c_void_p, C.c_int]
        L.pbf_get_options.argtypes = [C.c_void_p, C.POINTER(Options)]
        L.pbf_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.pbf_get_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.pbf_upload_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pbf_download_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_device_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.pbf_bind_device_buffers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.pbf_get_predicted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_get_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_get_cell_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_get_neighbour_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_get_lambda.argtypes = [C.c_void_p, C.c_void_p]
        L.pbf_get_vorticity.argtypes = [C.c_void_p, C.c_void_p]
        L.pbf_enable_timing.argtypes = [C.c_void_p, C.c_int]
        L.pbf_get_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.pbf_get_solver_kernel_timings.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.pbf_pick_particle.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_int32)]
        L.pbf_toggle_highlight.argtypes = [C.c_void_p, C.c_uint32]
        L.pbf_get_diagnostics.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pbf_scene_dam_break.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float, C.c_int,
                                          C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.pbf_sort_bits.argtypes = [C.POINTER(C.c_int32)]
        L.pbf_sort_passes.argtypes = [C.POINTER(C.c_int32)]
        L.pbf_register_gl_buffers.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
        L.pbf_unregister_gl_buffers.argtypes = [C.c_void_p]
        L.pbf_register_external_buffers.argtypes = [C.c_void_p, MAP_FN, UNMAP_FN, C.c_void_p]
        L.pbf_unregister_external_buffers.argtypes = [C.c_void_p]
        L.pbf_state_file_write.argtypes = [C.c_char_p, C.POINTER(StateInfo), C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_state_file_info.argtypes = [C.c_char_p, C.POINTER(StateInfo)]
        L.pbf_state_file_read.argtypes = [C.c_char_p, C.POINTER(StateInfo), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pbf_save_state.argtypes = [C.c_void_p, C.c_char_p]
        L.pbf_load_state.argtypes = [C.c_void_p, C.c_char_p]
        L.pbf_step_count.restype = C.c_uint64
        L.pbf_step_count.argtypes = [C.c_void_p]
        L.pbf_slab_unique_id.argtypes = [C.c_void_p]
        L.pbf_slab_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.pbf_slab_init_group.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_uint32]
        L.pbf_slab_p2p_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.pbf_slab_p2p_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pbf_slab_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pbf_slab_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.pbf_slab_download_highlight.argtypes = [C.c_void_p, C.c_void_p]
        L.pbf_slab_layer_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.pbf_slab_set_planes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.pbf_slab_phase_times.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.pbf_slab_step.argtypes = [C.c_void_p, C.c_int]
        L.pbf_slab_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_uint32), C.c_int]
        L.pbf_slab_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("pbf_b200 error %d: %s" % (rc, lib().pbf_last_error().decode()))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):      # torch tensor (host pinned or device, as the entry point requires)
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def default_params():
    p = Params()
    lib().pbf_default_params(C.byref(p))
    return p


def wpoly6(r, h):
    """SPH::Wpoly6 (reference src/SPH.cpp:159-164)."""
    return lib().pbf_wpoly6(r, h)


def sort_bits(grid):
    return lib().pbf_sort_bits((C.c_int32 * 3)(*grid))


def sort_passes(grid):
    """Onesweep passes the library needs for this grid's key bits (the reference needs sort_bits / 2 two-bit passes)."""
    return lib().pbf_sort_passes((C.c_int32 * 3)(*grid))


def write_state_file(path, pos, vel=None, highlight=None, grid=(128, 64, 128), wall=(16.0, 0.0, 16.0), ref_quirks=True,
                     params=None, steps=0, options=None):
    """Writes a state file from HOST arrays (no device needed); see include/pbf_c.h, "state files"."""
    pos = np.ascontiguousarray(pos, np.float32)
    vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
    highlight = None if highlight is None else np.ascontiguousarray(highlight, np.uint32)
    info = StateInfo(pos.shape[0], (C.c_int32 * 3)(*grid), (C.c_float * 3)(*wall), int(ref_quirks),
                     params if params is not None else default_params(), steps,
                     options if options is not None else Options(0, -1.0, 0))
    _check(lib().pbf_state_file_write(os.fsencode(path), C.byref(info), _ptr(pos), _ptr(vel), _ptr(highlight)))


def state_file_info(path):
    info = StateInfo()
    _check(lib().pbf_state_file_info(os.fsencode(path), C.byref(info)))
    return info


def read_state_file(path):
    """-> (info, pos, vel, highlight) as HOST arrays; raises on a bad magic, version, size or checksum."""
    info = state_file_info(path)
    n = info.num_particles
    pos, vel, hl = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32), np.empty(n, np.uint32)
    _check(lib().pbf_state_file_read(os.fsencode(path), C.byref(info), _ptr(pos), _ptr(vel), _ptr(hl), n))
    return info, pos, vel, hl


def dam_break(nx, ny, nz, origin=(32.5, 0.5, 32.5), spacing=0.94, mirror=False, seed=12345, id0=0):
    """Seeded block fill of Simulation::ResetParticleBuffer (reference src/Simulation.cpp:216-230)."""
    n = nx * ny * nz
    pos = np.empty((n, 4), np.float32)
    vel = np.empty((n, 4), np.float32)
    _check(lib().pbf_scene_dam_break(nx, ny, nz, (C.c_float * 3)(*origin), spacing, int(mirror), seed, id0,
                                     _ptr(pos), _ptr(vel)))
    return pos, vel


class SPH:
    """Mirror of the reference's SPH class (src/SPH.h): SPH(numparticles, gridsize) / Run() / parameter accessors.

    Buffers returned by GetPositionBuffer/GetVelocityBuffer/GetHighlightBuffer are DEVICE addresses (ints) of the
    by-id N x float4 / N x uint32 arrays -- the CUDA counterpart of the GL buffer names the reference returns.
    """

    def __init__(self, numparticles, gridsize=(128, 64, 128), wall=(16.0, 0.0, 16.0), ref_quirks=True,
                 device=-1, use_graph=True, capacity=0):
        self._h = C.c_void_p()
        cfg = Config(numparticles, capacity, (C.c_int32 * 3)(*gridsize), (C.c_float * 3)(*wall),
                     int(ref_quirks), device, int(use_graph))
        _check(lib().pbf_create(C.byref(cfg), C.byref(self._h)))
        self.numparticles = numparticles
        self.gridsize = tuple(gridsize)
        self.ncell = gridsize[0] * gridsize[1] * gridsize[2]

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:   # _lib is gone at interpreter exit
            _lib.pbf_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # --- parameters (src/SPH.h:57-222) ---------------------------------------------------------------
    def _get(self):
        p = Params()
        _check(lib().pbf_get_params(self._h, C.byref(p)))
        return p

    def _set(self, **kw):
        p = self._get()
        for k, v in kw.items():
            setattr(p, k, v)
        _check(lib().pbf_set_params(self._h, C.byref(p)))

    def set_params(self, p):
        _check(lib().pbf_set_params(self._h, C.byref(p)))

    def set_options(self, density_self_term=None, wall_restitution=None, full_support=None):
        """Opt-in corrections (not in the reference): self term in the density, velocity reflection at the walls, neighbour
        search over the whole kernel support (5 x 5 x 5 cells instead of 3 x 3 x 3)."""
        o = Options()
        _check(lib().pbf_get_options(self._h, C.byref(o)))
        if density_self_term is not None:
            o.density_self_term = int(bool(density_self_term))
        if wall_restitution is not None:
            o.wall_restitution = wall_restitution
        if full_support is not None:
            o.full_support = int(bool(full_support))
        _check(lib().pbf_set_options(self._h, C.byref(o)))

    def set_canonical_order(self, on=True):
        """Verification mode: one summation order on every code path, so that slab runs equal the single-domain run bit for bit."""
        _check(lib().pbf_set_canonical_order(self._h, int(bool(on))))

    def get_options(self):
        o = Options()
        _check(lib().pbf_get_options(self._h, C.byref(o)))
        return o

    def GetRestDensity(self): return 1.0 / self._get().one_over_rho_0
    def SetRestDensity(self, rho): self._set(one_over_rho_0=np.float32(1.0) / np.float32(rho))
    def GetCFMEpsilon(self): return self._get().epsilon
    def SetCFMEpsilon(self, v): self._set(epsilon=v)
    def GetGravity(self): return self._get().gravity
    def SetGravity(self, v): self._set(gravity=v)
    def GetTimestep(self): return self._get().timestep
    def SetTimestep(self, v): self._set(timestep=v)
    def GetTensileInstabilityK(self): return self._get().tensile_instability_k
    def SetTensileInstabilityK(self, v): self._set(tensile_instability_k=v)
    def GetTensileInstabilityScale(self): return self._get().tensile_instability_scale
    def SetTensileInstabilityScale(self, v): self._set(tensile_instability_scale=v)
    def GetXSPHViscosity(self): return self._get().xsph_viscosity_c
    def SetXSPHViscosity(self, v): self._set(xsph_viscosity_c=v)
    def GetVorticityEpsilon(self): return self._get().vorticity_epsilon
    def SetVorticityEpsilon(self, v): self._set(vorticity_epsilon=v)
    def GetNumSolverIterations(self): return self._get().num_solver_iterations
    def SetNumSolverIterations(self, k): self._set(num_solver_iterations=int(k))
    def IsVorticityConfinementEnabled(self): return bool(self._get().vorticity_confinement)
    def SetVorticityConfinementEnabled(self, flag): self._set(vorticity_confinement=int(bool(flag)))
    def SetExternalForce(self, state): self._set(external_force=int(bool(state)))
    Wpoly6 = staticmethod(wpoly6)

    # --- buffers --------------------------------------------------------------------------------------
    def _bufs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().pbf_device_buffers(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def GetPositionBuffer(self): return self._bufs()[0]
    def GetVelocityBuffer(self): return self._bufs()[1]
    def GetHighlightBuffer(self): return self._bufs()[2]

    def bind_device_buffers(self, pos=None, vel=None, highlight=None):
        _check(lib().pbf_bind_device_buffers(self._h, _ptr(pos), _ptr(vel), _ptr(highlight)))

    def upload(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        _check(lib().pbf_upload_state(self._h, _ptr(pos), _ptr(vel), pos.shape[0]))

    def download(self, highlight=False):
        n = self.numparticles
        pos = np.empty((n, 4), np.float32)
        vel = np.empty((n, 4), np.float32)
        hl = np.empty(n, np.uint32) if highlight else None
        _check(lib().pbf_download_state(self._h, _ptr(pos), _ptr(vel), _ptr(hl)))
        return (pos, vel, hl) if highlight else (pos, vel)

    # --- stepping ---------------------------------------------------------------------------------------
    def Run(self, nsteps=1):
        """SPH::Run (src/SPH.cpp:246-334)."""
        _check(lib().pbf_step(self._h, nsteps))

    def step_host(self, pos, vel, nsteps=1):
        """End-to-end call: HOST pos/vel in, one step, HOST pos/vel out (in place)."""
        _check(lib().pbf_step_host(self._h, _ptr(pos), _ptr(vel), nsteps))

    def save_state(self, path):
        """Current by-id state, parameters and step counter -> state file."""
        _check(lib().pbf_save_state(self._h, os.fsencode(path)))

    def load_state(self, path):
        _check(lib().pbf_load_state(self._h, os.fsencode(path)))

    @classmethod
    def from_state_file(cls, path, **kw):
        """A new SPH sized from the file's header, holding the file's state (resume)."""
        info = state_file_info(path)
        sph = cls(info.num_particles, tuple(info.grid), wall=tuple(info.wall), ref_quirks=bool(info.ref_quirks), **kw)
        sph.load_state(path)
        return sph

    @property
    def step_count(self): return lib().pbf_step_count(self._h)

    def register_gl_buffers(self, pos, vel, highlight):
        """CUDA-GL interop: run on the renderer's GL buffer objects (needs a current GL context)."""
        _check(lib().pbf_register_gl_buffers(self._h, pos, vel, highlight))

    def unregister_gl_buffers(self): _check(lib().pbf_unregister_gl_buffers(self._h))

    def register_external_buffers(self, map_fn, unmap_fn):
        """The map/unmap protocol of the GL interop for any other owner of the by-id buffers: map_fn(stream) returns the
        three device addresses (pos, vel, highlight), valid until unmap_fn(stream)."""
        def _map(user, stream, p, v, h):
            try:
                a, b, c = map_fn(stream)
                p[0], v[0], h[0] = a, b, c
                return 0
            except Exception:
                return 1

        def _unmap(user, stream):
            try:
                unmap_fn(stream)
                return 0
            except Exception:
                return 1
        self._ext_cb = (MAP_FN(_map), UNMAP_FN(_unmap))      # keep the thunks alive
        _check(lib().pbf_register_external_buffers(self._h, self._ext_cb[0], self._ext_cb[1], None))

    def unregister_external_buffers(self):
        _check(lib().pbf_unregister_external_buffers(self._h))
        self._ext_cb = None

    def sync(self): _check(lib().pbf_sync(self._h))
    def predict(self): _check(lib().pbf_predict(self._h))
    def sort(self): _check(lib().pbf_sort(self._h))
    def build_cells(self): _check(lib().pbf_build_cells(self._h))
    def highlight(self): _check(lib().pbf_highlight(self._h))
    def calc_lambda(self): _check(lib().pbf_calc_lambda(self._h))
    def update_positions(self): _check(lib().pbf_update_positions(self._h))
    def finalize(self): _check(lib().pbf_finalize(self._h))
    def vorticity(self): _check(lib().pbf_vorticity(self._h))

    def sort_pairs(self, keys_in, vals_in, keys_out, vals_out, n, bits):
        _check(lib().pbf_sort_pairs(self._h, _ptr(keys_in), _ptr(vals_in), _ptr(keys_out), _ptr(vals_out), n, bits))

    # --- read-back ------------------------------------------------------------------------------------------
    def get_predicted(self):
        n = self.numparticles
        rec = np.empty((n, 4), np.float32)
        keys = np.empty(n, np.uint32)
        _check(lib().pbf_get_predicted(self._h, _ptr(rec), _ptr(keys)))
        return rec, keys

    def get_sorted(self, records=True):
        n = self.numparticles
        keys = np.empty(n, np.uint32)
        perm = np.empty(n, np.uint32)
        rec = np.empty((n, 4), np.float32) if records else None
        _check(lib().pbf_get_sorted(self._h, _ptr(keys), _ptr(perm), _ptr(rec)))
        return keys, perm, rec

    def get_cell_ranges(self):
        start = np.empty(self.ncell, np.int32)
        end = np.empty(self.ncell, np.int32)
        _check(lib().pbf_get_cell_ranges(self._h, _ptr(start), _ptr(end)))
        return start, end

    def get_neighbour_runs(self):
        n = self.numparticles
        rs = np.empty((n, 9), np.int32)
        rc = np.empty((n, 9), np.int32)
        _check(lib().pbf_get_neighbour_runs(self._h, _ptr(rs), _ptr(rc)))
        return rs, rc

    def get_lambda(self):
        out = np.empty(self.numparticles, np.float32)
        _check(lib().pbf_get_lambda(self._h, _ptr(out)))
        return out

    def get_vorticity(self):
        out = np.empty(self.numparticles, np.float32)
        _check(lib().pbf_get_vorticity(self._h, _ptr(out)))
        return out

    def enable_timing(self, on=True): _check(lib().pbf_enable_timing(self._h, int(on)))

    def get_timings(self):
        ms = (C.c_float * 5)()
        _check(lib().pbf_get_timings(self._h, ms))
        return list(ms)

    def get_solver_kernel_timings(self):
        """(ms per calclambda launch, ms per updatepos launch), averaged over the last timed step's iterations."""
        a, b = C.c_float(), C.c_float()
        _check(lib().pbf_get_solver_kernel_timings(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def OutputTiming(self):
        """SPH::OutputTiming (src/SPH.cpp:218-240): same five phase labels."""
        names = ("Position prediction", "Sorting", "Neighbour cell search", "Solver", "Vorticity confinement")
        for name, ms in zip(names, self.get_timings()):
            print("%s: %g ms" % (name, ms))

    def pick_particle(self, origin, direction, radius=0.5):
        """Selection::GetParticle as a ray cast: id of the nearest particle sphere hit by the ray, or -1."""
        out = C.c_int32()
        _check(lib().pbf_pick_particle(self._h, (C.c_float * 3)(*origin), (C.c_float * 3)(*direction), radius, C.byref(out)))
        return out.value

    def toggle_highlight(self, particle_id):
        """The highlight-word update of Simulation::OnMouseDown (src/Simulation.cpp:182-186)."""
        _check(lib().pbf_toggle_highlight(self._h, particle_id))

    def diagnostics(self, density=True, kinetic=True):
        d, k = C.c_double(), C.c_double()
        _check(lib().pbf_get_diagnostics(self._h, C.byref(d) if density else None, C.byref(k) if kinetic else None))
        return d.value, k.value

    @property
    def kernel_launches(self): return lib().pbf_kernel_launches(self._h)

    def tile_stats(self):
        """(tiles, tiles on the shared-memory tiled sweep path) of the last build_cells / step."""
        a, b, why = C.c_uint32(0), C.c_uint32(0), (C.c_uint32 * 8)()
        _check(lib().pbf_get_tile_stats(self._h, C.byref(a), C.byref(b), why))
        self.tile_fallback_reasons = list(why)
        return a.value, b.value
    @property
    def stream(self): return lib().pbf_stream(self._h)
