"""ctypes front-end of the CPU oracle (oracle/pbf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (pbf_b200) never imports this module.
PARITY PINNED against the reference's own shaders compiled by g++ (oracle/ref.py, tests/test_oracle_ref.py); see the
header of pbf_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpbf_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pbf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Grid(C.Structure):
    _fields_ = [("gx", C.c_int), ("gy", C.c_int), ("gz", C.c_int),
                ("wall_x", C.c_float), ("wall_y", C.c_float), ("wall_z", C.c_float),
                ("ref_quirks", C.c_int)]


class Params(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("one_over_rho_0", "epsilon", "gravity", "timestep",
                                         "tensile_instability_k", "tensile_instability_scale",
                                         "xsph_viscosity_c", "vorticity_epsilon")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ora_wpoly6.restype = C.c_float
        L.ora_wpoly6.argtypes = [C.c_float, C.c_float]
        L.ora_sortbits.restype = C.c_int
        L.ora_num_threads.restype = C.c_int
        L.ora_key_of.restype = C.c_uint32
        L.ora_pack_run.restype = C.c_int32
        L.ora_kinetic_energy.restype = C.c_double
        L.ora_density_error.restype = C.c_double
        L.ora_sim_create.restype = C.c_void_p
        for name in ("sorted", "predicted", "lambda", "rho", "vorticity", "start", "end",
                     "run_start", "run_count", "skey"):
            getattr(L, "ora_sim_" + name).restype = C.c_void_p
            getattr(L, "ora_sim_" + name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_params():
    p = Params()
    lib().ora_default_params(C.byref(p))
    return p


def make_grid(gx=128, gy=64, gz=128, wall=(16.0, 0.0, 16.0), ref_quirks=1):
    return Grid(gx, gy, gz, wall[0], wall[1], wall[2], ref_quirks)


def dam_break(nx, ny, nz, origin=(32.5, 0.5, 32.5), spacing=0.94, mirror=False, seed=12345, id0=0):
    n = nx * ny * nz
    pos = np.zeros((n, 4), np.float32)
    vel = np.zeros((n, 4), np.float32)
    lib().ora_dam_break(nx, ny, nz, C.c_float(origin[0]), C.c_float(origin[1]), C.c_float(origin[2]),
                        C.c_float(spacing), int(mirror), C.c_uint32(seed), C.c_uint32(id0), _p(pos), _p(vel))
    return pos, vel


def predict(pos, vel, params, grid, extforce=False):
    n = pos.shape[0]
    rec = np.empty((n, 4), np.float32)
    lib().ora_predict(n, _p(pos), _p(vel), C.byref(params), C.byref(grid), int(extforce), _p(rec))
    return rec


def keys(rec, grid):
    n = rec.shape[0]
    k = np.empty(n, np.uint32)
    lib().ora_keys(n, _p(rec), C.byref(grid), _p(k))
    return k


def sortbits(grid):
    return lib().ora_sortbits(grid.gx, grid.gy, grid.gz)


def sort(rec, grid):
    n = rec.shape[0]
    out = np.empty((n, 4), np.float32)
    k = np.empty(n, np.uint32)
    lib().ora_sort(n, _p(rec), C.byref(grid), _p(out), _p(k))
    return out, k


def findcells(rec_sorted, grid, end=None):
    ncell = grid.gx * grid.gy * grid.gz
    start = np.empty(ncell, np.int32)
    if end is None:
        end = np.zeros(ncell, np.int32)
    lib().ora_findcells(rec_sorted.shape[0], _p(rec_sorted), C.byref(grid), _p(start), _p(end))
    return start, end


def neighbourcells(rec_sorted, grid, start, end):
    n = rec_sorted.shape[0]
    rs = np.empty((n, 9), np.int32)
    rc = np.empty((n, 9), np.int32)
    lib().ora_neighbourcells(n, _p(rec_sorted), C.byref(grid), _p(start), _p(end), _p(rs), _p(rc))
    return rs, rc


def calclambda(rec_sorted, rs, rc, params):
    n = rec_sorted.shape[0]
    lam = np.empty(n, np.float32)
    rho = np.empty(n, np.float32)
    lib().ora_calclambda(n, _p(rec_sorted), _p(rs), _p(rc), C.byref(params), _p(lam), _p(rho))
    return lam, rho


def updatepos(rec_sorted, rs, rc, lam, params, grid):
    n = rec_sorted.shape[0]
    out = np.empty((n, 4), np.float32)
    lib().ora_updatepos(n, _p(rec_sorted), _p(rs), _p(rc), _p(lam), C.byref(params), C.byref(grid), _p(out))
    return out


def update(rec_sorted, params, pos, vel):
    """In place on pos, vel (by id)."""
    lib().ora_update(rec_sorted.shape[0], _p(rec_sorted), C.byref(params), _p(pos), _p(vel))


def vorticity(rec_sorted, rs, rc, params, vel):
    """In place on vel (by id); returns |omega| per sorted index."""
    n = rec_sorted.shape[0]
    w = np.empty(n, np.float32)
    lib().ora_vorticity(n, _p(rec_sorted), _p(rs), _p(rc), C.byref(params), _p(vel), _p(w))
    return w


def highlight(rec_sorted, rs, rc, hl):
    lib().ora_highlight(rec_sorted.shape[0], _p(rec_sorted), _p(rs), _p(rc), _p(hl))


def kinetic_energy(vel):
    return lib().ora_kinetic_energy(vel.shape[0], _p(vel))


def density_error(rho, params):
    return lib().ora_density_error(rho.shape[0], _p(rho), C.byref(params))


class Sim:
    """Whole-step oracle (SPH::Run, src/SPH.cpp:246-334) with persistent scratch."""

    def __init__(self, n, grid):
        self.n, self.grid = n, grid
        self.h = C.c_void_p(lib().ora_sim_create(n, C.byref(grid)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_sim_destroy(self.h)
            self.h = None

    def step(self, pos, vel, params, iterations, vorticity=False, extforce=False, highlight=None):
        lib().ora_sim_step(self.h, C.byref(params), int(iterations), int(vorticity), int(extforce),
                           _p(pos), _p(vel), _p(highlight) if highlight is not None else None)

    def _view(self, name, dtype, shape):
        ptr = getattr(lib(), "ora_sim_" + name)(self.h)
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = (C.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    sorted = property(lambda s: s._view("sorted", np.float32, (s.n, 4)))
    predicted = property(lambda s: s._view("predicted", np.float32, (s.n, 4)))
    lam = property(lambda s: s._view("lambda", np.float32, (s.n,)))
    rho = property(lambda s: s._view("rho", np.float32, (s.n,)))
    vort = property(lambda s: s._view("vorticity", np.float32, (s.n,)))
    skey = property(lambda s: s._view("skey", np.uint32, (s.n,)))
    run_start = property(lambda s: s._view("run_start", np.int32, (s.n, 9)))
    run_count = property(lambda s: s._view("run_count", np.int32, (s.n, 9)))

    @property
    def start(self):
        g = self.grid
        return self._view("start", np.int32, (g.gx * g.gy * g.gz,))

    @property
    def end(self):
        g = self.grid
        return self._view("end", np.int32, (g.gx * g.gy * g.gz,))


def num_threads():
    return lib().ora_num_threads()


def set_num_threads(n):
    lib().ora_set_num_threads(int(n))
