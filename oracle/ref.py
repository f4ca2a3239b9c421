"""ctypes front-end of oracle/_ref/libpbf_ref.so: the reference's OWN compute shaders compiled verbatim by g++
(oracle/ref_harness.cpp, oracle/glsl_compat.h, oracle/ref_translate.py).

TEST INFRASTRUCTURE ONLY: this is what pins oracle/pbf_oracle.c -- and through it the CUDA path -- to the reference's source
text.  The library is built from /root/reference where that tree exists (this container); on a GPU box only the prebuilt
.so is used.  `available()` is False when there is neither."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libpbf_ref.so")
REFERENCE = os.environ.get("PBF_REFERENCE_TREE", "/root/reference")

ORDER_JACOBI, ORDER_AS_DISPATCHED = 0, 1


def build():
    """Compiles the reference's shaders (only possible where the reference tree is); returns the .so path or None."""
    if os.path.isdir(os.path.join(REFERENCE, "shaders", "sph")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REF=" + REFERENCE])
    return _SO if os.path.exists(_SO) else None


def available():
    return build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libpbf_ref.so is missing and %s is not there to build it from" % REFERENCE)
        L = C.CDLL(so)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int]
        for name in ("ref_destroy", "ref_predict", "ref_sort", "ref_find_cells", "ref_policy_define_last_end",
                     "ref_neighbour_cells", "ref_highlight", "ref_calclambda", "ref_update"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_updatepos.argtypes = [C.c_void_p, C.c_int]
        L.ref_vorticity.argtypes = [C.c_void_p, C.c_int]
        L.ref_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_params.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_extforce.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_oob_fetch.argtypes = [C.c_void_p, C.c_int]
        L.ref_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_numbits.restype = C.c_uint
        L.ref_numbits.argtypes = [C.c_void_p]
        L.ref_blocksum_levels.argtypes = [C.c_void_p]
        for name in ("ref_get_records", "ref_set_records", "ref_get_lambda", "ref_get_vorticity", "ref_get_neighbours"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefSim:
    """The reference's SPH object (src/SPH.h) on the CPU: same constructor arguments, same Run() sequence, its shaders."""

    def __init__(self, n, grid=(128, 64, 128)):
        self.n, self.grid = n, tuple(grid)
        self.h = C.c_void_p(lib().ref_create(n, *grid))
        if not self.h:
            raise ValueError("the reference needs a particle count that is a multiple of 512")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def set_params(self, P):
        """P: anything with the eight sphparams_t fields (e.g. oracle.Params)."""
        names = ("one_over_rho_0", "epsilon", "gravity", "timestep", "tensile_instability_k", "tensile_instability_scale",
                 "xsph_viscosity_c", "vorticity_epsilon")
        a = np.array([getattr(P, k) for k in names], np.float32)
        lib().ref_set_params(self.h, _p(a))

    def get_params(self):
        a = np.zeros(8, np.float32)
        lib().ref_get_params(self.h, _p(a))
        return a

    def set_extforce(self, on): lib().ref_set_extforce(self.h, int(bool(on)))
    def set_oob_fetch(self, v): lib().ref_set_oob_fetch(self.h, int(v))

    def upload(self, pos, vel, highlight=None):
        pos, vel = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(vel, np.float32)
        hl = None if highlight is None else np.ascontiguousarray(highlight, np.uint32)
        lib().ref_upload(self.h, _p(pos), _p(vel), _p(hl))

    def download(self):
        pos, vel, hl = np.empty((self.n, 4), np.float32), np.empty((self.n, 4), np.float32), np.empty(self.n, np.uint32)
        lib().ref_download(self.h, _p(pos), _p(vel), _p(hl))
        return pos, vel, hl

    def predict(self): lib().ref_predict(self.h)
    def sort(self): lib().ref_sort(self.h)
    def find_cells(self): lib().ref_find_cells(self.h)
    def policy_define_last_end(self): lib().ref_policy_define_last_end(self.h)
    def neighbour_cells(self): lib().ref_neighbour_cells(self.h)
    def highlight(self): lib().ref_highlight(self.h)
    def calclambda(self): lib().ref_calclambda(self.h)
    def updatepos(self, order=ORDER_JACOBI): lib().ref_updatepos(self.h, order)
    def update(self): lib().ref_update(self.h)
    def vorticity(self, order=ORDER_JACOBI): lib().ref_vorticity(self.h, order)

    def step(self, iterations, vorticity=False, order=ORDER_JACOBI, define_last_end=True):
        lib().ref_step(self.h, int(iterations), int(vorticity), int(order), int(define_last_end))

    @property
    def numbits(self): return lib().ref_numbits(self.h)
    @property
    def blocksum_levels(self): return lib().ref_blocksum_levels(self.h)

    def records(self):
        out = np.empty((self.n, 4), np.float32)
        lib().ref_get_records(self.h, _p(out))
        return out

    def set_records(self, rec):
        lib().ref_set_records(self.h, _p(np.ascontiguousarray(rec, np.float32)))

    def lam(self):
        out = np.empty(self.n, np.float32)
        lib().ref_get_lambda(self.h, _p(out))
        return out

    def vort(self):
        out = np.empty(self.n, np.float32)
        lib().ref_get_vorticity(self.h, _p(out))
        return out

    def neighbours(self):
        """The neighbour buffer as the reference stores it: per particle 3 x ivec4 = 12 words, 9 packed runs + 3 zeros."""
        out = np.empty((self.n, 12), np.int32)
        lib().ref_get_neighbours(self.h, _p(out))
        return out

    def packed_runs(self):
        w = self.neighbours()
        return w[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]], w[:, [3, 7, 11]]

    def grid_tables(self):
        nc = self.grid[0] * self.grid[1] * self.grid[2]
        start, end = np.empty(nc, np.int32), np.empty(nc, np.int32)
        lib().ref_get_grid(self.h, _p(start), _p(end))
        return start, end
