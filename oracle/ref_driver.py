"""oracle/ref_driver.py -- TEST INFRASTRUCTURE (never imported by the product path).

ctypes driver for the UNMODIFIED reference built by oracle/Makefile.ref into
oracle/_ref/libliggghts_ref.so.  It speaks the reference's own C API
(/root/reference/src/library.h:59-74) plus the read-only accessors of
oracle/ref_shim.cpp.  Used (a) in this container to generate tests/golden/*.npz and to pin
oracle/dem_oracle.c, (b) on the GPU box as the `--impl reference` CPU arm of bench.py.
The reference calls exit(1) on errors (error.cpp:160-186), so decks are trusted input.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libliggghts_ref.so")


def available():
    return os.path.exists(LIB)


class Ref:
    def __init__(self, log=None, lib=None, extra_args=()):
        """lib / extra_args: another build of the reference tree and more command-line switches (the `-suffix b200` shim
        builds of integration/Makefile)"""
        self.lib = C.CDLL(lib or LIB)
        L = self.lib
        L.lammps_open_no_mpi.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p)]
        L.lammps_close.argtypes = [C.c_void_p]
        L.lammps_command.argtypes = [C.c_void_p, C.c_char_p]
        L.lammps_command.restype = C.c_void_p
        L.lammps_extract_atom.argtypes = [C.c_void_p, C.c_char_p]
        L.lammps_extract_atom.restype = C.c_void_p
        L.lammps_extract_fix.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.lammps_extract_fix.restype = C.c_void_p
        L.lammps_get_natoms.argtypes = [C.c_void_p]
        for f in ("ref_nlocal", "ref_nghost", "ref_neigh_ncalls"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_ntimestep.argtypes = [C.c_void_p]
        L.ref_ntimestep.restype = C.c_long
        L.ref_pairlist_count.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_pairlist_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        args = [b"liggghts", b"-screen", b"/dev/null", b"-log", (log or "none").encode(), b"-echo", b"none"] + [a.encode() for a in extra_args]
        argv = (C.c_char_p * len(args))(*args)
        self.h = C.c_void_p()
        L.lammps_open_no_mpi(len(args), argv, C.byref(self.h))

    def close(self):
        if self.h:
            self.lib.lammps_close(self.h)
            self.h = None

    def cmd(self, text):
        for line in text.strip().splitlines():
            line = line.strip()
            if line and not line.startswith("#"):
                self.lib.lammps_command(self.h, line.encode())

    @property
    def nlocal(self):
        return self.lib.ref_nlocal(self.h)

    @property
    def neigh_builds(self):
        return self.lib.ref_neigh_ncalls(self.h)

    def _vec(self, name, n, ctype, cols):
        p = self.lib.lammps_extract_atom(self.h, name.encode())
        if cols == 1:
            arr = C.cast(p, C.POINTER(ctype))
            return np.array([arr[i] for i in range(n)])
        rows = C.cast(p, C.POINTER(C.POINTER(ctype)))
        # rows of a contiguous block (memory.h create2d): copy through the first row pointer
        base = C.cast(rows[0], C.POINTER(ctype * (n * cols)))
        return np.array(base.contents, dtype=np.float64).reshape(n, cols).copy()

    def atoms(self):
        """per-particle state sorted by tag"""
        n = self.nlocal
        out = {"tag": self._vec("id", n, C.c_int, 1).astype(np.int32)}
        out["type"] = self._vec("type", n, C.c_int, 1).astype(np.int32)
        for k in ("x", "v", "f", "omega", "torque"):
            out[k] = self._vec(k, n, C.c_double, 3)
        for k in ("radius", "rmass", "density"):
            out[k] = self._vec(k, n, C.c_double, 1).astype(np.float64)
        o = np.argsort(out["tag"], kind="stable")
        return {k: v[o] for k, v in out.items()}

    def fix_peratom_array(self, fix_id, ncols):
        """per-atom array of a fix (e.g. primitive-wall history 'history_<wallid>'), by tag"""
        n = self.nlocal
        tag = self._vec("id", n, C.c_int, 1)
        p = self.lib.lammps_extract_fix(self.h, fix_id.encode(), 1, 2, 0, 0)
        rows = C.cast(p, C.POINTER(C.POINTER(C.c_double)))
        a = np.array([[rows[i][c] for c in range(ncols)] for i in range(n)], dtype=np.float64).reshape(n, ncols)
        return a[np.argsort(tag, kind="stable")]

    def mesh_topology(self, mesh_id):
        L = self.lib
        L.ref_mesh_ntri.argtypes = [C.c_void_p, C.c_char_p]
        n = L.ref_mesh_ntri(self.h, mesh_id.encode())
        nodes = np.zeros((n, 3, 3)); ea = np.zeros((n, 3), np.int32); ca = np.zeros((n, 3), np.int32); nn = np.zeros(n, np.int32)
        L.ref_mesh_topology.argtypes = [C.c_void_p, C.c_char_p] + [C.c_void_p] * 4
        L.ref_mesh_topology(self.h, mesh_id.encode(), nodes.ctypes.data, ea.ctypes.data, ca.ctypes.data, nn.ctypes.data)
        return {"nodes": nodes, "edge_active": ea, "corner_active": ca, "nneighs": nn}

    def fix_vector(self, fix_id, n):
        """global vector of a fix (Fix::compute_vector(0..n-1)) through lammps_extract_fix(style 0, type 1)"""
        out = np.zeros(n)
        for i in range(n):
            p = self.lib.lammps_extract_fix(self.h, fix_id.encode(), 0, 1, i, 0)
            out[i] = C.cast(p, C.POINTER(C.c_double))[0]
            self.lib.lammps_free.argtypes = [C.c_void_p]
            self.lib.lammps_free(p)
        return out

    def compute_vector(self, compute_id, n):
        """global vector of a compute through lammps_extract_compute(style 0, type 1) (invokes Compute::compute_vector when it has
        not been invoked on this step)"""
        L = self.lib
        L.lammps_extract_compute.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.lammps_extract_compute.restype = C.c_void_p
        p = L.lammps_extract_compute(self.h, compute_id.encode(), 0, 1)
        return np.array([C.cast(p, C.POINTER(C.c_double))[i] for i in range(n)])

    def mesh_geometry(self, mesh_id):
        L = self.lib
        L.ref_mesh_ntri.argtypes = [C.c_void_p, C.c_char_p]
        n = L.ref_mesh_ntri(self.h, mesh_id.encode())
        ev = np.zeros((n, 3, 3)); en = np.zeros((n, 3, 3)); sn = np.zeros((n, 3)); ce = np.zeros((n, 3)); vn = np.zeros((n, 3, 3))
        L.ref_mesh_geometry.argtypes = [C.c_void_p, C.c_char_p] + [C.c_void_p] * 5
        L.ref_mesh_geometry(self.h, mesh_id.encode(), ev.ctypes.data, en.ctypes.data, sn.ctypes.data, ce.ctypes.data, vn.ctypes.data)
        return {"edge_vec": ev, "edge_norm": en, "surf_norm": sn, "center": ce, "v_node": vn}

    def mesh_contacts(self, mesh_id):
        """mesh contact rows of fix_contact_history_mesh sorted by (tag, triangle id)"""
        L = self.lib
        L.ref_mesh_contacts.argtypes = [C.c_void_p, C.c_char_p, C.c_int] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
        dn = C.c_int(0)
        n = L.ref_mesh_contacts(self.h, mesh_id.encode(), 0, None, None, None, C.byref(dn))
        tag = np.zeros(max(n, 1), np.int32); tri = np.zeros(max(n, 1), np.int32); hist = np.zeros((max(n, 1), max(dn.value, 1)))
        if n > 0:
            L.ref_mesh_contacts(self.h, mesh_id.encode(), n, tag.ctypes.data, tri.ctypes.data, hist.ctypes.data, C.byref(dn))
        tag, tri, hist = tag[:max(n, 0)], tri[:max(n, 0)], hist[:max(n, 0), :dn.value]
        o = np.lexsort((tri, tag))
        return {"tag": tag[o], "tri": tri[o], "hist": hist[o]}

    def pairs(self):
        """granular half list: canonical (tag_lo, tag_hi) sorted, flag, history rows with the
        sign convention of the (tag_lo -> first) orientation (all our history values are
        newtonflag=1 vectors: they flip sign when the pair is viewed from the other side,
        fix_contact_history.cpp:406-409)."""
        dn = C.c_int(0)
        n = self.lib.ref_pairlist_count(self.h, C.byref(dn))
        dnum = dn.value
        ti = np.zeros(n, np.int32); tj = np.zeros(n, np.int32)
        ii = np.zeros(n, np.int32); jj = np.zeros(n, np.int32)
        fl = np.zeros(n, np.int32); hist = np.zeros((n, max(dnum, 1)), np.float64)
        self.lib.ref_pairlist_dump(self.h, ti.ctypes.data, tj.ctypes.data, ii.ctypes.data, jj.ctypes.data,
                                   fl.ctypes.data, hist.ctypes.data if dnum else None)
        hist = hist[:, :dnum]
        swap = ti > tj
        lo = np.where(swap, tj, ti); hi = np.where(swap, ti, tj)
        hist = np.where(swap[:, None], -hist, hist)
        o = np.lexsort((hi, lo))
        return {"lo": lo[o], "hi": hi[o], "flag": fl[o], "hist": hist[o], "dnum": dnum}
