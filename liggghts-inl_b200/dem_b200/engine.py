import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

# every symbol declared in include/dem_b200.h
ABI_SYMBOLS = [
    "dem_create", "dem_destroy", "dem_nccl_unique_id", "dem_decomposition", "dem_last_error", "dem_version", "dem_set_option", "dem_set_units", "dem_set_box",
    "dem_set_ntypes", "dem_set_processors", "dem_set_neighbor", "dem_set_timestep", "dem_set_contact_distance_factor", "dem_set_property",
    "dem_set_pair_style", "dem_add_wall_primitive", "dem_set_gravity", "dem_set_freeze",
    "dem_set_integrate", "dem_set_extra_force", "dem_upload_particles", "dem_insert_particles", "dem_insert_step_begin", "dem_insert_step_end", "dem_setup", "dem_run", "dem_nlocal", "dem_download",
    "dem_pair_count", "dem_download_pairs", "dem_download_wall_history", "dem_get_stats",
    "dem_add_mesh", "dem_move_mesh", "dem_add_wall_mesh", "dem_download_mesh", "dem_mesh_force", "dem_mesh_contact_count", "dem_download_mesh_contacts",
    "dem_bond_counter", "dem_contact_count", "dem_download_contacts", "dem_trim_memory", "dem_brick_layout", "dem_deck_open", "dem_deck_close", "dem_deck_command", "dem_deck_file", "dem_deck_last_error", "dem_deck_warnings", "dem_deck_ntimestep", "dem_deck_output", "dem_deck_screen",
]


class DemError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("ntimestep", C.c_long), ("nbuilds", C.c_long), ("nlocal", C.c_long), ("nghost", C.c_long),
                ("npairs_full", C.c_long), ("ncontacts_full", C.c_long), ("kernel_launches", C.c_long),
                ("maxneigh", C.c_int), ("dnum", C.c_int), ("step_kernel_ms", C.c_double),
                ("step_kernel_calls", C.c_long)]


def library_path():
    return os.environ.get("DEM_B200_LIB", os.path.join(_ROOT, "libdem_b200.so"))


def load_library(path=None):
    path = path or library_path()
    if not os.path.exists(path):
        raise DemError("libdem_b200.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    return C.CDLL(path)


def read_stl(path):
    """ASCII or binary STL -> (ntri, 3, 3) float64 (the reference: input_mesh_tri.cpp:308-591)"""
    raw = open(path, "rb").read()
    head = raw[:512].lstrip().lower()
    if head.startswith(b"solid") and b"facet" in raw[:4096].lower():
        v = [[float(t) for t in line.split()[1:4]] for line in raw.decode("ascii", "replace").splitlines() if line.strip().startswith("vertex")]
        return np.asarray(v, np.float64).reshape(-1, 3, 3)
    n = int(np.frombuffer(raw[80:84], "<u4")[0])
    rec = np.frombuffer(raw[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return rec["v"].astype(np.float64)


def write_stl(path, nodes, name="mesh"):
    """ASCII STL with round-trip (%.17g) vertices"""
    with open(path, "w") as f:
        f.write("solid %s\n" % name)
        for t in np.asarray(nodes, np.float64).reshape(-1, 3, 3):
            f.write(" facet normal 0 0 0\n  outer loop\n")
            for v in t:
                f.write("   vertex %.17g %.17g %.17g\n" % tuple(v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid %s\n" % name)


def _strv(args):
    arr = (C.c_char_p * len(args))(*[str(a).encode() for a in args])
    return len(args), arr


class Engine:
    """One DEM engine on one GPU.  `lib`/`prefix` exist so the test-suite can drive the CPU
    oracle (tests only) through the very same call sequence."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id=None, stream=None, lib=None, prefix="dem_"):
        self._lib = lib if lib is not None else load_library()
        self._p = prefix
        self._h = C.c_void_p()
        f = self._fn("create")
        f.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        idbuf = C.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
        rc = f(C.byref(self._h), device, rank, nranks, idbuf, C.c_void_p(stream or 0))
        if rc != 0 or not self._h:
            msg = self._err() if self._h else "engine creation failed (no usable sm_100 GPU?)"
            raise DemError(msg)

    @staticmethod
    def nccl_unique_id(lib=None):
        """128-byte ncclUniqueId (call on rank 0, then broadcast to all ranks)"""
        lib = lib if lib is not None else load_library()
        buf = C.create_string_buffer(128)
        lib.dem_nccl_unique_id.argtypes = [C.c_void_p]
        if lib.dem_nccl_unique_id(buf) != 0:
            raise DemError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return bytes(buf.raw)

    def decomposition(self):
        pg = (C.c_int * 3)(); ml = (C.c_int * 3)(); lo = (C.c_double * 3)(); hi = (C.c_double * 3)()
        self._call("decomposition", [C.c_void_p] * 4, pg, ml, lo, hi)
        return list(pg), list(ml), list(lo), list(hi)

    # -- plumbing -----------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _err(self):
        f = self._fn("last_error"); f.restype = C.c_char_p; f.argtypes = [C.c_void_p]
        return (f(self._h) or b"").decode()

    def _call(self, name, argtypes, *args):
        f = self._fn(name); f.argtypes = [C.c_void_p] + argtypes; f.restype = C.c_int
        rc = f(self._h, *args)
        if rc != 0:
            raise DemError("%s%s failed (%d): %s" % (self._p, name, rc, self._err()))

    def close(self):
        if self._h:
            f = self._fn("destroy"); f.argtypes = [C.c_void_p]; f.restype = None
            f(self._h); self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def option(self, name, value):
        if self._p == "dem_":
            self._call("set_option", [C.c_char_p, C.c_double], name.encode(), float(value))

    # -- deck vocabulary ----------------------------------------------------------------
    def units(self, style):
        self._call("set_units", [C.c_char_p], style.encode())

    def box(self, lo, hi, periodic=(0, 0, 0)):
        self._call("set_box", [C.POINTER(C.c_double)] * 2 + [C.POINTER(C.c_int)],
                   (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_int * 3)(*[int(p) for p in periodic]))

    def ntypes(self, n):
        self._call("set_ntypes", [C.c_int], n)

    def processors(self, px, py, pz):
        self._call("set_processors", [C.c_int] * 3, px, py, pz)

    def neighbor(self, skin, every=1, delay=0, check=True):
        self._call("set_neighbor", [C.c_double, C.c_int, C.c_int, C.c_int], skin, every, delay, int(check))

    def timestep(self, dt):
        self._call("set_timestep", [C.c_double], dt)

    def property_global(self, name, kind, values):
        v = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).ravel()
        self._call("set_property", [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int],
                   name.encode(), kind.encode(), v.ctypes.data, v.size)

    def pair_style(self, text):
        """text as in the deck after `pair_style gran`, e.g. 'model hertz tangential history'"""
        n, a = _strv(text.split())
        self._call("set_pair_style", [C.c_int, C.POINTER(C.c_char_p)], n, a)

    def wall_primitive(self, wall_id, text):
        """text as in the deck after `fix ID all wall/gran`"""
        n, a = _strv(text.split())
        self._call("add_wall_primitive", [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)], wall_id.encode(), n, a)

    def mesh(self, mesh_id, atom_type, nodes, options=""):
        """`fix ID all mesh/surface file F type T [options]`; nodes = the file's triangles, shape (ntri, 3, 3)"""
        nd = np.ascontiguousarray(nodes, np.float64).reshape(-1, 9)
        n, a = _strv(options.split())
        self._call("add_mesh", [C.c_char_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.POINTER(C.c_char_p)],
                   mesh_id.encode(), int(atom_type), nd.ctypes.data, len(nd), n, a)

    def mesh_stl(self, mesh_id, atom_type, path, options=""):
        self.mesh(mesh_id, atom_type, read_stl(path), options)

    def move_mesh(self, mesh_id, text):
        """`fix ID all move/mesh mesh MESH linear vx vy vz`"""
        n, a = _strv(text.split())
        self._call("move_mesh", [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)], mesh_id.encode(), n, a)

    def wall_mesh(self, wall_id, text):
        """text as in the deck after `fix ID all wall/gran`: 'model ... mesh n_meshes N meshes id...'"""
        n, a = _strv(text.split())
        self._call("add_wall_mesh", [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)], wall_id.encode(), n, a)

    def gravity(self, magnitude, direction):
        self._call("set_gravity", [C.c_double, C.POINTER(C.c_double)], magnitude, (C.c_double * 3)(*direction))

    def freeze(self, groupbit):
        self._call("set_freeze", [C.c_int], groupbit)

    def integrate(self, groupbit=1):
        self._call("set_integrate", [C.c_int], groupbit)

    # -- particles ----------------------------------------------------------------------
    def upload(self, tag, type, x, radius, density, v=None, omega=None, mask=None):
        n = len(tag)
        self._keep = [np.ascontiguousarray(tag, np.int32), np.ascontiguousarray(type, np.int32),
                      None if mask is None else np.ascontiguousarray(mask, np.int32),
                      np.ascontiguousarray(x, np.float64), None if v is None else np.ascontiguousarray(v, np.float64),
                      None if omega is None else np.ascontiguousarray(omega, np.float64),
                      np.ascontiguousarray(radius, np.float64), np.ascontiguousarray(density, np.float64)]
        ptr = [a.ctypes.data if a is not None else None for a in self._keep]
        self._call("upload_particles", [C.c_long] + [C.c_void_p] * 8, n, *ptr)

    def contact_distance_factor(self, f):
        self._call("set_contact_distance_factor", [C.c_double], float(f))

    def insert(self, tag, type, x, radius, density, v=None, omega=None, mask=None):
        """particles added between two runs (create_atoms / fix insert/*): appended, existing contact history kept; call setup() next"""
        n = len(tag)
        keep = [np.ascontiguousarray(tag, np.int32), np.ascontiguousarray(type, np.int32),
                None if mask is None else np.ascontiguousarray(mask, np.int32),
                np.ascontiguousarray(x, np.float64), None if v is None else np.ascontiguousarray(v, np.float64),
                None if omega is None else np.ascontiguousarray(omega, np.float64),
                np.ascontiguousarray(radius, np.float64), np.ascontiguousarray(density, np.float64)]
        ptr = [a.ctypes.data if a is not None else None for a in keep]
        self._call("insert_particles", [C.c_long] + [C.c_void_p] * 8, n, *ptr)

    def extra_force(self, fix_id, kind, groupbit, values):
        """fix addforce (kind 0: fx fy fz) / fix viscous (kind 1: gamma) on a group; values=None removes the fix"""
        if values is None:
            self._call("set_extra_force", [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int], fix_id.encode(), int(kind), int(groupbit), None, -1)
            return
        v = np.ascontiguousarray(values, np.float64)
        self._call("set_extra_force", [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int], fix_id.encode(), int(kind), int(groupbit), v.ctypes.data, len(v))

    def insert_step_begin(self):
        """first half of the timestep in which fix insert/* creates particles (fix_insert.cpp:672-905): first half step of the
        existing particles; download('x') then returns the positions the reference's overlap check sees"""
        self._call("insert_step_begin", [])

    def insert_step_end(self, tag, type, x, radius, density, v=None, omega=None, mask=None):
        """second half: the particles appear, rebuild, forces, second half step (mass = density * volume as fix insert forms it)"""
        n = len(tag)
        keep = [np.ascontiguousarray(tag, np.int32), np.ascontiguousarray(type, np.int32),
                None if mask is None else np.ascontiguousarray(mask, np.int32),
                np.ascontiguousarray(x, np.float64), None if v is None else np.ascontiguousarray(v, np.float64),
                None if omega is None else np.ascontiguousarray(omega, np.float64),
                np.ascontiguousarray(radius, np.float64), np.ascontiguousarray(density, np.float64)]
        ptr = [a.ctypes.data if a is not None else None for a in keep]
        self._call("insert_step_end", [C.c_long] + [C.c_void_p] * 8, n, *ptr)

    def setup(self):
        self._call("setup", [])

    def run(self, nsteps):
        self._call("run", [C.c_long], int(nsteps))

    # -- read-back ----------------------------------------------------------------------
    @property
    def nlocal(self):
        f = self._fn("nlocal"); f.argtypes = [C.c_void_p]; f.restype = C.c_long
        return f(self._h)

    def download(self, field, out=None):
        """per-particle field ordered by tag.  `out`: optional caller-owned C-contiguous array of the right dtype/shape
        (e.g. a view of page-locked memory: the engine copies straight into the caller's buffer)"""
        n = self.nlocal
        if field in ("tag", "type", "mask"):
            shape, dt = (n,), np.int32
        elif field in ("radius", "rmass", "density"):
            shape, dt = (n,), np.float64
        else:
            shape, dt = (n, 3), np.float64
        if out is None:
            out = np.zeros(shape, dt)
        elif out.dtype != dt or out.shape != shape or not out.flags.c_contiguous:
            raise ValueError("download(%s): out must be a C-contiguous %s array of shape %s" % (field, np.dtype(dt).name, shape))
        self._call("download", [C.c_char_p, C.c_void_p, C.c_long], field.encode(), out.ctypes.data, n)
        return out

    def atoms(self, fields=("tag", "type", "x", "v", "f", "omega", "torque", "radius", "rmass")):
        return {k: self.download(k) for k in fields}

    def pairs(self):
        npairs = C.c_long(0); dnum = C.c_int(0)
        self._call("pair_count", [C.POINTER(C.c_long), C.POINTER(C.c_int)], C.byref(npairs), C.byref(dnum))
        n, d = npairs.value, dnum.value
        lo = np.zeros(n, np.int32); hi = np.zeros(n, np.int32); fl = np.zeros(n, np.int32)
        hist = np.zeros((n, max(d, 1)), np.float64)
        self._call("download_pairs", [C.c_void_p] * 4, lo.ctypes.data, hi.ctypes.data, fl.ctypes.data, hist.ctypes.data)
        return {"lo": lo, "hi": hi, "flag": fl, "hist": hist[:, :d], "dnum": d}

    def bond_counter(self):
        """compute bond/counter as the reference returns it between two runs (see include/dem_b200.h): [created, broken, created - broken (unsigned), 0, 0, 0]"""
        out = np.zeros(6, np.float64)
        self._call("bond_counter", [C.c_void_p], out.ctypes.data)
        return out

    def contacts(self):
        """(option contact_output) rows of the last force evaluation: own tag, partner tag, force and torque on the own particle"""
        n = C.c_long(0)
        self._call("contact_count", [C.POINTER(C.c_long)], C.byref(n))
        n = n.value
        tag = np.zeros(n, np.int32); partner = np.zeros(n, np.int32)
        f = np.zeros((n, 3)); t = np.zeros((n, 3))
        self._call("download_contacts", [C.c_void_p] * 4, tag.ctypes.data, partner.ctypes.data, f.ctypes.data, t.ctypes.data)
        return {"tag": tag, "partner": partner, "force": f, "torque": t}

    def wall_history(self, wall_id, dnum):
        n = self.nlocal
        out = np.zeros((n, max(dnum, 1)), np.float64)
        self._call("download_wall_history", [C.c_char_p, C.c_void_p, C.c_long], wall_id.encode(), out.ctypes.data, n)
        return out[:, :dnum]

    def mesh_field(self, mesh_id, field, ntri):
        """topology / geometry of a mesh: nodes, edge_vec, edge_norm (ntri,3,3) f64; surf_norm, center (ntri,3) f64; edge_active, corner_active (ntri,3) i32; obtuse, nneighs (ntri,) i32"""
        if field in ("nodes", "edge_vec", "edge_norm"):
            out = np.zeros((ntri, 3, 3), np.float64)
        elif field in ("surf_norm", "center"):
            out = np.zeros((ntri, 3), np.float64)
        elif field in ("edge_active", "corner_active"):
            out = np.zeros((ntri, 3), np.int32)
        else:
            out = np.zeros(ntri, np.int32)
        self._call("download_mesh", [C.c_char_p, C.c_char_p, C.c_void_p, C.c_long], mesh_id.encode(), field.encode(), out.ctypes.data, out.size)
        return out

    def mesh_force(self, mesh_id):
        """`f_<mesh>[1..9]` of a `fix mesh/surface/stress`: total force, total torque about the reference point, reference point"""
        out = np.zeros(9, np.float64)
        self._call("mesh_force", [C.c_char_p, C.c_void_p], mesh_id.encode(), out.ctypes.data)
        return out

    def mesh_contacts(self, mesh_id):
        """per-particle mesh contact rows (fix_contact_history_mesh): sorted by (tag, triangle id)"""
        n = C.c_long(0); dnum = C.c_int(0)
        self._call("mesh_contact_count", [C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_int)], mesh_id.encode(), C.byref(n), C.byref(dnum))
        tag = np.zeros(n.value, np.int32); tri = np.zeros(n.value, np.int32); hist = np.zeros((n.value, max(dnum.value, 1)), np.float64)
        self._call("download_mesh_contacts", [C.c_char_p] + [C.c_void_p] * 3, mesh_id.encode(), tag.ctypes.data, tri.ctypes.data, hist.ctypes.data)
        return {"tag": tag, "tri": tri, "hist": hist[:, :dnum.value]}

    def stats(self):
        s = Stats()
        self._call("get_stats", [C.POINTER(Stats)], C.byref(s))
        return s


def brick_layout(nranks, rank, lo, hi, periodic, procgrid=None, x=None, lib=None):
    """decomposition without an engine or a GPU: dict(pgrid, myloc, sublo, subhi, neigh[6], mine[n] or None)"""
    lib = lib if lib is not None else load_library()
    f = lib.dem_brick_layout
    f.restype = C.c_int
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_long, C.c_void_p, C.c_void_p]
    pg = (C.c_int * 3)(); ml = (C.c_int * 3)(); sl = (C.c_double * 3)(); sh = (C.c_double * 3)(); ng = (C.c_int * 6)()
    n = 0 if x is None else len(x)
    xs = None if x is None else np.ascontiguousarray(x, np.float64)
    mine = None if x is None else np.zeros(n, np.int32)
    rc = f(nranks, rank, (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_int * 3)(*[int(p) for p in periodic]),
           (C.c_int * 3)(*procgrid) if procgrid else None, pg, ml, sl, sh, ng, n,
           xs.ctypes.data if xs is not None else None, mine.ctypes.data if mine is not None else None)
    if rc != 0:
        raise DemError("dem_brick_layout failed (%d)" % rc)
    return dict(pgrid=list(pg), myloc=list(ml), sublo=list(sl), subhi=list(sh), neigh=list(ng), mine=mine)


class Deck:
    """Input-script front end over an Engine: the reference's `lammps.command()` / `lammps.file()` (python/liggghts.py)
    for the hot-path commands.  `lib`/`prefix` let the test-suite bind the same parser to the CPU oracle (tests only)."""

    def __init__(self, engine, lib=None, prefix=None):
        self.engine = engine
        self._lib = lib if lib is not None else engine._lib
        self._p = prefix if prefix is not None else engine._p + "deck_"
        self._h = C.c_void_p()
        f = getattr(self._lib, self._p + "open"); f.argtypes = [C.POINTER(C.c_void_p), C.c_void_p]; f.restype = C.c_int
        if f(C.byref(self._h), engine._h) != 0:
            raise DemError("deck_open failed")

    def _err(self):
        f = getattr(self._lib, self._p + "last_error"); f.restype = C.c_char_p; f.argtypes = [C.c_void_p]
        return (f(self._h) or b"").decode()

    def command(self, line):
        f = getattr(self._lib, self._p + "command"); f.argtypes = [C.c_void_p, C.c_char_p]; f.restype = C.c_int
        for one in line.splitlines():
            rc = f(self._h, one.encode())
            if rc != 0:
                raise DemError("%s (%d): %s" % (one.strip(), rc, self._err()))

    def file(self, path):
        f = getattr(self._lib, self._p + "file"); f.argtypes = [C.c_void_p, C.c_char_p]; f.restype = C.c_int
        rc = f(self._h, path.encode())
        if rc != 0:
            raise DemError("%s (%d): %s" % (path, rc, self._err()))

    @property
    def warnings(self):
        f = getattr(self._lib, self._p + "warnings"); f.restype = C.c_char_p; f.argtypes = [C.c_void_p]
        return (f(self._h) or b"").decode()

    @property
    def output(self):
        """thermo header and lines produced so far"""
        f = getattr(self._lib, self._p + "output"); f.restype = C.c_char_p; f.argtypes = [C.c_void_p]
        return (f(self._h) or b"").decode()

    @property
    def ntimestep(self):
        f = getattr(self._lib, self._p + "ntimestep"); f.restype = C.c_long; f.argtypes = [C.c_void_p]
        return f(self._h)

    def close(self):
        if self._h:
            f = getattr(self._lib, self._p + "close"); f.argtypes = [C.c_void_p]; f.restype = None
            f(self._h); self._h = C.c_void_p()
