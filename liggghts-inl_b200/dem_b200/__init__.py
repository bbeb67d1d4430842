"""dem_b200 -- host-side Python mirror of the C ABI in include/dem_b200.h.

Thin ctypes binding (no torch types cross the boundary).  The method names follow the
input-script vocabulary of the reference (pair_style, fix wall/gran, fix property/global,
neighbor, timestep, run ...) so that a deck maps 1:1 onto calls.  The CUDA library is
mandatory: importing works without it (so CPU-only hosts can inspect symbols), but
creating an Engine raises if libdem_b200.so or a B200 is missing -- there is no CPU path.
"""
from .engine import Engine, Deck, brick_layout, DemError, Stats, library_path, load_library, ABI_SYMBOLS, read_stl, write_stl  # noqa: F401
