"""nextdenovo_b200 -- B200-native read-overlap engine behind NextDenovo's ``minimap2-nd --step 1`` interface.

Host-side mirror (Python) of the reference's option handling and file formats on top of the C-ABI
library ``libndgpu.so`` (hand-written sm_100a CUDA).  There is no CPU compute path: importing the
compute API without the built library, or running it without a B200, raises.
"""
from .options import IdxOpt, MapOpt, set_opt, parse_num  # noqa: F401
from .formats import (read_2bit, write_2bit, read_fastx, read_seqfile, reads_from_ascii, Reads, decode_ovl8, decode_ovl10, encode_ovl8_py,
                      name_keys, x31_hash)  # noqa: F401

__all__ = ["IdxOpt", "MapOpt", "set_opt", "parse_num", "read_2bit", "write_2bit", "Reads",
           "decode_ovl8", "decode_ovl10", "encode_ovl8_py", "name_keys", "x31_hash", "Engine", "lib_path"]


def __getattr__(name):          # lazy: the ctypes binding loads the CUDA library on first use
    if name in ("Engine", "lib_path", "load_library", "NdgError"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
