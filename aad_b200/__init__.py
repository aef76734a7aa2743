"""aad_b200 -- B200-native (sm_100a CUDA) implementation of the AAD ADPCM encode/decode hot path.

The product is the C-ABI shared library ``aad_b200/libaad_b200.so`` (C host code + CUDA
kernels, built in-tree by ``__graft_entry__.build()`` / ``make -C aad_b200/csrc``).  This
package is only the Python face of that library for tests and ``bench.py``: ctypes bindings of
the reference's own encoder/decoder API (``capi``) and of the batch / device-resident
extension (``gpu``).  There is no Python or CPU implementation of the codec in here -- if the
library is missing, importing ``aad_b200.lib`` raises.
"""
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent
import os as _os

# AAD_B200_LIBRARY: load another build of the same library (kernel experiments); default = the in-tree build
LIBRARY_PATH = Path(_os.environ.get("AAD_B200_LIBRARY") or PACKAGE_DIR / "libaad_b200.so")

__all__ = ["PACKAGE_DIR", "LIBRARY_PATH", "load"]


def load():
    """Load libaad_b200.so and return (AADCApi, GpuApi).  Raises if it was not built."""
    if not LIBRARY_PATH.exists():
        raise ImportError(
            f"{LIBRARY_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C aad_b200/csrc` (there is no CPU fallback)")
    from .capi import AADCApi
    from .gpu import GpuApi
    api = AADCApi(LIBRARY_PATH)
    return api, GpuApi(api.lib)
