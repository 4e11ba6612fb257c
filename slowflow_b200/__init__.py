"""slowflow_b200 -- B200 (sm_100a) implementation of Slow Flow's variational refinement hot path.

The product is ``lib/libslowflow_gpu.so`` (hand-written CUDA behind the C ABI of
``include/slowflow_gpu.h``).  This Python package is the thin host-side mirror used by the tests,
``bench.py`` and ``__graft_entry__``: ctypes views of the reference's ``image_t`` /
``color_image_t`` / ``variational_params_t`` and wrappers with the reference's entry-point names.
There is no CPU fallback: importing works anywhere, calling a compute entry without the CUDA
library or without a GPU raises.
"""
from .image import Image, ColorImage, image_t, color_image_t  # noqa: F401
from .params import VariationalParams, MTParams, variational_params_default, mt_params_default  # noqa: F401
from .api import Context, variational, Variational_MT, load_library, library_path  # noqa: F401

__version__ = "0.1.0"
