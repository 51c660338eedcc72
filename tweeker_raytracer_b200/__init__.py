"""B200-native path-tracing core behind the rtigo3 host interface.

  core  -- ctypes binding of librtcore.so, the C ABI of the hand-written sm_100a kernels (include/rtc_core.h)
  host  -- ctypes binding of librtigo3host.so, the C++ Application / Raytracer / Device mirror of rtigo3

Both libraries are native code built by `make` (or __graft_entry__.build()).  There is no Python or CPU
rendering path: importing works anywhere, rendering needs a CUDA device and fails loudly without one.
"""
from . import core, host  # noqa: F401

__all__ = ["core", "host"]
