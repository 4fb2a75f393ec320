"""msc-futhark-ray-tracer_b200: B200-native replacement for the Futhark library of bryal/msc-futhark-ray-tracer.

The product is csrc/ (sm_100a CUDA behind the futhark_* C ABI of include/tracer.h, built as libtracer.so /
libtracer.a).  This package is the thin Python host side: ctypes bindings that mirror the reference's host
wrappers (demo-interactive/liblys.c, demo-save/src/wrapper.rs), the OBJ/MTL loader binding and scene helpers.
There is no CPU fallback: importing works anywhere, creating a Context needs a CUDA device."""
from .tracer import (Context, State, TracerError, build, lib_path, load_obj, KEY)  # noqa: F401
from . import scenes  # noqa: F401
