#!/usr/bin/env python
"""Test-suite launcher: runs a script (or `-m module`) with the product's Python binding pointed at the CPU SIMT-emulator
build of the kernels (tests/simt_emu/_build/libtracer_emu.so, path in LYS_EMU_LIB).

The product binding (msc-futhark-ray-tracer_b200/tracer.py) has no switch that could load a CPU-executing library: it only
ever opens the libtracer.so next to it.  This launcher is the one place that redirects it, by patching the module attribute
before anything is loaded, and it lives under tests/.

    LYS_EMU_LIB=<libtracer_emu.so> python tests/simt_emu/run_with_emu.py tools/gpu_parity_quick.py cornell
    LYS_EMU_LIB=<libtracer_emu.so> python tests/simt_emu/run_with_emu.py -m pytest tests/test_gpu_parity.py -m gpu -k soup
"""
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
lib = os.environ.get('LYS_EMU_LIB')
if not lib or not os.path.exists(lib):
    raise SystemExit('run_with_emu.py: LYS_EMU_LIB must point at the emulator build (tests/simt_emu/emu_build.py)')
tracer = importlib.import_module('msc-futhark-ray-tracer_b200.tracer')
tracer._SO = lib
args = sys.argv[1:]
if not args:
    raise SystemExit(__doc__)
if args[0] == '-m':
    sys.argv = args[1:]
    runpy.run_module(args[1], run_name='__main__', alter_sys=True)
else:
    sys.argv = args
    sys.path.insert(0, os.path.dirname(os.path.abspath(args[0])))
    runpy.run_path(args[0], run_name='__main__')
