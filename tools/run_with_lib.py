#!/usr/bin/env python
"""Development aid: run a tool script against another build of libtracer (tools/build_variants.sh), e.g.
    python tools/run_with_lib.py msc-futhark-ray-tracer_b200/variants/libtracer_minb12.so tools/bench_configs.py 4 5
The product binding has no library override; this launcher patches the module attribute before anything is loaded."""
import importlib
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib, script = os.path.abspath(sys.argv[1]), sys.argv[2]
assert os.path.exists(lib), lib
importlib.import_module('msc-futhark-ray-tracer_b200.tracer')._SO = lib
sys.argv = sys.argv[2:]
sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
runpy.run_path(script, run_name='__main__')
