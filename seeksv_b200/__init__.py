"""seeksv_b200: B200-native implementation of seeksv's hot path (getclip / getsv / somatic).

The product is the C-ABI shared library (include/seeksv_b200.h) and the `seeksv` CLI built from
seeksv_b200/csrc (CUDA, sm_100a) and seeksv_b200/host (C++). This Python package is only the loader
and thin ctypes bindings used by bench.py, the tests and multi-GPU runs under torch.distributed.
There is no CPU path: every entry point needs the CUDA library and a B200.
"""
from .lib import (SvbError, Bam, Context, cli_path, lib_path, load, plan_getsv, run_cli)  # noqa: F401
