"""crackling_b200 -- B200-native ISSL off-target scorer (drop-in for Crackling's isslScoreOfftargets).

The product is native: `crackling_b200/lib/libissl_cuda.so` (C ABI in include/issl_cuda.h,
hand-written sm_100a kernels in crackling_b200/csrc/) and the host program
`bin/isslScoreOfftargets`.  This package is the thin ctypes mirror of that C ABI used by the
tests and by bench.py; it adds no compute of its own and has no CPU fallback.
"""
from .binding import (  # noqa: F401
    IsslError, Index, Device, Sites, HostBuffer, replicate, score_multi, multi_chunk, lib, lib_path, build, pack_guides, unpack_guide, format_lines, method_code,
    local_mit_score, mit_table, triple_visits, triple_layout, device_count, cli_path, create_cli_path, extract_cli_path, LAYOUTS, METHODS,
)
