#!/usr/bin/env python3
"""score_once.py -- the smallest program that runs the hot kernel at full size: synthetic human-scale index, one batch
of guides scored a few times through the C ABI (no torch).  Used under ncu (B200_PROFILING.md) and for quick A/B timing:
prints the library's own CUDA-event times.

    python tools/score_once.py [--guides 100000] [--reps 3] [--max-dist 4] [--threshold 0] [--slice-width 8] [--sites N]
"""
import argparse
import json
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import crackling_b200 as cb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--guides", type=int, default=100_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-dist", type=int, default=4)
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--method", default="and")
    ap.add_argument("--slice-width", type=int, default=8)
    ap.add_argument("--sites", type=int, default=581_250_000)
    ap.add_argument("--layout", default="auto")
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--family-size", type=int, default=0)
    a = ap.parse_args()
    dev = cb.Device.synthetic(0, a.layout, seed=1, uniform_sites=a.sites, families=a.families, family_size=a.family_size,
                              seq_length=20, slice_width=a.slice_width)
    rng = np.random.default_rng(2)
    n_own = a.guides * 9 // 10
    guides = np.concatenate([dev.read_sites(rng.integers(0, dev.info["offtargetsCount"], n_own).astype(np.uint64)),
                             rng.integers(0, 1 << 40, a.guides - n_own, dtype=np.uint64)])
    rng.shuffle(guides)
    hg, hm, hc = cb.HostBuffer(guides.size, np.uint64), cb.HostBuffer(guides.size, np.float64), cb.HostBuffer(guides.size, np.float64)
    hg.array[:] = guides
    out = []
    for _ in range(a.reps):
        dev.score_into(hg.array, a.max_dist, a.threshold, a.method, hm.array, hc.array)
        st = dev.stats
        out.append({"scan_ms": round(st["scan_ms"], 4), "total_ms": round(st["total_ms"], 4), "hits": st["hits"],
                    "launches": st["launches"], "bucket_visits": st["bucket_visits"]})
    print(json.dumps({"guides": a.guides, "max_dist": a.max_dist, "threshold": a.threshold, "slice_width": a.slice_width,
                      "checksum": float(hm.array.sum() + hc.array.sum()), "reps": out}))
    dev.close()


if __name__ == "__main__":
    main()
