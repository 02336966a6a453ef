#!/usr/bin/env python3
"""config5_sweep.py -- BASELINE.json configs[4] / SURVEY.md 8d (C5): the slice-width sweep w = 4 / 8 / 10 x maxDist 2-5 on
the human-scale index, threshold 0, one JSON line per point with hits/guide, C_g (list entries the reference's loop visits
per guide), HBM GB/s of the scan kernel against the measured peak, guides/s, and -- on EVERY line -- the unmodified
reference binary timed on the same box's host cores on a sample of the same guides plus a byte-for-byte comparison of its
stdout with ours (full-size parity for w = 4 and w = 10, which the small fixtures cannot reach).

One index per slice width is built in HBM, written out once as a real .issl for the reference, and scored at the four
distances: the reference's index load (measured once per width with an empty guide file) is subtracted from its wall time.

    python tools/config5_sweep.py [--out gpurun_out/config5.jsonl] [--guides 100000] [--widths 8,4,10] [--dists 2,3,4,5]
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (guide generator, guide-file writer, peak, reference path)
import crackling_b200 as cb  # noqa: E402


def run_reference(exe, issl, gpath, max_dist, env):
    t0 = time.perf_counter()
    p = subprocess.run([exe, issl, gpath, str(max_dist), "0.0", "and"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return time.perf_counter() - t0, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/config5.jsonl")
    ap.add_argument("--guides", type=int, default=100_000)
    ap.add_argument("--small-guides", type=int, default=10_000, help="C5 as stated scores 10 k guides: timed as well")
    ap.add_argument("--sites", type=int, default=bench.HUMAN_SITES)
    ap.add_argument("--widths", default="8,4,10")
    ap.add_argument("--dists", default="2,3,4,5")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="aimed scoring time of the reference per point")
    a = ap.parse_args()
    peak, peak_src = bench.hbm_peak()
    exe = bench.reference_exe()
    cores = os.cpu_count() or 1
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    scratch = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    out = open(a.out, "w")
    for w in [int(x) for x in a.widths.split(",")]:
        t0 = time.perf_counter()
        dev = cb.Device.synthetic(0, "auto", seed=1, uniform_sites=a.sites, seq_length=20, slice_width=w)
        t_build = time.perf_counter() - t0
        info = dev.info
        layout = {1: "res32", 2: "sig64", 3: "gather", 4: "triple"}[info["layout"]]
        guides = bench.make_guides(dev, a.guides, seed=2)
        hg, hm, hc = cb.HostBuffer(guides.size, np.uint64), cb.HostBuffer(guides.size, np.float64), cb.HostBuffer(guides.size, np.float64)
        hg.array[:] = guides
        with tempfile.TemporaryDirectory(dir=scratch) as tmp:
            issl, epath, gpath = os.path.join(tmp, "index.issl"), os.path.join(tmp, "empty.txt"), os.path.join(tmp, "guides.txt")
            t_load = None
            if exe:
                dev.write_issl(issl)
                open(epath, "wb").close()
                t_load, _ = run_reference(exe, issl, epath, 4, env)      # index load only: fails on the empty guide file
            per_guide_s = None                                            # the reference's cost per guide, from the previous point
            for md in [int(x) for x in a.dists.split(",")]:
                reps = []
                for _ in range(a.reps):
                    t0 = time.perf_counter()
                    dev.score_into(hg.array, md, 0.0, "and", hm.array, hc.array)
                    wall = time.perf_counter() - t0
                    st = dev.stats
                    reps.append((wall, st))
                wall, st = min(reps[1:], key=lambda r: r[0])
                small = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    dev.score_into(hg.array[:a.small_guides], md, 0.0, "and", hm.array[:a.small_guides], hc.array[:a.small_guides])
                    small.append(time.perf_counter() - t0)
                dev.score_into(hg.array, md, 0.0, "and", hm.array, hc.array)
                st = dev.stats
                n = guides.size
                blk = info["triple_block_bytes"]
                if st["bucket_visits"]:
                    alg = st["bucket_visits"] * blk if blk else st["bucket_visits"] * 8 + st["streamed"] * 2
                    unit = "bucket visit"
                else:
                    alg = info["bytes_per_candidate"] * st["candidates"]
                    unit = "list entry"
                gbs = alg / (st["scan_ms"] / 1e3) / 1e9 if st["scan_ms"] > 0 else 0.0
                line = {"config": "C5", "slice_width": w, "max_dist": md, "threshold": 0.0, "method": "and", "sites": info["offtargetsCount"],
                        "hbm_layout": layout, "index_hbm_gb": round(info["hbm_bytes"] / 1e9, 2), "index_build_s": round(t_build, 2),
                        "guides": n, "value": n / (st["total_ms"] / 1e3), "unit": "guides/s", "ms_per_100k_guides": st["total_ms"] * 1e5 / n,
                        "e2e": {"value": n / wall, "unit": "guides/s", "how": "issl_score, pinned host guides -> pinned host scores, best of %d" % (a.reps - 1)},
                        "guides_10k": {"guides": a.small_guides, "e2e": a.small_guides / min(small), "unit": "guides/s"},
                        "hits_per_guide": st["hits"] / n, "candidates_per_guide_Cg": st["candidates"] / n,
                        "bucket_visits_per_guide": st["bucket_visits"] / n, "entries_read_per_guide": st["streamed"] / n,
                        "heavy_hits_per_guide": st["heavy_hits"] / n, "sorted_hits_per_guide": st["sorted_hits"] / n,
                        "gpu_launches": st["launches"],
                        "roofline": {"bound": "hbm", "unit_of_work": unit, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                     "peak_source": peak_src, "kernel_ms": st["scan_ms"], "kernel_share_of_call": st["scan_ms"] / st["total_ms"],
                                     "reference_equivalent_GBps": 4 * st["candidates"] / (st["scan_ms"] / 1e3) / 1e9 if st["scan_ms"] > 0 else None}}
                if exe:
                    # sample sized for ~cpu_seconds of reference scoring: the reference walks C_g list entries per guide at
                    # roughly 1.6e8 entries/s per core (first point), then from the previous point's own rate
                    if per_guide_s is None:
                        per_guide_s = st["candidates"] / n / (1.6e8 * cores)
                    k = int(max(32, min(1600, a.cpu_seconds / per_guide_s)))
                    bench.write_guide_file(gpath, guides[:k])
                    t_all, p = run_reference(exe, issl, gpath, md, env)
                    threads = cores
                    if p.returncode != 0 or not p.stdout:
                        # the reference's score-table lookup inserts missing masks into a SHARED hash map from inside the
                        # OpenMP region (isslScoreOfftargets.cpp:394, operator[]): with distances the table was not built for
                        # (w = 8: maxDist 5, w = 10: maxDist >= 4) concurrent insertions crash it; one thread is safe
                        line["reference_multithreaded"] = f"exit code {p.returncode} with {cores} OpenMP threads ({p.stderr.decode()[-120:].strip()!r}); rerun with OMP_NUM_THREADS=1"
                        # ~6 s of single-threaded scoring, so that the noise of the index load time (+-0.5 s of ~15 s) stays small
                        k, threads = int(max(24, min(400, 6.0 / (per_guide_s * cores)))), 1
                        bench.write_guide_file(gpath, guides[:k])
                        t_all, p = run_reference(exe, issl, gpath, md, dict(env, OMP_NUM_THREADS="1"))
                    t = max(t_all - t_load, 1e-6)
                    if threads == cores:
                        per_guide_s = t / k
                    ours = cb.format_lines(guides[:k], hm.array[:k], hc.array[:k], "and").splitlines()
                    theirs = p.stdout.splitlines()
                    same = sum(x == y for x, y in zip(ours, theirs)) if p.returncode == 0 else 0
                    line["cpu_baseline"] = {"value": k / t, "unit": "guides/s", "cores": threads, "kind": "reference",
                                            "sample": f"oracle/_ref/isslScoreOfftargets on the same index ({os.path.getsize(issl) / 1e9:.1f} GB .issl), first {k} guides, "
                                                      f"{threads} OpenMP thread(s) ({bench.cpu_model()}); scoring {t:.2f} s = wall {t_all:.1f} s minus {t_load:.1f} s index load"}
                    line["parity"] = f"{same}/{len(theirs)} output lines byte-identical to the reference's stdout at full index size"
                    line["speedup_e2e_vs_reference"] = (n / wall) / (k / t)
                out.write(json.dumps(line) + "\n")
                out.flush()
                print(json.dumps({k2: line[k2] for k2 in ("slice_width", "max_dist", "value", "hits_per_guide", "parity") if k2 in line}), file=sys.stderr)
        for b in (hg, hm, hc):
            b.close()
        dev.close()
    out.close()


if __name__ == "__main__":
    main()
