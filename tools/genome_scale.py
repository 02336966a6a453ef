#!/usr/bin/env python3
"""genome_scale.py -- the producers at genome scale: a seeded random genome (default 3.1 Gbp
in two FASTA records, 60 bases per line, the shape /root/reference/src/crackling/utils/extractOfftargets.py:64-109 reads)
-> device-side extractOfftargets (issl_sites_add_fasta: both strands, the tool's regexes, sort) -> device-side
isslCreateIndex (issl_device_create_from_sites, TRIPLE layout) -> 100 000 guides scored.  Reports Mbp/s of the
extraction, the index build time, and the site statistics next to those of the i.i.d. site generator bench.py uses
(issl_device_create_synthetic): distinct fraction, entries per visited bucket, hits per guide, guides/s.

    python tools/genome_scale.py [--bp 3100000000] [--guides 100000] > gpurun_out/genome_scale.json
"""
import argparse
import ctypes as C
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import crackling_b200 as cb  # noqa: E402


def write_genome(path: str, bp: int, seed: int, line: int = 60) -> float:
    """Two records (97 % and 3 % of the bases; the reference tool's single-file path needs at least two)."""
    t0 = time.perf_counter()
    rng = np.random.Generator(np.random.PCG64(seed))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for name, n in (("chrA", bp - bp // 32), ("chrB", bp // 32)):
            f.write(f">{name} synthetic, seed {seed}\n".encode())
            left = n
            while left:
                rows = min(left // line, 4_000_000)
                if rows == 0:
                    f.write(lut[rng.integers(0, 4, left, dtype=np.uint8)].tobytes() + b"\n")
                    break
                block = np.empty((rows, line + 1), dtype=np.uint8)
                block[:, :line] = lut[rng.integers(0, 4, rows * line, dtype=np.uint8)].reshape(rows, line)
                block[:, line] = 10
                block.tofile(f)
                left -= rows * line
    return time.perf_counter() - t0


def score_stats(dev, guides, label):
    hg, hm, hc = cb.HostBuffer(guides.size, np.uint64), cb.HostBuffer(guides.size, np.float64), cb.HostBuffer(guides.size, np.float64)
    hg.array[:] = guides
    best, reps = None, []
    for _ in range(4):
        dev.score_into(hg.array, 4, 0.0, "and", hm.array, hc.array)
        st = dev.stats
        reps.append({"scan_ms": round(st["scan_ms"], 3), "total_ms": round(st["total_ms"], 3), "launches": st["launches"], "sorted_hits": st["sorted_hits"], "heavy_hits": st["heavy_hits"]})
        if best is None or st["scan_ms"] < best["scan_ms"]:
            best = st
    n = guides.size
    info = dev.info
    out = {"index": label, "sites_distinct": info["offtargetsCount"], "sites_total": info["seqCount"],
           "distinct_fraction": info["offtargetsCount"] / max(info["seqCount"], 1), "index_hbm_gb": round(info["hbm_bytes"] / 1e9, 2),
           "guides": n, "guides_per_s": n / (best["total_ms"] / 1e3), "scan_ms": best["scan_ms"], "hits_per_guide": best["hits"] / n,
           "bucket_visits_per_guide": best["bucket_visits"] / n, "entries_per_visited_bucket": best["streamed"] / max(best["bucket_visits"], 1),
           "candidates_per_guide": best["candidates"] / n, "reps": reps, "checksum": float(hm.array.sum() + hc.array.sum())}
    ll = dev.list_lengths.astype(np.float64)
    ll = ll[ll > 0]
    out["list_length_max_over_mean"] = float(ll.max() / ll.mean())
    for b in (hg, hm, hc):
        b.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=3_100_000_000)
    ap.add_argument("--guides", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--scratch", default="/dev/shm")
    ap.add_argument("--skip-iid", action="store_true")
    a = ap.parse_args()
    path = os.path.join(a.scratch, f"genome_{os.getpid()}.fa")
    out = {"genome_bp": a.bp, "records": 2, "line_length": 60}
    try:
        out["write_fasta_s"] = round(write_genome(path, a.bp, a.seed), 2)
        out["fasta_gb"] = os.path.getsize(path) / 1e9
        text = np.memmap(path, dtype=np.uint8, mode="r")
        sites = cb.Sites(0)
        t0 = time.perf_counter()
        cb.binding._check(cb.lib().issl_sites_add_fasta(sites._h, C.cast(text.ctypes.data, C.c_char_p), text.size, 1))
        t_extract = time.perf_counter() - t0
        n_sites = sites.count
        out["extract"] = {"seconds": round(t_extract, 3), "Mbp_per_s": a.bp / 1e6 / t_extract, "sites": n_sites, "characters": sites.characters,
                          "sites_per_bp": n_sites / a.bp, "expected_sites_per_bp_iid": 2 * 3 / 64,
                          "what": "issl_sites_add_fasta: host text (page cache) -> HBM in chunks, both strands matched on the device "
                                  "(ref extractOfftargets.py:23-24, :64-109), keys kept in HBM"}
        del text
        t0 = time.perf_counter()
        dev = cb.Device.from_sites(sites, 8, "auto")
        out["index_from_sites_s"] = round(time.perf_counter() - t0, 3)
        out["index_from_sites"] = "issl_device_create_from_sites: sort, collapse into occurrence counts, slice lists, ten sub-bucket copies + blocks (ref isslCreateIndex.cpp:184-252)"
        sites.close()
        guides = bench.make_guides(dev, a.guides, seed=2)
        out["from_genome"] = score_stats(dev, guides, "random genome -> extractOfftargets -> index, all on the device")
        dev.close()
    finally:
        if os.path.exists(path):
            os.unlink(path)
    if a.skip_iid:
        print(json.dumps(out))
        return
    t0 = time.perf_counter()
    dev = cb.Device.synthetic(0, "auto", seed=1, uniform_sites=bench.HUMAN_SITES)
    out["synthetic_build_s"] = round(time.perf_counter() - t0, 3)
    out["iid_sites"] = score_stats(dev, bench.make_guides(dev, a.guides, seed=2), "bench.py's generator: 581.25 M i.i.d. sites, first base A/C/G")
    dev.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
