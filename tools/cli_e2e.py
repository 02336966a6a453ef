#!/usr/bin/env python3
"""Process-level drop-in comparison on the GPU box: the same .issl file and guide file through
(a) bin/isslScoreOfftargets (ours) and (b) oracle/_ref/isslScoreOfftargets (the unmodified reference).

Builds the synthetic index in HBM, serialises it to a real .issl in /dev/shm (the reference's byte
format), writes the guide file, runs both executables the way Crackling.py:767-778 does
(`binary issl guides maxDist threshold method > out`), and diffs their stdout.

    python tools/cli_e2e.py [--sites N] [--guides G] [--ref-guides R] [--threshold T] [--gpus K]
"""
import argparse, json, os, pathlib, subprocess, sys, tempfile, time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import crackling_b200 as cb          # noqa: E402
from bench import make_guides        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=581_250_000)
    ap.add_argument("--guides", type=int, default=100_000)
    ap.add_argument("--ref-guides", type=int, default=1600)
    ap.add_argument("--threshold", default="0")
    ap.add_argument("--method", default="and")
    ap.add_argument("--max-dist", type=int, default=4)
    ap.add_argument("--slice-width", type=int, default=8)
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--family-size", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    out = {"sites": a.sites, "guides": a.guides, "threshold": a.threshold, "method": a.method, "gpus": a.gpus}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        dev = cb.Device.synthetic(0, "auto", seed=1, uniform_sites=a.sites, families=a.families, family_size=a.family_size,
                                  slice_width=a.slice_width)
        guides = make_guides(dev, a.guides, seed=2)
        t0 = time.perf_counter()
        dev.write_issl(f"{tmp}/index.issl")
        out["write_issl_s"] = time.perf_counter() - t0
        out["issl_gb"] = os.path.getsize(f"{tmp}/index.issl") / 1e9
        dev.close()
        text = b"".join(cb.unpack_guide(int(s)).encode() + b"\n" for s in guides)
        pathlib.Path(f"{tmp}/guides.txt").write_bytes(text)
        pathlib.Path(f"{tmp}/guides_ref.txt").write_bytes(text[:21 * a.ref_guides])
        args = [str(a.max_dist), a.threshold, a.method]
        env = dict(os.environ, ISSL_TIMING="1", ISSL_GPUS=str(a.gpus))
        t0 = time.perf_counter()
        p = subprocess.run([str(cb.cli_path()), f"{tmp}/index.issl", f"{tmp}/guides.txt", *args], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=env)
        out["ours_wall_s"] = time.perf_counter() - t0
        out["ours_rc"] = p.returncode
        out["ours_stderr"] = p.stderr.decode()[-1500:]
        ours = p.stdout
        ref = pathlib.Path(ROOT / "oracle" / "_ref" / "isslScoreOfftargets")
        if ref.exists() and a.ref_guides:
            t0 = time.perf_counter()
            q = subprocess.run([str(ref), f"{tmp}/index.issl", f"{tmp}/guides_ref.txt", *args], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE)
            out["ref_wall_s"] = time.perf_counter() - t0
            out["ref_guides"] = a.ref_guides
            theirs = q.stdout.splitlines()
            mine = ours.splitlines()[:len(theirs)]
            out["identical_lines"] = sum(x == y for x, y in zip(mine, theirs))
            out["compared_lines"] = len(theirs)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
