#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs oracle/_ref, i.e. `make -C oracle ref`, which compiles
/root/reference/src/ISSL/*.cpp with the reference Makefile's flags).  For every case it writes
    tests/golden/<case>/offtargets.txt   sorted off-target text fed to the reference isslCreateIndex
    tests/golden/<case>/guides.txt       guide file fed to the reference isslScoreOfftargets
    tests/golden/<case>/expected.json    sha256 + header of the reference-built .issl, the
                                         reference's stdout for every (maxDist, threshold, method)
                                         run, and the hit tuples (guide, id, dist, occ) dumped by
                                         the instrumented reference build for a subset of runs
The .issl images themselves are not committed (the w=4/w=5 score tables alone are megabytes);
tests rebuild them with the oracle's restatement of isslCreateIndex and check the sha256.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import pathlib
import sys
import tempfile

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import issl_testdata as td          # noqa: E402
from oracle import oracle           # noqa: E402

METHODS = ["mit", "cfd", "and", "or", "avg", "bogus"]

CASES = {
    # name: (offtarget kwargs, guide kwargs, seq_length, slice_width, maxDists, thresholds, hit-dump runs)
    "w8_families": (dict(seed=101, n_random=3000, n_families=16, family_size=60), dict(seed=201, n=96), 20, 8,
                    [0, 1, 2, 3, 4, 5], [0, 75, 50.5], [(4, 0, "and"), (4, 75, "and"), (2, 0, "mit"), (5, 0, "cfd")]),
    "w10_truncated": (dict(seed=102, n_random=2500, n_families=10, family_size=50), dict(seed=202, n=64), 20, 10,
                      [0, 2, 3, 4], [0, 75], [(3, 0, "and"), (4, 0, "and")]),
    "w4_wide": (dict(seed=103, n_random=600, n_families=6, family_size=30), dict(seed=203, n=32), 20, 4,
                [2, 4, 5], [0, 75], [(4, 0, "and")]),
    "w5_unaligned": (dict(seed=104, n_random=800, n_families=6, family_size=30), dict(seed=204, n=32), 20, 5,
                     [3, 4], [0, 60], [(4, 0, "and")]),
    "dense_repeats": (dict(seed=105, n_random=200, n_families=4, family_size=400, max_sub_rate=0.08, dup_frac=0.3),
                      dict(seed=205, n=48, frac_exact=0.6, frac_mut=0.35), 20, 8,
                      [4], [0, 1, 20, 75, 99, -1, 100, 150], [(4, 0, "and"), (4, 20, "or"), (4, 20, "avg")]),
    "tiny": (dict(seed=106, n_random=3, n_families=0, family_size=0, dup_frac=0.0, low_complexity=0),
             dict(seed=206, n=5, frac_exact=0.8, frac_mut=0.2), 20, 8, [4], [0], [(4, 0, "and")]),
}


def main() -> int:
    if not oracle.have_ref():
        print("oracle/_ref is missing: run `make -C oracle ref` in the build container", file=sys.stderr)
        return 1
    for name, (okw, gkw, L, w, dists, thresholds, hitruns) in CASES.items():
        d = HERE / name
        d.mkdir(exist_ok=True)
        text = td.make_offtargets(seq_length=L, **okw)
        guides = td.make_guides(offtargets=text, seq_length=L, **gkw)
        (d / "offtargets.txt").write_bytes(text)
        (d / "guides.txt").write_bytes(guides)
        with tempfile.TemporaryDirectory() as tmp:
            issl = f"{tmp}/index.issl"
            oracle.ref_create_index(str(d / "offtargets.txt"), L, w, issl)
            img = pathlib.Path(issl).read_bytes()
            exp = {
                "seq_length": L, "slice_width": w,
                "issl_sha256": hashlib.sha256(img).hexdigest(), "issl_bytes": len(img),
                "header": oracle.header(img),
                "runs": [], "hits": [],
            }
            for md in dists:
                for thr in thresholds:
                    for m in METHODS:
                        rc, out = oracle.ref_score(issl, str(d / "guides.txt"), md, thr, m, threads=3)
                        exp["runs"].append({"maxDist": md, "threshold": thr, "method": m, "returncode": rc,
                                            "stdout": out.decode()})
            for md, thr, m in hitruns:
                rc, out, hits = oracle.ref_score(issl, str(d / "guides.txt"), md, thr, m, threads=1, hits=True)
                exp["hits"].append({"maxDist": md, "threshold": thr, "method": m,
                                    "tuples": hits.tolist()})
        (d / "expected.json").write_text(json.dumps(exp, indent=0))
        nh = sum(len(h["tuples"]) for h in exp["hits"])
        print(f"{name}: sites={exp['header']['offtargetsCount']} runs={len(exp['runs'])} hit tuples={nh}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
