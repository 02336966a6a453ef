#!/usr/bin/env python3
"""Full-size differential soak on the GPU box: the CUDA path against the oracle (oracle/issl_oracle.c, the CPU
restatement pinned to the compiled reference) on a human-scale synthetic index with planted repeat families, over
methods x thresholds x maxDist.  Raw fp64 scores must be bit-identical; early exits must agree.  Test infrastructure:
the oracle is used as the checker only.

    python tools/soak_parity.py [--sites N] [--guides G] > gpurun_out/soak.json
"""
import argparse, json, os, pathlib, sys, tempfile, time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import crackling_b200 as cb          # noqa: E402
from bench import make_guides        # noqa: E402
from oracle import oracle            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=560_000_000)
    ap.add_argument("--families", type=int, default=3000)
    ap.add_argument("--family-size", type=int, default=700)
    ap.add_argument("--guides", type=int, default=4000, help="batch scored by the GPU (>= 2368 so that every guide owns a CTA)")
    ap.add_argument("--check", type=int, default=320, help="guides of the batch also scored by the oracle, per combination")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--slice-width", type=int, default=8, help="4 and 10 as well: the oracle walks 31x more list entries per guide at w = 4, so use --check 32")
    ap.add_argument("--dists", default="2,3,4,5")
    ap.add_argument("--methods", default="and,or,avg,mit,cfd")
    ap.add_argument("--thresholds", default="0,50,75,90")
    a = ap.parse_args()
    dev = cb.Device.synthetic(0, "auto", seed=a.seed, uniform_sites=a.sites, families=a.families, family_size=a.family_size,
                              max_sub_rate=0.12, slice_width=a.slice_width)
    guides = make_guides(dev, a.guides, seed=11, families=a.families, family_frac=0.3, index_seed=a.seed)
    out = {"sites": dev.info["offtargetsCount"], "slice_width": a.slice_width, "layout": dev.info["layout"], "guides": int(guides.size), "checked_per_combo": a.check,
           "combos": [], "mismatches": 0}
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        path = os.path.join(tmp, "index.issl")
        dev.write_issl(path)
        img = np.fromfile(path, dtype=np.uint8)
        os.unlink(path)
        rng = np.random.default_rng(3)
        t0 = time.perf_counter()
        for md in [int(x) for x in a.dists.split(",")]:
            for method in a.methods.split(","):
                for thr in [float(x) for x in a.thresholds.split(",")]:
                    mit, cfd = dev.score(guides, md, thr, method)
                    st = dev.stats
                    pick = np.sort(rng.choice(guides.size, a.check, replace=False))
                    want = oracle.score(img, guides[pick], md, thr, method, threads=0)
                    bad = 0
                    if mit is not None:
                        bad += int((mit[pick].view(np.uint64) != want["mit"].view(np.uint64)).sum())
                    if cfd is not None:
                        bad += int((cfd[pick].view(np.uint64) != want["cfd"].view(np.uint64)).sum())
                    out["mismatches"] += bad
                    out["combos"].append({"maxDist": md, "method": method, "threshold": thr, "hits_per_guide": st["hits"] / guides.size,
                                          "early_exits": st["early_exits"], "heavy_hits": st["heavy_hits"], "bit_mismatches": bad})
        out["seconds"] = time.perf_counter() - t0
    dev.close()
    print(json.dumps(out))
    return 1 if out["mismatches"] else 0


if __name__ == "__main__":
    sys.exit(main())
