#!/usr/bin/env python3
"""Config 1 of BASELINE.json, produced end to end by the UNMODIFIED reference tools:
synthetic 5 Mbp genome -> reference extractOfftargets -> reference isslCreateIndex (l=20, w=8)
-> 10 000 guides -> reference isslScoreOfftargets (MIT, maxDist 4; also `and`, and threshold 75).
Only hashes and a few head lines are committed (the site list is 19.7 MB, the .issl 45 MB);
tests regenerate the inputs from the seeds with the oracle's restatements and compare hashes.

Run in the build container: needs /root/reference (Python extractor) and oracle/_ref.
"""
import hashlib, json, os, pathlib, subprocess, sys, tempfile

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import issl_testdata as td            # noqa: E402
from oracle import oracle, extract    # noqa: E402

GENOME_SEED, GUIDE_SEED, N_GUIDES = 20261018, 9, 10_000
RUNS = [(4, 0, "mit"), (4, 75, "mit"), (4, 0, "and"), (4, 75, "and"), (3, 0, "cfd")]


def sha(b): return hashlib.sha256(b).hexdigest()


def main():
    records = td.make_genome(GENOME_SEED)
    guides = td.sample_guides(GUIDE_SEED, extract.guide_candidates(records), N_GUIDES)
    with tempfile.TemporaryDirectory() as tmp:
        td.write_fasta(f"{tmp}/genome.fa", records)
        env = dict(os.environ, PYTHONPATH="/root/reference/src")
        subprocess.run([sys.executable, "-m", "crackling.utils.extractOfftargets", f"{tmp}/sites.txt", f"{tmp}/genome.fa"],
                       check=True, env=env, stdout=subprocess.DEVNULL)
        sites = pathlib.Path(f"{tmp}/sites.txt").read_bytes()
        assert extract.extract_offtargets(records) == sites, "oracle extraction differs from the reference tool"
        oracle.ref_create_index(f"{tmp}/sites.txt", 20, 8, f"{tmp}/index.issl")
        img = pathlib.Path(f"{tmp}/index.issl").read_bytes()
        assert oracle.create_index(sites, 20, 8) == img
        pathlib.Path(f"{tmp}/guides.txt").write_bytes(guides)
        exp = {"genome_seed": GENOME_SEED, "guide_seed": GUIDE_SEED, "n_guides": N_GUIDES,
               "sites_sha256": sha(sites), "sites_lines": len(sites) // 21, "issl_sha256": sha(img), "issl_bytes": len(img),
               "header": oracle.header(img), "guides_sha256": sha(guides), "runs": []}
        for md, thr, m in RUNS:
            rc, out = oracle.ref_score(f"{tmp}/index.issl", f"{tmp}/guides.txt", md, thr, m)
            exp["runs"].append({"maxDist": md, "threshold": thr, "method": m, "returncode": rc, "stdout_sha256": sha(out),
                                "lines": out.count(b"\n"), "head": out.decode().splitlines()[:5]})
    (HERE / "c1").mkdir(exist_ok=True)
    (HERE / "c1" / "expected.json").write_text(json.dumps(exp, indent=1))
    print(json.dumps({k: exp[k] for k in ("sites_lines", "issl_bytes", "header")}))


if __name__ == "__main__":
    main()
