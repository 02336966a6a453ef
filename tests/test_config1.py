"""BASELINE.json config 1 at full size: synthetic 5 Mbp genome -> extractOfftargets ->
isslCreateIndex (l=20, w=8) -> 10 000 guides, MIT maxDist 4.  tests/golden/c1/expected.json holds
the hashes of what the UNMODIFIED reference tools produced for these seeds
(tests/golden/make_golden_c1.py); inputs are regenerated here from the seeds.

CPU: the oracle's restatements reproduce the reference's site list, .issl image and stdout.
GPU: the drop-in executable prints byte-identical stdout for every run."""
import hashlib
import json
import pathlib
import subprocess

import pytest

import crackling_b200 as cb
import issl_testdata as td
from oracle import extract, oracle

EXP = json.loads((pathlib.Path(__file__).parent / "golden" / "c1" / "expected.json").read_text())
_cache = {}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def inputs():
    if not _cache:
        records = td.make_genome(EXP["genome_seed"])
        sites = extract.extract_offtargets(records)
        guides = td.sample_guides(EXP["guide_seed"], extract.guide_candidates(records), EXP["n_guides"])
        _cache.update(sites=sites, guides=guides, img=oracle.create_index(sites, 20, 8))
    return _cache


def test_oracle_reproduces_reference_tools_on_config1():
    c = inputs()
    assert len(c["sites"]) // 21 == EXP["sites_lines"] and sha(c["sites"]) == EXP["sites_sha256"]
    assert sha(c["guides"]) == EXP["guides_sha256"]
    assert len(c["img"]) == EXP["issl_bytes"] and sha(c["img"]) == EXP["issl_sha256"]
    run = EXP["runs"][0]                                  # MIT, maxDist 4, threshold 0: the config as named
    rc, out = oracle.cli(c["img"], c["guides"], run["maxDist"], run["threshold"], run["method"], threads=0)
    assert rc == 0 and out.count(b"\n") == run["lines"] and sha(out) == run["stdout_sha256"]


@pytest.mark.gpu
def test_host_program_matches_reference_on_config1(tmp_path):
    c = inputs()
    (tmp_path / "index.issl").write_bytes(c["img"])
    (tmp_path / "guides.txt").write_bytes(c["guides"])
    for run in EXP["runs"]:
        p = subprocess.run([str(cb.cli_path()), str(tmp_path / "index.issl"), str(tmp_path / "guides.txt"),
                            str(run["maxDist"]), str(run["threshold"]), run["method"]],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == run["returncode"], p.stderr
        assert p.stdout.decode().splitlines()[:5] == run["head"]
        assert sha(p.stdout) == run["stdout_sha256"], (run["maxDist"], run["threshold"], run["method"])
