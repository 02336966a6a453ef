"""Off-target site extraction: the device-side extractOfftargets against
(a) tests/golden/extract/* -- messy FASTA inputs and what the UNMODIFIED reference tool wrote for them
(tests/golden/make_golden_extract.py), (b) the oracle's restatement on fresh inputs, (c) BASELINE.json
config 1 at full size (tests/golden/c1: the real tool's site list for the 5 Mbp genome)."""
import hashlib
import json
import os
import pathlib
import subprocess

import numpy as np
import pytest

import crackling_b200 as cb
import issl_testdata as td
from oracle import extract, oracle

GOLDEN = pathlib.Path(__file__).parent / "golden" / "extract"
CASES = sorted(p.name for p in GOLDEN.iterdir() if (p / "expected.json").exists())
C1 = json.loads((pathlib.Path(__file__).parent / "golden" / "c1" / "expected.json").read_text())


def sha(b):
    return hashlib.sha256(b).hexdigest()


def load(name):
    exp = json.loads((GOLDEN / name / "expected.json").read_text())
    return exp, [(GOLDEN / name / f).read_bytes() for f in exp["inputs"]]


def keys_to_text(keys: np.ndarray) -> bytes:
    shifts = np.arange(19, -1, -1, dtype=np.uint64) * np.uint64(2)
    codes = ((keys[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    lines = np.concatenate([np.frombuffer(b"ACGT", dtype=np.uint8)[codes], np.full((keys.size, 1), 10, dtype=np.uint8)], axis=1)
    return lines.tobytes()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reads_fasta_like_the_reference_tool(name):
    exp, inputs = load(name)
    out = extract.extract_from_inputs(inputs)
    lines = out.decode().splitlines()
    assert len(lines) == exp["sites_lines"] and len(set(lines)) == exp["distinct"]
    assert lines[:5] == exp["head"] and lines[-5:] == exp["tail"]
    assert sha(out) == exp["sites_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [None, 4096, 97])
@pytest.mark.parametrize("name", CASES)
def test_device_extraction_matches_reference_tool(name, chunk, tmp_path):
    """Whole-buffer and tiny-chunk passes (windows spanning chunk cuts, headers at chunk edges) give the same file."""
    exp, inputs = load(name)
    if chunk:
        os.environ["ISSL_EXTRACT_CHUNK"] = str(chunk)
    try:
        sites = cb.Sites(0)
    finally:
        os.environ.pop("ISSL_EXTRACT_CHUNK", None)
    for raw in inputs:
        sites.add_fasta(raw, single_input=len(inputs) == 1)
    assert sites.count == exp["sites_lines"]
    out = tmp_path / "sites.txt"
    sites.write_text(out)
    got = out.read_bytes()
    assert got.decode().splitlines()[:5] == exp["head"]
    assert sha(got) == exp["sites_sha256"]
    assert keys_to_text(sites.keys()) == got
    # straight into the index, no text file: same .issl as isslCreateIndex on the tool's output
    for w in (8, 10):
        dev = cb.Device.from_sites(sites, w)
        dev.write_issl(tmp_path / "from_sites.issl")
        assert (tmp_path / "from_sites.issl").read_bytes() == oracle.create_index(got, 20, w)
        dev.close()
    sites.close()


@pytest.mark.gpu
def test_extract_program_is_a_drop_in(tmp_path):
    exp, inputs = load("several_files")
    d = tmp_path / "genome"
    d.mkdir()
    for k, raw in enumerate(inputs):
        (d / f"in{k}.fa").write_bytes(raw)
    exe = str(cb.extract_cli_path())
    # a list of files, with the tool's two optional flags
    r = subprocess.run([exe, str(tmp_path / "a.txt"), *[str(d / f"in{k}.fa") for k in range(len(inputs))], "--threads", "4",
                        "--maxOpenFiles=100"], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert sha((tmp_path / "a.txt").read_bytes()) == exp["sites_sha256"]
    assert b"Found 5003 targets" in r.stdout or b"Found %d targets" % exp["sites_lines"] in r.stdout
    # a directory
    r = subprocess.run([exe, str(tmp_path / "b.txt"), str(d)], capture_output=True)
    assert r.returncode == 0 and sha((tmp_path / "b.txt").read_bytes()) == exp["sites_sha256"]
    # one multi-FASTA file, plus the index in the same run
    exp1, in1 = load("single_messy")
    (tmp_path / "one.fa").write_bytes(in1[0])
    r = subprocess.run([exe, str(tmp_path / "c.txt"), str(tmp_path / "one.fa"), "--index", str(tmp_path / "c.issl"), "--slice-width", "8"],
                       capture_output=True)
    assert r.returncode == 0, r.stderr
    text = (tmp_path / "c.txt").read_bytes()
    assert sha(text) == exp1["sites_sha256"]
    assert (tmp_path / "c.issl").read_bytes() == oracle.create_index(text, 20, 8)
    r = subprocess.run([exe, str(tmp_path / "x.txt")], capture_output=True)
    assert r.returncode == 2 and r.stderr.startswith(b"usage:")
    r = subprocess.run([exe, str(tmp_path / "x.txt"), str(tmp_path / "missing.fa")], capture_output=True)
    assert r.returncode == 1 and b"Cannot read" in r.stderr


@pytest.mark.gpu
def test_config1_genome_at_full_size(tmp_path):
    """5 Mbp genome of BASELINE.json config 1: the site list equals the real tool's (sha256 committed), and the
    index built from it on the device equals the real isslCreateIndex's."""
    records = td.make_genome(C1["genome_seed"])
    td.write_fasta(tmp_path / "genome.fa", records)
    raw = (tmp_path / "genome.fa").read_bytes()
    sites = cb.Sites(0)
    sites.add_fasta(raw, single_input=True)
    assert sites.count == C1["sites_lines"] and sites.characters == sum(len(r) for r in records)
    sites.write_text(tmp_path / "sites.txt")
    assert sha((tmp_path / "sites.txt").read_bytes()) == C1["sites_sha256"]
    dev = cb.Device.from_sites(sites, 8)
    dev.write_issl(tmp_path / "index.issl")
    img = (tmp_path / "index.issl").read_bytes()
    assert len(img) == C1["issl_bytes"] and sha(img) == C1["issl_sha256"]
    dev.close()
    sites.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fresh_random_fasta_against_oracle(seed, tmp_path):
    rng = np.random.default_rng(seed)
    recs = []
    for k in range(int(rng.integers(2, 6))):
        n = int(rng.integers(10, 40000))
        s = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .04]).tobytes().decode()
        width = int(rng.integers(1, 120))
        eol = ["\n", "\r\n"][int(rng.integers(0, 2))]
        recs.append(f">r{k}{eol}" + "".join(s[i:i + width] + eol for i in range(0, n, width)))
    raw = "".join(recs).encode()
    want = extract.extract_from_inputs([raw])
    os.environ["ISSL_EXTRACT_CHUNK"] = str(int(rng.integers(64, 5000)))
    try:
        sites = cb.Sites(0)
    finally:
        del os.environ["ISSL_EXTRACT_CHUNK"]
    sites.add_fasta(raw, True)
    sites.write_text(tmp_path / "s.txt")
    assert (tmp_path / "s.txt").read_bytes() == want
    sites.close()
