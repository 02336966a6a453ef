"""Guide-side pre-filters of the pipeline: the oracle against the reference-made vectors
(CPU), the device pass against the oracle (GPU)."""
import json
import pathlib

import numpy as np
import pytest

import crackling_b200 as cb
from oracle import prefilters as pf

GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "prefilters.json").read_text())


def test_oracle_matches_reference_vectors():
    assert len(GOLDEN) >= 400
    seen = 0
    for v in GOLDEN:
        f, at = pf.filters(v["target23"])
        assert f == v["flags"] and at == v["at"]
        seen |= f
    assert seen == 15, "every filter must be exercised by the vectors"


def test_oracle_at_percentage_matches_reference_helper():
    ref = pathlib.Path("/root/reference/src/crackling/Helpers.py")
    if not ref.exists():
        pytest.skip("reference tree not present (GPU box)")
    import sys
    sys.path.insert(0, "/root/reference/src")
    from crackling.Helpers import AT_percentage
    rng = np.random.default_rng(5)
    for _ in range(500):
        s = "".join(rng.choice(list("ACGTN"), 20))
        assert pf.at_percentage(s) == AT_percentage(s)


@pytest.mark.gpu
def test_device_filters_match_oracle():
    from conftest import golden_case
    case = golden_case("w8_families")
    dev = cb.Device.from_index(cb.Index(case.issl), 0, "res32")
    rng = np.random.default_rng(6)
    targets = [v["target23"] for v in GOLDEN]
    targets += ["".join(rng.choice(list("ACGT"), 23)) for _ in range(5000)]
    text = ("\n".join(targets) + "\n").encode()
    flags, at, packed = dev.guide_filters(text)
    for i, t in enumerate(targets):
        f, a = pf.filters(t)
        assert flags[i] == f and at[i] == a, t
        assert int(packed[i]) == pf.encode20(t)
    assert np.array_equal(packed, cb.pack_guides(("\n".join(t[:20] for t in targets) + "\n").encode()))
    # the survivors go straight to the scorer, no temp file in between (ref Crackling.py:747-778)
    keep = packed[flags == 0][:64]
    mit, cfd = dev.score(keep, 4, 0, "and")
    assert mit.size == keep.size and np.all(mit > 0) and np.all(cfd > 0)
    with pytest.raises(cb.IsslError):
        dev.guide_filters(text[:-1])
    dev.close()


DUP_GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "duplicates.json").read_text())


def test_oracle_duplicates_match_reference_loop():
    """oracle.prefilters.duplicates against what the reference's own loop (Crackling.py:231-241, executed by
    tests/golden/make_duplicates_golden.py) recorded, dropped and called not unique."""
    assert sum(c["numDuplicateGuides"] for c in DUP_GOLDEN) > 2000
    for c in DUP_GOLDEN:
        flags, later, seqs = pf.duplicates(c["targets"])
        assert later == c["numDuplicateGuides"] and seqs == len(c["duplicateGuides"])
        assert [i for i, f in enumerate(flags) if not f & pf.FILTER_DUPLICATE] == c["recorded"]
        dup = set(c["duplicateGuides"])
        assert all(bool(f & pf.FILTER_NOT_UNIQUE) == (t in dup) for f, t in zip(flags, c["targets"]))


@pytest.mark.gpu
def test_device_duplicates_match_oracle():
    """issl_guide_duplicates (pack, stable sort, run heads) against the oracle: the reference's vectors, and 300 000
    targets drawn from a pool small enough that most sequences repeat."""
    from conftest import golden_case
    case = golden_case("w8_families")
    dev = cb.Device.from_index(cb.Index(case.issl), 0, "res32")
    rng = np.random.default_rng(8)
    pool = ["".join(rng.choice(list("ACGT"), 23)) for _ in range(60_000)]
    big = [pool[i] for i in rng.integers(0, len(pool), 300_000)]
    for targets in [c["targets"] for c in DUP_GOLDEN] + [big]:
        text = ("\n".join(targets) + "\n").encode() if targets else b""
        flags, later, seqs = dev.guide_duplicates(text)
        want, wlater, wseqs = pf.duplicates(targets)
        assert (later, seqs) == (wlater, wseqs)
        assert np.array_equal(flags, np.array(want, dtype=np.uint8))
    with pytest.raises(cb.IsslError):
        dev.guide_duplicates(b"ACGT\n")
    dev.close()


@pytest.mark.gpu
def test_pipeline_page_without_temp_files(tmp_path):
    """crackling_b200.pipeline.OfftargetScorer.score_page replaces Crackling.py:737-786 (guide file, subprocess, output
    file): the dictionary it returns is what the pipeline would have parsed from the reference binary's output."""
    import subprocess
    from conftest import golden_case
    from crackling_b200.pipeline import OfftargetScorer, FILTER_DUPLICATE
    case = golden_case("w8_families")
    issl = tmp_path / "index.issl"
    issl.write_bytes(bytes(case.issl))
    rng = np.random.default_rng(9)
    guides20 = [l.decode() for l in case.guides.split(b"\n") if l]
    targets = [g + "".join(rng.choice(list("ACGT"), 1)) + "GG" for g in guides20]
    targets += targets[:7]                                    # duplicates: the pipeline would have dropped these
    with OfftargetScorer(str(issl)) as scorer:
        flags, at, later, seqs = scorer.candidate_flags(targets)
        assert later == 7 and seqs == 7 and np.count_nonzero(flags & FILTER_DUPLICATE) == 7
        page = [t for t, f in zip(targets, flags) if not f & FILTER_DUPLICATE]
        for method, thr in (("and", 0.0), ("mit", 75.0), ("cfd", 50.0)):
            got = scorer.score_page(page, 4, thr, method)
            # what the reference pipeline does: guide file -> binary -> output file -> dict (Crackling.py:747-786)
            gfile, ofile = tmp_path / "guides.txt", tmp_path / "out.txt"
            gfile.write_text("".join(t[0:20] + "\n" for t in page))
            exe = pathlib.Path(__file__).parent.parent / "oracle" / "_ref" / "isslScoreOfftargets"
            if not exe.exists():
                exe = cb.cli_path()                           # (GPU box without the reference build: our own drop-in)
            with open(ofile, "w") as fo:
                subprocess.run([str(exe), str(issl), str(gfile), "4", str(thr), method], stdout=fo, check=True)
            want = {}
            for f in [x.split("\t") for x in ofile.read_text().splitlines(True)]:
                if len(f) == 3:
                    want[f[0]] = {"MIT": float(f[1].strip()), "CFD": float(f[2].strip())}
            assert got == want and len(got) == len(set(t[:20] for t in page))
