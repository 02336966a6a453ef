"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
and through the drop-in host program, against (a) the committed output of the unmodified
reference and (b) the oracle on fresh seeded inputs.  Bars: stdout byte-identical; hit tuples
identical including order; raw fp64 scores bit-identical (the accumulation order of the reference
is reproduced), which is stronger than the 1e-9 relative tolerance BASELINE.json asks for."""
import os
import subprocess

import numpy as np
import pytest

import crackling_b200 as cb
import issl_testdata as td
from conftest import GOLDEN_CASES, golden_case
from oracle import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9   # BASELINE.json north_star: <= 1e-9 relative on aggregated MIT/CFD scores


def fmt_stdout(guides_packed, mit, cfd, seq_length=20):
    """ref isslScoreOfftargets.cpp:514-527"""
    lines = []
    for i, s in enumerate(guides_packed):
        a = "%f" % mit[i] if mit is not None else "-1"
        b = "%f" % cfd[i] if cfd is not None else "-1"
        lines.append(f"{cb.unpack_guide(int(s), seq_length)}\t{a}\t{b}\n")
    return "".join(lines)


def layouts_for(case):
    w, L = case.slice_width, case.seq_length
    res32 = w % 2 == 0 and 2 * L - min(w, 8) <= 32
    return (["res32"] if res32 else []) + ["sig64", "gather"]


_devices = {}


def device_for(name, layout):
    key = (name, layout)
    if key not in _devices:
        case = golden_case(name)
        _devices[key] = cb.Device.from_index(cb.Index(case.issl), 0, layout)
    return _devices[key]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_stdout_matches_reference_all_layouts(name):
    case = golden_case(name)
    guides = cb.pack_guides(case.guides, case.seq_length)
    for layout in layouts_for(case):
        dev = device_for(name, layout)
        assert dev.info["layout"] == cb.LAYOUTS[layout]
        for run in case.expected["runs"]:
            mit, cfd = dev.score(guides, run["maxDist"], run["threshold"], run["method"])
            got = fmt_stdout(guides, mit, cfd, case.seq_length)
            assert got == run["stdout"], (name, layout, run["maxDist"], run["threshold"], run["method"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hit_tuples_match_instrumented_reference(name):
    case = golden_case(name)
    guides = cb.pack_guides(case.guides, case.seq_length)
    for layout in layouts_for(case):
        dev = device_for(name, layout)
        for h in case.expected["hits"]:
            _, _, hits = dev.score_hits(guides, h["maxDist"], h["threshold"], h["method"])
            want = np.array(h["tuples"], dtype=np.int64).reshape(-1, 4)
            assert np.array_equal(hits, want), (name, layout, h["maxDist"], h["threshold"], h["method"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_raw_scores_bit_identical_to_oracle(name):
    case = golden_case(name)
    guides = cb.pack_guides(case.guides, case.seq_length)
    dev = device_for(name, "auto")
    for method in ("mit", "cfd", "and", "or", "avg"):
        for thr in (0, 75, 33.3):
            for md in (0, 2, 4, 6):
                want = oracle.score(case.issl, guides, md, thr, method, threads=1)
                mit, cfd = dev.score(guides, md, thr, method)
                for got, ref in ((mit, want["mit"]), (cfd, want["cfd"])):
                    if got is None:
                        continue
                    assert np.allclose(got, ref, rtol=REL_TOL, atol=0)
                    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (name, method, thr, md)


@pytest.mark.parametrize("name", ["w8_families", "w10_truncated", "w5_unaligned"])
def test_host_program_stdout_byte_identical(name, tmp_path):
    case = golden_case(name)
    (tmp_path / "i.issl").write_bytes(case.issl)
    (tmp_path / "g.txt").write_bytes(case.guides)
    for run in case.expected["runs"][::5]:
        p = subprocess.run([str(cb.cli_path()), str(tmp_path / "i.issl"), str(tmp_path / "g.txt"), str(run["maxDist"]),
                            str(run["threshold"]), run["method"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == run["returncode"], p.stderr
        assert p.stdout.decode() == run["stdout"]


def test_host_program_on_several_gpus(tmp_path):
    """Index replicated per GPU, guides split into contiguous ranges, output in input order."""
    n = cb.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    case = golden_case("w8_families")
    (tmp_path / "i.issl").write_bytes(case.issl)
    (tmp_path / "g.txt").write_bytes(case.guides)
    for run in case.expected["runs"][::17]:
        env = dict(os.environ, ISSL_GPUS=str(n), ISSL_TIMING="1")
        p = subprocess.run([str(cb.cli_path()), str(tmp_path / "i.issl"), str(tmp_path / "g.txt"), str(run["maxDist"]),
                            str(run["threshold"]), run["method"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert p.returncode == 0, p.stderr
        assert p.stdout.decode() == run["stdout"]
        if run["method"] != "bogus":
            assert p.stderr.decode().count("[issl] gpu ") == n


@pytest.mark.parametrize("seed,w,n_random,families,fsize", [(11, 8, 200_000, 40, 300), (12, 10, 120_000, 20, 200),
                                                            (13, 4, 20_000, 10, 100), (14, 8, 1, 0, 0)])
def test_fresh_inputs_against_oracle(seed, w, n_random, families, fsize):
    text = td.make_offtargets(seed, n_random=n_random, n_families=families, family_size=fsize)
    guides_txt = td.make_guides(seed + 100, text, n=1500)
    img = oracle.create_index(text, 20, w)
    guides = cb.pack_guides(guides_txt)
    dev = cb.Device.from_index(cb.Index(img), 0, "auto")
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("or", 40, 3), ("avg", 60, 5), ("mit", 0, 2), ("cfd", 90, 4)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=(thr == 0))
        mit, cfd = dev.score(guides, md, thr, method)
        if mit is not None:
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64))
        if cfd is not None:
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64))
        if thr == 0:   # the unit of work: list entries visited (isslScoreOfftargets.cpp:344)
            assert dev.stats["candidates"] == int(want["candidates"].sum())
    mit, cfd, hits = dev.score_hits(guides, 4, 0, "and")
    want = oracle.score(img, guides, 4, 0, "and", threads=1, want_hits=True)["hits"]
    assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1))
    dev.close()


@pytest.mark.parametrize("max_group", [1, 4, 8, 32])
@pytest.mark.parametrize("w", [8, 10])
def test_guide_groups_and_bitsliced_path(max_group, w):
    """Many guides per slice list: exercises the register groups (2/4/8) and, for max_group 32, the
    bit-sliced blocks of 9..32 guides; every grouping must give the oracle's exact answer."""
    text = td.make_offtargets(31, n_random=60_000, n_families=30, family_size=400, max_sub_rate=0.12)
    img = oracle.create_index(text, 20, w)
    rng = np.random.default_rng(32)
    roots = td.pack_guides(td.make_guides(33, text, n=40, frac_exact=1.0, frac_mut=0.0))
    guides = []
    for r in roots:                                   # 75 variants per root with 0-3 substitutions
        for _ in range(75):
            g = int(r)
            for pos in rng.choice(20, size=int(rng.integers(0, 4)), replace=False):
                g ^= int(rng.integers(1, 4)) << (2 * int(pos))
            guides.append(g)
    guides = np.array(guides, dtype=np.uint64)
    os.environ["ISSL_MAX_GROUP"] = str(max_group)
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "auto")
    finally:
        del os.environ["ISSL_MAX_GROUP"]
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("avg", 50, 3), ("or", 0, 0), ("and", 0, 5)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (max_group, method, thr, md)
        assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (max_group, method, thr, md)
        if thr == 0:
            st = dev.stats
            assert st["candidates"] == int(want["candidates"].sum())
            if max_group == 1:
                assert st["streamed"] == st["candidates"]
            elif max_group == 32 and md <= 4:
                assert st["candidates"] > 4 * st["streamed"]          # blocks of up to 32 guides share a chunk
            else:
                assert st["candidates"] > 2 * st["streamed"]
    _, _, hits = dev.score_hits(guides, 4, 0, "and")
    want = oracle.score(img, guides, 4, 0, "and", threads=1, want_hits=True)["hits"]
    assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1))
    dev.close()


def test_edge_cases():
    case = golden_case("w8_families")
    dev = device_for("w8_families", "auto")
    guides = cb.pack_guides(case.guides)
    # zero guides; one guide; maxDist < 0 (nothing can match); maxDist 20 (everything matches -> survivor buffer grows)
    mit, cfd = dev.score(guides[:0], 4, 0, "and")
    assert mit.size == 0 and cfd.size == 0
    for md in (-1, 20):
        want = oracle.score(case.issl, guides[:7], md, 0, "and")
        mit, cfd = dev.score(guides[:7], md, 0, "and")
        assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64))
        assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64))
    # duplicates in the batch are scored independently (ref: every line is scored)
    dup = np.concatenate([guides[:5], guides[:5]])
    mit, _ = dev.score(dup, 4, 0, "mit")
    assert np.array_equal(mit[:5], mit[5:])
    # NaN / negative / huge thresholds follow the same arithmetic as the reference (maximum_sum, :326)
    for thr in (float("nan"), -5.0, 1e300, 1e-300):
        want = oracle.score(case.issl, guides, 4, thr, "or")
        mit, cfd = dev.score(guides, 4, thr, "or")
        assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), thr
        assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), thr


def test_small_internal_batches_give_same_answer():
    case = golden_case("dense_repeats")
    guides = cb.pack_guides(case.guides)
    want = oracle.score(case.issl, guides, 4, 20, "and")
    os.environ["ISSL_BATCH"] = "7"
    try:
        dev = cb.Device.from_index(cb.Index(case.issl), 0, "auto")
    finally:
        del os.environ["ISSL_BATCH"]
    mit, cfd = dev.score(guides, 4, 20, "and")
    assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64))
    assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64))
    assert dev.stats["early_exits"] > 0
    dev.close()


def test_corrupt_lists_are_refused():
    case = golden_case("w8_families")
    img = bytearray(case.issl)
    # swap two adjacent entries of the first list: ids no longer ascending
    h = case.expected["header"]
    off = 48 + 16 * h["scoresCount"] + 8 * h["offtargetsCount"] + 8 * h["sliceCount"] * 256
    sizes = np.frombuffer(bytes(img[off - 8 * h["sliceCount"] * 256:off]), dtype=np.uint64)
    first = int(np.argmax(sizes >= 2))
    pos = off + 8 * int(sizes[:first].sum())
    img[pos:pos + 8], img[pos + 8:pos + 16] = img[pos + 8:pos + 16], img[pos:pos + 8]
    with pytest.raises(cb.IsslError) as e:
        cb.Device.from_index(cb.Index(bytes(img)), 0, "auto")
    assert e.value.code == 6


@pytest.mark.parametrize("w,layout", [(8, "res32"), (8, "gather"), (10, "res32"), (4, "sig64")])
def test_synthetic_index_is_what_the_reference_builder_would_write(w, layout, tmp_path):
    dev = cb.Device.synthetic(0, layout, seed=5, uniform_sites=30_000, families=8, family_size=400, max_sub_rate=0.1,
                              slice_width=w)
    path = tmp_path / "synth.issl"
    dev.write_issl(path)
    img = path.read_bytes()
    hd = oracle.header(img)
    assert hd["rc"] == 0 and hd["seqCount"] == 30_000 + 8 * 400 and hd["offtargetsCount"] <= hd["seqCount"]
    # rebuild the sorted text file from (signature, occurrences) and run the oracle's isslCreateIndex
    off = 48 + 16 * hd["scoresCount"]
    sigs = np.frombuffer(img[off:off + 8 * hd["offtargetsCount"]], dtype=np.uint64)
    ent = np.frombuffer(img[off + 8 * hd["offtargetsCount"] + 8 * hd["sliceCount"] * (1 << w):], dtype=np.uint64)
    occ = np.zeros(hd["offtargetsCount"], dtype=np.int64)
    first = ent[:hd["offtargetsCount"]]
    occ[(first & np.uint64(0xFFFFFFFF)).astype(np.int64)] = (first >> np.uint64(32)).astype(np.int64)
    lines = [cb.unpack_guide(int(s)).encode() + b"\n" for s in sigs]
    assert lines == sorted(lines) and len(set(lines)) == len(lines)
    text = b"".join(l * int(k) for l, k in zip(lines, occ))
    assert oracle.create_index(text, 20, w) == img
    # and it scores like the oracle does on that image
    rng = np.random.default_rng(3)
    guides = np.concatenate([dev.read_sites(rng.integers(0, hd["offtargetsCount"], 300).astype(np.uint64)),
                             rng.integers(0, 1 << 40, 100, dtype=np.uint64)])
    want = oracle.score(img, guides, 4, 0, "and", threads=0)
    mit, cfd = dev.score(guides, 4, 0, "and")
    assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64))
    assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64))
    dev.close()


def test_family_roots_match_the_device_generator(tmp_path):
    """bench.py recomputes the planted families' roots on the host (config 4 guides)."""
    import bench
    dev = cb.Device.synthetic(0, "auto", seed=7, uniform_sites=2000, families=6, family_size=50, max_sub_rate=0.0)
    dev.write_issl(tmp_path / "f.issl")
    img = (tmp_path / "f.issl").read_bytes()
    hd = oracle.header(img)
    off = 48 + 16 * hd["scoresCount"]
    sigs = np.frombuffer(img[off:off + 8 * hd["offtargetsCount"]], dtype=np.uint64)
    ent = np.frombuffer(img[off + 8 * hd["offtargetsCount"] + 8 * hd["sliceCount"] * 256:], dtype=np.uint64)[:hd["offtargetsCount"]]
    occ = dict(zip((ent & np.uint64(0xFFFFFFFF)).tolist(), (ent >> np.uint64(32)).tolist()))
    for root in bench.family_roots(7, 6):
        where = np.flatnonzero(sigs == root)
        assert where.size == 1 and occ[int(where[0])] >= 50
    dev.close()
