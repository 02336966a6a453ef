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
    triple = w in (8, 4, 10) and L == 20
    return (["triple"] if triple else []) + (["res32"] if res32 else []) + ["sig64", "gather"]


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
    """Index on one GPU from the file, replicas by peer copy, guides handed out in chunks, output in input order."""
    n = cb.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    case = golden_case("w8_families")
    (tmp_path / "i.issl").write_bytes(case.issl)
    (tmp_path / "g.txt").write_bytes(case.guides)
    for fanout in ("1", "0"):
        for run in case.expected["runs"][::17]:
            env = dict(os.environ, ISSL_GPUS=str(n), ISSL_TIMING="1", ISSL_CHUNK="7", ISSL_FANOUT=fanout)
            p = subprocess.run([str(cb.cli_path()), str(tmp_path / "i.issl"), str(tmp_path / "g.txt"), str(run["maxDist"]),
                                str(run["threshold"]), run["method"]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            assert p.returncode == 0, p.stderr
            assert p.stdout.decode() == run["stdout"]
            if run["method"] != "bogus":
                assert f"[issl] {n} gpu(s), chunks of 7 guides" in p.stderr.decode()


def test_score_multi_chunks_on_one_device():
    """issl_score_multi's chunking (the product's multi-GPU entry point) with a single device: any chunk size gives
    issl_score's answer, in input order, and the per-device count adds up."""
    case = golden_case("w8_families")
    guides = np.tile(cb.pack_guides(case.guides, case.seq_length), 5)
    dev = device_for("w8_families", "auto")
    for method, thr in (("and", 0.0), ("mit", 75.0), ("avg", 50.0)):
        mit0, cfd0 = dev.score(guides, 4, thr, method)
        for chunk in (1, 7, 64, guides.size, 0):
            mit, cfd = np.full(guides.size, -5.0), np.full(guides.size, -5.0)
            st, per = cb.score_multi([dev], guides, 4, thr, method, mit, cfd, chunk)
            assert np.array_equal(mit.view(np.uint64), mit0.view(np.uint64))
            if cfd0 is not None:
                assert np.array_equal(cfd.view(np.uint64), cfd0.view(np.uint64))
            else:
                assert np.all(cfd == -5.0)          # a column the method skips is left untouched
            assert st["guides"] == guides.size and int(per.sum()) == guides.size


def test_replicas_and_score_multi_on_several_gpus():
    """issl_device_clone + issl_score_multi: replicas made by peer copy score exactly like the original, and one call
    spread over all GPUs gives the single-GPU answer whatever the chunk size; pinned host buffers."""
    n = cb.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    text = td.make_offtargets(31, n_random=150_000, n_families=30, family_size=200)
    img = oracle.create_index(text, 20, 8)
    first = cb.Device.from_index(cb.Index(img), 0, "auto")
    devs = cb.replicate(first, list(range(n)))
    assert [d.info["cuda_device"] for d in devs] == list(range(n))
    assert all(d.info["hbm_bytes"] == first.info["hbm_bytes"] and d.info["layout"] == first.info["layout"] for d in devs)
    guides = td.pack_guides(td.make_guides(32, text, n=3000))
    hg, hm, hc = cb.HostBuffer(guides.size, np.uint64), cb.HostBuffer(guides.size, np.float64), cb.HostBuffer(guides.size, np.float64)
    hg.array[:] = guides
    for method, thr, md in (("and", 0.0, 4), ("and", 75.0, 4), ("or", 60.0, 3), ("cfd", 0.0, 5)):
        want = oracle.score(img, guides, md, thr, method, threads=0)
        for d in devs[1:]:
            mit, cfd = d.score(guides, md, thr, method)
            for got, ref in ((mit, want["mit"]), (cfd, want["cfd"])):
                assert got is None or np.array_equal(got.view(np.uint64), ref.view(np.uint64))
        for chunk in (97, 1000, 0):
            hm.array[:] = -1.0; hc.array[:] = -1.0
            st, per = cb.score_multi(devs, hg.array, md, thr, method, hm.array, hc.array, chunk)
            if method != "cfd":
                assert np.array_equal(hm.array.view(np.uint64), want["mit"].view(np.uint64))
            if method != "mit":
                assert np.array_equal(hc.array.view(np.uint64), want["cfd"].view(np.uint64))
            assert int(per.sum()) == guides.size
            if chunk == 97:
                assert np.count_nonzero(per) == n, per     # every GPU took part
    with pytest.raises(cb.IsslError):
        cb.score_multi([devs[0], devs[0]], hg.array, 4, 0.0, "and", hm.array, hc.array)
    with pytest.raises(cb.IsslError):
        devs[0].clone(0)
    for d in devs:
        d.close()


def test_out_of_memory_batches_are_halved_and_layout_falls_back(monkeypatch):
    """ISSL_ERR_NOMEM inside a batch halves the batch and repeats it (hooks simulate the failure); an index whose
    sub-bucket copies do not fit under ISSL_LAYOUT_AUTO is scored through its slice lists -- same results either way."""
    case = golden_case("w8_families")
    guides = np.tile(cb.pack_guides(case.guides, case.seq_length), 40)
    dev = device_for("w8_families", "auto")
    want = dev.score(guides, 4, 75.0, "and")
    monkeypatch.setenv("ISSL_TEST_NOMEM_ABOVE", "1500")
    got = dev.score(guides, 4, 75.0, "and")
    assert dev.stats["guides"] == guides.size
    monkeypatch.delenv("ISSL_TEST_NOMEM_ABOVE")
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    monkeypatch.setenv("ISSL_TEST_TRIPLE_NOMEM", "1")
    lists = cb.Device.from_index(cb.Index(case.issl), 0, "auto")
    monkeypatch.delenv("ISSL_TEST_TRIPLE_NOMEM")
    assert lists.info["layout"] == cb.LAYOUTS["res32"]
    for a, b in zip(lists.score(guides, 4, 75.0, "and"), want):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    with pytest.raises(cb.IsslError) as e:      # an explicit request is not silently downgraded
        monkeypatch.setenv("ISSL_TEST_TRIPLE_NOMEM", "1")
        cb.Device.from_index(cb.Index(case.issl), 0, "triple")
    assert e.value.code == 7
    lists.close()


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
        dev = cb.Device.from_index(cb.Index(img), 0, "res32")
    finally:
        del os.environ["ISSL_MAX_GROUP"]
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("avg", 50, 3), ("or", 0, 0), ("and", 0, 5), ("and", 0, 6),
                            ("mit", 0, 7), ("cfd", 0, 8)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (max_group, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (max_group, method, thr, md)
        if thr == 0:
            st = dev.stats
            assert st["candidates"] == int(want["candidates"].sum())
            if max_group == 1:
                assert st["streamed"] == st["candidates"]
            elif max_group == 32 and md <= 7:
                assert st["candidates"] > 4 * st["streamed"]          # blocks of up to 32 guides share a chunk
            else:
                assert st["candidates"] > 2 * st["streamed"]
    _, _, hits = dev.score_hits(guides, 4, 0, "and")
    want = oracle.score(img, guides, 4, 0, "and", threads=1, want_hits=True)["hits"]
    assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1))
    dev.close()


@pytest.mark.parametrize("triple_max", [6, 3])
def test_triple_layout_sub_bucket_scan(triple_max):
    """ISSL_LAYOUT_TRIPLE reads only the sub-buckets that can hold a hit (issl_triple_visits) and must still give
    the oracle's exact answer -- scores bit-identical, hit tuples identical including order, early exits at the same
    survivor -- on dense families (many hits per guide, many sites per bucket), for every maxDist it serves, and
    hand over to the RES32 list scan above ISSL_TRIPLE_MAXDIST."""
    text = td.make_offtargets(41, n_random=150_000, n_families=40, family_size=500, max_sub_rate=0.15)
    img = oracle.create_index(text, 20, 8)
    rng = np.random.default_rng(42)
    roots = td.pack_guides(td.make_guides(43, text, n=60, frac_exact=1.0, frac_mut=0.0))
    guides = []
    for r in roots:
        for _ in range(20):
            g = int(r)
            for pos in rng.choice(20, size=int(rng.integers(0, 5)), replace=False):
                g ^= int(rng.integers(1, 4)) << (2 * int(pos))
            guides.append(g)
    guides = np.concatenate([np.array(guides, dtype=np.uint64), rng.integers(0, 1 << 40, 300, dtype=np.uint64)])
    os.environ["ISSL_TRIPLE_MAXDIST"] = str(triple_max)
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "triple")
    finally:
        del os.environ["ISSL_TRIPLE_MAXDIST"]
    assert dev.info["layout"] == cb.LAYOUTS["triple"]
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("avg", 50, 3), ("or", 0, 0), ("mit", 0, 1), ("cfd", 0, 2),
                            ("and", 0, 5), ("or", 30, 5), ("and", 0, 6), ("and", 60, 6), ("mit", 0, 7), ("cfd", 0, 9)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (method, thr, md)
        st = dev.stats
        if thr == 0:
            assert st["candidates"] == int(want["candidates"].sum())
        if md <= triple_max:
            visits, _ = cb.triple_visits(md)
            assert st["bucket_visits"] > 0 and st["streamed"] < st["candidates"]
            if thr == 0:
                assert st["bucket_visits"] == visits.size * guides.size
        else:
            assert st["bucket_visits"] == 0
    for md, thr in ((4, 0), (5, 0), (4, 75), (2, 0)):
        _, _, hits = dev.score_hits(guides, md, thr, "and")
        want = oracle.score(img, guides, md, thr, "and", threads=1, want_hits=True)["hits"]
        assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1)), (md, thr)
    dev.close()


@pytest.mark.parametrize("fuse,blocks,flush,hit_cap", [(2, "64", "0", None), (2, "32", "0", None), (1, "64", "0", None),
                                                       (0, "0", "0", None), (2, "0", "0", None), (2, "64", "1", None),
                                                       (1, "32", "1", None), (0, "64", "1", None),
                                                       (2, "64", "0", "3000"), (1, "64", "1", "3000"), (0, "32", "0", "3000"),
                                                       (2, "64", "0", "noflag"), (1, "32", "1", "noflag"), (0, "64", "0", "noflag"),
                                                       (2, "32", "1", None), (2, "64", "1", "3000"), (2, "64", "1", "noflag"),
                                                       (2, "64", "1", "noheavy"), (2, "0", "1", None)])
def test_triple_large_batch_tails(fuse, blocks, flush, hit_cap):
    """With at least 148 x 16 guides in a batch every guide gets one CTA, and the scan's tails come into play: 2 = the
    CTA sorts, scores and accumulates the guide's hits itself (fused), 1 = per-guide segments finished by
    k_score_segments, 0 = general pipeline; guides with more than 512 hits (dense families here) take the general
    pipeline, mixed with the others in one batch -- except with fuse = 2 and flush = 1 on a blocked copy, where the
    scan variant that empties a full record list in the middle of the scan keeps such a guide's hits as sort keys, and
    the CTA sorts and finishes it too ("noheavy" switches that off).  Blocked (bit-sliced sub-blocks of 31/62 entries, with buckets that
    do not fit) and contiguous copies; with hit_cap the survivor buffers start far too small, so that every call is
    launched again after an overflow (the fused tail must then start from the same per-guide state).  Everything
    must stay bit-identical to the oracle, early exits included."""
    text = td.make_offtargets(51, n_random=120_000, n_families=30, family_size=900, max_sub_rate=0.12)
    img = oracle.create_index(text, 20, 8)
    rng = np.random.default_rng(52)
    roots = td.pack_guides(td.make_guides(53, text, n=400, frac_exact=1.0, frac_mut=0.0))
    guides = []
    for r in roots:
        for _ in range(6):
            g = int(r)
            for pos in rng.choice(20, size=int(rng.integers(0, 4)), replace=False):
                g ^= int(rng.integers(1, 4)) << (2 * int(pos))
            guides.append(g)
    guides = np.concatenate([np.array(guides, dtype=np.uint64), rng.integers(0, 1 << 40, 600, dtype=np.uint64)])
    assert guides.size >= 148 * 16
    os.environ["ISSL_TRIPLE_FUSE"] = str(fuse)
    os.environ["ISSL_TRIPLE_BLOCKS"] = blocks
    os.environ["ISSL_TRIPLE_FLUSH"] = flush
    if hit_cap == "noflag":     # the path of indexes with >= 2^31 sites: no occurrence flag in the stored ids
        os.environ["ISSL_TRIPLE_OCCFLAG"] = "0"
    elif hit_cap == "noheavy":
        os.environ["ISSL_TRIPLE_HEAVY"] = "0"
    elif hit_cap:
        os.environ["ISSL_HIT_CAP"] = hit_cap
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "triple")
    finally:
        del os.environ["ISSL_TRIPLE_FUSE"], os.environ["ISSL_TRIPLE_BLOCKS"], os.environ["ISSL_TRIPLE_FLUSH"]
        os.environ.pop("ISSL_HIT_CAP", None)
        os.environ.pop("ISSL_TRIPLE_OCCFLAG", None)
        os.environ.pop("ISSL_TRIPLE_HEAVY", None)
    assert dev.info["triple_block_bytes"] == 2 * int(blocks)
    in_kernel = fuse == 2 and flush == "1" and blocks != "0" and hit_cap != "noheavy"
    seen_big = False
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("or", 40, 3), ("avg", 55, 5), ("mit", 0, 2), ("cfd", 20, 4)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (fuse, blocks, flush, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (fuse, blocks, flush, method, thr, md)
        st = dev.stats
        if thr == 0:
            assert st["candidates"] == int(want["candidates"].sum())
            visits, _ = cb.triple_visits(md)
            assert st["bucket_visits"] == visits.size * guides.size
        if (method, thr, md) == ("and", 0, 4):
            hits = oracle.score(img, guides, 4, 0, "and", threads=1, want_hits=True)["hits"]
            per_guide = np.bincount(hits["guide"].astype(np.int64), minlength=guides.size)
            assert st["hits"] == per_guide.sum()
            seen_big = per_guide.max() > 512 and (per_guide <= 512).sum() > 1000
            if in_kernel:      # every guide above the list's capacity was sorted and finished by its own CTA
                assert st["sorted_hits"] == 0 and st["heavy_hits"] >= per_guide[per_guide > 512].sum()
            elif fuse == 2:
                assert st["heavy_hits"] == 0 and st["sorted_hits"] >= per_guide[per_guide > 512].sum()
    assert seen_big, "the fixture should mix guides above and below the per-CTA hit capacity"
    dev.close()


@pytest.mark.parametrize("w,blocks", [(8, "64"), (8, "32"), (4, "64")])
def test_heavy_guides_are_sorted_and_finished_inside_the_scan_kernel(w, blocks):
    """Guides with thousands of hits (maxDist 5 and 6, families of near-copies with many exact duplicates): the scan's
    flush variant turns full record lists into 64-bit sort keys (ordering slice, site text rank, occurrences), the CTA
    radix-sorts them in its own chunks of global memory and accumulates them in the reference's order (ref
    isslScoreOfftargets.cpp:330-344, :466-502) -- no device-wide sort.  Long bucket remainders (a family's root bucket
    holds thousands of entries) are shared by the whole CTA.  Bit-identical to the oracle, early exits included; the
    survivor buffers also start far too small once, so that the launch is repeated."""
    text = td.make_offtargets(91, n_random=90_000, n_families=6, family_size=5000, max_sub_rate=0.07)
    img = oracle.create_index(text, 20, w)
    rng = np.random.default_rng(92)
    roots = td.pack_guides(td.make_guides(93, text, n=300, frac_exact=1.0, frac_mut=0.0))
    guides = np.concatenate([roots.repeat(7), rng.integers(0, 1 << 40, 300, dtype=np.uint64)])
    for i in range(roots.size * 7):
        for pos in rng.choice(20, size=int(rng.integers(0, 3)), replace=False):
            guides[i] ^= np.uint64(int(rng.integers(1, 4)) << (2 * int(pos)))
    assert guides.size >= 148 * 16
    for hit_cap in (None, "5000"):
        os.environ["ISSL_TRIPLE_BLOCKS"] = blocks
        if hit_cap:
            os.environ["ISSL_HIT_CAP"] = hit_cap
        try:
            dev = cb.Device.from_index(cb.Index(img), 0, "triple")
        finally:
            del os.environ["ISSL_TRIPLE_BLOCKS"]
            os.environ.pop("ISSL_HIT_CAP", None)
        if w == 4:
            runs = (("and", 0, 4), ("or", 45, 4), ("avg", 70, 4))     # above maxDist 4 sliceWidth 4 takes its slice lists
        elif hit_cap:
            runs = (("and", 0, 5), ("or", 45, 5), ("avg", 70, 4))
        else:
            runs = (("and", 0, 5), ("and", 0, 6), ("or", 45, 5), ("avg", 70, 6), ("mit", 30, 4), ("cfd", 0, 4), ("and", 75, 4))
        for method, thr, md in runs:
            want = oracle.score(img, guides, md, thr, method, threads=0)
            mit, cfd = dev.score(guides, md, thr, method)       # the first call on a handle learns that hits are many ...
            mit, cfd = dev.score(guides, md, thr, method)       # ... and the second runs the flush variant
            st = dev.stats
            if method != "cfd":
                assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (w, blocks, hit_cap, method, thr, md)
            if method != "mit":
                assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (w, blocks, hit_cap, method, thr, md)
            if thr == 0:    # with an early exit fewer hits are found, and the handle may not expect heavy guides
                assert st["sorted_hits"] == 0, (w, blocks, hit_cap, method, thr, md)
                assert st["heavy_hits"] > 100 * 1000, (w, blocks, hit_cap, method, thr, md, st)
        dev.close()


@pytest.mark.parametrize("case", ["sorted", "sorted_ids", "shuffled", "w4_sorted", "w4_shuffled"])
def test_fused_tail_orders_hits_by_site_text_or_by_id(case):
    """The reference accumulates a slice's hits in ascending id (isslScoreOfftargets.cpp:330-344).  On an index whose
    sites are in text order (what extractOfftargets + isslCreateIndex produce) ids are text ranks, and the fused tail
    orders hits by their own signature without looking ids up; sites that occur more than once sit first in their
    bucket and are the only ones whose count is fetched.  An index built from UNSORTED text (ids in file order, equal
    sites under several ids) must be recognised at load and take the id path; ISSL_SITE_ORDER=0 forces that path.
    Early exits make every printed digit depend on the order, so bit-identical scores prove it."""
    w = 4 if case.startswith("w4") else 8
    text = td.make_offtargets(81, n_random=110_000, n_families=30, family_size=600, max_sub_rate=0.1)
    if case.endswith("shuffled"):
        lines = np.frombuffer(text, dtype=np.uint8).reshape(-1, 21).copy()
        np.random.default_rng(82).shuffle(lines, axis=0)
        text = lines.tobytes()
    img = oracle.create_index(text, 20, w)
    rng = np.random.default_rng(83)
    roots = td.pack_guides(td.make_guides(84, text, n=450, frac_exact=1.0, frac_mut=0.0))
    guides = np.concatenate([roots.repeat(5), rng.integers(0, 1 << 40, 300, dtype=np.uint64)])
    for i in range(roots.size * 5):
        for pos in rng.choice(20, size=int(rng.integers(0, 3)), replace=False):
            guides[i] ^= np.uint64(int(rng.integers(1, 4)) << (2 * int(pos)))
    assert guides.size >= 148 * 16
    os.environ["ISSL_TRIPLE_BLOCKS"] = "32"
    if case == "sorted_ids":
        os.environ["ISSL_SITE_ORDER"] = "0"
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "triple")
    finally:
        del os.environ["ISSL_TRIPLE_BLOCKS"]
        os.environ.pop("ISSL_SITE_ORDER", None)
    assert dev.info["triple_hit_bytes"] == (0 if case in ("sorted", "w4_sorted") else 28)
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("or", 40, 3), ("avg", 55, 4), ("mit", 30, 2), ("cfd", 20, 4)):
        want = oracle.score(img, guides, md, thr, method, threads=0)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (case, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (case, method, thr, md)
    dev.close()


@pytest.mark.parametrize("fuse,blocks", [(2, "64"), (1, "32"), (0, "0"), (2, "0")])
def test_triple_layout_slice_width_4(fuse, blocks):
    """sliceWidth 4 (ten 2-base slices) under TRIPLE: up to maxDist 4 the same sub-buckets are read, but hits are
    accumulated in the order of the lowest exactly matching 2-base slice, and the early exit takes effect in the
    ordered accumulation (one wave); from maxDist 5 on a hit may match on a 2-base slice without matching on any whole
    byte, and the buckets whose three key bytes all differ are read as well (issl_triple_visits_w4).  Large batch, so
    that the fused tail, the segment kernel and the general pipeline (guides with more than 512 hits) all run."""
    text = td.make_offtargets(61, n_random=100_000, n_families=25, family_size=800, max_sub_rate=0.12)
    img = oracle.create_index(text, 20, 4)
    rng = np.random.default_rng(62)
    roots = td.pack_guides(td.make_guides(63, text, n=400, frac_exact=1.0, frac_mut=0.0))
    guides = []
    for r in roots:
        for _ in range(6):
            g = int(r)
            for pos in rng.choice(20, size=int(rng.integers(0, 4)), replace=False):
                g ^= int(rng.integers(1, 4)) << (2 * int(pos))
            guides.append(g)
    guides = np.concatenate([np.array(guides, dtype=np.uint64), rng.integers(0, 1 << 40, 400, dtype=np.uint64)])
    os.environ["ISSL_TRIPLE_FUSE"] = str(fuse)
    os.environ["ISSL_TRIPLE_BLOCKS"] = blocks
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "auto")
    finally:
        del os.environ["ISSL_TRIPLE_FUSE"], os.environ["ISSL_TRIPLE_BLOCKS"]
    assert dev.info["layout"] == cb.LAYOUTS["triple"]
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("or", 40, 3), ("avg", 55, 4), ("mit", 0, 2), ("cfd", 20, 4), ("and", 30, 5), ("and", 0, 5),
                            ("or", 60, 6)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (fuse, blocks, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (fuse, blocks, method, thr, md)
        st = dev.stats
        if thr == 0:
            assert st["candidates"] == int(want["candidates"].sum())
        assert st["bucket_visits"] > 0   # the sub-bucket scan serves every maxDist up to ISSL_TRIPLE_MAXDIST (6)
    _, _, hits = dev.score_hits(guides[:600], 4, 0, "and")
    want = oracle.score(img, guides[:600], 4, 0, "and", threads=1, want_hits=True)["hits"]
    assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1))
    dev.close()


@pytest.mark.parametrize("fuse,blocks,flush", [(2, "64", "0"), (2, "32", "1"), (1, "32", "0"), (0, "0", "0"), (2, "0", "0")])
def test_triple_layout_slice_width_10(fuse, blocks, flush):
    """sliceWidth 10 (four 5-base slices) under TRIPLE.  The builder truncates slice values to 8 bits (ref
    isslCreateIndex.cpp:228): list (i, v) holds the sites that agree with v on the first FOUR bases of slice i, and a guide
    only finds such a list when the fifth base of its own slice is A.  The sub-bucket copies are built from permuted
    signatures (the four 4-base units, then the four fifth bases), a guide keeps a hit only behind an open gate, and the
    hit is accumulated under the lowest such slice -- one wave, the early exit takes effect in the ordered accumulation.
    Guides are mutated family members with every combination of gates, and plain random ones."""
    text = td.make_offtargets(71, n_random=100_000, n_families=25, family_size=800, max_sub_rate=0.12)
    img = oracle.create_index(text, 20, 10)
    rng = np.random.default_rng(72)
    roots = td.pack_guides(td.make_guides(73, text, n=400, frac_exact=1.0, frac_mut=0.0))
    guides = []
    for r in roots:
        for k in range(6):
            g = int(r)
            for pos in rng.choice(20, size=int(rng.integers(0, 4)), replace=False):
                g ^= int(rng.integers(1, 4)) << (2 * int(pos))
            for i in range(4):                      # open (A) or close the gate of every slice at random
                if rng.random() < 0.6:
                    g &= ~(3 << (10 * i + 8))
            guides.append(g)
    guides = np.concatenate([np.array(guides, dtype=np.uint64), rng.integers(0, 1 << 40, 400, dtype=np.uint64)])
    os.environ["ISSL_TRIPLE_FUSE"] = str(fuse)
    os.environ["ISSL_TRIPLE_BLOCKS"] = blocks
    os.environ["ISSL_TRIPLE_FLUSH"] = flush
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "auto")
    finally:
        del os.environ["ISSL_TRIPLE_FUSE"], os.environ["ISSL_TRIPLE_BLOCKS"], os.environ["ISSL_TRIPLE_FLUSH"]
    assert dev.info["layout"] == cb.LAYOUTS["triple"]
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("or", 40, 3), ("avg", 55, 5), ("mit", 0, 2), ("cfd", 20, 4), ("and", 30, 6)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (fuse, blocks, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (fuse, blocks, method, thr, md)
        st = dev.stats
        if thr == 0:
            assert st["candidates"] == int(want["candidates"].sum())
        assert st["bucket_visits"] > 0
    _, _, hits = dev.score_hits(guides[:600], 4, 0, "and")
    want = oracle.score(img, guides[:600], 4, 0, "and", threads=1, want_hits=True)["hits"]
    assert np.array_equal(hits, np.stack([want[k].astype(np.int64) for k in ("guide", "id", "dist", "occ")], axis=1))
    dev.close()


@pytest.mark.parametrize("waves", ["0", "1", "2"])
@pytest.mark.parametrize("layout", ["triple", "res32"])
def test_early_exit_wave_policies(waves, layout):
    """With an early exit the slices are scanned in waves (guides that left are dropped from later waves), in one
    launch, or adaptively (single-slice waves until few guides leave, then the rest at once).  The printed score of an
    early-exited guide depends on exactly where it stopped: every policy must reproduce the reference's exit points."""
    text = td.make_offtargets(71, n_random=100_000, n_families=25, family_size=700, max_sub_rate=0.1)
    img = oracle.create_index(text, 20, 8)
    rng = np.random.default_rng(72)
    roots = td.pack_guides(td.make_guides(73, text, n=500, frac_exact=1.0, frac_mut=0.0))
    guides = np.concatenate([roots.repeat(5), rng.integers(0, 1 << 40, 500, dtype=np.uint64)])
    for i in range(guides.size):                      # 0-2 substitutions per guide
        for pos in rng.choice(20, size=int(rng.integers(0, 3)), replace=False):
            guides[i] ^= np.uint64(int(rng.integers(1, 4)) << (2 * int(pos)))
    os.environ["ISSL_WAVES"] = waves
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, layout)
    finally:
        del os.environ["ISSL_WAVES"]
    exits = 0
    for method, thr in (("and", 75), ("or", 60), ("avg", 90), ("mit", 30), ("cfd", 99.5), ("and", 5)):
        want = oracle.score(img, guides, 4, thr, method, threads=0)
        mit, cfd = dev.score(guides, 4, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (waves, layout, method, thr)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (waves, layout, method, thr)
        exits += dev.stats["early_exits"]
    assert exits > 1000
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


@pytest.mark.parametrize("w,layout", [(8, "triple"), (8, "res32"), (8, "gather"), (10, "res32"), (4, "sig64"), (4, "triple")])
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


# ---- the device-side isslCreateIndex ------------------------------------

@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_device_builder_writes_the_reference_issl(name, tmp_path):
    """issl_device_create_from_text + issl_device_write_issl == what the unmodified reference
    isslCreateIndex wrote for the same text (sha256 committed by tests/golden/make_golden.py)."""
    import hashlib
    case = golden_case(name)
    dev = cb.Device.from_text(case.offtargets, case.seq_length, case.slice_width)
    out = tmp_path / "built.issl"
    dev.write_issl(out)
    img = out.read_bytes()
    assert len(img) == case.expected["issl_bytes"]
    assert hashlib.sha256(img).hexdigest() == case.expected["issl_sha256"]
    # the handle scores without going through the file
    guides = cb.pack_guides(case.guides, case.seq_length)
    run = next(r for r in case.expected["runs"] if r["method"] == "and" and r["maxDist"] == 4)
    mit, cfd = dev.score(guides, run["maxDist"], run["threshold"], run["method"])
    assert fmt_stdout(guides, mit, cfd, case.seq_length) == run["stdout"]
    dev.close()


def test_create_index_program_is_a_drop_in(tmp_path):
    """bin/isslCreateIndex: same argv, same stdout lines, byte-identical output file, same refusals."""
    case = golden_case("w8_families")
    src, out = tmp_path / "offtargets.txt", tmp_path / "o.issl"
    src.write_bytes(case.offtargets)
    r = subprocess.run([str(cb.create_cli_path()), str(src), "20", "8", str(out)], capture_output=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == (b"Finished counting occurrences, now constructing index...\n"
                        b"Finished constructing index, now precalculating scores...\n"
                        b"Finished calculating scores, now preparing to write to disk...\n"
                        b"Writing to disk...\nDone.\n")
    assert r.stderr.startswith(b"Number of sequences: %d\n" % (len(case.offtargets) // 21))
    assert out.read_bytes() == case.issl
    bad = tmp_path / "bad.txt"
    bad.write_bytes(case.offtargets[:-3])
    r = subprocess.run([str(cb.create_cli_path()), str(bad), "20", "8", str(out)], capture_output=True)
    assert r.returncode == 1 and b"not a multiple of the expected line length (21)" in r.stderr
    r = subprocess.run([str(cb.create_cli_path()), str(src), "33", "8", str(out)], capture_output=True)
    assert r.returncode == 1 and b"greater than 32" in r.stderr
    r = subprocess.run([str(cb.create_cli_path()), str(src)], capture_output=True)
    assert r.returncode == 1 and r.stderr.startswith(b"Usage: ")


@pytest.mark.parametrize("w", [8, 10, 4])
def test_device_builder_on_messy_text(w):
    """Non-ACGT characters pack as A but still split runs (the reference compares the TEXT, isslCreateIndex.cpp:192);
    unsorted input, long runs and a one-line file behave as in the reference's restatement."""
    rng = np.random.default_rng(40 + w)
    codes = rng.integers(0, 4, (6000, 20))
    lines = ["".join("ACGT"[c] for c in row) for row in codes]
    lines += [lines[5]] * 700 + ["N" * 20, "A" * 20, "N" * 20, "ANNNNNNNNNNNNNNNNNNN"] + [lines[7][:10] + "acgtnRYKMS"] * 3
    sorted_text = "".join(s + "\n" for s in sorted(lines)).encode()
    shuffled = list(lines)
    rng.shuffle(shuffled)
    unsorted_text = "".join(s + "\n" for s in shuffled).encode()
    for text in (sorted_text, unsorted_text, b"ACGTACGTACGTACGTACGT\n"):
        want = oracle.create_index(text, 20, w)
        dev = cb.Device.from_text(text, 20, w)
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            dev.write_issl(os.path.join(d, "m.issl"))
            got = open(os.path.join(d, "m.issl"), "rb").read()
        dev.close()
        assert got == want


# ---- the resident scorer -------------------------------------------------

def _server_request(sock_path, op):
    """issl_wire.h Request with no payload (ping = 2, shutdown = 3); returns the Response status."""
    import socket
    import struct
    req = struct.pack("<8sIidiiQII16i", b"ISSLREQ1", op, 0, 0.0, 0, 0, 0, 0, 0, *([0] * 16))
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
        s.settimeout(60)
        s.connect(str(sock_path))
        s.sendall(req)
        rsp = b""
        while len(rsp) < 12:
            chunk = s.recv(4096)
            if not chunk:
                break
            rsp += chunk
    assert rsp[:8] == b"ISSLRSP1"
    return struct.unpack("<i", rsp[8:12])[0]


def test_resident_server_is_invisible_to_the_caller(tmp_path):
    """ISSL_SERVER: same stdout as the in-process path; the second call finds the index resident; a rewritten
    index file is reloaded; the server never holds the caller's stdout open; shutdown removes the socket."""
    import time
    a, b = golden_case("w8_families"), golden_case("w10_truncated")
    idx, ga, gb = tmp_path / "index.issl", tmp_path / "ga.txt", tmp_path / "gb.txt"
    idx.write_bytes(a.issl)
    ga.write_bytes(a.guides)
    gb.write_bytes(b.guides)
    sock = tmp_path / "issl.sock"
    env = dict(os.environ, ISSL_SERVER=str(sock), ISSL_SERVER_AUTOSTART="1", ISSL_SERVER_LOG=str(tmp_path / "server.log"),
               ISSL_SERVER_IDLE_S="120", ISSL_TIMING="1", ISSL_GPUS="1")
    exe = str(cb.cli_path())

    def run(guides, md, thr, method):
        return subprocess.run([exe, str(idx), str(guides), str(md), str(thr), method], capture_output=True, env=env, timeout=300)
    try:
        runs = [r for r in a.expected["runs"] if r["method"] in ("and", "mit", "avg")][:6]
        for k, run_ in enumerate(runs):
            r = run(ga, run_["maxDist"], run_["threshold"], run_["method"])
            assert r.returncode == 0, r.stderr
            assert r.stdout.decode() == run_["stdout"]
            assert (b"index resident" if k else b"index loaded") in r.stderr, r.stderr
        assert _server_request(sock, 2) == 0
        # the file changes under the same path: the server must notice
        time.sleep(0.01)
        idx.write_bytes(b.issl)
        run_ = next(r for r in b.expected["runs"] if r["method"] == "and" and r["maxDist"] == 4)
        r = run(gb, run_["maxDist"], run_["threshold"], run_["method"])
        assert r.returncode == 0 and r.stdout.decode() == run_["stdout"] and b"index loaded" in r.stderr
        # errors come back on stderr with exit 1, nothing on stdout
        bad = tmp_path / "bad.issl"
        bad.write_bytes(a.issl[:1000])
        r = subprocess.run([exe, str(bad), str(ga), "4", "0", "and"], capture_output=True, env=env, timeout=300)
        assert r.returncode == 1 and r.stdout == b""
    finally:
        if sock.exists():
            assert _server_request(sock, 3) == 0
    for _ in range(100):
        if not sock.exists():
            break
        time.sleep(0.1)
    assert not sock.exists()
    # without a server and without autostart the program scores in-process and says so
    env2 = dict(env, ISSL_SERVER_AUTOSTART="0")
    idx.write_bytes(a.issl)
    r = subprocess.run([exe, str(idx), str(ga), "4", "0", "and"], capture_output=True, env=env2, timeout=300)
    run_ = next(r_ for r_ in a.expected["runs"] if r_["method"] == "and" and r_["maxDist"] == 4 and float(r_["threshold"]) == 0)
    assert r.returncode == 0 and r.stdout.decode() == run_["stdout"] and b"scoring in-process" in r.stderr


@pytest.mark.parametrize("flush", ["0", None])
def test_triple_skewed_index_many_overflowing_buckets(flush):
    """Indexes made by the reference's extractor are not uniform: it takes the reverse strand's site from the wrong end of
    its match (extractOfftargets.py:97-106), so half of the sites end in AG / GG and the sub-buckets keyed on those values
    are several times fuller than the mean.  Here: 60 000 uniform sites + 150 000 sites over a two-letter alphabet ending in
    [AG]G, blocks of 31 slots: a guide of the second kind meets ~130 buckets that overflow their block -- more than the 64
    a CTA notes in shared memory, so the rest go through the global bitmap (non-flush scan) or are finished whenever the list
    fills (flush scan).  Scores bit-identical to the oracle either way."""
    rng = np.random.default_rng(71)
    F = rng.integers(0, 4, (60_000, 20), dtype=np.uint8)
    F[:, 0] = rng.integers(0, 3, 60_000)
    R = rng.integers(0, 2, (150_000, 20), dtype=np.uint8)
    R[:, 18] = rng.integers(0, 2, 150_000) * 2
    R[:, 19] = 2
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    lines = sorted(bytes(r) for r in letters[np.concatenate([F, R, R[:500]])])      # (some sites occur twice)
    text = b"".join(l + b"\n" for l in lines)
    img = oracle.create_index(text, 20, 8)
    pick = np.concatenate([R[rng.integers(0, len(R), 1700)], F[rng.integers(0, len(F), 500)],
                           rng.integers(0, 4, (300, 20), dtype=np.uint8)]).copy()
    for row in pick[: 1200]:
        for pos in rng.choice(20, size=int(rng.integers(0, 3)), replace=False):
            row[pos] = (row[pos] + rng.integers(1, 4)) % 4
    guides = td.pack_guides(b"".join(bytes(r) + b"\n" for r in letters[pick]))
    os.environ["ISSL_TRIPLE_BLOCKS"] = "32"
    if flush is not None:
        os.environ["ISSL_TRIPLE_FLUSH"] = flush
    try:
        dev = cb.Device.from_index(cb.Index(img), 0, "triple")
    finally:
        del os.environ["ISSL_TRIPLE_BLOCKS"]
        os.environ.pop("ISSL_TRIPLE_FLUSH", None)
    assert dev.info["triple_block_bytes"] == 64
    for method, thr, md in (("and", 0, 4), ("and", 75, 4), ("mit", 0, 3), ("or", 50, 4), ("cfd", 0, 2)):
        want = oracle.score(img, guides, md, thr, method, threads=0, want_candidates=True)
        mit, cfd = dev.score(guides, md, thr, method)
        if method != "cfd":
            assert np.array_equal(mit.view(np.uint64), want["mit"].view(np.uint64)), (flush, method, thr, md)
        if method != "mit":
            assert np.array_equal(cfd.view(np.uint64), want["cfd"].view(np.uint64)), (flush, method, thr, md)
    st = dev.stats
    assert st["bucket_visits"] > 0
    dev.close()
