"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the
header declares, host-side logic (guide packing, method names, builder arithmetic, .issl
validation) agrees with the oracle, and the host program keeps the reference's CLI contract.
No compute call is made (there is no GPU here and the product has no CPU path)."""
import pathlib
import re
import subprocess

import numpy as np
import pytest

import crackling_b200 as cb
import issl_testdata as td
from conftest import golden_case
from oracle import oracle

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "issl_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(issl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = cb.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libissl_cuda.so does not export {n}"
    assert set(L._issl_symbols) == set(names), "binding.py and issl_cuda.h disagree"
    assert L.issl_abi_version() == 4


def test_format_lines_prints_what_printf_prints():
    """issl_format_lines replaces the reference's printf loop (isslScoreOfftargets.cpp:514-527): same digits as "%f"."""
    rng = np.random.default_rng(5)
    n = 200_000
    guides = rng.integers(0, 1 << 40, n, dtype=np.uint64)
    # scores as the scorer produces them (10000 / (100 + sum)) plus awkward values: ties at the sixth decimal,
    # tiny, huge, negative zero, infinities, NaN
    mit = 10000.0 / (100.0 + rng.exponential(50.0, n))
    cfd = 10000.0 / (100.0 + rng.exponential(5.0, n))
    mit[:12] = [100.0, 0.0, -0.0, 0.0000005, 0.0000015, 0.0000025, 1e-300, 1e300, 99.9999995, np.inf, -np.inf, np.nan]
    cfd[:6] = [0.5, 1.0000005, 2.5e-7, 123456789.123456789, 1e22, 5e-324]
    for method, (a, b) in {"and": (mit, cfd), "mit": (mit, None), "cfd": (None, cfd), "bogus": (None, None)}.items():
        got = cb.format_lines(guides, mit, cfd, method)
        want = b"".join(b"%s\t%s\t%s\n" % (cb.unpack_guide(int(g)).encode(), (b"%f" % a[i]) if a is not None else b"-1",
                                           (b"%f" % b[i]) if b is not None else b"-1") for i, g in enumerate(guides[:3000]))
        assert got.startswith(want), method
        assert got.count(b"\n") == n
    assert cb.format_lines(guides[:0], mit[:0], cfd[:0], "and") == b""


def test_multi_gpu_chunks():
    """issl_multi_chunk: about eight chunks per device, between 65 536 guides and one internal batch."""
    assert cb.multi_chunk(1000, 1) == 1000
    assert cb.multi_chunk(1000, 8) == 65536
    assert cb.multi_chunk(10_000_000, 8) % 4096 == 0 and 150_000 < cb.multi_chunk(10_000_000, 8) < 170_000
    assert cb.multi_chunk(1 << 30, 2) == 1 << 20
    with pytest.raises(cb.IsslError):
        cb.score_multi([], np.zeros(1, np.uint64), 4, 0.0, "and", np.zeros(1), np.zeros(1))


def test_pack_unpack_and_methods():
    g = td.make_guides(7, td.make_offtargets(7, n_random=50, n_families=1, family_size=5), n=33)
    got = cb.pack_guides(g)
    assert np.array_equal(got, td.pack_guides(g))
    assert [int(x) for x in got] == [oracle.encode(g[i * 21:i * 21 + 20]) for i in range(33)]
    assert cb.unpack_guide(got[0]) == g[:20].decode()
    assert cb.pack_guides(b"NNNNACGTACGTACGTACGT\n")[0] == cb.pack_guides(b"AAAAACGTACGTACGTACGT\n")[0]
    with pytest.raises(cb.IsslError):
        cb.pack_guides(b"ACGT\n" * 3 + b"A", 4)
    for name, code in cb.METHODS.items():
        assert cb.method_code(name) == code
    assert cb.method_code("MIT") == 0 and cb.method_code("") == 0


@pytest.mark.parametrize("w", [4, 5, 8, 10])
def test_mit_table_is_bit_identical_to_reference_builder(w):
    # the oracle image is itself pinned to the reference isslCreateIndex (test_oracle_golden.py)
    img = oracle.create_index(b"ACGTACGTACGTACGTACGT\n", 20, w)
    masks, scores, header_count = cb.mit_table(20, w)
    assert header_count == oracle.header(img)["scoresCount"]
    pairs = np.frombuffer(img[48:48 + 16 * header_count], dtype=np.uint64).reshape(-1, 2)
    assert np.array_equal(pairs[:, 0], masks)
    assert np.array_equal(pairs[:, 1], scores.view(np.uint64))
    assert cb.local_mit_score(0x1) == 100.0 and cb.local_mit_score(0x5) == 5.219780219780221


def test_index_validation():
    case = golden_case("w8_families")
    img = case.issl
    ix = cb.Index(img)
    assert ix.info == {k: v for k, v in case.expected["header"].items() if k != "rc"}
    for cut, text in [(40, "header invalid"), (48 + 100, "header invalid"), (len(img) - 8, "slice contents"),
                      (48 + 16 * 6195 + 64, "off-target sequences")]:
        with pytest.raises(cb.IsslError) as e:
            cb.Index(img[:cut])
        assert e.value.code == 2 and text in e.value.message
    # a list-size table that does not sum to one entry per site per slice is refused
    bad = bytearray(img)
    off = 48 + 16 * 6195 + 8 * ix.info["offtargetsCount"]
    bad[off:off + 8] = (int.from_bytes(bad[off:off + 8], "little") + 1).to_bytes(8, "little")
    with pytest.raises(cb.IsslError) as e:
        cb.Index(bytes(bad))
    assert e.value.code == 6


def run_cli(*args):
    return subprocess.run([str(cb.cli_path()), *map(str, args)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)


def test_cli_contract_without_scoring(tmp_path):
    case = golden_case("tiny")
    (tmp_path / "i.issl").write_bytes(case.issl)
    (tmp_path / "g.txt").write_bytes(case.guides)
    p = run_cli()
    assert p.returncode == 1 and p.stdout == b"" and b"Usage:" in p.stderr and b"[issltable] [query file] [max distance]" in p.stderr
    p = run_cli(tmp_path / "i.issl", tmp_path / "g.txt", 4, 0)          # the reference would crash; we refuse
    assert p.returncode == 1 and p.stdout == b""
    # unknown method: every line SEQ\t-1\t-1 (isslScoreOfftargets.cpp:122-143, :517-525) -- nothing is scored
    p = run_cli(tmp_path / "i.issl", tmp_path / "g.txt", 4, 0, "bogus")
    want = [r for r in case.expected["runs"] if r["method"] == "bogus"][0]
    assert p.returncode == 0 and p.stdout.decode() == want["stdout"]
    # guide file not a multiple of the line length: same three stderr lines, exit 1
    (tmp_path / "bad.txt").write_bytes(case.guides[:-3])
    p = run_cli(tmp_path / "i.issl", tmp_path / "bad.txt", 4, 0, "mit")
    assert p.returncode == 1 and p.stdout == b""
    assert p.stderr.decode().splitlines()[0] == "Error: query file is not a multiple of the expected line length (21)"
    (tmp_path / "empty.txt").write_bytes(b"")
    p = run_cli(tmp_path / "i.issl", tmp_path / "empty.txt", 4, 0, "mit")
    assert p.returncode == 1 and b"Failed to read in query file." in p.stderr
    (tmp_path / "trunc.issl").write_bytes(case.issl[:30])
    p = run_cli(tmp_path / "trunc.issl", tmp_path / "g.txt", 4, 0, "mit")
    assert p.returncode == 1 and b"Error reading index: header invalid" in p.stderr and p.stdout == b""


def test_no_gpu_means_loud_failure_not_fallback(tmp_path):
    if cb.device_count() > 0:
        pytest.skip("a GPU is present")
    case = golden_case("tiny")
    with pytest.raises(cb.IsslError) as e:
        cb.Device.from_index(cb.Index(case.issl))
    assert e.value.code == 4
    (tmp_path / "i.issl").write_bytes(case.issl)
    (tmp_path / "g.txt").write_bytes(case.guides)
    p = run_cli(tmp_path / "i.issl", tmp_path / "g.txt", 4, 0, "mit")
    assert p.returncode == 1 and p.stdout == b"" and b"no CPU fallback" in p.stderr


def test_product_never_touches_the_oracle():
    for path in list((ROOT / "crackling_b200").rglob("*.py")) + list((ROOT / "crackling_b200" / "csrc").glob("*")) + \
            [ROOT / "include" / "issl_cuda.h"]:
        if path.is_file() and path.suffix in (".py", ".h", ".cpp", ".cu", ".cuh"):
            text = path.read_text(errors="replace")
            assert "issl_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
    out = subprocess.run(["ldd", str(cb.lib_path())], stdout=subprocess.PIPE).stdout.decode()
    assert "oracle" not in out


# --- ISSL_LAYOUT_TRIPLE: the sub-bucket visit table (host arithmetic, no GPU) -------------------------------

_HAM4 = np.array([sum(1 for f in range(4) if (x >> (2 * f)) & 3) for x in range(256)])


def test_triple_layout_tables():
    """The ten triples are all 3-subsets of the five slices; resp(E) always contains E (|E| <= 3) or is contained in
    it; every pair is the {M, H} of exactly one triple and every single slice the H of its triple, which is what puts
    the sub-buckets a guide reads next to each other (issl_triple_tables.h)."""
    import itertools
    lay, resp = cb.triple_layout()
    sets = [frozenset(int(x) for x in row[:3]) for row in lay]
    assert sorted(map(sorted, sets)) == [list(c) for c in itertools.combinations(range(5), 3)]
    for t, row in enumerate(lay):
        assert sorted(int(x) for x in row) == [0, 1, 2, 3, 4] and row[3] < row[4]
    assert resp[0] == 15
    for E in range(1, 32):
        Eset = frozenset(s for s in range(5) if E >> s & 1)
        T = sets[resp[E]]
        assert Eset <= T if len(Eset) <= 3 else T <= Eset
        if len(Eset) == 1:
            assert lay[resp[E]][2] in Eset                       # the exact slice is the high key byte
        if len(Eset) == 2:
            assert Eset == {int(lay[resp[E]][1]), int(lay[resp[E]][2])}   # the varied slice is the low key byte
    assert len({int(resp[(1 << i) | (1 << j)]) for i in range(5) for j in range(i + 1, 5)}) == 10


def test_triple_visit_table_counts():
    # 10 exact triples; + 10 pairs x {x: 1 <= ham4(x) <= D-2}; + 5 singles x {(xj, xk): ham4 sum <= D-2}
    for d, want in ((-1, 0), (0, 10), (2, 10), (3, 130), (4, 1390), (5, 8950), (6, 37300)):
        v, wave = cb.triple_visits(d)
        assert v.size == want and wave[0] == 0 and wave[5] == want
        assert np.unique(v & 0x0FFFFFFF).size == v.size, "a bucket would be read twice"
        assert np.all(np.diff(wave.astype(np.int64)) >= 0)


@pytest.mark.parametrize("max_dist", [0, 1, 3, 4, 5])
def test_triple_visits_cover_every_reference_hit_exactly_once(max_dist):
    """Brute force over planted neighbourhoods: the sites found through the visit table, filtered by the
    kernel's rule resp(E) == triple, are exactly the sites the reference scores (in a looked-up list, i.e.
    some slice matches exactly, and dist <= maxDist; isslScoreOfftargets.cpp:330-390), each exactly once and
    attributed to the lowest matching slice."""
    rng = np.random.default_rng(100 + max_dist)
    visits, wave = cb.triple_visits(max_dist)
    lay, resp = cb.triple_layout()
    by_wave = np.searchsorted(wave[1:], np.arange(visits.size), side="right")

    def slices(s):
        return [(s >> (8 * k)) & 0xFF for k in range(5)]

    def key_of(sl, t):
        return sl[lay[t][0]] | sl[lay[t][1]] << 8 | sl[lay[t][2]] << 16

    for g in (int(x) for x in rng.integers(0, 1 << 40, 4)):
        sites = {g}
        for _ in range(1500):
            s = g
            for p in rng.choice(20, int(rng.integers(0, 8)), replace=False):
                s ^= int(rng.integers(1, 4)) << (2 * int(p))
            sites.add(s)
        gs = slices(g)
        buckets = [dict() for _ in range(10)]
        for s in sites:
            sl = slices(s)
            for t in range(10):
                buckets[t].setdefault(key_of(sl, t), []).append(s)
        want = {}
        for s in sites:
            sl = slices(s)
            E = [k for k in range(5) if sl[k] == gs[k]]
            if E and sum(_HAM4[a ^ b] for a, b in zip(sl, gs)) <= max_dist:
                want[s] = min(E)
        got = {}
        for e, w in zip(visits.tolist(), by_wave.tolist()):
            t, budget, pat = (e >> 24) & 15, e >> 28, e & 0xFFFFFF
            comp = [int(lay[t][3]), int(lay[t][4])]
            for s in buckets[t].get(key_of(gs, t) ^ pat, []):
                sl = slices(s)
                if sum(_HAM4[sl[c] ^ gs[c]] for c in comp) > budget:
                    continue   # the kernel's fast-path filter must never drop a true hit (checked by got == want)
                E = sum(1 << k for k in range(5) if sl[k] == gs[k])
                if E and sum(_HAM4[a ^ b] for a, b in zip(sl, gs)) <= max_dist and resp[E] == t:
                    assert s not in got, "hit produced twice"
                    assert (E & -E).bit_length() - 1 == w, "hit attributed to the wrong slice wave"
                    got[s] = w
        assert got == want


def test_built_library_carries_the_tuned_scan_kernel():
    """The shipped libissl_cuda.so is the sm_100a build DESIGN.md 4 describes: the headline kernel reads a sub-block with
    two 32-byte loads (LDG.256, no L1 allocation, evict-first in L2), holds 48 registers (ten CTAs per SM), spills nothing
    and contains no tensor-core or TMA instruction -- and no kernel is built for any other architecture."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not pathlib.Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    lib = str(cb.binding.LIB)
    elf = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    usage = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    m = re.search(r"Function _ZN4issl21k_scan_triple_blockedILi2ELb1ELb0ELb0EEEvNS_10TripleArgsE:\s*\n\s*REG:(\d+) STACK:(\d+)", usage)
    assert m, "the headline instantiation k_scan_triple_blocked<2, true, false, false> is missing"
    assert (int(m.group(1)), int(m.group(2))) == (48, 0)
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN4issl21k_scan_triple_blockedILi2ELb1ELb0ELb0EEEvNS_10TripleArgsE", lib],
                          capture_output=True, text=True, check=True).stdout
    assert len(re.findall(r"LDG\.E\.NA\.EFL2\.256", sass)) == 2
    assert "IMAD.HI" in sass and "LOP3.LUT" in sass
    for absent in ("UTMALDG", "UTCMMA", "HMMA", "STL", "LDL"):
        assert absent not in sass, absent
