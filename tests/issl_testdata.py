"""Seeded synthetic inputs for the ISSL scorer tests (numpy only, no reference needed).

`make_offtargets` returns the sorted, LF-terminated fixed-width text file that the reference's
isslCreateIndex consumes (isslCreateIndex.cpp:138-207): uniform random sites plus planted
near-repeat families and exact duplicates, so that small indexes still produce many hits,
multi-slice duplicates and occurrence counts > 1.  `make_guides` returns the guide file the
scorer consumes (isslScoreOfftargets.cpp:275-294).
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _to_lines(codes: np.ndarray) -> list[bytes]:
    return [BASES[row].tobytes() for row in codes]


def make_offtargets(seed: int, n_random: int = 2000, n_families: int = 12, family_size: int = 40,
                    max_sub_rate: float = 0.15, dup_frac: float = 0.05, seq_length: int = 20,
                    low_complexity: int = 8) -> bytes:
    rng = np.random.default_rng(seed)
    sites = [rng.integers(0, 4, size=(n_random, seq_length), dtype=np.uint8)]
    for _ in range(n_families):
        root = rng.integers(0, 4, size=seq_length, dtype=np.uint8)
        rate = rng.uniform(0.0, max_sub_rate)
        fam = np.tile(root, (family_size, 1))
        sub = rng.random(fam.shape) < rate
        fam = np.where(sub, (fam + rng.integers(1, 4, size=fam.shape, dtype=np.uint8)) % 4, fam).astype(np.uint8)
        sites.append(fam)
    if low_complexity:
        lc = np.zeros((low_complexity, seq_length), dtype=np.uint8)
        for k in range(low_complexity):
            a, b = rng.integers(0, 4, size=2)
            lc[k, 0::2], lc[k, 1::2] = a, b
            flip = rng.integers(0, seq_length)
            lc[k, flip] = (lc[k, flip] + 1) % 4
        sites.append(lc)
    allsites = np.concatenate(sites, axis=0)
    ndup = int(len(allsites) * dup_frac)
    if ndup:
        pick = rng.integers(0, len(allsites), size=ndup)
        reps = rng.integers(1, 5, size=ndup)
        allsites = np.concatenate([allsites, np.repeat(allsites[pick], reps, axis=0)], axis=0)
    lines = sorted(_to_lines(allsites))
    return b"".join(l + b"\n" for l in lines)


def make_guides(seed: int, offtargets: bytes, n: int = 64, seq_length: int = 20,
                frac_exact: float = 0.4, frac_mut: float = 0.4) -> bytes:
    """Guides: exact copies of sites, sites with 1-4 substitutions, and uniform random 20-mers."""
    rng = np.random.default_rng(seed)
    line = seq_length + 1
    nsites = len(offtargets) // line
    arr = np.frombuffer(offtargets, dtype=np.uint8).reshape(nsites, line)[:, :seq_length]
    lut = np.zeros(256, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    codes = lut[arr]
    out = []
    for _ in range(n):
        u = rng.random()
        if u < frac_exact and nsites:
            g = codes[rng.integers(0, nsites)].copy()
        elif u < frac_exact + frac_mut and nsites:
            g = codes[rng.integers(0, nsites)].copy()
            k = rng.integers(1, 5)
            pos = rng.choice(seq_length, size=k, replace=False)
            g[pos] = (g[pos] + rng.integers(1, 4, size=k, dtype=np.uint8)) % 4
        else:
            g = rng.integers(0, 4, size=seq_length, dtype=np.uint8)
        out.append(BASES[g].tobytes())
    return b"".join(g + b"\n" for g in out)


def pack_guides(guide_file: bytes, seq_length: int = 20) -> np.ndarray:
    """2-bit packing of a guide file, numpy restatement of isslScoreOfftargets.cpp:63-71."""
    line = seq_length + 1
    n = len(guide_file) // line
    arr = np.frombuffer(guide_file, dtype=np.uint8).reshape(n, line)[:, :seq_length]
    lut = np.zeros(256, dtype=np.uint64)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    codes = lut[arr]
    shifts = (np.arange(seq_length, dtype=np.uint64) * np.uint64(2))
    return np.bitwise_or.reduce(codes << shifts, axis=1).astype(np.uint64) if n else np.zeros(0, dtype=np.uint64)


def make_genome(seed: int, lengths=(4_800_000, 200_000)) -> list[bytes]:
    """i.i.d. uniform ACGT records (config 1 of BASELINE.json: a 5 Mbp bacterial-size genome; two
    records because the reference extractor fails on a single-record FASTA)."""
    rng = np.random.default_rng(seed)
    return [BASES[rng.integers(0, 4, size=n, dtype=np.uint8)].tobytes() for n in lengths]


def write_fasta(path, records: list[bytes], width: int = 70) -> None:
    with open(path, "wb") as f:
        for k, rec in enumerate(records):
            f.write(b">record%d\n" % k)
            for o in range(0, len(rec), width):
                f.write(rec[o:o + width] + b"\n")


def sample_guides(seed: int, pool: np.ndarray, n: int) -> bytes:
    """n guides drawn without replacement from a pool of 20-mers (rows of letters)."""
    rng = np.random.default_rng(seed)
    pick = rng.choice(pool.shape[0], size=n, replace=False)
    lines = np.concatenate([pool[pick], np.full((n, 1), ord("\n"), dtype=np.uint8)], axis=1)
    return lines.tobytes()
