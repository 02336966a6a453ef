"""numpy restatement of the reference's off-target extraction.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/crackling/utils/extractOfftargets.py: the two look-ahead regexes at
:23-24, the slicing at :97-106 -- including this snapshot's quirk that the REVERSE-strand site is
rc(match[0:20]), the reverse complement of the FIRST 20 characters of the 23-character match --
and the global sort at :112-191 (every record sorted, then merged: one globally sorted file).
Pinned by tests/golden/c1 (sha256 of the real tool's output for the same genome).
"""
from __future__ import annotations

import numpy as np

A, C, G, T = (ord(c) for c in "ACGT")
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP[a] = b


def _windows(seq: np.ndarray, starts: np.ndarray, width: int) -> np.ndarray:
    return seq[starts[:, None] + np.arange(width)[None, :]]


def extract_sites(records: list[bytes]) -> np.ndarray:
    """All sites (n x 20 uint8 letters) of all records, unsorted."""
    out = []
    for rec in records:
        s = np.frombuffer(rec.upper(), dtype=np.uint8)
        n = s.size
        if n < 23:
            continue
        i = np.arange(n - 22)
        # forward: [ACG][ACGT]{19}[ACGT][AG]G -> match[0:20]
        valid = np.ones(n - 22, dtype=bool)
        acgt = (s == A) | (s == C) | (s == G) | (s == T)
        bad = np.concatenate([[0], np.cumsum(~acgt)])
        allowed = (bad[i + 23] - bad[i]) == 0                      # all 23 characters are ACGT
        fwd = allowed & (s[i] != T) & ((s[i + 21] == A) | (s[i + 21] == G)) & (s[i + 22] == G)
        out.append(_windows(s, i[fwd], 20))
        # reverse: C[CT][ACGT][ACGT]{19}[TGC] -> rc(match[0:20])
        rev = allowed & (s[i] == C) & ((s[i + 1] == C) | (s[i + 1] == T)) & (s[i + 22] != A)
        w = _windows(s, i[rev], 20)
        out.append(_COMP[w[:, ::-1]])
        del valid
    return np.concatenate(out, axis=0) if out else np.zeros((0, 20), dtype=np.uint8)


def extract_offtargets(records: list[bytes]) -> bytes:
    """The sorted, LF-terminated text file the reference tool writes."""
    sites = extract_sites(records)
    if sites.shape[0] == 0:
        return b""
    order = np.lexsort(sites.T[::-1])
    lines = np.concatenate([sites[order], np.full((sites.shape[0], 1), ord("\n"), dtype=np.uint8)], axis=1)
    return lines.tobytes()


def guide_candidates(records: list[bytes]) -> np.ndarray:
    """First 20 bases of every forward [ACGT]{21}GG site (mirrors Crackling.py:153) -- guide pool."""
    out = []
    for rec in records:
        s = np.frombuffer(rec.upper(), dtype=np.uint8)
        i = np.arange(s.size - 22)
        ok = (s[i + 21] == G) & (s[i + 22] == G)
        out.append(_windows(s, i[ok], 20))
    return np.concatenate(out, axis=0)


# ---- FASTA reading (the tool's two input paths) -------------------------------------------------

def _lines(raw: bytes):
    """Lines as Python's open(path, 'r') yields them: universal newlines, terminators translated to LF."""
    import io
    return io.TextIOWrapper(io.BytesIO(raw), encoding="utf-8", newline=None)


def read_single_input(raw: bytes) -> list[bytes]:
    """One input file: explodeMultiFastaFile (extractOfftargets.py:26-62) writes each record to its own file
    (header, then every other line strip()ped and upper-cased, concatenated), and processingNode (:73-90) reads
    that file back.  The tool indexes line[0] of the stripped line, so a blank line -- or sequence before the
    first header -- makes it crash; here blank lines are skipped and leading sequence forms a record."""
    records, cur = [], None
    for line in _lines(raw):
        line = line.strip()
        if not line:
            continue
        if line[0] == ">":
            cur = []
            records.append(cur)
        else:
            if cur is None:
                cur = []
                records.append(cur)
            cur.append(line.upper().strip())
    return ["".join(r).rstrip().upper().encode() for r in records]


def read_one_of_several_inputs(raw: bytes) -> list[bytes]:
    """Several input files: processingNode (:73-90) reads each directly.  Records are keyed by header text
    (line[1:]), a repeated header starts its record afresh (:83), lines are rstrip()ped only (:89)."""
    seqs, header = {}, None
    for line in _lines(raw):
        if line[0] == ">":
            header = line[1:]
            seqs[header] = []
        else:
            if header not in seqs:
                seqs[header] = []
            seqs[header].append(line.rstrip().upper())
    return ["".join(v).encode() for v in seqs.values()]


def extract_from_inputs(inputs: list[bytes]) -> bytes:
    """What `extractOfftargets.py output inputs...` writes for these file contents."""
    if len(inputs) == 1:
        records = read_single_input(inputs[0])
    else:
        records = [r for raw in inputs for r in read_one_of_several_inputs(raw)]
    return extract_offtargets(records)
