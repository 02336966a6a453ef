"""The scorer's call site in Crackling's pipeline, without its temp files.

`/root/reference/src/crackling/Crackling.py:737-786` writes every page of candidate guides to a text file, shells out
to `isslScoreOfftargets` (which reloads the whole index), and parses the output file back into
`targetsScored[seq20] = {'MIT': ..., 'CFD': ...}`.  `OfftargetScorer` is what a maintainer puts in place of that block:
the index goes to HBM once, a page is a list of target23 strings in and the same dictionary out -- with the values the
pipeline would have read from the file (the reference prints "%f", Crackling.py parses it with float(): six decimals),
so everything after :786 (thresholds at :795-830, the output CSV) behaves identically.

`candidate_flags` is the sequence-only part of what comes before (duplicates :211-240 / :291-296, G20 :318-326,
leading T :336-346, AT % :356-368, TTTT :378-384) as device passes over the same text.

This module adds no compute of its own: everything runs in libissl_cuda.so (no CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import binding as cb

FILTER_G20, FILTER_LEADING_T, FILTER_AT, FILTER_TTTT, FILTER_DUPLICATE, FILTER_NOT_UNIQUE = 1, 2, 4, 8, 16, 32


def _text23(targets23) -> bytes:
    text = ("\n".join(targets23) + "\n").encode() if len(targets23) else b""
    if len(text) != 24 * len(targets23):
        raise ValueError("every target must be 23 characters (20-mer + PAM), as Crackling.py:151-165 yields them")
    return text


class OfftargetScorer:
    """Create once per run (per GPU set), reuse for every page -- replaces Crackling.py:737-786."""

    def __init__(self, issl_path: str, cuda_devices=(0,), layout: str = "auto"):
        self.index = cb.Index(issl_path)
        first = cb.Device.from_index(self.index, int(cuda_devices[0]), layout)
        # further GPUs get a replica by peer copy (issl_device_clone), not another upload from the host
        self.devices = cb.replicate(first, [int(d) for d in cuda_devices]) if len(cuda_devices) > 1 else [first]
        self.seq_length = int(self.index.info["seqLength"])

    def close(self):
        for d in self.devices:
            d.close()
        self.devices = []
        self.index.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def candidate_flags(self, targets23):
        """(flags, AT %, numDuplicateGuides, len(duplicateGuides)): ISSL_FILTER_* bits per target, in input order."""
        text = _text23(targets23)
        if not text:
            return np.zeros(0, np.uint8), np.zeros(0), 0, 0
        flags, at, _ = self.devices[0].guide_filters(text)
        dup, later, seqs = self.devices[0].guide_duplicates(text)
        return flags | dup, at, later, seqs

    def score_arrays(self, targets23, max_dist: int, threshold: float, method: str):
        """(packed 20-mers, mit, cfd) as the library returns them (full precision); a column the method skips is None."""
        text = _text23(targets23)
        n = len(targets23)
        if n == 0:
            return np.zeros(0, np.uint64), None, None
        _, _, packed = self.devices[0].guide_filters(text)
        if len(self.devices) == 1:
            mit, cfd = self.devices[0].score(packed, max_dist, threshold, method)
            return packed, mit, cfd
        hg, hm, hc = cb.HostBuffer(n, np.uint64), cb.HostBuffer(n, np.float64), cb.HostBuffer(n, np.float64)
        try:
            hg.array[:] = packed
            cb.score_multi(self.devices, hg.array, max_dist, threshold, method, hm.array, hc.array)
            code = cb.method_code(method)
            mit = hm.array.copy() if code != cb.METHODS["cfd"] else None
            cfd = hc.array.copy() if code != cb.METHODS["mit"] else None
        finally:
            for b in (hg, hm, hc):
                b.close()
        return packed, mit, cfd

    def score_page(self, targets23, max_dist: int, threshold: float, method: str) -> dict:
        """targetsScored of Crackling.py:780-786: {target23[0:20]: {'MIT': float, 'CFD': float}} for one page."""
        packed, mit, cfd = self.score_arrays(targets23, max_dist, threshold, method)
        if packed.size == 0:
            return {}
        # the values the pipeline reads are the printed ones: same bytes as the reference's stdout, parsed as :783-785 does
        lines = cb.format_lines(packed, mit, cfd, method, self.seq_length).split(b"\n")
        out = {}
        for line in lines:
            f = line.split(b"\t")
            if len(f) == 3:
                out[f[0].decode()] = {"MIT": float(f[1]), "CFD": float(f[2])}
        return out
