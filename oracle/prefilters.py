"""Test infrastructure (never imported by the product): CPU restatement of Crackling's sequence-only guide filters,
/root/reference/src/crackling/Crackling.py:312-384, which sit between guide extraction (:151-165) and the scorer
call (:747-778).  Pinned to the reference where it can be imported: AT_percentage is checked against
crackling.Helpers.AT_percentage (Helpers.py:21-27) by tests/test_prefilters.py when /root/reference is present, and
the committed vectors in tests/golden/prefilters.json were produced by tests/golden/make_prefilter_golden.py running
that function and the literal conditions below."""

FILTER_G20, FILTER_LEADING_T, FILTER_AT, FILTER_TTTT = 1, 2, 4, 8


def at_percentage(seq: str) -> float:
    """Helpers.py:21-27"""
    total = 0.0
    length = float(len(seq))
    for c in seq:
        if c in "AT":
            total += 1
    return 100.0 * total / length


def filters(target23: str):
    """(flags of the filters the target fails, AT % of the 20-mer)"""
    f = 0
    if target23[19] != 'G':                                           # Crackling.py:320
        f |= FILTER_G20
    if (target23[-2:] == 'GG' and target23[0] == 'T') or \
       (target23[:2] == 'CC' and target23[-1] == 'A'):               # Crackling.py:338-339
        f |= FILTER_LEADING_T
    at = at_percentage(target23[0:20])                                # Crackling.py:358
    if at < 20 or at > 65:                                            # Crackling.py:360
        f |= FILTER_AT
    if 'TTTT' in target23:                                            # Crackling.py:380
        f |= FILTER_TTTT
    return f, at


def encode20(target23: str) -> int:
    """What the scorer packs from the line Crackling.py:750-751 writes (isslScoreOfftargets.cpp:63-71, :99-102)."""
    lut = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
    return sum(lut.get(c, 0) << (2 * j) for j, c in enumerate(target23[0:20]))


FILTER_DUPLICATE, FILTER_NOT_UNIQUE = 16, 32


def duplicates(targets23):
    """Crackling.py:211-240 over the targets in discovery order, and what :291-296 later derives from it:
    (flags per target, numDuplicateGuides, len(duplicateGuides))."""
    candidateGuides = set()
    duplicateGuides = set()
    numDuplicateGuides = 0
    later = []
    for guide in targets23:
        if guide not in candidateGuides:                              # Crackling.py:221
            candidateGuides.add(guide)
            later.append(False)
        else:
            duplicateGuides.add(guide)                                # Crackling.py:228
            numDuplicateGuides += 1
            later.append(True)
    flags = [(FILTER_DUPLICATE if l else 0) | (FILTER_NOT_UNIQUE if g in duplicateGuides else 0)   # :291 `if row[0] in duplicateGuides`
             for g, l in zip(targets23, later)]
    return flags, numDuplicateGuides, len(duplicateGuides)
