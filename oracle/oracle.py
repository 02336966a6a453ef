"""ctypes front-end of the CPU oracle (oracle/issl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (crackling_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libissl_oracle.so"
REF_DIR = HERE / "_ref"

METHODS = {"unknown": 0, "mit": 1, "cfd": 2, "and": 3, "or": 4, "avg": 5}

_lib = None


def build(force: bool = False) -> pathlib.Path:
    """Compile the C restatement (gcc) -- building the checker is not using it."""
    src = HERE / "issl_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    return LIB_PATH


def build_ref() -> bool:
    """Compile the unmodified reference into oracle/_ref when /root/reference is present."""
    subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)
    return have_ref()


def have_ref() -> bool:
    return all((REF_DIR / n).exists() for n in ("isslScoreOfftargets", "isslCreateIndex"))


def ref_binary(name: str) -> str:
    return str(REF_DIR / name)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        u8p, u64p, u32p, i32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint64, C.c_uint32, C.c_int32, C.c_double))
        L.oracle_score_issl.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_double,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.oracle_score_issl.restype = C.c_int
        L.oracle_cli.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_double, C.c_char_p,
                                 C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.oracle_cli.restype = C.c_int
        L.oracle_create_index.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.oracle_create_index.restype = C.c_int
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_sequence_to_signature.argtypes = [C.c_char_p, C.c_size_t]
        L.oracle_sequence_to_signature.restype = C.c_uint64
        L.oracle_sscore.argtypes = [C.c_uint64, C.c_size_t]
        L.oracle_sscore.restype = C.c_double
        L.oracle_issl_header.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_issl_header.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        del u8p, u64p, u32p, i32p, f64p
        _lib = L
    return _lib


def _img(img) -> np.ndarray:
    if isinstance(img, (bytes, bytearray, memoryview)):
        img = np.frombuffer(img, dtype=np.uint8)
    if isinstance(img, (str, os.PathLike)):
        img = np.fromfile(img, dtype=np.uint8)
    return np.ascontiguousarray(img, dtype=np.uint8)


def encode(seq: str | bytes, seq_length: int | None = None) -> int:
    b = seq.encode() if isinstance(seq, str) else seq
    return int(lib().oracle_sequence_to_signature(b, len(b) if seq_length is None else seq_length))


def encode_guides(guides, seq_length: int = 20) -> np.ndarray:
    return np.array([encode(g, seq_length) for g in guides], dtype=np.uint64)


def header(img) -> dict:
    a = _img(img)
    out = np.zeros(6, dtype=np.uint64)
    rc = lib().oracle_issl_header(a.ctypes.data, a.size, out.ctypes.data)
    keys = ("offtargetsCount", "seqLength", "seqCount", "sliceWidth", "sliceCount", "scoresCount")
    d = {k: int(x) for k, x in zip(keys, out)}
    d["rc"] = rc
    return d


def score(img, guides: np.ndarray, max_dist: int, threshold: float, method: str | int,
          threads: int = 1, want_hits: bool = False, want_candidates: bool = False):
    """Restatement of the scoring loop.  Returns dict(mit, cfd[, hits][, candidates]).

    hits is a structured array (guide, id, dist, occ) in encounter order (threads == 1)."""
    a = _img(img)
    g = np.ascontiguousarray(guides, dtype=np.uint64)
    n = g.size
    m = METHODS[method] if isinstance(method, str) else int(method)
    mit = np.empty(n, dtype=np.float64)
    cfd = np.empty(n, dtype=np.float64)
    cand = np.zeros(n, dtype=np.uint64) if want_candidates else None
    count = C.c_size_t(0)
    L = lib()

    def call(cap, hg, hi, hd, ho):
        return L.oracle_score_issl(a.ctypes.data, a.size, g.ctypes.data, n, max_dist, float(threshold), m, threads,
                                   mit.ctypes.data, cfd.ctypes.data,
                                   hg.ctypes.data if hg is not None else None,
                                   hi.ctypes.data if hi is not None else None,
                                   hd.ctypes.data if hd is not None else None,
                                   ho.ctypes.data if ho is not None else None,
                                   cap, C.byref(count) if want_hits else None,
                                   cand.ctypes.data if cand is not None else None)

    out = {}
    if want_hits:
        rc = call(0, None, None, None, None)
        if rc:
            raise RuntimeError(f"oracle_score_issl failed rc={rc}")
        cap = count.value
        hg = np.zeros(max(cap, 1), dtype=np.uint64); hi = np.zeros(max(cap, 1), dtype=np.uint32)
        hd = np.zeros(max(cap, 1), dtype=np.int32); ho = np.zeros(max(cap, 1), dtype=np.uint32)
        rc = call(cap, hg, hi, hd, ho)
        hits = np.zeros(cap, dtype=[("guide", "u8"), ("id", "u4"), ("dist", "i4"), ("occ", "u4")])
        hits["guide"], hits["id"], hits["dist"], hits["occ"] = hg[:cap], hi[:cap], hd[:cap], ho[:cap]
        out["hits"] = hits
    else:
        rc = call(0, None, None, None, None)
    if rc:
        raise RuntimeError(f"oracle_score_issl failed rc={rc}")
    out["mit"], out["cfd"] = mit, cfd
    if cand is not None:
        out["candidates"] = cand
    return out


def cli(img, guide_file: bytes, max_dist: int, threshold: float, method: str, threads: int = 1):
    """main() on memory buffers -> (exit status, stdout bytes)."""
    a = _img(img)
    hd = header(a)
    n = len(guide_file) // (hd["seqLength"] + 1) if hd["seqLength"] < 1000 else 0
    cap = max(1, n) * (hd["seqLength"] + 800) + 64
    buf = C.create_string_buffer(cap)
    out_len = C.c_size_t(0)
    rc = lib().oracle_cli(a.ctypes.data, a.size, guide_file, len(guide_file), max_dist, float(threshold),
                          method.encode(), threads, buf, cap, C.byref(out_len))
    return rc, buf.raw[:out_len.value]


def create_index(text: bytes, seq_length: int, slice_width: int) -> bytes:
    """Restatement of isslCreateIndex -> .issl image bytes."""
    p = C.c_void_p()
    n = C.c_size_t(0)
    rc = lib().oracle_create_index(text, len(text), seq_length, slice_width, C.byref(p), C.byref(n))
    if rc:
        raise RuntimeError(f"oracle_create_index failed rc={rc}")
    try:
        return C.string_at(p.value, n.value)
    finally:
        lib().oracle_free(p)


def sscore(mask: int, seq_length: int = 20) -> float:
    return float(lib().oracle_sscore(mask, seq_length))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# ---------------------------------------------------------------------------------------------
# the real reference, when oracle/_ref holds its binaries
# ---------------------------------------------------------------------------------------------
def ref_create_index(text_path: str, seq_length: int, slice_width: int, out_path: str) -> None:
    subprocess.run([ref_binary("isslCreateIndex"), text_path, str(seq_length), str(slice_width), out_path],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def ref_score(issl_path: str, guides_path: str, max_dist: int, threshold, method: str,
              threads: int | None = None, hits: bool = False):
    """Runs the reference scorer; returns (returncode, stdout bytes[, hit tuples])."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    exe = ref_binary("isslScoreOfftargets_hits" if hits else "isslScoreOfftargets")
    p = subprocess.run([exe, issl_path, guides_path, str(max_dist), str(threshold), method],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    if not hits:
        return p.returncode, p.stdout
    rows = [tuple(int(x) for x in ln.split(b"\t")[1:]) for ln in p.stderr.splitlines() if ln.startswith(b"HIT\t")]
    arr = np.array(rows, dtype=np.int64).reshape(-1, 4)
    return p.returncode, p.stdout, arr
