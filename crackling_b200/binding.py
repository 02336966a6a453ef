"""ctypes mirror of include/issl_cuda.h (same names, same argument meaning, same error behaviour)."""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "crackling_b200" / "lib" / "libissl_cuda.so"
if os.environ.get("ISSL_CUDA_LIB"):   # A/B timing of kernel variants built into another directory (tools/ab_build.sh); never a fallback
    LIB = pathlib.Path(os.environ["ISSL_CUDA_LIB"]).resolve()
CLI = ROOT / "bin" / "isslScoreOfftargets"
CREATE_CLI = ROOT / "bin" / "isslCreateIndex"
EXTRACT_CLI = ROOT / "bin" / "extractOfftargets"

METHODS = {"unknown": 0, "mit": 1, "cfd": 2, "and": 3, "or": 4, "avg": 5}
LAYOUTS = {"auto": 0, "res32": 1, "sig64": 2, "gather": 3, "triple": 4}


class IsslError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[issl status {code}] {message}")
        self.code = code
        self.message = message


class _Info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("offtargetsCount", "seqLength", "seqCount", "sliceWidth", "sliceCount", "scoresCount")]


class _DeviceInfo(C.Structure):
    _fields_ = [("cuda_device", C.c_int), ("layout", C.c_int), ("bytes_per_candidate", C.c_uint32),
                ("hbm_bytes", C.c_uint64), ("list_entries", C.c_uint64), ("info", _Info),
                ("triple_block_bytes", C.c_uint32), ("triple_hit_bytes", C.c_uint32)]


class _SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("uniform_sites", C.c_uint64), ("families", C.c_uint32), ("family_size_min", C.c_uint32),
                ("family_size_max", C.c_uint32), ("max_sub_rate", C.c_double), ("low_complexity_fraction", C.c_double),
                ("seqLength", C.c_uint32), ("sliceWidth", C.c_uint32)]


class _Stats(C.Structure):
    _fields_ = [("guides", C.c_uint64), ("candidates", C.c_uint64), ("hits", C.c_uint64),
                ("scan_launches", C.c_uint64), ("launches", C.c_uint64), ("scan_ms", C.c_double),
                ("total_ms", C.c_double), ("early_exits", C.c_uint64), ("streamed", C.c_uint64),
                ("bucket_visits", C.c_uint64), ("heavy_hits", C.c_uint64), ("sorted_hits", C.c_uint64), ("heavy_ms", C.c_double)]


_lib = None


def build(force: bool = False) -> pathlib.Path:
    """Compile libissl_cuda.so and bin/isslScoreOfftargets in-tree (nvcc, sm_100a)."""
    if force:
        subprocess.run(["make", "-s", "-C", str(ROOT), "clean"], check=True)
    subprocess.run(["make", "-s", "-C", str(ROOT), "all"], check=True)
    return LIB


def lib_path() -> pathlib.Path:
    return LIB


def cli_path() -> pathlib.Path:
    return CLI


def create_cli_path() -> pathlib.Path:
    return CREATE_CLI


def extract_cli_path() -> pathlib.Path:
    return EXTRACT_CLI


def lib() -> C.CDLL:
    """Loads the in-tree shared library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise IsslError(-1, f"{LIB} is missing: run `make` (or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(str(LIB))
        vp, sz, u64, i, d = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_double
        pp = C.POINTER(C.c_void_p)
        sigs = {
            "issl_index_open": ([C.c_char_p, pp], i),
            "issl_index_from_memory": ([vp, sz, pp], i),
            "issl_index_info": ([vp, C.POINTER(_Info)], i),
            "issl_index_close": ([vp], None),
            "issl_pack_guides": ([C.c_char_p, sz, sz, vp], i),
            "issl_unpack_guide": ([u64, sz, C.c_char_p], None),
            "issl_method_from_string": ([C.c_char_p], i),
            "issl_format_lines": ([vp, vp, vp, sz, sz, i, vp, sz], sz),
            "issl_device_count": ([], i),
            "issl_device_create": ([vp, i, i, pp], i),
            "issl_device_create_synthetic": ([i, i, u64, u64, C.c_uint32, C.c_uint32, d, C.c_uint32, C.c_uint32, pp], i),
            "issl_device_create_synthetic_ex": ([i, i, vp, pp], i),
            "issl_device_list_lengths": ([vp, vp, sz], sz),
            "issl_device_create_from_text": ([C.c_char_p, sz, C.c_uint32, C.c_uint32, i, i, pp], i),
            "issl_sites_create": ([i, pp], i),
            "issl_sites_destroy": ([vp], None),
            "issl_sites_add_fasta": ([vp, C.c_char_p, sz, i], i),
            "issl_sites_count": ([vp, C.POINTER(u64), C.POINTER(u64)], i),
            "issl_sites_write_text": ([vp, C.c_char_p], i),
            "issl_sites_read_keys": ([vp, u64, u64, vp], i),
            "issl_device_create_from_sites": ([vp, C.c_uint32, i, pp], i),
            "issl_device_clone": ([vp, i, pp], i),
            "issl_device_get_info": ([vp, C.POINTER(_DeviceInfo)], i),
            "issl_device_destroy": ([vp], None),
            "issl_device_write_issl": ([vp, C.c_char_p], i),
            "issl_device_read_sites": ([vp, vp, u64, vp], i),
            "issl_score": ([vp, vp, sz, i, d, i, vp, vp], i),
            "issl_score_device": ([vp, vp, sz, i, d, i, vp, vp, vp], i),
            "issl_score_hits": ([vp, vp, sz, i, d, i, vp, vp, vp, vp, vp, vp, sz, C.POINTER(sz)], i),
            "issl_last_stats": ([vp, C.POINTER(_Stats)], i),
            "issl_score_multi": ([vp, sz, vp, sz, i, d, i, vp, vp, sz, C.POINTER(_Stats), vp], i),
            "issl_multi_chunk": ([sz, sz], sz),
            "issl_host_alloc": ([sz, pp], i),
            "issl_host_free": ([vp], None),
            "issl_guide_filters": ([vp, C.c_char_p, sz, vp, vp, vp], i),
            "issl_guide_duplicates": ([vp, C.c_char_p, sz, vp, C.POINTER(u64), C.POINTER(u64)], i),
            "issl_local_mit_score": ([u64, sz], d),
            "issl_mit_table": ([sz, sz, vp, vp, sz, C.POINTER(u64)], sz),
            "issl_triple_visits": ([C.c_int, vp, sz, vp], sz),
            "issl_triple_visits_w4": ([C.c_int, vp, sz, vp], sz),
            "issl_triple_layout": ([vp, vp], None),
            "issl_last_error": ([], C.c_char_p),
            "issl_abi_version": ([], i),
        }
        for name, (argtypes, restype) in sigs.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = argtypes, restype
        L._issl_symbols = tuple(sigs)
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise IsslError(rc, lib().issl_last_error().decode(errors="replace"))


def method_code(name) -> int:
    return int(name) if isinstance(name, int) else lib().issl_method_from_string(str(name).encode())


def device_count() -> int:
    return lib().issl_device_count()


def pack_guides(text: bytes, seq_length: int = 20) -> np.ndarray:
    n = len(text) // (seq_length + 1)
    out = np.zeros(n, dtype=np.uint64)
    _check(lib().issl_pack_guides(text, len(text), seq_length, out.ctypes.data))
    return out


def unpack_guide(sig: int, seq_length: int = 20) -> str:
    buf = C.create_string_buffer(seq_length + 1)
    lib().issl_unpack_guide(int(sig), seq_length, buf)
    return buf.raw[:seq_length].decode()


def format_lines(guides: np.ndarray, mit, cfd, method, seq_length: int = 20) -> bytes:
    """issl_format_lines: the reference's stdout lines for these guides and scores."""
    g = np.ascontiguousarray(guides, dtype=np.uint64)
    m = method_code(method)
    mit = None if mit is None else np.ascontiguousarray(mit, dtype=np.float64)
    cfd = None if cfd is None else np.ascontiguousarray(cfd, dtype=np.float64)
    args = (g.ctypes.data, mit.ctypes.data if mit is not None else None, cfd.ctypes.data if cfd is not None else None,
            g.size, seq_length, m)
    need = lib().issl_format_lines(*args, None, 0)
    buf = C.create_string_buffer(max(need, 1))
    n = lib().issl_format_lines(*args, buf, need)
    return buf.raw[:n]


def local_mit_score(mask: int, seq_length: int = 20) -> float:
    return float(lib().issl_local_mit_score(int(mask), seq_length))


def mit_table(seq_length: int, slice_width: int):
    count = C.c_uint64(0)
    lib().issl_mit_table(seq_length, slice_width, None, None, 0, C.byref(count))
    masks = np.zeros(count.value, dtype=np.uint64)
    scores = np.zeros(count.value, dtype=np.float64)
    n = lib().issl_mit_table(seq_length, slice_width, masks.ctypes.data, scores.ctypes.data, count.value, C.byref(count))
    return masks[:n], scores[:n], int(count.value)


def triple_visits(max_dist: int):
    """(entries, waveStart[6]) of issl_triple_visits: pattern24 | triple << 24 | budget << 28, ordered by slice."""
    wave = np.zeros(6, dtype=np.uint32)
    n = lib().issl_triple_visits(int(max_dist), None, 0, wave.ctypes.data)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    n = lib().issl_triple_visits(int(max_dist), out.ctypes.data, n, wave.ctypes.data)
    return out[:n], wave


def triple_layout():
    """(slices[10][5], resp[32]) of issl_triple_layout."""
    sl = np.zeros(50, dtype=np.uint8)
    resp = np.zeros(32, dtype=np.uint8)
    lib().issl_triple_layout(sl.ctypes.data, resp.ctypes.data)
    return sl.reshape(10, 5), resp


def _info_dict(s) -> dict:
    return {n: int(getattr(s, n)) for n, _ in s._fields_}


class Index:
    """A parsed .issl image (issl_index_open / issl_index_from_memory)."""

    def __init__(self, source):
        self._h = C.c_void_p()
        self._keep = None
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else source
            n = buf.size
            aligned = np.zeros((n + 7) // 8 + 1, dtype=np.uint64)        # 8-byte aligned private copy
            aligned.view(np.uint8)[:n] = buf
            self._keep = aligned
            _check(lib().issl_index_from_memory(aligned.ctypes.data, n, C.byref(self._h)))
        else:
            _check(lib().issl_index_open(str(source).encode(), C.byref(self._h)))

    @property
    def info(self) -> dict:
        s = _Info()
        _check(lib().issl_index_info(self._h, C.byref(s)))
        return _info_dict(s)

    def close(self):
        if self._h:
            lib().issl_index_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Device:
    """An index resident in one GPU's HBM (issl_device_create / issl_device_create_synthetic)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_index(cls, index: Index, cuda_device: int = 0, layout: str | int = "auto") -> "Device":
        h = C.c_void_p()
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        _check(lib().issl_device_create(index._h, cuda_device, lay, C.byref(h)))
        return cls(h)

    @classmethod
    def synthetic(cls, cuda_device: int = 0, layout: str | int = "auto", seed: int = 1, uniform_sites: int = 1 << 20,
                  families: int = 0, family_size: int = 0, max_sub_rate: float = 0.15, seq_length: int = 20,
                  slice_width: int = 8) -> "Device":
        h = C.c_void_p()
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        _check(lib().issl_device_create_synthetic(cuda_device, lay, seed, uniform_sites, families, family_size,
                                                  float(max_sub_rate), seq_length, slice_width, C.byref(h)))
        return cls(h)

    @classmethod
    def synthetic_ex(cls, cuda_device: int = 0, layout: str | int = "auto", seed: int = 1, uniform_sites: int = 1 << 20,
                     families: int = 0, family_size_min: int = 0, family_size_max: int = 0, max_sub_rate: float = 0.15,
                     low_complexity_fraction: float = 0.0, seq_length: int = 20, slice_width: int = 8) -> "Device":
        """issl_device_create_synthetic_ex: log-uniform family sizes and low-complexity tracts (BASELINE.json configs[3])."""
        h = C.c_void_p()
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        spec = _SynthSpec(seed, uniform_sites, families, family_size_min, family_size_max, float(max_sub_rate),
                          float(low_complexity_fraction), seq_length, slice_width)
        _check(lib().issl_device_create_synthetic_ex(cuda_device, lay, C.byref(spec), C.byref(h)))
        return cls(h)

    @property
    def list_lengths(self) -> np.ndarray:
        """Lengths of the slice lists, slice-major (issl_device_list_lengths)."""
        n = lib().issl_device_list_lengths(self._h, None, 0)
        out = np.zeros(n, dtype=np.uint64)
        lib().issl_device_list_lengths(self._h, out.ctypes.data, n)
        return out

    @classmethod
    def from_text(cls, text: bytes, seq_length: int = 20, slice_width: int = 8, cuda_device: int = 0,
                  layout: str | int = "auto") -> "Device":
        """issl_device_create_from_text: the device-side isslCreateIndex."""
        h = C.c_void_p()
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        _check(lib().issl_device_create_from_text(text, len(text), seq_length, slice_width, cuda_device, lay, C.byref(h)))
        return cls(h)

    @classmethod
    def from_sites(cls, sites: "Sites", slice_width: int = 8, layout: str | int = "auto") -> "Device":
        """issl_device_create_from_sites: extracted sites -> index, on the device."""
        h = C.c_void_p()
        lay = LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        _check(lib().issl_device_create_from_sites(sites._h, slice_width, lay, C.byref(h)))
        return cls(h)

    def clone(self, cuda_device: int) -> "Device":
        """issl_device_clone: a replica of this index on another GPU, copied over NVLink."""
        h = C.c_void_p()
        _check(lib().issl_device_clone(self._h, int(cuda_device), C.byref(h)))
        return Device(h)

    @property
    def info(self) -> dict:
        s = _DeviceInfo()
        _check(lib().issl_device_get_info(self._h, C.byref(s)))
        d = {"cuda_device": s.cuda_device, "layout": s.layout, "bytes_per_candidate": s.bytes_per_candidate,
             "hbm_bytes": int(s.hbm_bytes), "list_entries": int(s.list_entries),
             "triple_block_bytes": int(s.triple_block_bytes), "triple_hit_bytes": int(s.triple_hit_bytes)}
        d.update(_info_dict(s.info))
        return d

    def guide_filters(self, text23: bytes):
        """(flags, AT %, packed 20-mers) of n lines of 23 characters + LF (issl_guide_filters)."""
        n = len(text23) // 24
        flags = np.zeros(n, dtype=np.uint8)
        at = np.zeros(n, dtype=np.float64)
        packed = np.zeros(n, dtype=np.uint64)
        _check(lib().issl_guide_filters(self._h, text23, len(text23), flags.ctypes.data, at.ctypes.data, packed.ctypes.data))
        return flags, at, packed

    def guide_duplicates(self, text23: bytes):
        """(flags, numDuplicateGuides, len(duplicateGuides)) of n lines of 23 characters + LF (issl_guide_duplicates)."""
        n = len(text23) // 24
        flags = np.zeros(max(n, 1), dtype=np.uint8)
        later, seqs = C.c_uint64(0), C.c_uint64(0)
        _check(lib().issl_guide_duplicates(self._h, text23, len(text23), flags.ctypes.data, C.byref(later), C.byref(seqs)))
        return flags[:n], int(later.value), int(seqs.value)

    @property
    def stats(self) -> dict:
        s = _Stats()
        _check(lib().issl_last_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def score(self, guides: np.ndarray, max_dist: int, threshold: float, method):
        """issl_score on host arrays -> (mit, cfd); a column the method skips is None."""
        g = np.ascontiguousarray(guides, dtype=np.uint64)
        m = method_code(method)
        mit = np.full(g.size, np.nan)
        cfd = np.full(g.size, np.nan)
        _check(lib().issl_score(self._h, g.ctypes.data, g.size, int(max_dist), float(threshold), m,
                                mit.ctypes.data, cfd.ctypes.data))
        return (mit if m in (1, 3, 4, 5) else None), (cfd if m in (2, 3, 4, 5) else None)

    def score_into(self, guides: np.ndarray, max_dist: int, threshold: float, method, mit: np.ndarray, cfd: np.ndarray):
        """issl_score with caller-owned (e.g. pinned) host buffers."""
        _check(lib().issl_score(self._h, guides.ctypes.data, guides.size, int(max_dist), float(threshold),
                                method_code(method), mit.ctypes.data, cfd.ctypes.data))

    def score_device(self, d_guides: int, n: int, max_dist: int, threshold: float, method, d_mit: int, d_cfd: int,
                     stream: int = 0):
        """issl_score_device on raw device pointers (e.g. torch tensors' data_ptr()) and a cudaStream_t."""
        _check(lib().issl_score_device(self._h, d_guides, n, int(max_dist), float(threshold), method_code(method),
                                       d_mit, d_cfd, stream))

    def score_hits(self, guides: np.ndarray, max_dist: int, threshold: float, method):
        g = np.ascontiguousarray(guides, dtype=np.uint64)
        m = method_code(method)
        mit = np.full(g.size, np.nan)
        cfd = np.full(g.size, np.nan)
        count = C.c_size_t(0)
        _check(lib().issl_score_hits(self._h, g.ctypes.data, g.size, int(max_dist), float(threshold), m,
                                     mit.ctypes.data, cfd.ctypes.data, None, None, None, None, 0, C.byref(count)))
        cap = max(1, count.value)
        hg = np.zeros(cap, dtype=np.uint64); hi = np.zeros(cap, dtype=np.uint32)
        hd = np.zeros(cap, dtype=np.int32); ho = np.zeros(cap, dtype=np.uint32)
        _check(lib().issl_score_hits(self._h, g.ctypes.data, g.size, int(max_dist), float(threshold), m,
                                     mit.ctypes.data, cfd.ctypes.data, hg.ctypes.data, hi.ctypes.data, hd.ctypes.data,
                                     ho.ctypes.data, cap, C.byref(count)))
        n = count.value
        hits = np.stack([hg[:n].astype(np.int64), hi[:n].astype(np.int64), hd[:n].astype(np.int64),
                         ho[:n].astype(np.int64)], axis=1)
        return mit, cfd, hits

    def read_sites(self, site_ids: np.ndarray) -> np.ndarray:
        ids = np.ascontiguousarray(site_ids, dtype=np.uint64)
        out = np.zeros(ids.size, dtype=np.uint64)
        _check(lib().issl_device_read_sites(self._h, ids.ctypes.data, ids.size, out.ctypes.data))
        return out

    def write_issl(self, path: str):
        _check(lib().issl_device_write_issl(self._h, str(path).encode()))

    def close(self):
        if self._h:
            lib().issl_device_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def replicate(first: Device, cuda_devices) -> list:
    """Replicas of `first` on every device of cuda_devices (which starts with first's own device), fanned out as a
    binary tree of peer copies: 0 -> 1, then 0 -> 2 and 1 -> 3, ... -- log2(n) rounds instead of n - 1 copies out of one GPU."""
    import threading
    devs = [first] + [None] * (len(cuda_devices) - 1)
    have = 1
    while have < len(devs):
        jobs = [(src, have + src) for src in range(have) if have + src < len(devs)]
        errors = []

        def run(src, dst):
            try:
                devs[dst] = devs[src].clone(cuda_devices[dst])
            except Exception as e:     # re-raised below, on the caller's thread
                errors.append(e)
        threads = [threading.Thread(target=run, args=j) for j in jobs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            for d in devs[1:]:
                if d is not None:
                    d.close()
            raise errors[0]
        have += len(jobs)
    return devs


def multi_chunk(n: int, n_devs: int) -> int:
    return int(lib().issl_multi_chunk(n, n_devs))


def score_multi(devs, guides: np.ndarray, max_dist: int, threshold: float, method, mit: np.ndarray, cfd: np.ndarray,
                chunk: int = 0):
    """issl_score_multi: one call over several GPUs (dynamic chunks, one host thread per device).
    Returns (stats summed over devices, guides scored per device)."""
    handles = (C.c_void_p * len(devs))(*[d._h for d in devs])
    st = _Stats()
    per = np.zeros(len(devs), dtype=np.uint64)
    _check(lib().issl_score_multi(handles, len(devs), guides.ctypes.data, guides.size, int(max_dist), float(threshold),
                                  method_code(method), mit.ctypes.data if mit is not None else None,
                                  cfd.ctypes.data if cfd is not None else None, int(chunk), C.byref(st), per.ctypes.data))
    return {n: getattr(st, n) for n, _ in st._fields_}, per


class HostBuffer:
    """issl_host_alloc: pinned, portable host memory viewed as a numpy array."""

    def __init__(self, n: int, dtype):
        self.dtype = np.dtype(dtype)
        self._p = C.c_void_p()
        _check(lib().issl_host_alloc(n * self.dtype.itemsize, C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(n * self.dtype.itemsize, 1),))[:n * self.dtype.itemsize].view(self.dtype)

    def close(self):
        if self._p:
            self.array = None
            lib().issl_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sites:
    """issl_sites: the device-side extractOfftargets."""

    def __init__(self, cuda_device: int = 0):
        self._h = C.c_void_p()
        _check(lib().issl_sites_create(cuda_device, C.byref(self._h)))

    def add_fasta(self, text: bytes, single_input: bool = True):
        _check(lib().issl_sites_add_fasta(self._h, text, len(text), 1 if single_input else 0))

    @property
    def count(self) -> int:
        n, c = C.c_uint64(), C.c_uint64()
        _check(lib().issl_sites_count(self._h, C.byref(n), C.byref(c)))
        return n.value

    @property
    def characters(self) -> int:
        n, c = C.c_uint64(), C.c_uint64()
        _check(lib().issl_sites_count(self._h, C.byref(n), C.byref(c)))
        return c.value

    def write_text(self, path):
        _check(lib().issl_sites_write_text(self._h, str(path).encode()))

    def keys(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.count - first if n is None else n
        out = np.empty(n, dtype=np.uint64)
        _check(lib().issl_sites_read_keys(self._h, first, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if self._h:
            lib().issl_sites_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
