#!/usr/bin/env python3
"""bench.py -- guides scored per second by the ISSL off-target scorer on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OpenMP CPU scorer, same workload
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1: one rank per GPU)

Workload (BASELINE.json: the metric "guides scored/sec (MIT+CFD, <=4 mm) at 1/2/4/8 B200" as configs[2] states it): a
synthetic human-scale index -- 581.25 M uniform NGG sites (3.1 Gbp x 2 strands x 3/32), l = 20, w = 8, built directly
in HBM by issl_device_create_synthetic -- and 10 000 000 guides IN TOTAL (90 % drawn from the index's own sites, 10 %
uniform random 20-mers), scored with method `and` (MIT + CFD), maxDist 4, threshold 0 (no early exit: every candidate
of every guide is visited).  One step = one pass of the scorer over all the guides.  Multi-GPU = STRONG scaling: the
index is replicated, the same 10 M guides are partitioned into N contiguous ranges, one rank per GPU; there is no
collective while scoring (guides are independent); torch.distributed carries the barrier, the max-over-ranks of the
step time and, in the end-to-end leg, the gather of the score ranges to rank 0 (NCCL send/recv over NVLink).
`config1_100k` repeats the measurement on configs[1]'s batch (the first 100 000 guides, one GPU) -- round 1's headline.
`--guides G` gives every GPU its own G guides instead (weak scaling; configs 4 and 5 via tools/run_configs.sh).

`value`  : guides/s with guides and outputs resident in HBM (issl_score_device), CUDA events, max over ranks.
`e2e`    : guides/s from HOST buffers to HOST buffers.  N = 1: issl_score with pinned host arrays (H2D of the guides and
           D2H of both score columns inside).  N > 1: every rank copies its range of guides from pinned host memory,
           scores, the score ranges are gathered to rank 0's GPU and copied to rank 0's pinned host arrays, in input order.
`one_process` (N > 1, rank 0): the PRODUCT's own multi-GPU path, as bin/isslScoreOfftargets runs it -- one process, the
           index built on GPU 0 and replicated to the other GPUs by peer copies over NVLink (issl_device_clone, timed as
           index_fanout_s), then issl_score_multi: chunks of guides handed out dynamically to one host thread per GPU,
           pinned host arrays in and out.  `multi_gpu_cli_parity`: bin/isslScoreOfftargets with ISSL_GPUS=N on a prefix of
           the same guides and the same index written out as a .issl, its stdout compared with the in-process scores
           (all lines) and with the unmodified reference binary's stdout (a strided sample that touches every GPU's chunks).
`roofline`: the dominant kernel, timed by the library's own CUDA events on the launching stream, against the
           measured HBM copy bandwidth in MEASURED_PEAKS.json.  Default layout (triple): the
           bucket scan k_scan_triple_blocked; unit = one (guide, sub-bucket) visit = one aligned read of the
           bucket's block (128 B at human scale), plus the bytes gathered per hit.  `frac_occupied_bytes` counts only
           the 2-byte entries the visited buckets hold (no padding); `dram_frac` uses the DRAM bytes ncu measured for the
           kernel.  `reference_equivalent` restates the same time as SURVEY.md 8d defines it: 4 B (inline-residual
           layout) x list entries the reference's loop would visit.  With --layout res32|sig64|gather: k_scan.
`cpu_baseline`: the unmodified reference binary (oracle/_ref/isslScoreOfftargets, built from
           /root/reference by oracle/Makefile) on the same index written out as a real .issl file and
           a bounded prefix of the same guides, all host cores; scoring time = wall time minus the
           wall time of the same command with an empty guide file (index load).  Baseline only.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

HUMAN_SITES = 581_250_000          # 3.1e9 bp * 2 strands * 3/32 sites per position
TOTAL_GUIDES = 10_000_000          # configs[2]: 10 M guides, partitioned at 1/2/4/8 GPUs
CONFIG1_GUIDES = 100_000           # configs[1]: 100 k guides on one GPU
MAX_DIST = 4
THRESHOLD = 0.0
METHOD = "and"
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=HUMAN_SITES, help="uniform synthetic sites before duplicate collapsing")
    ap.add_argument("--total-guides", type=int, default=TOTAL_GUIDES, help="guides in all, partitioned over the GPUs (strong scaling)")
    ap.add_argument("--guides", type=int, default=None, help="guides PER GPU instead (weak scaling)")
    ap.add_argument("--no-one-process", action="store_true", help="N > 1: skip the one-process multi-GPU leg (issl_score_multi)")
    ap.add_argument("--no-cli-parity", action="store_true", help="N > 1: skip bin/isslScoreOfftargets with ISSL_GPUS=N")
    ap.add_argument("--cli-guides", type=int, default=1 << 21, help="guides handed to the host program in the parity leg")
    ap.add_argument("--layout", default="auto", choices=["auto", "triple", "res32", "sig64", "gather"])
    ap.add_argument("--slice-width", type=int, default=8)
    ap.add_argument("--method", default=METHOD)
    ap.add_argument("--max-dist", type=int, default=MAX_DIST)
    ap.add_argument("--threshold", type=float, default=THRESHOLD)
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--family-size", type=int, default=0)
    ap.add_argument("--family-size-max", type=int, default=0,
                    help="> --family-size: family sizes are log-uniform in [--family-size, this] (config 4 as SURVEY.md 8d states it)")
    ap.add_argument("--low-complexity", type=float, default=0.0,
                    help="fraction of further sites overlapping poly-A/T and dinucleotide tracts (config 4: 0.01)")
    ap.add_argument("--family-guides", type=float, default=0.0,
                    help="fraction of guides drawn from the planted families' roots (0-2 substitutions), config 4")
    ap.add_argument("--cpu-guides", type=int, default=0, help="guides in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-group", type=int, default=32, choices=[1, 2, 4, 8, 32],
                    help="guides sharing one streamed list chunk (1 = pure HBM streaming, one guide per scan item)")
    ap.add_argument("--scratch", default=None, help="directory for the .issl handed to the reference (default /dev/shm)")
    return ap.parse_args()


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, same as mix64() in csrc/issl_kernels.cuh."""
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def family_roots(seed: int, families: int) -> np.ndarray:
    """Roots of the planted near-repeat families, recomputed as k_synth_sites derives them."""
    f = np.arange(families, dtype=np.uint64)
    with np.errstate(over="ignore"):
        def rng3(a):
            return _mix64(_mix64(_mix64(np.full(families, seed, dtype=np.uint64)) ^ np.uint64(a)) ^ (f * np.uint64(0xD6E8FEB86659FD93)))
        root = (rng3(3) & np.uint64((1 << 40) - 1) & ~np.uint64(3)) | (rng3(4) % np.uint64(3))
    return root


def make_guides(dev, n: int, seed: int, families: int = 0, family_frac: float = 0.0, index_seed: int = 1) -> np.ndarray:
    """90 % of the guides are sites of the index itself, 10 % uniform random 20-mers;
    with family_frac > 0 that fraction is drawn from the planted families instead (config 4)."""
    rng = np.random.default_rng(seed)
    n_fam = int(n * family_frac) if families else 0
    n_own = ((n - n_fam) * 9) // 10
    own = dev.read_sites(rng.integers(0, dev.info["offtargetsCount"], n_own).astype(np.uint64))
    rnd = rng.integers(0, 1 << 40, n - n_fam - n_own, dtype=np.uint64)
    parts = [own, rnd]
    if n_fam:
        fam = family_roots(index_seed, families)[rng.integers(0, families, n_fam)].copy()
        for _ in range(2):                                   # up to two substitutions per guide
            pos = rng.integers(0, 20, n_fam).astype(np.uint64) * np.uint64(2)
            sub = rng.integers(0, 4, n_fam).astype(np.uint64)           # 0 = no change
            fam ^= sub << pos
        parts.append(fam)
    g = np.concatenate(parts)
    rng.shuffle(g)
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_candidate(layout_name: str):
    """dram bytes per streamed list entry of k_scan from the committed ncu --set full capture, if any."""
    p = ROOT / "profiles" / "scan_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(layout_name)
        except ValueError:
            return None
    return None


# ---------------------------------------------------------------------------------------------
# the reference on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def reference_exe() -> str | None:
    exe = ROOT / "oracle" / "_ref" / "isslScoreOfftargets"
    return str(exe) if exe.exists() else None


def time_reference(issl_path: str, guides: np.ndarray, max_dist: int, threshold: float, method: str, workdir: str):
    """Wall time of the unmodified reference scorer on `guides`, minus its index-load time
    (same command, empty guide file: the reference loads the index, then fails on the empty file)."""
    import crackling_b200 as cb
    gpath, epath = os.path.join(workdir, "guides.txt"), os.path.join(workdir, "empty.txt")
    write_guide_file(gpath, guides)
    open(epath, "wb").close()
    exe = reference_exe()
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)      # the pipeline never sets a thread count (Crackling.py:767-775)
    args = [str(max_dist), repr(float(threshold)), method]
    t0 = time.perf_counter()
    subprocess.run([exe, issl_path, epath, *args], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    p = subprocess.run([exe, issl_path, gpath, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    t_all = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"reference scorer failed: {p.stderr.decode()[-300:]}")
    return max(t_all - t_load, 1e-9), t_load, p.stdout


def time_port(issl_path: str, guides: np.ndarray, max_dist: int, threshold: float, method: str):
    from oracle import oracle
    img = np.fromfile(issl_path, dtype=np.uint8)
    t0 = time.perf_counter()
    oracle.score(img, guides, max_dist, threshold, method, threads=0)
    return time.perf_counter() - t0


def fmt_lines(guides: np.ndarray, mit: np.ndarray, cfd: np.ndarray, method: str = "and") -> bytes:
    """The reference's output lines (isslScoreOfftargets.cpp:514-527)."""
    import crackling_b200 as cb
    return cb.format_lines(guides, mit, cfd, method)


def write_guide_file(path: str, guides: np.ndarray):
    """20 bases + LF per guide: format_lines with an unknown method prints "SEQ\\t-1\\t-1", the sequence is its first 20 bytes."""
    import crackling_b200 as cb
    lines = np.frombuffer(cb.format_lines(guides, None, None, "none"), dtype=np.uint8).reshape(-1, 27)
    out = np.empty((lines.shape[0], 21), dtype=np.uint8)
    out[:, :20] = lines[:, :20]
    out[:, 20] = 10
    out.tofile(path)


def cpu_baseline(dev, guides: np.ndarray, args, n_sample: int, gpu_mit=None, gpu_cfd=None) -> dict:
    cores = os.cpu_count() or 1
    parity = None
    scratch = args.scratch or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    with tempfile.TemporaryDirectory(dir=scratch) as tmp:
        issl_path = os.path.join(tmp, "index.issl")
        dev.write_issl(issl_path)
        sample = guides[:n_sample]
        if reference_exe():
            t, t_load, ref_stdout = time_reference(issl_path, sample, args.max_dist, args.threshold, args.method, tmp)
            kind = "reference"
            if gpu_mit is not None and args.method in ("and", "or", "avg"):
                ours = fmt_lines(sample, gpu_mit[:n_sample], gpu_cfd[:n_sample], args.method).splitlines()
                theirs = ref_stdout.splitlines()
                same = sum(a == b for a, b in zip(ours, theirs))
                parity = f"{same}/{len(theirs)} output lines byte-identical to the reference's stdout at full index size"
            note = (f"oracle/_ref/isslScoreOfftargets (unmodified reference, g++ -O3 -fopenmp -mpopcnt) on the same index "
                    f"written as a {os.path.getsize(issl_path) / 1e9:.1f} GB .issl, first {n_sample} guides of the batch, "
                    f"OpenMP default threads = {cores} cores ({cpu_model()}); scoring {t:.1f} s = wall minus {t_load:.1f} s index load")
        else:
            t = time_port(issl_path, sample, args.max_dist, args.threshold, args.method)
            kind = "port"
            note = f"oracle C port (oracle/issl_oracle.c), first {n_sample} guides, {cores} OpenMP threads, {t:.1f} s"
    out = {"value": n_sample / t, "unit": "guides/s", "cores": cores, "kind": kind, "sample": note}
    if parity:
        out["parity"] = parity
    return out


# ---------------------------------------------------------------------------------------------
def roofline_of(args, info, layout_name, acc, ms_total, clk, n_steps_guides):
    """The dominant kernel against the HBM roofline, from the library's counters summed over the timed steps."""
    peak, peak_src = hbm_peak()
    bpc = info["bytes_per_candidate"]
    scan_launches = max(acc["scan_launches"], 1)
    per_launch_ms = acc["scan_ms"] / scan_launches
    sm_hz = (clk["sm_mhz"] or 1965.0) * 1e6
    candidates, hits, streamed, bucket_visits = acc["candidates"], acc["hits"], acc["streamed"], acc["bucket_visits"]
    if layout_name == "triple" and bucket_visits > 0:
        # unit = one (guide, sub-bucket) visit.  Blocked copy: one aligned read of the bucket's block; otherwise an
        # offset pair (8 B) + the bucket's residuals (2 B each).  Every hit adds what the tail gathers for it
        # (hit_bytes: 28 when offsets and ids are looked up, 0 when the site itself orders the hits).  
        blk = info["triple_block_bytes"]
        hit_bytes = info.get("triple_hit_bytes", 28)
        alg = (bucket_visits * blk if blk else bucket_visits * 8 + streamed * 2) + hits * hit_bytes
        occupied = streamed * 2 + hits * hit_bytes
        secs = acc["scan_ms"] / 1e3
        achieved = alg / secs / 1e9
        tpv = ncu_traffic_per_candidate("triple_per_visit")
        ref_equiv = 4 * candidates / secs / 1e9
        return {"bound": "hbm", "kernel": info.get("scan_kernel", "k_scan_triple_blocked" if blk else "k_scan_triple"), "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "frac_occupied_bytes": occupied / secs / 1e9 / peak,
                "dram_frac": (tpv * bucket_visits / secs / 1e9 / peak) if tpv else None,
                "unit_of_work": "bucket visit", "algorithmic_bytes_per_visit": blk if blk else None,
                "algorithmic_bytes_per_hit": hit_bytes, "visits_per_launch": bucket_visits / scan_launches,
                "visits_per_guide": bucket_visits / max(n_steps_guides, 1),
                "bucket_entries_per_guide": streamed / max(n_steps_guides, 1),
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": acc["scan_ms"] / ms_total,
                "traffic": (tpv * bucket_visits / scan_launches) if tpv else None,
                "traffic_source": ("profiles/scan_traffic.json: ncu --set full dram read+write bytes of one scan "
                                   "launch / its bucket visits, x visits per launch here") if tpv else None,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "reference_equivalent": {"GB/s": ref_equiv, "x_hbm_peak": ref_equiv / peak,
                                         "definition": "4 B x list entries the reference's loop visits (SURVEY.md 8d, RES32 "
                                                       "layout) over the same kernel time",
                                         "candidates_per_launch": candidates / scan_launches,
                                         "entries_read_per_candidate": streamed / max(candidates, 1)}}
    per_launch_bytes = bpc * candidates / scan_launches
    achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
    tpc = ncu_traffic_per_candidate(layout_name)
    roofline = {"bound": "hbm", "kernel": f"k_scan<{layout_name}>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_candidate": bpc, "candidates_per_launch": candidates / scan_launches,
                "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": acc["scan_ms"] / ms_total,
                "traffic": (tpc * streamed / scan_launches) if tpc else None,
                "traffic_source": ("profiles/scan_traffic.json: ncu --set full dram bytes per list entry streamed x entries "
                                   "streamed per launch (each chunk is read once per guide group)") if tpc else None,
                "frac_of_nominal_8TBps": achieved / 8000.0}
    # With list reuse (max_group > 1) a chunk read from HBM once serves up to 32 guides, so the figure above --
    # algorithmic bytes of the reference's per-guide walk over time -- legitimately exceeds the HBM peak
    # (SURVEY.md 8d).  What then bounds the kernel is an integer pipe.  Both views are reported.
    pairs_per_s = candidates / scan_launches / (per_launch_ms / 1e3)
    streamed_gbs = bpc * streamed / scan_launches / (per_launch_ms / 1e3) / 1e9
    if args.max_group == 32 and layout_name in ("res32", "triple") and args.max_dist <= 7:
        # bit-sliced path: 40 ALU-pipe instructions (16 nibble extractions + 24 adder/threshold LOP3) per 32 pairs,
        # ALU pipe = 64 lanes/clk/SM  ->  51.2 pairs/clk/SM
        pipe, per_clk = "alu (bit-sliced: 1.25 ALU-pipe instr per pair at 64 lanes/clk/SM x 148 SMs)", 64 / 1.25
    else:
        pipe, per_clk = "xu (one POPC per pair at 16 lanes/clk/SM x 148 SMs)", 16.0
    roofline.update({"reuse": candidates / max(streamed, 1), "streamed_gbs": streamed_gbs,
                     "streamed_frac_of_hbm_peak": streamed_gbs / peak,
                     "pipe_bound": {"pipe": pipe, "achieved_pairs_per_s": pairs_per_s,
                                    "peak_pairs_per_s": 148 * per_clk * sm_hz, "frac": pairs_per_s / (148 * per_clk * sm_hz),
                                    "sm_mhz_used": sm_hz / 1e6}})
    return roofline


def wait_for_free_hbm(devices, need_bytes: int, timeout_s: float = 120.0) -> bool:
    """The other ranks have left: wait until their GPUs have room for a replica of the index."""
    import torch
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < timeout_s:
        if all(torch.cuda.mem_get_info(d)[0] >= need_bytes for d in devices):
            return True
        time.sleep(0.25)
    return False


def one_process_leg(args, dev, guides: np.ndarray, world: int, want_mit, want_cfd) -> dict:
    """The product's own multi-GPU path (what bin/isslScoreOfftargets does): one process, replicas by peer copy,
    issl_score_multi with pinned host arrays.  Runs on rank 0 once the other ranks have released their GPUs."""
    import torch
    import crackling_b200 as cb
    out = {}
    need = int(dev.info["hbm_bytes"] * 1.02) + (2 << 30)
    if not wait_for_free_hbm(range(1, world), need):
        return {"skipped": "the other ranks' GPUs did not free up in time"}
    t0 = time.perf_counter()
    devs = cb.replicate(dev, list(range(world)))
    out["index_fanout_s"] = round(time.perf_counter() - t0, 3)
    out["index_fanout"] = (f"{dev.info['hbm_bytes'] / 1e9:.1f} GB per replica, binary tree of issl_device_clone peer copies "
                           f"(GPU 0 -> {world - 1} more GPUs), {dev.info['hbm_bytes'] * (world - 1) / 1e9 / max(out['index_fanout_s'], 1e-9):.0f} GB/s in aggregate")
    n = guides.size
    hg, hm, hc = cb.HostBuffer(n, np.uint64), cb.HostBuffer(n, np.float64), cb.HostBuffer(n, np.float64)
    hg.array[:] = guides
    for _ in range(max(args.warmup, 1)):
        cb.score_multi(devs, hg.array, args.max_dist, args.threshold, args.method, hm.array, hc.array)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, per = cb.score_multi(devs, hg.array, args.max_dist, args.threshold, args.method, hm.array, hc.array)
    dt = time.perf_counter() - t0
    out.update({"value": n * args.steps / dt, "unit": "guides/s", "ms_per_step": dt / args.steps * 1e3,
                "what": "issl_score_multi, host to host (pinned): dynamic chunks, one host thread per GPU, one process",
                "chunk_guides": cb.multi_chunk(n, world), "guides_per_device_last_step": [int(x) for x in per],
                "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(n * 16)})
    if want_mit is not None:
        same = bool(np.array_equal(hm.array.view(np.uint64), want_mit.view(np.uint64)) and
                    np.array_equal(hc.array.view(np.uint64), want_cfd.view(np.uint64)))
        out["parity"] = ("bit-identical to the rank-per-GPU scores gathered on rank 0" if same
                         else "DIFFERS from the rank-per-GPU scores")
    mit_all, cfd_all = hm.array.copy(), hc.array.copy()
    for d in devs[1:]:
        d.close()
    for b in (hg, hm, hc):
        b.close()
    torch.cuda.synchronize()
    return out, mit_all, cfd_all


def cli_parity_leg(args, dev, guides: np.ndarray, world: int, mit: np.ndarray, cfd: np.ndarray) -> str:
    """bin/isslScoreOfftargets with ISSL_GPUS=N (untimed): every line against the in-process scores, and a strided
    sample (it touches every GPU's chunks) against the unmodified reference binary's stdout."""
    import crackling_b200 as cb
    n = min(args.cli_guides, guides.size)
    scratch = args.scratch or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    with tempfile.TemporaryDirectory(dir=scratch) as tmp:
        issl_path, gpath = os.path.join(tmp, "index.issl"), os.path.join(tmp, "guides.txt")
        dev.write_issl(issl_path)
        dev.close()                      # the host program needs the HBM
        write_guide_file(gpath, guides[:n])
        env = dict(os.environ, ISSL_GPUS=str(world), ISSL_TIMING="1")
        cli_args = [str(args.max_dist), repr(float(args.threshold)), args.method]
        p = subprocess.run([str(cb.cli_path()), issl_path, gpath, *cli_args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        if p.returncode != 0:
            return f"host program failed: {p.stderr.decode()[-300:]}"
        timing = [l for l in p.stderr.decode().splitlines() if l.startswith("[issl]")]
        got = p.stdout.splitlines()
        want = fmt_lines(guides[:n], mit[:n], cfd[:n], args.method).splitlines()
        same = sum(a == b for a, b in zip(got, want)) if len(got) == len(want) else 0
        msg = f"{same}/{len(want)} lines of bin/isslScoreOfftargets (ISSL_GPUS={world}) identical to the in-process scores"
        if reference_exe():
            k = min(n, 100 * (os.cpu_count() or 1), 1600)
            idx = (np.arange(k, dtype=np.int64) * (n // k))
            _, _, ref_stdout = time_reference(issl_path, guides[idx], args.max_dist, args.threshold, args.method, tmp)
            theirs = ref_stdout.splitlines()
            ok = sum(got[i] == t for i, t in zip(idx, theirs)) if len(got) == n else 0
            msg += f"; {ok}/{len(theirs)} lines identical to the unmodified reference's stdout (every {n // k}th guide)"
        if timing:
            msg += " | " + " | ".join(timing[-2:])
    return msg


def main() -> int:
    args = parse_args()
    # Only the JSON line may reach stdout: libraries (NCCL's version banner, ...) are sent to stderr.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference" and rank != 0:
        return 0            # rank 0 alone runs the CPU reference

    import torch
    import crackling_b200 as cb
    from crackling_b200.sharding import max_over_ranks, shard_bounds

    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: this benchmark has no CPU path"})
        return 1
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    os.environ["ISSL_MAX_GROUP"] = str(args.max_group)
    t_build = time.perf_counter()
    if args.family_size_max > args.family_size or args.low_complexity > 0:
        dev = cb.Device.synthetic_ex(local_rank, args.layout, seed=1, uniform_sites=args.sites, families=args.families,
                                     family_size_min=args.family_size, family_size_max=max(args.family_size_max, args.family_size),
                                     max_sub_rate=0.15, low_complexity_fraction=args.low_complexity, seq_length=20,
                                     slice_width=args.slice_width)
    else:
        dev = cb.Device.synthetic(local_rank, args.layout, seed=1, uniform_sites=args.sites, families=args.families,
                                  family_size=args.family_size, max_sub_rate=0.15, seq_length=20, slice_width=args.slice_width)
    t_build = time.perf_counter() - t_build
    info = dev.info
    layout_name = {1: "res32", 2: "sig64", 3: "gather", 4: "triple"}[info["layout"]]
    strong = args.guides is None
    # strong scaling: every rank draws the same guides and scores its own contiguous range of them
    total = args.total_guides if strong else args.guides * world
    if strong:
        all_guides = make_guides(dev, total, seed=3, families=args.families, family_frac=args.family_guides)
        lo, hi = shard_bounds(total, world, rank)
        guides = all_guides[lo:hi]
    else:
        all_guides = None
        guides = make_guides(dev, args.guides, seed=2 + rank, families=args.families, family_frac=args.family_guides)
        lo, hi = rank * args.guides, (rank + 1) * args.guides
    repeats = ""
    if args.families:
        sizes = (f"{args.family_size}-{args.family_size_max} copies (log-uniform)" if args.family_size_max > args.family_size
                 else f"{args.family_size} copies")
        repeats = f" + {args.families} near-repeat families of {sizes}, per-base substitution rate 0-15 %"
    if args.low_complexity > 0:
        repeats += f" + {args.low_complexity:.1%} low-complexity (poly-A/T, dinucleotide) tract sites"
    workload = (f"synthetic {'human' if args.sites == HUMAN_SITES else 'genome'}-scale index: {args.sites} uniform NGG sites{repeats} -> {info['offtargetsCount']} distinct, "
                f"l=20 w={args.slice_width}, " +
                (f"{total} guides in all partitioned over the GPUs (BASELINE.json configs[2])" if strong
                 else f"{args.guides} guides per GPU") +
                f" (90% index sites, 10% random), method {args.method}, maxDist {args.max_dist}, threshold {args.threshold:g}")
    config = {"workload": workload, "sites": info["offtargetsCount"], "global_guides": total,
              "method": args.method, "max_dist": args.max_dist,
              "threshold": args.threshold, "slice_width": args.slice_width, "hbm_layout": layout_name,
              "families": args.families, "family_size": args.family_size, "family_size_max": args.family_size_max,
              "low_complexity": args.low_complexity, "family_guides": args.family_guides,
              "max_group": args.max_group,
              "parallelism": f"replicated index, guides partitioned x{args.gpus}, no collective while scoring",
              "l2": ("inputs larger than L2 (126 MB): every step reads its sub-buckets (~180 KB per guide, random 128-byte blocks "
                     "out of a 21 GB copy; ~45 MB of slice lists per guide with the list-scan layouts)")}
    # measured, not configured (the arms build different copies of the same index): beside config, like index_build_s
    index_hbm_gb = round(info["hbm_bytes"] / 1e9, 2)
    scaling = "strong" if strong else "weak"

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        cores = os.cpu_count() or 1
        n_sample = args.cpu_guides or min(guides.size, 100 * cores)
        n_total = n_sample * (args.steps + args.warmup)
        reps = -(-n_total // guides.size)
        pool = np.tile(guides, reps)[:n_total]
        scratch = args.scratch or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
        with tempfile.TemporaryDirectory(dir=scratch) as tmp:
            issl_path = os.path.join(tmp, "index.issl")
            dev.write_issl(issl_path)
            dev.close()
            if reference_exe():
                t, t_load, _ = time_reference(issl_path, pool, args.max_dist, args.threshold, args.method, tmp)
                kind = "reference"
            else:
                t, t_load = time_port(issl_path, pool, args.max_dist, args.threshold, args.method), 0.0
                kind = "port"
        per_step = t / (args.steps + args.warmup)
        v = n_sample / per_step
        note = (f"{kind}: one process over {args.steps + args.warmup} x {n_sample} guides ({args.warmup} warm-up + {args.steps} "
                f"timed steps' worth; the reference reloads its index per process, so steps share one invocation), "
                f"{cores} OpenMP threads on {cpu_model()} (the reference's sample config assumes 128, config.ini:104; the pipeline "
                f"never passes a thread count), scoring {t:.1f} s after subtracting {t_load:.1f} s index load")
        emit(({"impl": "reference", "metric": "guides scored/sec (MIT+CFD, <=4 mm)", "value": v, "unit": "guides/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
               "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u64 popcount + f64 scores",
               "data": "synthetic", "config": config, "index_build_s": round(t_build, 2), "index_hbm_gb": index_hbm_gb,
               "host": {"cores": cores, "cpu": cpu_model()},
               "cpu_baseline": {"value": v, "unit": "guides/s", "cores": cores, "kind": kind, "sample": note},
               "e2e": {"value": v, "unit": "guides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ our arm
    n = guides.size
    stream = torch.cuda.current_stream()
    keys = ("scan_ms", "scan_launches", "launches", "candidates", "hits", "streamed", "bucket_visits", "early_exits", "heavy_hits", "heavy_ms")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure_device(g: np.ndarray, steps: int, warmup: int, sample_clocks: bool):
        """K steps with guides and outputs resident in HBM; CUDA events on torch's current stream (the one the library launches on)."""
        m = g.size
        d_g = torch.from_numpy(g.view(np.int64)).cuda()
        d_m = torch.zeros(m, dtype=torch.float64, device="cuda")
        d_c = torch.zeros(m, dtype=torch.float64, device="cuda")

        def step():
            dev.score_device(d_g.data_ptr(), m, args.max_dist, args.threshold, args.method, d_m.data_ptr(), d_c.data_ptr(), stream.cuda_stream)
        for _ in range(warmup):
            step()
        acc = dict.fromkeys(keys, 0)
        barrier()
        clock_window = "timed region"
        with ClockSampler(local_rank) as clocks:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                step()
                st = dev.stats
                for k in keys:
                    acc[k] += st[k]
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            # nvidia-smi cannot sample faster than every ~100 ms and the timed region may be shorter than that: keep the
            # same step running (untimed) under the sampler until it has had 0.6 s of this load to look at
            t_obs = time.perf_counter()
            while sample_clocks and ms < 600.0 and time.perf_counter() - t_obs < 0.6:
                step()
                clock_window = "timed region + 0.6 s of the same step, untimed (the region is shorter than the sampler's period)"
        return ms, acc, dict(clocks.summary(), window=clock_window), d_m.cpu().numpy(), d_c.cpu().numpy()

    ms_total, acc, clk, dev_mit, dev_cfd = measure_device(guides, args.steps, args.warmup, True)
    ms_per_step = max_over_ranks(ms_total, dist, "cuda") / args.steps
    value = total / (ms_per_step / 1e3)

    # ---- end to end: host buffers in, host buffers out
    h_guides = torch.from_numpy(guides.view(np.int64)).pin_memory()
    if world == 1:
        h_mit = torch.zeros(n, dtype=torch.float64).pin_memory()
        h_cfd = torch.zeros(n, dtype=torch.float64).pin_memory()
        hg, hm, hc = h_guides.numpy().view(np.uint64), h_mit.numpy(), h_cfd.numpy()
        dev.score_into(hg, args.max_dist, args.threshold, args.method, hm, hc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dev.score_into(hg, args.max_dist, args.threshold, args.method, hm, hc)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        full_mit, full_cfd = hm, hc
        e2e_how = "issl_score: pinned host guides -> pinned host scores"
    else:
        # every rank: H2D of its range, scoring, its two score ranges sent to rank 0's GPU (NCCL over NVLink, the only
        # exchange the path has); rank 0: D2H of all scores into pinned host arrays, in input order
        d_g = torch.empty(n, dtype=torch.int64, device="cuda")
        d_out = torch.zeros(2, total if rank == 0 else n, dtype=torch.float64, device="cuda")
        h_out = torch.zeros(2, total, dtype=torch.float64).pin_memory() if rank == 0 else None
        mine = d_out[:, lo:hi] if rank == 0 else d_out
        d_m, d_c = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")

        def e2e_step():
            d_g.copy_(h_guides, non_blocking=True)
            dev.score_device(d_g.data_ptr(), n, args.max_dist, args.threshold, args.method, d_m.data_ptr(), d_c.data_ptr(), stream.cuda_stream)
            mine[0].copy_(d_m); mine[1].copy_(d_c)
            if rank == 0:
                ops = []
                recv = []
                for r in range(1, world):
                    rlo, rhi = shard_bounds(total, world, r) if strong else (r * args.guides, (r + 1) * args.guides)
                    buf = torch.empty(2, rhi - rlo, dtype=torch.float64, device="cuda")
                    recv.append((rlo, rhi, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, r))
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                for rlo, rhi, buf in recv:
                    d_out[:, rlo:rhi].copy_(buf)
                h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
            else:
                for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), 0)]):
                    w.wait()
                torch.cuda.synchronize()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        full_mit, full_cfd = (h_out[0].numpy(), h_out[1].numpy()) if rank == 0 else (None, None)
        hm, hc = (full_mit[lo:hi], full_cfd[lo:hi]) if rank == 0 else (d_m.cpu().numpy(), d_c.cpu().numpy())
        e2e_how = ("per rank: pinned host guides -> GPU, issl_score_device; score ranges sent to rank 0's GPU (NCCL send/recv), "
                   "rank 0: D2H into pinned host arrays in input order")
    e2e_value = total * args.steps / max_over_ranks(e2e_s, dist, "cuda")
    assert np.array_equal(hm, dev_mit) and np.array_equal(hc, dev_cfd), "host and device paths disagree"

    if rank != 0:
        dev.close()
        del dev
        torch.cuda.empty_cache()
        dist.barrier()
        dist.destroy_process_group()
        return 0

    roofline = roofline_of(args, info, layout_name, acc, ms_total, clk, args.steps * n)
    triple_scan = layout_name == "triple" and acc["bucket_visits"] > 0
    result = {"metric": "guides scored/sec (MIT+CFD, <=4 mm)", "value": value, "unit": "guides/s", "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
              "scaling": scaling, "vs_baseline": None,
              "dtype": "u16 bit-sliced compare (LOP3), f64 scores" if triple_scan else "u32/u64 xor+popcount, f64 scores", "data": "synthetic",
              "config": config, "index_build_s": round(t_build, 2), "index_hbm_gb": index_hbm_gb, "clocks": clk,
              "e2e": {"value": e2e_value, "unit": "guides/s", "h2d_bytes_per_step": int(total * 8), "d2h_bytes_per_step": int(total * 16),
                      "how": e2e_how},
              "gpu_launches": int(acc["launches"]), "scan_launches": int(acc["scan_launches"]),
              "ms_per_100k_guides": ms_per_step * 1e5 / total * world,
              "hits_per_guide": acc["hits"] / max(args.steps * n, 1), "candidates_per_guide": acc["candidates"] / max(args.steps * n, 1),
              "early_exit_fraction": acc["early_exits"] / max(args.steps * n, 1),
              "roofline": roofline}

    if args.families or args.low_complexity > 0:
        # the slice-list lengths the repeats produce (the reference streams whole lists)
        ll = dev.list_lengths.astype(np.float64)
        ll = ll[ll > 0] if (ll > 0).any() else ll
        edges = [0, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 26, 1 << 40]
        hist, _ = np.histogram(ll, bins=edges)
        result["list_lengths"] = {"lists": int(ll.size), "mean": float(ll.mean()), "max": float(ll.max()), "max_over_mean": float(ll.max() / ll.mean()),
                                  "p50": float(np.percentile(ll, 50)), "p99": float(np.percentile(ll, 99)),
                                  "histogram": {f"<{e}": int(c) for e, c in zip(edges[1:], hist)}}
        result["heavy_hits_per_guide"] = acc.get("heavy_hits", 0) / max(args.steps * n, 1)

    if world == 1 and strong and n > CONFIG1_GUIDES:
        # configs[1]: the first 100 000 guides alone, one GPU (round 1's headline workload)
        g1 = guides[:CONFIG1_GUIDES]
        ms1, acc1, _, m1, c1 = measure_device(g1, args.steps, args.warmup, False)
        assert np.array_equal(m1, dev_mit[:CONFIG1_GUIDES]) and np.array_equal(c1, dev_cfd[:CONFIG1_GUIDES])
        hg1 = h_guides.numpy().view(np.uint64)[:CONFIG1_GUIDES]
        dev.score_into(hg1, args.max_dist, args.threshold, args.method, hm[:CONFIG1_GUIDES], hc[:CONFIG1_GUIDES])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dev.score_into(hg1, args.max_dist, args.threshold, args.method, hm[:CONFIG1_GUIDES], hc[:CONFIG1_GUIDES])
        t1 = (time.perf_counter() - t0) / args.steps
        r1 = roofline_of(args, info, layout_name, acc1, ms1, clk, args.steps * CONFIG1_GUIDES)
        result["config1_100k"] = {"workload": "BASELINE.json configs[1]: the first 100 000 of the same guides, one call",
                                  "value": CONFIG1_GUIDES / (ms1 / args.steps / 1e3), "unit": "guides/s", "ms_per_step": ms1 / args.steps,
                                  "e2e": CONFIG1_GUIDES / t1, "roofline_frac": r1["frac"], "kernel_ms_per_launch": r1["kernel_ms_per_launch"]}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if strong and not args.no_one_process:
            try:
                leg = one_process_leg(args, dev, all_guides, world, full_mit, full_cfd)
                if isinstance(leg, tuple):
                    result["one_process"], full_mit, full_cfd = leg
                else:
                    result["one_process"] = leg
            except Exception as e:
                result["one_process"] = {"failed": str(e)[-300:]}
        if strong and not args.no_cli_parity:
            try:
                result["multi_gpu_cli_parity"] = cli_parity_leg(args, dev, all_guides, world, full_mit, full_cfd)
            except Exception as e:
                result["multi_gpu_cli_parity"] = f"failed: {str(e)[-300:]}"

    if world == 1 and not args.no_cpu_baseline:
        try:
            n_sample = args.cpu_guides or min(n, 100 * (os.cpu_count() or 1))
            result["cpu_baseline"] = cpu_baseline(dev, guides, args, n_sample, hm, hc)
        except Exception as e:   # the baseline must not take the GPU number down with it
            result["cpu_baseline"] = {"value": None, "unit": "guides/s", "cores": os.cpu_count(), "kind": "reference",
                                      "sample": f"failed: {e}"}
    emit(result)
    dev.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
