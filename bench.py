#!/usr/bin/env python3
"""bench.py -- guides scored per second by the ISSL off-target scorer on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's OpenMP CPU scorer, same workload
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1: one rank per GPU)

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): a synthetic
human-scale index -- 581.25 M uniform NGG sites (3.1 Gbp x 2 strands x 3/32), l = 20, w = 8,
built directly in HBM by issl_device_create_synthetic -- and 100 000 guides per GPU (90 % drawn
from the index's own sites, 10 % uniform random 20-mers), scored with method `and` (MIT + CFD),
maxDist 4, threshold 0 (no early exit: every candidate of every guide is visited).  One step = one
pass of the scorer over the whole guide batch.  Multi-GPU: the index is replicated, every rank
scores its own 100 000 guides, there is no data-path collective (weak scaling); torch.distributed
is used only for the barrier and the max-over-ranks of the step time.

`value`  : guides/s with guides and outputs resident in HBM (issl_score_device), CUDA events.
`e2e`    : guides/s through issl_score with pinned HOST buffers (H2D of guides and D2H of both score
           columns inside the timed region), wall clock around the blocking calls.
`roofline`: the dominant kernel, timed by the library's own CUDA events on the launching stream, against the
           measured HBM copy bandwidth in MEASURED_PEAKS.json.  Default layout (triple): the
           bucket scan k_scan_triple_blocked; unit = one (guide, sub-bucket) visit = one aligned read of the
           bucket's block (128 B at human scale), plus 28 B per hit (offset pair, id, 16-byte record).
           `reference_equivalent` restates the same time as SURVEY.md 8d defines it: 4 B (inline-residual
           layout) x list entries the reference's loop would visit -- far above the HBM peak, because the
           sub-bucket index lets a guide skip ~99.6 % of its five slice lists.  With --layout res32|sig64|gather:
           k_scan, bytes/candidate of the layout x list entries visited.
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
GUIDES_PER_GPU = 100_000
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
    ap.add_argument("--guides", type=int, default=GUIDES_PER_GPU, help="guides per GPU")
    ap.add_argument("--layout", default="auto", choices=["auto", "triple", "res32", "sig64", "gather"])
    ap.add_argument("--slice-width", type=int, default=8)
    ap.add_argument("--method", default=METHOD)
    ap.add_argument("--max-dist", type=int, default=MAX_DIST)
    ap.add_argument("--threshold", type=float, default=THRESHOLD)
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--family-size", type=int, default=0)
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
def reference_exe() -> str | None:
    exe = ROOT / "oracle" / "_ref" / "isslScoreOfftargets"
    return str(exe) if exe.exists() else None


def time_reference(issl_path: str, guides: np.ndarray, max_dist: int, threshold: float, method: str, workdir: str):
    """Wall time of the unmodified reference scorer on `guides`, minus its index-load time
    (same command, empty guide file: the reference loads the index, then fails on the empty file)."""
    import crackling_b200 as cb
    gpath, epath = os.path.join(workdir, "guides.txt"), os.path.join(workdir, "empty.txt")
    with open(gpath, "wb") as f:
        f.write(b"".join(cb.unpack_guide(int(s)).encode() + b"\n" for s in guides))
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


def fmt_lines(guides: np.ndarray, mit: np.ndarray, cfd: np.ndarray) -> bytes:
    """The reference's output lines (isslScoreOfftargets.cpp:514-527) for method and/or/avg."""
    import crackling_b200 as cb
    return b"".join(b"%s\t%s\t%s\n" % (cb.unpack_guide(int(s)).encode(), b"%f" % m, b"%f" % c)
                    for s, m, c in zip(guides, mit, cfd))


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
                ours = fmt_lines(sample, gpu_mit[:n_sample], gpu_cfd[:n_sample]).splitlines()
                theirs = ref_stdout.splitlines()
                same = sum(a == b for a, b in zip(ours, theirs))
                parity = f"{same}/{len(theirs)} output lines byte-identical to the reference's stdout at full index size"
            note = (f"oracle/_ref/isslScoreOfftargets (unmodified reference, g++ -O3 -fopenmp -mpopcnt) on the same index "
                    f"written as a {os.path.getsize(issl_path) / 1e9:.1f} GB .issl, first {n_sample} guides of the batch, "
                    f"OpenMP default threads = {cores} cores; scoring {t:.1f} s = wall minus {t_load:.1f} s index load")
        else:
            t = time_port(issl_path, sample, args.max_dist, args.threshold, args.method)
            kind = "port"
            note = f"oracle C port (oracle/issl_oracle.c), first {n_sample} guides, {cores} OpenMP threads, {t:.1f} s"
    out = {"value": n_sample / t, "unit": "guides/s", "cores": cores, "kind": kind, "sample": note}
    if parity:
        out["parity"] = parity
    return out


# ---------------------------------------------------------------------------------------------
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
    dev = cb.Device.synthetic(local_rank, args.layout, seed=1, uniform_sites=args.sites, families=args.families,
                              family_size=args.family_size, max_sub_rate=0.15, seq_length=20, slice_width=args.slice_width)
    t_build = time.perf_counter() - t_build
    info = dev.info
    layout_name = {1: "res32", 2: "sig64", 3: "gather", 4: "triple"}[info["layout"]]
    guides = make_guides(dev, args.guides, seed=2 + rank, families=args.families, family_frac=args.family_guides)
    workload = (f"synthetic human-scale index: {args.sites} uniform NGG sites -> {info['offtargetsCount']} distinct, "
                f"l=20 w={args.slice_width}, {args.guides} guides/GPU (90% index sites, 10% random), method {args.method}, "
                f"maxDist {args.max_dist}, threshold {args.threshold:g}")
    config = {"workload": workload, "sites": info["offtargetsCount"], "guides_per_gpu": args.guides,
              "global_guides": args.guides * world, "method": args.method, "max_dist": args.max_dist,
              "threshold": args.threshold, "slice_width": args.slice_width, "hbm_layout": layout_name,
              "families": args.families, "family_size": args.family_size, "family_guides": args.family_guides,
              "max_group": args.max_group,
              "index_hbm_gb": round(info["hbm_bytes"] / 1e9, 2), "index_build_s": round(t_build, 2),
              "parallelism": f"replicated index, guides partitioned x{world}, no collective",
              "l2": ("inputs larger than L2 (126 MB): every step reads its sub-buckets (~180 KB per guide, random 128-byte blocks "
                     "out of a 21 GB copy; ~45 MB of slice lists per guide with the list-scan layouts)")}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        n_sample = args.cpu_guides or min(args.guides, 100 * (os.cpu_count() or 1))
        total = n_sample * (args.steps + args.warmup)
        reps = -(-total // guides.size)
        pool = np.tile(guides, reps)[:total]
        scratch = args.scratch or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
        cores = os.cpu_count() or 1
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
                f"{cores} OpenMP threads, scoring {t:.1f} s after subtracting {t_load:.1f} s index load")
        emit(({"impl": "reference", "metric": "guides scored/sec (MIT+CFD, <=4 mm)", "value": v, "unit": "guides/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64 scores",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "guides/s", "cores": cores, "kind": kind, "sample": note},
                          "e2e": {"value": v, "unit": "guides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ our arm
    n = guides.size
    d_guides = torch.from_numpy(guides.view(np.int64)).cuda()
    d_mit = torch.zeros(n, dtype=torch.float64, device="cuda")
    d_cfd = torch.zeros(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step_device():
        dev.score_device(d_guides.data_ptr(), n, args.max_dist, args.threshold, args.method, d_mit.data_ptr(),
                         d_cfd.data_ptr(), stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    scan_ms, scan_launches, launches, candidates, hits, streamed, bucket_visits, early_exits = 0.0, 0, 0, 0, 0, 0, 0, 0
    barrier()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
            st = dev.stats
            scan_ms += st["scan_ms"]; scan_launches += st["scan_launches"]; launches += st["launches"]
            candidates += st["candidates"]; hits += st["hits"]; streamed += st["streamed"]
            bucket_visits += st["bucket_visits"]; early_exits += st["early_exits"]
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        # nvidia-smi cannot sample faster than every ~100 ms and the timed region may be shorter than that: keep the
        # same step running (untimed) under the sampler until it has had 0.6 s of this load to look at
        t_obs = time.perf_counter()
        clock_window = "timed region"
        while ms_total < 600.0 and time.perf_counter() - t_obs < 0.6:
            step_device()
            clock_window = "timed region + 0.6 s of the same step, untimed (the region is shorter than the sampler's period)"
    from crackling_b200.sharding import max_over_ranks
    ms_total_max = max_over_ranks(ms_total, dist, "cuda")
    ms_per_step = ms_total_max / args.steps
    value = args.guides * world / (ms_per_step / 1e3)

    # end to end through the host-buffer entry point (pinned host memory)
    h_guides = torch.from_numpy(guides.view(np.int64)).pin_memory()
    h_mit = torch.zeros(n, dtype=torch.float64).pin_memory()
    h_cfd = torch.zeros(n, dtype=torch.float64).pin_memory()
    hg, hm, hc = (t.numpy() for t in (h_guides, h_mit, h_cfd))
    hg = hg.view(np.uint64)
    dev.score_into(hg, args.max_dist, args.threshold, args.method, hm, hc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev.score_into(hg, args.max_dist, args.threshold, args.method, hm, hc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    e2e_value = args.guides * world * args.steps / max_over_ranks(e2e_s, dist, "cuda")
    assert np.array_equal(hm, d_mit.cpu().numpy()) and np.array_equal(hc, d_cfd.cpu().numpy()), "host and device paths disagree"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    bpc = info["bytes_per_candidate"]
    per_launch_ms = scan_ms / max(scan_launches, 1)
    clk = clocks.summary()
    sm_hz = (clk["sm_mhz"] or 1965.0) * 1e6
    triple_scan = layout_name == "triple" and bucket_visits > 0
    if triple_scan:
        # unit = one (guide, sub-bucket) visit.  Blocked copy: one aligned read of the bucket's block; otherwise an
        # offset pair (8 B) + the bucket's residuals (2 B each).  Every hit adds an offset pair, an id and a 16-byte
        # record.  
        blk = info["triple_block_bytes"]
        alg = (bucket_visits * blk if blk else bucket_visits * 8 + streamed * 2) + hits * 28
        per_launch_bytes = alg / max(scan_launches, 1)
        achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
        tpv = ncu_traffic_per_candidate("triple_per_visit")
        ref_equiv = 4 * candidates / max(scan_launches, 1) / (per_launch_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_scan_triple_blocked" if blk else "k_scan_triple", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "unit_of_work": "bucket visit", "algorithmic_bytes_per_visit": blk if blk else None,
                    "algorithmic_bytes_per_hit": 28, "visits_per_launch": bucket_visits / max(scan_launches, 1),
                    "visits_per_guide": bucket_visits / max(args.steps * n, 1),
                    "bucket_entries_per_guide": streamed / max(args.steps * n, 1),
                    "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": scan_ms / ms_total,
                    "traffic": (tpv * bucket_visits / max(scan_launches, 1)) if tpv else None,
                    "traffic_source": ("profiles/scan_traffic.json: ncu --set full dram read+write bytes of one k_scan_triple "
                                       "launch / its bucket visits, x visits per launch here") if tpv else None,
                    "frac_of_nominal_8TBps": achieved / 8000.0,
                    "reference_equivalent": {"GB/s": ref_equiv, "x_hbm_peak": ref_equiv / peak,
                                             "definition": "4 B x list entries the reference's loop visits (SURVEY.md 8d, RES32 "
                                                           "layout) over the same kernel time",
                                             "candidates_per_launch": candidates / max(scan_launches, 1),
                                             "entries_read_per_candidate": streamed / max(candidates, 1)}}
    else:
        per_launch_bytes = bpc * candidates / max(scan_launches, 1)
        achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
        tpc = ncu_traffic_per_candidate(layout_name)
        roofline = {"bound": "hbm", "kernel": f"k_scan<{layout_name}>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_candidate": bpc, "candidates_per_launch": candidates / max(scan_launches, 1),
                    "kernel_ms_per_launch": per_launch_ms, "kernel_share_of_step": scan_ms / ms_total,
                    "traffic": (tpc * streamed / max(scan_launches, 1)) if tpc else None,
                    "traffic_source": ("profiles/scan_traffic.json: ncu --set full dram bytes per list entry streamed x entries "
                                       "streamed per launch (each chunk is read once per guide group)") if tpc else None,
                    "frac_of_nominal_8TBps": achieved / 8000.0}
        # With list reuse (max_group > 1) a chunk read from HBM once serves up to 32 guides, so the figure above --
        # algorithmic bytes of the reference's per-guide walk over time -- legitimately exceeds the HBM peak
        # (SURVEY.md 8d).  What then bounds the kernel is an integer pipe.  Both views are reported.
        pairs_per_s = candidates / max(scan_launches, 1) / (per_launch_ms / 1e3)
        streamed_gbs = bpc * streamed / max(scan_launches, 1) / (per_launch_ms / 1e3) / 1e9
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

    result = {"metric": "guides scored/sec (MIT+CFD, <=4 mm)", "value": value, "unit": "guides/s", "n_gpus": world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None,
              "dtype": "u16 bit-sliced compare (LOP3), f64 scores" if triple_scan else "u32/u64 xor+popcount, f64 scores", "data": "synthetic",
              "config": config, "clocks": dict(clocks.summary(), window=clock_window),
              "e2e": {"value": e2e_value, "unit": "guides/s", "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(n * 16)},
              "gpu_launches": int(launches), "scan_launches": int(scan_launches),
              "hits_per_guide": hits / max(args.steps * n, 1), "candidates_per_guide": candidates / max(args.steps * n, 1),
              "early_exit_fraction": early_exits / max(args.steps * n, 1),
              "roofline": roofline}

    if world == 1 and not args.no_cpu_baseline:
        try:
            n_sample = args.cpu_guides or min(n, 100 * (os.cpu_count() or 1))
            for t in (d_guides, d_mit, d_cfd):
                del t
            result["cpu_baseline"] = cpu_baseline(dev, guides, args, n_sample, hm, hc)
        except Exception as e:   # the baseline must not take the GPU number down with it
            result["cpu_baseline"] = {"value": None, "unit": "guides/s", "cores": os.cpu_count(), "kind": "reference",
                                      "sample": f"failed: {e}"}
    emit(result)
    dev.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
