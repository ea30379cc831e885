#!/usr/bin/env python
"""bench.py — oligo-frequency-vector throughput on B200 (BASELINE.json metric: Gbases/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic sequences.  Default workload is
BASELINE.json configs[1]: 10 M x 150 bp reads, k=5 canonical, f32 normalised rows (512 columns).
  value     : Gbases/s with inputs and outputs resident in HBM (device entry point of the C ABI)
  e2e       : Gbases/s through the host-buffer C-ABI call (pinned host buffers, H2D + D2H inside)
  roofline  : algorithmic bytes / device time of the step vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the oracle's restatement of the reference algorithm (C + OpenMP, all host cores) on a
                 bounded sample of the same workload.  The reference itself is Rust and cannot be
                 built in this image, so kind = "port".
N > 1 (torchrun): every rank runs the same per-GPU workload on its own GPU (weak scaling, rows are
independent so there is no collective on the data path); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (n, length spec, k, canonical, out dtype, norm)
    "reads150_k5": dict(n=10_000_000, length=150, k=5, dtype="f32", norm=1, seed=20250001,
                        desc="10M x 150bp short reads, k=5 canonical f32 normalised (BASELINE configs[1])"),
    "reads150_k3": dict(n=10_000_000, length=150, k=3, dtype="f32", norm=1, seed=20250001,
                        desc="10M x 150bp short reads, k=3 canonical f32 normalised (the CLI's default k; probe, not a BASELINE config)"),
    "reads150_k4": dict(n=10_000_000, length=150, k=4, dtype="f32", norm=1, seed=20250001,
                        desc="10M x 150bp short reads, k=4 canonical f32 normalised (probe, not a BASELINE config)"),
    "reads150_k6": dict(n=2_000_000, length=150, k=6, dtype="f32", norm=1, seed=20250001,
                        desc="2M x 150bp short reads, k=6 canonical f32 normalised (probe, not a BASELINE config)"),
    "reads150_k7": dict(n=500_000, length=150, k=7, dtype="f32", norm=1, seed=20250001,
                        desc="500k x 150bp short reads, k=7 canonical f32 normalised (probe, not a BASELINE config)"),
    "reads10k_k7": dict(n=1_000_000, length=10_000, k=7, dtype="f32", norm=1, seed=20250002,
                        desc="1M x 10kbp long reads, k=7 canonical f32 normalised (BASELINE configs[2])"),
    "reads10k_k6": dict(n=1_000_000, length=10_000, k=6, dtype="f32", norm=1, seed=20250002,
                        desc="1M x 10kbp long reads, k=6 canonical f32 normalised (occupancy probe, not a BASELINE config)"),
    "reads10k_k5": dict(n=1_000_000, length=10_000, k=5, dtype="f32", norm=1, seed=20250002,
                        desc="1M x 10kbp long reads, k=5 canonical f32 normalised (occupancy probe, not a BASELINE config)"),
    "contigs_k4": dict(n=20_000, length="contigs", k=4, dtype="f32", norm=1, seed=20250003,
                       desc="20k metagenome contigs 1-500kbp with N runs, k=4 (BASELINE configs[3])"),
    "reads10k_k8": dict(n=100_000, length=10_000, k=8, dtype="u32", norm=0, seed=20250004,
                        desc="100k x 10kbp, k=8 canonical u32 counts (BASELINE configs[4] i)"),
    "reads100k_k10": dict(n=2_000, length=100_000, k=10, dtype="u32", norm=0, seed=20250005,
                          desc="2k x 100kbp, k=10 canonical u32 counts, global-atomic path (configs[4] ii)"),
    "reads100k_k10_f32": dict(n=2_000, length=100_000, k=10, dtype="f32", norm=1, seed=20250005,
                              desc="2k x 100kbp, k=10 canonical f32 normalised (probe of the in-place normalisation, "
                                   "not a BASELINE config)"),
    "reads100k_k9": dict(n=8_000, length=100_000, k=9, dtype="u32", norm=0, seed=20250006,
                         desc="8k x 100kbp, k=9 canonical u32 counts (probe, not a BASELINE config)"),
}
DT = {"u32": (0, np.uint32, 4), "f32": (1, np.float32, 4), "f64": (2, np.float64, 8)}


def dim_of(k: int) -> int:
    return 4 ** k // 2 if k % 2 else (4 ** k + 4 ** (k // 2)) // 2


def make_workload(spec: dict, scale: float, device):
    """Synthetic bases/offsets on the GPU (torch RNG, seeded).  Returns (bases u8, offsets i64) tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(spec["seed"])
    n = max(32, int(spec["n"] * scale))
    letters = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    if spec["length"] == "contigs":
        # log-uniform lengths in [1e3, 5e5]; per-contig GC in [0.3, 0.7]; N runs; sporadic IUPAC; soft-masking
        u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
        lengths = torch.exp(u * (np.log(5e5) - np.log(1e3)) + np.log(1e3)).to(torch.int64)
        offsets = torch.zeros(n + 1, dtype=torch.int64, device=device)
        offsets[1:] = torch.cumsum(lengths, 0)
        total = int(offsets[-1])
        seq_id = torch.repeat_interleave(torch.arange(n, device=device), lengths)
        gc = 0.3 + 0.4 * torch.rand(n, generator=g, device=device)
        r = torch.rand(total, generator=g, device=device)
        gcb = gc[seq_id]
        is_gc = r < gcb
        half = torch.rand(total, generator=g, device=device) < 0.5
        code = torch.where(is_gc, torch.where(half, 1, 2), torch.where(half, 0, 3))
        bases = letters[code]
        del r, gcb, is_gc, half, code, seq_id
        # N runs: Poisson(2) per contig, log-uniform length [1, 1e4]
        nruns = torch.poisson(torch.full((n,), 2.0, device=device), generator=g).to(torch.int64)
        rid = torch.repeat_interleave(torch.arange(n, device=device), nruns)
        rl = torch.exp(torch.rand(len(rid), generator=g, device=device) * np.log(1e4)).to(torch.int64)
        rs = offsets[rid] + (torch.rand(len(rid), generator=g, device=device, dtype=torch.float64)
                             * lengths[rid].double()).to(torch.int64)
        re_ = torch.minimum(rs + rl, offsets[rid + 1])
        delta = torch.zeros(total + 1, dtype=torch.int32, device=device)
        delta.index_add_(0, rs, torch.ones_like(rs, dtype=torch.int32))
        delta.index_add_(0, re_, -torch.ones_like(re_, dtype=torch.int32))
        inrun = torch.cumsum(delta[:-1], 0) > 0
        bases[inrun] = ord("N")
        del delta, inrun
        iupac = torch.tensor(list(b"RYKMSW"), dtype=torch.uint8, device=device)
        m = torch.rand(total, generator=g, device=device) < 1e-4
        bases[m] = iupac[torch.randint(0, 6, (int(m.sum()),), generator=g, device=device)]
        # 5 % soft-masked: lower-case stretches of 1 kbp blocks
        blk = torch.rand((total + 1023) // 1024, generator=g, device=device) < 0.05
        low = torch.repeat_interleave(blk, 1024)[:total]
        bases = torch.where(low & (bases != ord("N")), bases | 0x20, bases)
        return bases.contiguous(), offsets
    L = int(spec["length"])
    total = n * L
    bases = torch.empty(total, dtype=torch.uint8, device=device)
    step = 1 << 28
    for a in range(0, total, step):
        b = min(total, a + step)
        bases[a:b] = letters[torch.randint(0, 4, (b - a,), generator=g, device=device)]
    offsets = torch.arange(n + 1, dtype=torch.int64, device=device) * L
    if L <= 1000:   # 1 % of reads get one N at a uniform position
        pick = torch.nonzero(torch.rand(n, generator=g, device=device) < 0.01).flatten()
        pos = torch.randint(0, L, (len(pick),), generator=g, device=device)
        bases[pick * L + pos] = ord("N")
    else:           # 0.1 % of positions N, in runs of geometric mean length 10
        nrun = int(total * 0.001 / 10)
        st = torch.randint(0, total, (nrun,), generator=g, device=device)
        ln = torch.distributions.Geometric(torch.tensor(0.1, device=device)).sample((nrun,)).to(torch.int64) + 1
        for j in range(int(ln.max())):
            m = ln > j
            bases[torch.clamp(st[m] + j, max=total - 1)] = ord("N")
    return bases, offsets


def algorithmic_bytes(total_bases: int, n: int, dim: int, esize: int) -> int:
    """ASCII bases + offset index + one output element per column."""
    return total_bases + (n + 1) * 8 + n * dim * esize


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(index: int):
    """Pin this rank's host threads (and therefore its page-locked buffers, first touch) to the CPUs NVML
    reports as local to the GPU, so 8 ranks do not push their D2H traffic across the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)[0], len(cpus)
    except Exception:
        return None
    return None


def cpu_sample(spec: dict, bases_h: np.ndarray, offsets_h: np.ndarray, budget_reads: int):
    """Time the oracle's port of the reference path on a bounded prefix of the workload."""
    from oracle import oracle as O
    n = min(len(offsets_h) - 1, budget_reads)
    offs = np.ascontiguousarray(offsets_h[: n + 1]).astype(np.uint64)
    nb = int(offs[-1])
    thr = host_threads()
    O.baseline_batch(bases_h[:nb], offs, spec["k"], True, spec["norm"], thr)  # warm-up pass (allocator, page faults), as in --impl reference
    t0 = time.perf_counter()
    _, used = O.baseline_batch(bases_h[:nb], offs, spec["k"], True, spec["norm"], thr)
    dt = time.perf_counter() - t0
    return nb / dt / 1e9, used, n, dt


def run_reference(args, spec, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; Rust reference not buildable here)."""
    if rank != 0:
        return
    rng = np.random.default_rng(spec["seed"])
    L = spec["length"] if spec["length"] != "contigs" else 80_000
    # size the per-step sample for ~3 s of all-core CPU work
    from oracle import oracle as O
    cores = host_threads()
    # ~3 s of all-core work per step at ~0.01 Gbases/s/core (measured: 0.15 Gbases/s on 16 cores)
    reads = max(64, int(min(spec["n"] * args.scale, 3.0 * cores * 0.01e9 / L)))
    lengths = np.full(reads, L, dtype=np.uint64)
    offsets = np.zeros(reads + 1, dtype=np.uint64)
    np.cumsum(lengths, out=offsets[1:])
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(offsets[-1]), dtype=np.uint8)]
    for _ in range(args.warmup):
        O.baseline_batch(bases, offsets, spec["k"], True, spec["norm"], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, used = O.baseline_batch(bases, offsets, spec["k"], True, spec["norm"], cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = int(offsets[-1]) / dt / 1e9
    sample = f"{reads} sequences x {L} bp per step ({int(offsets[-1])} bases)"
    line = {
        "impl": "reference", "metric": "oligo_vectors_throughput", "value": val, "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": spec["desc"], "k": spec["k"], "canonical": True},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sequences_per_s": reads / dt,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="reads150_k5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's sequence count")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the file-level (FASTQ -> text) driver sample")
    ap.add_argument("--force-path", type=int, default=None)
    ap.add_argument("--seq-threads", type=int, default=None)
    ap.add_argument("--global-wave-mb", type=int, default=None)
    ap.add_argument("--dense-odd", type=int, default=None)
    ap.add_argument("--even-rank", type=int, default=None)
    ap.add_argument("--global-steps-per-warp", type=int, default=None)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="any ktb_oligo_set_option key (repeatable)")
    args = ap.parse_args()
    spec = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, spec, rank, world)
        return

    import torch
    import torch.distributed as dist
    from kmertools_b200 import OligoComputer, HostBuffer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kmertools_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    k = spec["k"]
    code, npdt, esize = DT[spec["dtype"]]
    dim = dim_of(k)
    oc = OligoComputer(k, device=local)
    if args.force_path is not None:
        oc.set_option("force_path", args.force_path)
    if args.seq_threads is not None:
        oc.set_option("seq_threads", args.seq_threads)
    if args.global_steps_per_warp is not None:
        oc.set_option("global_steps_per_warp", args.global_steps_per_warp)
    if args.even_rank is not None:
        oc.set_option("even_rank", args.even_rank)
    if args.dense_odd is not None:
        oc.set_option("dense_odd", args.dense_odd)
    if args.global_wave_mb is not None:
        oc.set_option("global_wave_bytes", args.global_wave_mb << 20)
    for kv in args.opt:
        key, _, val = kv.partition("=")
        oc.set_option(key, int(val))
    bases, offsets = make_workload(spec, args.scale, dev)
    n = offsets.numel() - 1
    total_bases = int(offsets[-1])
    tdt = {"u32": torch.int32, "f32": torch.float32, "f64": torch.float64}[spec["dtype"]]
    out = torch.empty((n, dim), dtype=tdt, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        oc.vectorise_device(bases.data_ptr(), offsets.data_ptr(), n, total_bases, out.data_ptr(),
                            norm_mode=spec["norm"], mins=True, out_dtype=code, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    launches_per_step = oc.stats()["launches"]
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the timed region is long enough for several nvidia-smi samples: repeat the K-step block if needed
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    # keep the GPU under the same load a little longer so the clock sampler sees it (not timed)
    t_end = time.time() + 1.0
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * total_bases / (ms_per_step * 1e-3) / 1e9

    # ---- sanity: size-independent property on the full output (rows sum to 1 or 0)
    if spec["norm"]:
        s = out[: min(n, 200_000)].sum(dim=1, dtype=torch.float64)
        ok = bool(torch.all(((s - 1).abs() < 1e-3) | (s == 0)))
    else:
        ok = True

    # ---- roofline of the step's dominant kernel
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    B = algorithmic_bytes(total_bases, n, dim, esize)
    achieved = B / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = ROOT / "profiles" / "ncu_dram_traffic.json"   # dram__bytes_read+write of the dominant kernel (one ncu --set full capture)
    if tpath.exists():
        ent = json.loads(tpath.read_text()).get(args.workload)
        if ent:
            traffic = ent["dram_bytes_per_sequence"] * n
            traffic_src = f"{ent['kernel']}: {ent['capture']}, scaled per sequence ({ent['summary']})"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_step": B,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- e2e through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        class _Pageable:   # same interface as HostBuffer when page-locking this much memory is refused
            def __init__(self, shape, dt):
                self.array = np.empty(shape, dtype=dt)

            def free(self):
                self.array = None

        host_kind = "pinned (ktb_host_alloc)"
        try:
            hb = HostBuffer((total_bases,), np.uint8)
            ho = HostBuffer((n + 1,), np.uint64)
            hout = HostBuffer((n, dim), npdt)
        except Exception as exc:   # e.g. 8 ranks x 22 GB of pinned memory on a small host
            host_kind = f"pageable (pinned allocation failed: {exc})"
            hb, ho, hout = _Pageable((total_bases,), np.uint8), _Pageable((n + 1,), np.uint64), _Pageable((n, dim), npdt)
        hb.array[:] = bases.cpu().numpy()
        ho.array[:] = offsets.cpu().numpy().astype(np.uint64)
        oc.vectorise_packed(hb.array, ho.array, norm_mode=spec["norm"], mins=True, dtype=npdt, out=hout.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            oc.vectorise_packed(hb.array, ho.array, norm_mode=spec["norm"], mins=True, dtype=npdt, out=hout.array)
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        st = oc.stats()
        e2e = {"value": world * total_bases / dt / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
               "ms_per_step": dt * 1e3, "kernel_ms": st["kernel_ms"], "h2d_ms": st["h2d_ms"], "d2h_ms": st["d2h_ms"],
               "host_memory": host_kind, "numa_binding": numa}
        # parity spot check against the device-path result
        same = bool(np.array_equal(hout.array[:1000], out[:1000].cpu().numpy()))
        e2e["matches_device_path"] = same
        keep_h = (hb, ho)
        hout.free()
    else:
        keep_h = None

    # ---- CPU baseline on rank 0 (bounded sample)
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))   # the CPU arm uses every host core
        except Exception:
            pass
        if keep_h is not None:
            bh, oh = keep_h[0].array, keep_h[1].array
        else:
            bh, oh = bases.cpu().numpy(), offsets.cpu().numpy().astype(np.uint64)
        cores = host_threads()
        mean_len = max(1, total_bases // n)
        budget = max(64, int(cores * 0.01e9 * 3 / mean_len))   # same bounded sample as --impl reference (~3 s)
        gb, used, ns, dt = cpu_sample(spec, bh, oh, budget)
        # the oracle as checker (SURVEY §8d "parity checks in every run"): the device-path rows of the first
        # sequences of the workload against the CPU restatement, counts exact, f32 = (float)(f64 quotient)
        from oracle import oracle as O
        m = int(min(n, 10_000, max(1, 50_000_000 // mean_len), max(1, (1 << 28) // (dim * 8))))
        po = np.ascontiguousarray(oh[: m + 1]).astype(np.uint64)
        want, _ = O.vectorise_batch(bh[: int(po[-1])], po, k, True, spec["norm"])
        got = out[:m].cpu().numpy().view(npdt)
        parity = {"sequences": m, "exact": bool(np.array_equal(got, want.astype(npdt))),
                  "rule": "u32 counts equal; f32 rows equal (float)(oracle f64 row), 0 ulp"}
        cpu = {"value": gb, "unit": "Gbases/s", "cores": used, "kind": "port",
               "sample": f"first {ns} sequences of the workload, {dt:.2f} s, C+OpenMP restatement of the reference "
                         f"(Rust reference not buildable in this image)"}

    # ---- file-level driver sample (rank 0, N=1): FASTQ in, text rows out, host parse time broken out
    cli = None
    if rank == 0 and world == 1 and not args.no_cli and spec["length"] != "contigs" and spec["norm"]:
        import tempfile
        from kmertools_b200 import io as kio
        L = int(spec["length"])
        m = max(1, min(n, int(1.0e9 / (dim * 9))))          # ~1 GB of text
        hb_np = bases[: m * L].cpu().numpy().reshape(m, L)
        with tempfile.TemporaryDirectory() as td:
            fq = os.path.join(td, "sample.fq")
            rec = np.empty((m, 2 * L + 8), dtype=np.uint8)   # "@r\n" + seq + "\n+\n" + qual + "\n"
            rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
            rec[:, 3:3 + L] = hb_np
            rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
            rec[:, 6 + L:6 + 2 * L] = ord("I")
            rec[:, 6 + 2 * L] = ord("\n")
            rec[:, 7 + 2 * L] = ord("\n")
            rec = rec[:, :7 + 2 * L]
            np.ascontiguousarray(rec).tofile(fq)
            kio.comp_oligo(fq, os.path.join(td, "warm.kmers"), k=k)      # warm-up (page cache, pinned buffers)
            stc = kio.comp_oligo(fq, os.path.join(td, "out.kmers"), k=k)
        cli = {"records": int(stc["records"]), "bases": int(stc["bases"]), "text_bytes": int(stc["bytes_written"]),
               "total_ms": stc["total_ms"], "host_parse_ms": stc["parse_ms"], "gpu_wait_ms": stc["gpu_wait_ms"],
               "write_ms": stc["write_ms"], "gbases_per_s": stc["bases"] / stc["total_ms"] / 1e6,
               "what": "kmertools comp oligo on a FASTQ sample of the workload (tmpfs/ disk I/O included), text "
                       "formatted on the GPU"}

    if rank == 0:
        line = {
            "metric": "oligo_vectors_throughput", "value": value, "unit": "Gbases/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": spec["dtype"],
            "data": "synthetic",
            "config": {"workload": args.workload, "desc": spec["desc"], "k": k, "canonical": True, "dim": dim,
                       "sequences_per_gpu": n, "bases_per_gpu": total_bases,
                       "l2": "inputs+outputs per step exceed L2 (126 MB)" if B > 4 * 126e6 else "small working set"},
            "sequences_per_s": world * n / (ms_per_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "cli": cli, "clocks": clocks,
            "gpu_launches": int(launches_per_step * args.steps), "rows_sum_to_one": ok, "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
