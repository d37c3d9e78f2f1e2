#!/usr/bin/env python
"""bench.py -- reads/s classified on the BASELINE.json configs[1] workload.

  python bench.py --gpus N --steps K --warmup W          our arm (CUDA path through the C ABI)
  python bench.py --impl reference ...                   the CPU arm: the oracle restatement of Slacken's algorithm on
                                                         all host threads (the Scala/Spark reference cannot run here:
                                                         no JVM in the image)
A step = one pass of the classify hot path over one batch of synthetic reads. `value` is timed with CUDA events
with the reads resident in HBM; `e2e` goes through slk_classify_batch with pinned HOST buffers (H2D + D2H inside).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_workload as bw  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        return [l for t, l in self.lines if t0 - 0.03 <= t <= t1 + 0.05]

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarize(lines):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v == "Active":
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(args, n_reads):
    """dram__bytes_read + dram__bytes_write of ONE launch of the classify kernel, from the committed ncu --set full capture
    (profiles/r01_traffic.json); None when the run is not the configuration that was profiled."""
    if args.traffic is not None:
        return args.traffic
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        t = json.load(open(p))
        if int(t["reads_per_launch"]) == int(n_reads):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_library(w, oracle, parents, genome_taxa, threads):
    """The reference's build path on the CPU: synthetic genomes -> removeInvalid -> super-mers -> LCA records."""
    t0 = time.perf_counter()
    lib = oracle.Library(oracle.params(k=w.k, m=w.m, spaces=w.spaces), parents, int(w.total_bases / 2.9))
    per = max(1, min(w.n_genomes, (64 << 20) // w.genome_len))
    for g0 in range(0, w.n_genomes, per):
        g1 = min(w.n_genomes, g0 + per)
        bases = oracle.synth_genome(w.gseed, g0 * w.genome_len, (g1 - g0) * w.genome_len)
        off = (np.arange(g1 - g0 + 1, dtype=np.int64) * w.genome_len)
        lib.add_sequences(bases, off, genome_taxa[g0:g1])
    return lib, time.perf_counter() - t0


def run_reference(args, w):
    from oracle import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    threads = oracle.max_threads()
    parents, ranks, names, genome_taxa = bw.taxonomy(w)
    log(f"[reference] building the {w.total_bases/1e9:.2f} Gbp library on {threads} host threads ...")
    lib, tb = cpu_library(w, oracle, parents, genome_taxa, threads)
    log(f"[reference] library: {len(lib)} records in {tb:.1f} s")
    sample = min(w.n_reads, args.cpu_sample)
    off = np.arange(sample + 1, dtype=np.int64) * w.read_len
    times = []
    for step in range(args.warmup + args.steps):
        first = (step * sample) % max(1, w.n_reads - sample + 1)
        reads = oracle.synth_reads(w.gseed, w.rseed, w.n_genomes, w.genome_len, first, sample, w.read_len)
        t0 = time.perf_counter()
        lib.classify(reads, off, confidence=w.confidence, min_hit_groups=w.min_hit_groups, threads=threads, with_hits=True)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": "reads/sec classified (150bp)", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": w.name, "sample_reads_per_step": sample, "library_records": len(lib)},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} reads per step, {len(times)} steps, full {len(lib)}-record library built on the CPU in {tb:.0f} s"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of Slacken's algorithm (oracle/), not Spark: the reference needs a JVM, absent from this image"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def build_gpu_library(ctx, tax, params, w, genome_taxa):
    from slacken_b200 import LibraryBuilder
    import ctypes as C
    from slacken_b200._lib import check
    per = max(1, min(w.n_genomes, (256 << 20) // w.genome_len))
    d_bases = ctx.dev_alloc(per * w.genome_len)
    d_off = ctx.dev_alloc((per + 1) * 8)
    d_tax = ctx.dev_alloc(per * 4)
    ctx.sync()
    t0 = time.perf_counter()
    b = LibraryBuilder(ctx, tax, params, expected_bases=w.total_bases)
    for g0 in range(0, w.n_genomes, per):
        g1 = min(w.n_genomes, g0 + per)
        n = (g1 - g0) * w.genome_len
        check(ctx._L.slk_synth_genome_dev(ctx.h, w.gseed, g0 * w.genome_len, n, C.c_void_p(d_bases)))
        ctx.h2d(d_off, np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len))
        ctx.h2d(d_tax, genome_taxa[g0:g1])
        b.add_dev(d_bases, d_off, d_tax, g1 - g0, n)
    index = b.finish()
    b.close()
    ctx.sync()
    dt = time.perf_counter() - t0
    for p in (d_bases, d_off, d_tax):
        ctx.dev_free(p)
    return index, dt


def run_ours(args, w):
    import ctypes as C
    from slacken_b200 import Classifier, DeviceTimer, GpuContext, IndexParams, ReportCounts, Taxonomy
    from slacken_b200._lib import check
    from slacken_b200.host import DETAIL_DTYPE, HIT_DTYPE, ClassifiedBatch, PackedReads, block_offsets

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from slacken_b200.dist import bind_to_gpu_numa
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa(local)   # before any pinned buffer exists
    ctx = GpuContext(local)
    parents, ranks, names, genome_taxa = bw.taxonomy(w)
    tax = Taxonomy(ctx, parents, ranks, names)
    params = IndexParams(k=w.k, m=w.m, spaces=w.spaces)
    index, t_build = build_gpu_library(ctx, tax, params, w, genome_taxa)
    if rank == 0:
        log(f"[ours] library: {len(index)} records from {w.total_bases/1e9:.2f} Gbp in {t_build:.2f} s "
            f"({w.total_bases/t_build/1e9:.2f} Gbases/s incl. synthetic genome generation)")

    # this rank's shard of reads (weak scaling: every GPU classifies n_reads reads of its own)
    n, L = w.n_reads, w.read_len
    first = rank * n
    d_reads = ctx.dev_alloc(n * L)
    check(ctx._L.slk_synth_reads_dev(ctx.h, w.gseed, w.rseed, w.n_genomes, w.genome_len, first, n, L, C.c_void_p(d_reads)))
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    d_off = ctx.dev_alloc(off.nbytes)
    ctx.h2d(d_off, off)
    # stage 1 on the device: the same reads as 2-bit packed blocks + ambiguity masks (the input form the north star names)
    boff = block_offsets(off)
    n_blocks = int(boff[-1])
    d_boff, d_codes, d_mask, d_len = ctx.dev_alloc(boff.nbytes), ctx.dev_alloc(n_blocks * 8), ctx.dev_alloc(n_blocks * 4), ctx.dev_alloc(n * 4)
    ctx.h2d(d_boff, boff)
    ctx.pack_reads_dev(d_reads, d_off, n, d_boff, d_codes, d_mask, d_len)
    t0 = time.perf_counter()
    ctx.pack_reads_dev(d_reads, d_off, n, d_boff, d_codes, d_mask, d_len)
    t_pack = time.perf_counter() - t0
    cls = Classifier(index)
    counts = ReportCounts(ctx, tax, 1)
    cls.attach_counts(counts, 0)
    hits_cap = cls.hits_bound(n, n * L, False)
    d_taxon, d_flags = ctx.dev_alloc(n * 4), ctx.dev_alloc(n)
    d_detail, d_hits, d_used = ctx.dev_alloc(n * DETAIL_DTYPE.itemsize), ctx.dev_alloc(hits_cap * 8), ctx.dev_alloc(8)

    counts_t = None
    if dist is not None:
        import torch

        class _Wrap:  # zero-copy view of the device counter matrix for the NCCL all-reduce
            __cuda_array_interface__ = {"shape": (tax.size,), "typestr": "<i8", "data": (counts.device_ptr(), False), "version": 2}
        counts_t = torch.as_tensor(_Wrap(), device=f"cuda:{local}")

    def device_step_ascii():
        cls.classify_dev(d_reads, d_off, 0, 0, n, d_taxon, d_flags, d_detail, d_hits, hits_cap, d_used,
                         confidence=w.confidence, min_hit_groups=w.min_hit_groups)

    def device_step():
        cls.classify_packed_dev(d_codes, d_mask, d_boff, d_len, 0, 0, 0, 0, n, d_taxon, d_flags, d_detail, d_hits, hits_cap,
                                d_used, confidence=w.confidence, min_hit_groups=w.min_hit_groups)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: inputs resident in HBM, CUDA events on the launch stream
    for _ in range(args.warmup):
        device_step()
    cls.sync()
    p0, h0 = cls.stats()
    l0 = cls.launches
    counts.reset()
    timer = DeviceTimer(cls)
    barrier()
    tw0 = time.perf_counter()
    timer.start()
    for _ in range(args.steps):
        device_step()
    timer.stop()
    ms = timer.elapsed_ms()
    if counts_t is not None:   # report aggregation across ranks: one all-reduce of the counter vector
        import torch
        t_ar = time.perf_counter()
        dist.all_reduce(counts_t)
        torch.cuda.synchronize()
        ms += 1e3 * (time.perf_counter() - t_ar)
    barrier()
    tw1 = time.perf_counter()
    ms = max_over_ranks(ms)
    p1, h1 = cls.stats()
    launches = cls.launches - l0
    probes_per_read = (p1 - p0) / (args.steps * n)
    hits_per_read = (h1 - h0) / (args.steps * n)
    clocks = ClockSampler.summarize(sampler.window(tw0, tw1))
    value = world * n * args.steps / (ms / 1e3)
    used = np.zeros(1, dtype=np.uint64)
    ctx.d2h(used, d_used)
    assert int(used[0]) <= hits_cap
    rep = counts.fetch(0)
    total_reads_counted = int(rep.sum())

    # roofline of the dominant (only) kernel of the step: the fused classify kernel
    S, H = probes_per_read, hits_per_read
    bytes_per_read = 12.0 * n_blocks / n + 8 + 4 + 32.0 * S + (4 + 1 + 24) + 8.0 * H   # packed input
    achieved = bytes_per_read * n * args.steps / (ms / 1e3) / 1e9 / 1.0  # per GPU: every rank runs the same launch
    peak, peak_src = measured_peak()
    # the same launch on ASCII-resident input (stage 1 fused into the kernel), for comparison
    for _ in range(2):
        device_step_ascii()
    cls.sync()
    t_a = DeviceTimer(cls)
    t_a.start()
    for _ in range(args.steps):
        device_step_ascii()
    t_a.stop()
    ms_ascii = max_over_ranks(t_a.elapsed_ms())
    cls.attach_counts(None)
    roofline = {"bound": "hbm", "kernel": "classify_kernel<5,true,true>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": measured_traffic(args, n),
                "algorithmic_bytes_per_launch": bytes_per_read * n,
                "bytes_per_read": bytes_per_read, "probes_per_read": S, "merged_hits_per_read": H,
                "probe_sectors_gbs": 32.0 * S * n * args.steps / (ms / 1e3) / 1e9,
                "lookups_per_s": S * n * args.steps / (ms / 1e3),
                "frac_of_measured_request_ceiling": S * n * args.steps / (ms / 1e3) / 42.8e9,
                "random_gather_ceiling": "a random 32-byte sector costs the B200 a whole 128-byte DRAM line, and the chip serves at most ~43 G "
                                         "random line requests/s to a plain lookup kernel (profiles/r01_probe_microbench.md, DESIGN.md section 3): "
                                         "one random sector per lookup cannot exceed ~0.21 of the copy-bandwidth roofline"}

    # ---- e2e: the host-buffer entry point, pinned host memory, H2D + D2H inside the timed region
    h_reads = ctx.pinned(n * L, np.uint8)
    ctx.d2h(h_reads, d_reads)
    h_off = ctx.pinned(n + 1, np.uint64)
    h_off[:] = off
    e2e_cap = 16 * n
    out = ClassifiedBatch(ctx.pinned(n, np.int32), ctx.pinned(n, np.uint8), ctx.pinned(n, DETAIL_DTYPE), ctx.pinned(e2e_cap, HIT_DTYPE))
    hp = PackedReads(ctx.pinned(n_blocks, np.uint64), ctx.pinned(n_blocks, np.uint32), ctx.pinned(n + 1, np.uint64), ctx.pinned(n, np.uint32))
    ctx.d2h(hp.codes, d_codes); ctx.d2h(hp.mask, d_mask); ctx.d2h(hp.len, d_len)
    hp.boff[:] = boff

    def timed_e2e(fn):
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt), t0, time.perf_counter()

    e2e_s, te0, te1 = timed_e2e(lambda: cls.classify_packed(hp, confidence=w.confidence, min_hit_groups=w.min_hit_groups, out=out))
    clocks_e2e = ClockSampler.summarize(sampler.window(te0, te1))
    e2e = {"value": world * n * args.steps / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": int(hp.nbytes),
           "d2h_bytes_per_step": int(out.taxon.nbytes + out.flags.nbytes + out.detail.nbytes + out.hits_used * 8),
           "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": "slk_classify_batch_packed: pinned HOST buffers holding 2-bit packed reads + ambiguity masks (the host-side "
                  "packing is the Scala driver's batching work and is outside the timed region), per-read hit lists on",
           "clocks": clocks_e2e}
    r_s, _, _ = timed_e2e(lambda: cls.classify_packed(hp, confidence=w.confidence, min_hit_groups=w.min_hit_groups,
                                                      per_read_output=False, out=out))
    e2e_report = {"value": world * n * args.steps / r_s, "unit": "reads/s", "h2d_bytes_per_step": int(hp.nbytes),
                  "d2h_bytes_per_step": int(out.taxon.nbytes + out.flags.nbytes + out.detail.nbytes), "ms_per_step": 1e3 * r_s / args.steps,
                  "api": "slk_classify_batch_packed without per-read hit lists (the reference's --nodetailed mode, "
                         "slacken/Classifier.scala:259-410): taxon, flags and lengths per read come back, no hits"}
    a_s, _, _ = timed_e2e(lambda: cls.classify(h_reads, h_off, confidence=w.confidence, min_hit_groups=w.min_hit_groups, out=out))
    e2e_ascii = {"value": world * n * args.steps / a_s, "unit": "reads/s", "h2d_bytes_per_step": int(h_reads.nbytes + h_off.nbytes),
                 "ms_per_step": 1e3 * a_s / args.steps, "api": "slk_classify_batch: pinned HOST buffers holding ASCII reads"}
    sampler.stop()

    line = {"metric": "reads/sec classified (150bp)", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": w.name, "reads_per_gpu_per_step": n, "library_records": len(index),
                       "library": "replicated per GPU", "input": "2-bit packed reads + ambiguity mask (72 B/read)", "l2": "inputs (reads 1.5 GB + table) are far larger than L2; no flush needed",
                       "host_cpus_bound_to_gpu_numa_node": len(numa_cpus),
                       "classified_fraction": float((rep.sum() - rep[0]) / max(1, rep.sum())),
                       "reads_counted_in_report": total_reads_counted},
            "probes_per_s": value * S, "clocks": clocks, "e2e": e2e, "e2e_report_only": e2e_report, "e2e_ascii_input": e2e_ascii,
            "value_ascii_input": {"value": world * n * args.steps / (ms_ascii / 1e3), "unit": "reads/s", "ms_per_step": ms_ascii / args.steps,
                                  "note": "same launch with ASCII reads resident in HBM (stage 1 fused into the kernel)"},
            "encode_kernel": {"ms": 1e3 * t_pack, "reads_per_s": n / t_pack, "gbs": (L + 8 + 12.0 * n_blocks / n + 4) * n / t_pack / 1e9,
                              "note": "stage 1 alone (pack_reads_kernel: ASCII -> 2-bit blocks + mask), wall clock around a synchronous call"},
            "gpu_launches": int(launches), "roofline": roofline,
            "build": {"seconds": t_build, "gbases_per_s": w.total_bases / t_build / 1e9, "records": len(index)}}

    # ---- cpu_baseline: the oracle on the host cores, bounded sample, rank 0 at N=1 only
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)   # the CPU leg gets every host core again
        from oracle import oracle
        threads = oracle.max_threads()
        t0 = time.perf_counter()
        id1, tx = index.records(sort=False)
        olib = oracle.Library(oracle.params(k=w.k, m=w.m, spaces=w.spaces), parents, len(id1))
        olib.add_records(id1, tx)
        del id1, tx
        t_lib = time.perf_counter() - t0
        sample = min(n, args.cpu_sample)
        o64 = off[:sample + 1].astype(np.int64)
        t0 = time.perf_counter()
        res, _, _, _ = olib.classify(h_reads[:sample * L], o64, confidence=w.confidence, min_hit_groups=w.min_hit_groups,
                                     threads=threads, with_hits=True)
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(res["taxon"], out.taxon[:sample]))
        line["cpu_baseline"] = {"value": sample / dt, "unit": "reads/s", "cores": threads, "kind": "port",
                                "sample": f"first {sample} reads of the same batch, {dt:.1f} s; library = the {len(olib)} records "
                                          f"of the GPU build loaded into the oracle's CPU table in {t_lib:.0f} s",
                                "taxa_equal_to_gpu_on_sample": same}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=None, help="override reads per GPU per step (default: configs[1], 10M)")
    ap.add_argument("--genomes", type=int, default=None)
    ap.add_argument("--genome-len", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="reads in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from an ncu --set full capture")
    args = ap.parse_args()
    w = bw.Workload()
    if args.reads:
        w.n_reads = args.reads
    if args.genomes:
        w.n_genomes = args.genomes
    if args.genome_len:
        w.genome_len = args.genome_len
    if args.reads or args.genomes or args.genome_len:
        w.name = (f"REDUCED synthetic {w.n_reads} x {w.read_len}bp reads vs {w.total_bases/1e9:.3f} Gbp library "
                  f"({w.n_genomes} x {w.genome_len}), k35/m31/s7")
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
