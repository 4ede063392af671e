#!/usr/bin/env python
"""bench.py — pair-HMM GCUPS of the B200 engine on BASELINE.json's (reads x haplotypes) workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3] [--reads R] [--haps H]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path (HaplotypeLikelihoodArray::populate semantics, one mapping position per pair, naive
shortcut disabled so that CPU and GPU do identical DP work) over one synthetic batch. Metric: banded DP cell-updates/s,
cells per alignment = 2*(L+band)*band (simd_pair_hmm.hpp:271). Prints ONE JSON line on rank 0.

  value  device-resident inputs → device-resident double matrix (+ gather to rank 0 when N > 1), max over ranks
  e2e    the same call through the C ABI with pinned HOST buffers: H2D of the batch and D2H of the matrix inside the timed region
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement"
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", help="BASELINE config name (octopus_b200.synth.CONFIGS)")
    ap.add_argument("--reads", type=int, default=None, help="override reads per GPU (debug)")
    ap.add_argument("--haps", type=int, default=None)
    ap.add_argument("--cpu-sample-reads", type=int, default=None, help="reads in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # non-headline modes (diagnostics; the headline is the default: one position per pair, DP for every pair, no flank state)
    ap.add_argument("--flank", default=None, help="LHS,RHS flank sizes: exercises the traceback + flank-discount path")
    ap.add_argument("--shortcut", action="store_true", help="enable the reference's naive shortcut (reference behaviour)")
    ap.add_argument("--map", action="store_true", help="candidate positions from the device k-mer mapper (reference behaviour)")
    ap.add_argument("--band", type=int, default=None, help="override the config's band (wide-band diagnostics)")
    ap.add_argument("--hap-len", type=int, default=None, help="override the haplotype length")
    ap.add_argument("--read-lens", default=None, help="override the read lengths, comma separated")
    ap.add_argument("--int-scores", action="store_true", help="HaplotypeLikelihoodModel::Config::use_int_scores (reference: int32 lanes)")
    return ap.parse_args()


METRIC = "pair-HMM GCUPS (DP cell-updates/s)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def issue_roofline(kernel_gcups, reads, band, clocks):
    lens = np.diff(np.asarray(reads.off))
    mean_ratio = float(((lens + band) / lens).mean())
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * mhz * 1e6 * (64.0 / 3.0) * mean_ratio / 1e9
    return {"bound": "alu-pipe (DPX packed-16)", "achieved": kernel_gcups, "peak": peak, "unit": "GCUPS", "frac": kernel_gcups / peak,
            "alu_instr_per_cell_pair": 6, "alu_thread_instr_per_clk_per_sm": 64, "sm_mhz": mhz}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads the CPU arm may use: the cgroup CPU quota when there is one (a 128-thread box may grant a pod only a
    few CPUs' worth of time), else the online CPU count."""
    n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def calibrated_sample(haps, reads, band, threads, target_s):
    """Number of reads whose (reads x all haplotypes) CPU pass takes about target_s seconds (measured on a small probe)."""
    probe = min(reads.n, max(64, 16 * threads))
    cpu_reference_run(haps, reads, band, probe, threads)
    g, dt, _, _, _ = cpu_reference_run(haps, reads, band, probe, threads)
    per_read = max(dt / probe, 1e-7)
    return int(max(probe, min(reads.n, target_s / per_read)))


def reference_batch_dict(haps, reads):
    return dict(read_bases=reads.bases, read_quals=reads.quals, read_off=reads.off, hap_seq=haps.seq,
                hap_mask_fwd=haps.snv_mask_fwd, hap_prior_fwd=haps.snv_prior_fwd, hap_mask_rev=haps.snv_mask_rev,
                hap_prior_rev=haps.snv_prior_rev, hap_gap_open=haps.gap_open, hap_gap_extend=haps.gap_extend, hap_off=haps.off)


def cpu_reference_run(haps, reads, band, n_sample_reads, threads):
    """Time the reference's own SIMD kernel (oracle/_ref; AVX2 build = what BASELINE names) — or, where that build is
    absent, the C port — on the first n_sample_reads reads x all haplotypes, one mapping position per pair.
    Returns (gcups, seconds, kind, isa_name, cells)."""
    from oracle.oracle import COracle, RefKernel, available_ref_isas
    n = min(n_sample_reads, reads.n)
    H = haps.n
    lens = np.diff(reads.off[:n + 1])
    cells = int((2 * (lens + band) * band).sum()) * H
    isas = available_ref_isas()
    if isas:
        isa = isas[0]   # the widest build this host runs == what the reference's -march=native build would select (AVX2 for band 16)
        k = RefKernel(isa)
        batch = reference_batch_dict(haps, reads)
        ridx = np.repeat(np.arange(n, dtype=np.int32), H)
        hidx = np.tile(np.arange(H, dtype=np.int32), n)
        woff = np.repeat((reads.begin[:n] - band).astype(np.int32), H)
        rev = np.repeat(reads.reverse[:n].astype(bool), H)
        parts = []
        for strand in (False, True):
            sel = rev == strand
            if sel.any():
                parts.append((strand, np.ascontiguousarray(ridx[sel]), np.ascontiguousarray(hidx[sel]), np.ascontiguousarray(woff[sel])))
        t0 = time.perf_counter()
        for strand, a, b, c in parts:
            k.align_batch(band, batch, a, b, c, nuc_prior=2, nthreads=threads, strand_rev=strand)
        dt = time.perf_counter() - t0
        return cells / dt / 1e9, dt, "reference", "%s<%d,short> (%s build)" % (k.name(band), band, isa), cells
    from octopus_b200.batch import ReadBlock
    a = int(reads.off[n])
    sub = ReadBlock(reads.off[:n + 1], reads.bases[:a], reads.quals[:a], reads.mapq[:n], reads.reverse[:n], reads.begin[:n])
    o = COracle()
    t0 = time.perf_counter()
    o.populate(band, haps, sub, dp_only=True)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, dt, "port", "scalar C restatement", cells


def workload_name(config, R, lens, H, hap_len, band):
    return ("%s per GPU: %d reads (L=%s) x %d haplotypes (%d bp), band=%d, one mapping position per pair, "
            "naive shortcut disabled (every pair runs the DP), mapq mixing on, double [H][R] out" % (config, R, lens, H, hap_len, band))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from octopus_b200 import synth
    cfg = synth.CONFIGS[args.config]
    threads = host_threads()
    # a bounded sample of the workload per step: sized so that the whole --steps/--warmup run takes about a minute
    haps, reads, band = synth.make_batch(args.config, n_reads=min(cfg["n_reads"], 200_000), n_haps=args.haps)
    n_sample = args.cpu_sample_reads or calibrated_sample(haps, reads, band, threads, 60.0 / max(1, args.steps + args.warmup))
    for _ in range(max(1, args.warmup)):
        cpu_reference_run(haps, reads, band, n_sample, threads)
    t_total, cells_total, info = 0.0, 0, None
    for _ in range(args.steps):
        g, dt, kind, name, cells = cpu_reference_run(haps, reads, band, n_sample, threads)
        t_total += dt; cells_total += cells; info = (kind, name)
    value = cells_total / t_total / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            # the same workload as the GPU arm (same generator, shapes, band and mode); each step times a bounded sample of it
            "config": {"workload": workload_name(args.config, args.reads or cfg["n_reads"], "/".join(map(str, cfg["read_lens"])), haps.n, cfg["hap_len"], band),
                       "sample_reads_per_step": n_sample,
                       "mode": {"flank_state": None, "naive_shortcut": False, "kmer_mapper": False}},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": info[0], "sample": "%d reads x %d haplotypes per step; %s" % (n_sample, haps.n, info[1])},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = synth.CONFIGS[args.config]
    # weak scaling: every rank owns its own batch of the named shape (its own regions' reads), haplotypes replicated
    read_lens = tuple(int(x) for x in args.read_lens.split(",")) if args.read_lens else None
    haps, reads, band = synth.make_batch(args.config, n_reads=args.reads, n_haps=args.haps, seed=cfg["seed"] + 1000 * rank,
                                         band=args.band, hap_len=args.hap_len, read_lens=read_lens)
    H, R = haps.n, reads.n
    cells = synth.total_cells(haps, reads, band)
    model_cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, disable_naive_shortcut=not args.shortcut, map_positions=args.map,
                                                use_int_scores=args.int_scores)
    flank_state = tuple(int(x) for x in args.flank.split(",")) if args.flank else None
    eng = PairHMMEngine(local)
    d_haps, d_reads = haps.to_device(dev), reads.to_device(dev)
    d_out = torch.empty((H, R), dtype=torch.float64, device=dev)

    recv = [torch.empty_like(d_out) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step():
        eng.populate(model_cfg, d_haps, d_reads, flank_state=flank_state, out=d_out)
        if world > 1:
            return shard.gather_slabs(d_out, world, rank, recv=recv)        # NCCL gather of the per-rank matrices to rank 0
        return d_out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp_ms, launches = [], 0
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step()
        dp_ms.append(eng.last_dp_kernel_ms())
        launches += eng.launch_count()
    ev1.record()
    sync_all()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # e2e: pinned host buffers through the C ABI, H2D + D2H inside the timed region
    p_haps, p_reads = haps.pin(), reads.pin()
    p_out_t = torch.empty((H, R), dtype=torch.float64).pin_memory()
    p_out = p_out_t.numpy()
    for _ in range(2):
        eng.populate(model_cfg, p_haps, p_reads, flank_state=flank_state, out=p_out)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.populate(model_cfg, p_haps, p_reads, flank_state=flank_state, out=p_out)
    e1.record()
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = float(ems.item())
    h2d = sum(int(a.nbytes) for a in haps.arrays().values()) + sum(int(a.nbytes) for a in reads.arrays().values())
    d2h = H * R * 8

    if rank == 0:
        value = cells * world * args.steps / (total_ms / 1e3) / 1e9
        e2e = cells * world * args.steps / (e2e_ms / 1e3) / 1e9
        kernel_ms = float(np.mean(dp_ms))
        # algorithmic HBM bytes of the dominant kernel (k_populate_fast) per launch, SURVEY.md §8(d):
        # 4 B per pair (integer score out) + the read row half-words (2 B per read base, read once per read pair)
        # + both strand column tables (16 B per haplotype base)
        alg_bytes = 4 * H * R + 2 * int(reads.off[-1]) + 16 * int(haps.off[-1])
        peak, peak_src = peaks()
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the DP kernel per step, from the committed ncu --set full captures
        # (profiles/r01f_populate_fast_c3_ncu_raw.csv: C3's single launch — 0.51 GB task words + 0.51 GB best[] read for the
        # atomicMin + tables / rows in, 0.57 GB best[] out; profiles/r01c_*: C2), default sizes only
        traffic = {("C3", 1_000_000, 128): 1.401e9 + 0.569e9, ("C2", 100_000, 64): 84.1e6 + 3.0e6}.get((args.config, R, H))
        line = {
            "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": workload_name(args.config, R, (args.read_lens or "/".join(map(str, cfg["read_lens"]))).replace(",", "/"), H, args.hap_len or cfg["hap_len"], band),
                       "alignments_per_step": H * R * world, "cells_per_step": cells * world,
                       "l2": "inputs+outputs (%.0f MB) larger than the 126 MB L2" % ((h2d + d2h + 4 * H * R) / 1e6),
                       "parallelism": "reads sharded over %d rank(s), haplotypes replicated, NCCL gather to rank 0" % world,
                       "mode": {"flank_state": flank_state, "naive_shortcut": bool(args.shortcut), "kmer_mapper": bool(args.map)}},
            "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "k_populate_fast<%d>" % band, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "the path is integer-issue bound, not HBM bound (SURVEY.md F3); see DESIGN.md for the issue-rate roofline",
                         "kernel_gcups": cells / (kernel_ms / 1e3) / 1e9},
            # The binding roofline: the packed cell costs 6 ALU-pipe instructions per 2 cells and the ALU pipe
            # issues 64 thread-instructions/clk/SM (profiles/r01_ubench_int.txt). In the reference's cell count 2(L+B)B per
            # alignment (the column sweep itself touches 2LB cells) the peak is 148 SMs x clock x 64/3 x (L+B)/L.
            "issue_roofline": issue_roofline(cells / (kernel_ms / 1e3) / 1e9, reads, band, clocks),
        }
        if not args.no_cpu_baseline:
            threads = host_threads()
            n_sample = args.cpu_sample_reads or calibrated_sample(haps, reads, band, threads, 12.0)
            g, dt, kind, name, _ = cpu_reference_run(haps, reads, band, n_sample, threads)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind,
                                    "sample": "first %d reads x %d haplotypes of the same batch, %.1f s; %s" % (n_sample, H, dt, name)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
