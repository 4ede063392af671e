"""Diagnostic: host-buffer populate (pipelined path) timing on C3; PHMM_TRACE=1 prints the per-chunk host timeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine, synth

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C3"
haps, reads, band = synth.make_batch(cfgname)
H, R = haps.n, reads.n
cells = synth.total_cells(haps, reads, band)
cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, disable_naive_shortcut=True, map_positions=False)
eng = PairHMMEngine(0)
p_haps, p_reads = haps.pin(), reads.pin()
out = torch.empty((H, R), dtype=torch.float64).pin_memory().numpy()
for _ in range(2):
    eng.populate(cfg, p_haps, p_reads, out=out)
torch.cuda.synchronize()
ts = []
for i in range(4):
    if i == 3: sys.stderr.write("=== TRACED CALL ===\n"); sys.stderr.flush()
    t0 = time.perf_counter(); eng.populate(cfg, p_haps, p_reads, out=out); ts.append((time.perf_counter() - t0) * 1e3)
print("chunk_pairs=%s e2e ms: %s  -> %.0f GCUPS" % (os.environ.get("PHMM_CHUNK_PAIRS", "default"), ["%.1f" % t for t in ts], cells / (min(ts) / 1e3) / 1e9))
d_haps, d_reads = haps.to_device(torch.device("cuda", 0)), reads.to_device(torch.device("cuda", 0))
d_out = torch.empty((H, R), dtype=torch.float64, device="cuda")
for _ in range(2): eng.populate(cfg, d_haps, d_reads, out=d_out)
torch.cuda.synchronize(); t0 = time.perf_counter(); eng.populate(cfg, d_haps, d_reads, out=d_out); torch.cuda.synchronize()
print("device-resident ms: %.1f  dp kernel ms %.1f" % ((time.perf_counter() - t0) * 1e3, eng.last_dp_kernel_ms()))
