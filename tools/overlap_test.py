"""Does one context's device->host copy overlap another context's kernels?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import parity_checks as pc, synth
from nanosim_b200 import _lib as L
from nanosim_b200.reference_fasta import PackedReference

ref = PackedReference.from_records(synth.ecoli5m())
a, cm, t = pc.make_engine("guppy", ref, fastq=True, seed=11)
b = a.clone()
N = 200000
for e in (a, b):
    e.simulate(L.NS_KIND_ALIGNED, 0, N); e.simulate(L.NS_KIND_ALIGNED, 0, N)
info = a.info
cap = int(info.seq_bytes)
bufs = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
meta = torch.empty(64 * N, dtype=torch.uint8, pin_memory=True)

def fetch_loop(k):
    for _ in range(k):
        a.fetch_into(bufs[0].data_ptr(), bufs[1].data_ptr(), meta.data_ptr())
def sim_loop(k):
    for i in range(k):
        b.simulate(L.NS_KIND_ALIGNED, i * N, N)

def timed(fn, *args):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(*args); torch.cuda.synchronize(); return time.perf_counter() - t0

tf = timed(fetch_loop, 4); ts = timed(sim_loop, 8)
print("fetch alone: %.1f ms per batch (%.1f GB/s)   simulate alone: %.1f ms per batch" % (tf / 4 * 1e3, 2 * cap * 4 / tf / 1e9, ts / 8 * 1e3))
def both():
    th = [threading.Thread(target=fetch_loop, args=(4,)), threading.Thread(target=sim_loop, args=(8,))]
    [x.start() for x in th]; [x.join() for x in th]
tb = timed(both)
print("4 fetches + 8 simulates concurrently: %.1f ms (serial sum would be %.1f ms)" % (tb * 1e3, (tf + ts) * 1e3))
