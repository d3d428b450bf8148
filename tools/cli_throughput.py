"""Wall-clock of the CLI end to end (model load, simulation, formatting, file writes) on the 5 Mb synthetic reference."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from nanosim_b200 import simulator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
extra = sys.argv[2:]
tmp = tempfile.mkdtemp(prefix="cli_tp_")
ref = os.path.join(tmp, "ecoli5m.fa")
synth.ecoli5m(ref)
t0 = time.time()
simulator.main(["genome", "-rg", ref, "-c", os.path.join(ROOT, "nanosim_b200", "data", "guppy_fab49712_plusq.npz"), "-n", str(n),
                "-o", os.path.join(tmp, "sim"), "--fastq", "-t", "32", "--seed", "1"] + extra)
dt = time.time() - t0
sz = {f: os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp) if f.startswith("sim")}
bases = (sz.get("sim_aligned_reads.fastq", 0) + sz.get("sim_unaligned_reads.fastq", 0)) / 2
print("CLI %d reads %s: %.1f s wall, ~%.2f Gbases, %.2f Gbases/s; files %s" % (n, extra, dt, bases / 1e9, bases / 1e9 / dt, {k: round(v / 1e9, 2) for k, v in sz.items()}))
