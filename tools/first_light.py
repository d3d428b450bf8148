"""First-light GPU run: build engine, simulate, verify edit scripts bit-exactly, print rates vs the model's training rates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import parity_checks as pc
import synth
from nanosim_b200 import _lib as L
from nanosim_b200.reference_fasta import PackedReference

t0 = time.time()
ref = PackedReference.from_records(synth.ecoli5m())
print("ref built", time.time() - t0, flush=True)
for model, fastq, chim in (("guppy", True, False), ("guppy", False, False), ("dorado", True, True)):
    eng, cm, t = pc.make_engine(model, ref, fastq=fastq, chimeric=chim, seed=11)
    for n in (2000, 50000, 50000, 262144, 262144):
        t1 = time.time()
        info = eng.simulate(L.NS_KIND_ALIGNED, 0, n)
        t2 = time.time()
        b = eng.fetch(want_ops=(n == 2000))
        t3 = time.time()
        print(model, "fastq" if fastq else "fasta", "chim" if chim else "", "n", n, "bases", info.total_bases, "ops", info.n_ops, "pieces", info.n_pieces,
              "ms total %.2f setup %.2f plan %.2f scan %.2f script %.2f emit %.2f" % (info.ms_total, info.ms_setup, info.ms_plan, info.ms_scan, info.ms_script, info.ms_emit),
              "wall sim %.3f fetch %.3f" % (t2 - t1, t3 - t2), "Gbases/s kernels %.2f" % (info.total_bases / info.ms_total / 1e6), flush=True)
        if n == 2000:
            nb = pc.check_edit_scripts(b, ref, fastq)
            s = pc.batch_stats(b, ref, fastq)
            print("  verified", nb, "bases; rates", pc.rates(s), "attempts max", int(b.reads["attempts"].max()),
                  "mean len", s["aligned_bases"] / s["n_aligned"], "strand R frac", s["strand_R_aligned"] / s["n_aligned"], flush=True)
            print("  training rates:", cm.text.get("error_rate.tsv", "").replace("\n", " | "))
            st = eng.op_stats()
            print("  op_stats events", st["events"], "host events", s["events"])
    info = eng.simulate(L.NS_KIND_UNALIGNED, 0, 5000)
    info = eng.simulate(L.NS_KIND_UNALIGNED, 0, 30000)
    info = eng.simulate(L.NS_KIND_UNALIGNED, 0, 30000)
    b = eng.fetch(want_ops=True)
    nb = pc.check_edit_scripts(b, ref, fastq, max_reads=500)
    print("  unaligned: bases", info.total_bases, "ops", info.n_ops, "ms %.2f plan %.2f script %.2f emit %.2f" % (info.ms_total, info.ms_plan, info.ms_script, info.ms_emit),
          "verified", nb, "attempts max", int(b.reads["attempts"].max()), "mean len", info.total_bases / info.n_reads, flush=True)
    eng.close()
print("FIRST LIGHT OK")
