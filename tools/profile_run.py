"""Small fixed workload for ncu captures: guppy FASTQ on the 5 Mb synthetic reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import parity_checks as pc
import synth
from nanosim_b200 import _lib as L
from nanosim_b200.reference_fasta import PackedReference

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
ref = PackedReference.from_records(synth.ecoli5m())
eng, cm, t = pc.make_engine("guppy", ref, fastq=True, seed=11)
for rep in range(2):
    info = eng.simulate(L.NS_KIND_ALIGNED, 0, n)
    b = eng.fetch()
    print("aligned", info.total_bases, "setup %.2f plan %.2f scan %.2f script %.2f emit %.2f total %.2f" % (info.ms_setup, info.ms_plan, info.ms_scan, info.ms_script, info.ms_emit, info.ms_total),
          "flagged", int((b.reads["flags"] & 1).sum()), "ops used", int(b.pieces["n_ops"].sum()), "of", info.n_ops)
    if nu:
        info = eng.simulate(L.NS_KIND_UNALIGNED, 0, nu)
        bu = eng.fetch()
        print("unaligned", info.total_bases, "plan %.2f scan %.2f script %.2f emit %.2f total %.2f" % (info.ms_plan, info.ms_scan, info.ms_script, info.ms_emit, info.ms_total),
              "flagged", int((bu.reads["flags"] & 1).sum()), "attempts>0", int((bu.reads["attempts"] > 0).sum()), "max len", int(bu.reads["seq_len"].max()),
              "ops/base %.3f" % (bu.pieces["n_ops"].sum() / max(1, bu.pieces["ref_len"].sum())))
