"""Small batches in every mode, meant to run under `compute-sanitizer --tool memcheck` (and racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import parity_checks as pc
from conftest import GOLDEN
from nanosim_b200 import _lib as L
from nanosim_b200.reference_fasta import PackedReference, MetaReference, read_abundance, read_expression, read_polya_list, POLYA_SCALE

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
mini = PackedReference.from_fasta(os.path.join(GOLDEN, "mini_ref.fa"))
circ = PackedReference.from_fasta(os.path.join(GOLDEN, "mini_circular.fa"))
runs = [("guppy fastq", lambda: pc.make_engine("guppy", mini, fastq=True, seed=3)[0]),
        ("guppy fasta", lambda: pc.make_engine("guppy", mini, fastq=False, seed=4)[0]),
        ("dorado chimeric hp6", lambda: pc.make_engine("dorado", mini, fastq=True, chimeric=True, kmer_bias=6, seed=5)[0]),
        ("dorado circular", lambda: pc.make_engine("dorado", circ, fastq=True, seed=6, circular=True, max_len=4000)[0]),
        ("guppy perfect", lambda: pc.make_engine("guppy", mini, fastq=True, perfect=True, seed=7)[0]),
        ("guppy scripted unaligned", lambda: pc.make_engine("guppy", mini, fastq=True, seed=8, unaligned_scripts=True)[0])]
for name, mk in runs:
    eng = mk()
    for kind in (L.NS_KIND_ALIGNED, L.NS_KIND_UNALIGNED):
        if "perfect" in name and kind == L.NS_KIND_UNALIGNED:
            continue
        info = eng.simulate(kind, 10, n)
        b = eng.fetch(want_ops=True)
        print(name, "kind", kind, "bases", info.total_bases, "ops", info.n_ops, flush=True)
    eng.close()
T = os.path.join(GOLDEN, "trx")
ref = PackedReference.from_fasta(os.path.join(T, "transcripts.fa"))
chrom, w = read_expression(os.path.join(T, "expression.tsv"), ref)
eng = pc.make_trx_engine(ref, chrom, w, read_polya_list(os.path.join(T, "polya.txt"), ref), fastq=True, seed=9, polya_scale=POLYA_SCALE["guppy"])[0]
for kind in (L.NS_KIND_ALIGNED, L.NS_KIND_UNALIGNED):
    info = eng.simulate(kind, 0, n)
    eng.fetch(want_ops=True)
    print("transcriptome kind", kind, "bases", info.total_bases, flush=True)
eng.close()
from conftest import meta_fixture
meta_fixture()
M = os.path.join(GOLDEN, "meta")
mref = MetaReference.from_genome_list(os.path.join(M, "genome_list_local.tsv"), os.path.join(M, "dna_type.tsv"))
numbers, samples = read_abundance(os.path.join(M, "abundance.tsv"), mref.species)
eng = pc.make_meta_engine(mref, samples[0], fastq=True, chimeric=True, seed=10)[0] if hasattr(pc, "make_meta_engine") else None
if eng is not None:
    for kind in (L.NS_KIND_ALIGNED, L.NS_KIND_UNALIGNED):
        info = eng.simulate(kind, 0, n)
        eng.fetch(want_ops=True)
        print("metagenome kind", kind, "bases", info.total_bases, flush=True)
    eng.close()
print("sanitize run complete")
