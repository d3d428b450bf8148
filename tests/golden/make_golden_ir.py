"""Golden vectors for the INTRON-RETENTION branch of transcriptome mode (simulator.py:114-191, 404-453, 1156-1192) from
the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_ir.py /tmp/models

HTSeq and pysam are absent here; oracle/ref_shim.py stands in for the three calls the branch makes
(GFF_Reader, GenomicInterval, Fastafile.fetch) following their documented behaviour, everything else -- update_structure,
extract_read_pos, the IR part of simulation_aligned_transcriptome -- is the reference's own code.

Writes the fixture tests/golden/ir/ (2-chromosome genome, GFF3 with exon/intron features, spliced transcripts, expression
profile, polyA list, IR Markov model) and tests/golden/vectors_ir.json."""
import hashlib
import json
import multiprocessing as mp
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

DRNA = "human_NA12878_dRNA_Bham1_guppy"
DORADO = "human_giab_hg002_sub1M_kitv14_dorado_v3.2.1"
IR = os.path.join(HERE, "ir")
COMP = str.maketrans("ACGT", "TGCA")


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def make_fixture():
    os.makedirs(IR, exist_ok=True)
    rng = np.random.default_rng(47)
    genome = {"chr1": 36000, "chr2": 24000}
    seqs = {k: np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes().decode() for k, n in genome.items()}
    with open(os.path.join(IR, "genome.fa"), "w") as f:
        for k, s in seqs.items():
            f.write(">%s test chromosome\n" % k)
            for j in range(0, len(s), 70):
                f.write(s[j:j + 70] + "\n")
    gff = ["##gff-version 3"]
    trx = []
    cursor = {"chr1": 500, "chr2": 300}
    for i in range(16):
        chrom = "chr1" if i % 3 else "chr2"
        if i == 13:
            chrom = "chrUn"                                   # annotated on a chromosome the genome file does not have
        strand = "+" if i % 2 == 0 else "-"
        n_exon = int(rng.integers(2, 7)) if i != 5 else 1     # one single-exon transcript (no intron)
        tid = "ENST%011d.%d" % (2000 + i, 1 + i % 2)
        start = cursor.get(chrom, 1000) + int(rng.integers(50, 400))
        feats, pos, exon_seqs = [], start, []
        for e in range(n_exon):
            el = int(rng.integers(90, 700))
            feats.append(("exon", pos, pos + el - 1))
            if chrom in seqs:
                exon_seqs.append(seqs[chrom][pos - 1:pos - 1 + el])
            else:
                exon_seqs.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, el)].tobytes().decode())
            pos += el
            if e + 1 < n_exon:
                il = int(rng.integers(70, 500))
                feats.append(("intron", pos, pos + il - 1))
                pos += il
        if chrom in cursor:
            cursor[chrom] = pos
        gff.append("%s\tTEST\ttranscript\t%d\t%d\t.\t%s\t.\tID=transcript%d;Parent=gene%d;transcript_id=%s" % (chrom, start, pos - 1, strand, i, i, tid))
        for ty, a, b in feats:
            src = "TEST" if ty == "exon" else "."
            extra = ";exon_number=1" if ty == "exon" else ""
            gff.append("%s\t%s\t%s\t%d\t%d\t.\t%s\t.\tParent=transcript%d;transcript_id=%s%s" % (chrom, src, ty, a, b, strand, i, tid, extra))
        spliced = "".join(exon_seqs)
        if strand == "-":
            spliced = spliced.translate(COMP)[::-1]
        trx.append((tid, spliced))
    gff.append("chr1\tTEST\tgene\t100\t400\t.\t+\t.\tID=gene99;gene_id=G99")     # a feature type the loader ignores
    with open(os.path.join(IR, "annotation.gff3"), "w") as f:
        f.write("\n".join(gff) + "\n")
    with open(os.path.join(IR, "transcripts.fa"), "w") as f, open(os.path.join(IR, "expression.tsv"), "w") as e, \
            open(os.path.join(IR, "polya.txt"), "w") as pa:
        e.write("target_id\test_counts\ttpm\n")
        for i, (tid, s) in enumerate(trx):
            f.write(">" + tid + " gene=G%d\n" % i)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + "\n")
            e.write("%s\t%.2f\t%s\n" % (tid, 10.0, float(np.round(rng.lognormal(2, 1.0), 4))))
            if i % 2 == 0:
                pa.write(tid + "\n")
    with open(os.path.join(IR, "IR_markov_model"), "w") as f:
        f.write("succedent\tno_IR\tIR\nstart\t0.6\t0.4\nno_IR\t0.7\t0.3\nIR\t0.5\t0.5\n")


def model_dir(models_dir):
    aug = tempfile.mkdtemp(prefix="drna_ir_")
    base = os.path.join(models_dir, DRNA)
    for fn in os.listdir(base):
        if not fn.endswith(".gff3") and "IR_markov" not in fn:
            os.symlink(os.path.join(base, fn), os.path.join(aug, fn))
    os.symlink(os.path.join(models_dir, DORADO, "training_base_qualities_model_parameters.tsv"),
               os.path.join(aug, "training_base_qualities_model_parameters.tsv"))
    os.symlink(os.path.join(IR, "IR_markov_model"), os.path.join(aug, "training_IR_markov_model"))
    os.symlink(os.path.join(IR, "annotation.gff3"), os.path.join(aug, "training_added_intron_final.gff3"))
    return os.path.join(aug, "training")


def main(models_dir):
    make_fixture()
    prefix = model_dir(models_dir)
    sim = ref_shim.load_reference_module()
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    sys.stdout = devnull
    sim.read_profile(os.path.join(IR, "genome.fa"), [1000], prefix, False, "transcriptome", None,
                     ref_t=os.path.join(IR, "transcripts.fa"), dna_type="linear", model_ir=True,
                     polya=os.path.join(IR, "polya.txt"), exp=os.path.join(IR, "expression.tsv"), homopolymer=False, fastq=True)
    sys.stdout = stdout
    out = {"structure": {k: [list(x) for x in v] for k, v in sorted(sim.dict_ref_structure.items())},
           "ir_model": {k: sorted([a, b, v2] for (a, b), v2 in v.items()) for k, v in sim.IR_markov_model.items()}}
    ups = []
    for s, key in enumerate(sorted(sim.dict_ref_structure)):
        seed_all(300 + s)
        flag, st = sim.update_structure(sim.dict_ref_structure[key], sim.IR_markov_model)
        rec = {"seed": 300 + s, "key": key, "flag": bool(flag), "types": [x[0] for x in st]}
        if flag and key in sim.seq_len:
            ref_len = sim.seq_len[key]
            length = max(1, ref_len // 3)
            ivs, retain, ir_list = sim.extract_read_pos(length, ref_len, st, key in sim.trx_with_polya)
            rec["extract"] = {"length": length, "ref_len": ref_len, "ivs": [[iv.chrom, iv.start, iv.end, iv.strand] for iv in ivs],
                              "retain": bool(retain), "ir_list": [list(x) for x in ir_list]}
        ups.append(rec)
    out["update_extract"] = ups
    runs = []
    tmp = tempfile.mkdtemp(prefix="golden_ir_")
    for i, cf in enumerate([dict(fastq=True, polya=True, basecaller="guppy", uracil=False, n=120),
                            dict(fastq=False, polya=False, basecaller=None, uracil=False, n=80)]):
        sim.total_simulated = mp.Value("i", 0, lock=True)
        seed_all(340 + i)
        o_reads, o_err = os.path.join(tmp, "r%d" % i), os.path.join(tmp, "e%d" % i)
        sys.stdout = devnull
        sim.simulation_aligned_transcriptome(True, o_reads, o_err, None, cf["basecaller"], cf["n"],
                                             os.path.join(IR, "polya.txt") if cf["polya"] else None, cf["fastq"], False, cf["uracil"])
        sys.stdout = stdout
        reads, err = open(o_reads).read(), open(o_err).read()
        heads = [l for l in reads.split("\n") if l[:1] in "@>" and "_aligned_" in l]
        runs.append({"cfg": cf, "seed": 340 + i, "reads_md5": md5(reads), "err_md5": md5(err), "n_lines": reads.count("\n"),
                     "n_ir": sum("_RetainedIntron_" in h for h in heads), "ir_headers": [h for h in heads if "_RetainedIntron_" in h][:6]})
    out["runs"] = runs
    with open(os.path.join(HERE, "vectors_ir.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote vectors_ir.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()},
          [(r["n_lines"], r["n_ir"]) for r in runs])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/models")
