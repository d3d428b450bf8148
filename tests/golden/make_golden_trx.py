"""Golden vectors for TRANSCRIPTOME mode (--no_model_ir) from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_trx.py /tmp/models

Writes the fixture tests/golden/trx/ (80 transcripts, expression profile, polyA list) and tests/golden/vectors_trx.json
with seeded outputs of make_cdf, select_nearest_kde2d, extract_read_trx, extract_read("transcriptome") and whole
simulation_aligned_transcriptome / simulation_unaligned loops (dRNA Bham1 guppy model + dorado quality table)."""
import hashlib
import json
import multiprocessing as mp
import os
import random
import shutil
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
TRX = os.path.join(HERE, "trx")


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def make_fixture():
    os.makedirs(TRX, exist_ok=True)
    rng = np.random.default_rng(31)
    with open(os.path.join(TRX, "transcripts.fa"), "w") as f, open(os.path.join(TRX, "expression.tsv"), "w") as e, \
            open(os.path.join(TRX, "polya.txt"), "w") as pa:
        e.write("target_id\test_counts\ttpm\n")
        for i in range(80):
            n = int(np.clip(rng.lognormal(7.3, 0.6), 300, 9000))
            tid = "ENST%011d.%d" % (1000 + i, 1 + i % 3)
            s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes().decode()
            f.write(">" + tid + " gene=G%d\n" % i)
            for j in range(0, n, 60):
                f.write(s[j:j + 60] + "\n")
            tpm = 0.0 if i % 17 == 5 else float(np.round(rng.lognormal(2, 1.5), 4))
            if i % 23 != 7:                                  # a few transcripts are missing from the expression profile
                e.write("%s\t%.2f\t%s\n" % (tid, tpm * 3, tpm))
            if i % 2 == 0:
                pa.write(tid + "\n")
        e.write("ENST99999999999.1\t5.0\t12.5\n")            # expressed but not in the reference


def main(models_dir):
    make_fixture()
    aug = tempfile.mkdtemp(prefix="drna_aug_")
    base = os.path.join(models_dir, DRNA)
    for fn in os.listdir(base):
        if not fn.endswith(".gff3"):
            os.symlink(os.path.join(base, fn), os.path.join(aug, fn))
    os.symlink(os.path.join(models_dir, DORADO, "training_base_qualities_model_parameters.tsv"),
               os.path.join(aug, "training_base_qualities_model_parameters.tsv"))
    prefix = os.path.join(aug, "training")
    sim = ref_shim.load_reference_module()
    devnull = open(os.devnull, "w")
    stdout = sys.stdout

    def profile(per=False, polya=True):
        sys.stdout = devnull
        sim.read_profile("", [1000], prefix, per, "transcriptome", None, ref_t=os.path.join(TRX, "transcripts.fa"),
                         dna_type="linear", model_ir=False, polya=os.path.join(TRX, "polya.txt") if polya else None,
                         exp=os.path.join(TRX, "expression.tsv"), homopolymer=False, fastq=True)
        sys.stdout = stdout

    out = {}
    profile()
    out["ecdf"] = {"n": len(sim.ecdf_length_list), "head": [[a, b] for a, b in sim.ecdf_length_list[:5]],
                   "weights_head": sim.ecdf_weight_list[:5], "weight_sum": float(sum(sim.ecdf_weight_list)),
                   "tail": [[a, b] for a, b in sim.ecdf_length_list[-3:]]}
    out["numbers"] = {"aligned": sim.number_aligned_l, "unaligned": sim.number_unaligned_l}
    seed_all(900)
    sample = sim.get_length_kde(sim.kde_aligned_2d, 400, False, False)
    out["nearest"] = [[int(L), int(sim.select_nearest_kde2d(sample, L))] for L in (350, 900, 1500, 2600, 5000, 8000)]
    ex = []
    for s, key, ln, pa in ((1, "ENST00000001000", 200, True), (2, "ENST00000001002", 250, False), (3, "ENST00000001004", 280, True)):
        seed_all(910 + s)
        ln = min(ln, sim.seq_len[key] - 1)
        seq, pos, retain = sim.extract_read_trx(key, ln, pa)
        ex.append({"seed": 910 + s, "key": key, "length": ln, "polya": pa, "pos": pos, "retain": bool(retain), "md5": md5(seq)})
    for s, ln in ((4, 400), (5, 3000)):
        seed_all(910 + s)
        seq, name = sim.extract_read("transcriptome", ln)
        ex.append({"seed": 910 + s, "length": ln, "name": name, "md5": md5(seq)})
    out["extract"] = ex
    runs = []
    tmp = tempfile.mkdtemp(prefix="golden_trx_")
    cfgs = [dict(fastq=True, per=False, polya=True, basecaller="guppy", uracil=False, n=80),
            dict(fastq=False, per=False, polya=False, basecaller=None, uracil=True, n=60),
            dict(fastq=True, per=True, polya=True, basecaller="albacore", uracil=False, n=40)]
    for i, cf in enumerate(cfgs):
        profile(cf["per"], cf["polya"])
        sim.total_simulated = mp.Value("i", 0, lock=True)
        seed_all(940 + i)
        o_reads, o_err = os.path.join(tmp, "r%d" % i), os.path.join(tmp, "e%d" % i)
        sys.stdout = devnull
        sim.simulation_aligned_transcriptome(False, o_reads, o_err, None, cf["basecaller"], cf["n"],
                                             os.path.join(TRX, "polya.txt") if cf["polya"] else None, cf["fastq"], cf["per"],
                                             cf["uracil"])
        sys.stdout = stdout
        reads, err = open(o_reads).read(), open(o_err).read()
        runs.append({"cfg": cf, "seed": 940 + i, "reads_md5": md5(reads), "err_md5": md5(err), "n_lines": reads.count("\n"),
                     "first_header": reads.split("\n")[0]})
    profile()
    for i, fq in enumerate((False, True)):
        sim.total_simulated = mp.Value("i", 0, lock=True)
        seed_all(960 + i)
        o_reads = os.path.join(tmp, "u%d" % i)
        sys.stdout = devnull
        sim.simulation_unaligned("transcriptome", 50, sim.max_chrom, None, None, o_reads, fq, 20, False)
        sys.stdout = stdout
        reads = open(o_reads).read()
        runs.append({"cfg": dict(fastq=fq, n=20, unaligned=True), "seed": 960 + i, "reads_md5": md5(reads),
                     "n_lines": reads.count("\n"), "first_header": reads.split("\n")[0], "max_l": int(sim.max_chrom)})
    shutil.rmtree(tmp)
    shutil.rmtree(aug)
    out["runs"] = runs
    with open(os.path.join(HERE, "vectors_trx.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote vectors_trx.json", os.path.getsize(os.path.join(HERE, "vectors_trx.json")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/models")
