"""Run-level golden histograms of the UNMODIFIED reference CLI in metagenome and transcriptome mode (build container
only; same protocol as make_golden_runs.py: chunks of ``-t 8`` runs through oracle/ref_shim.py, reduced by
tests/run_stats.py, accumulated into tests/golden/ref_stats_<config>.json).

    python tests/golden/make_golden_runs_modes.py meta_even_fastq_chimeric 200000 20000 [/tmp/models]
    python tests/golden/make_golden_runs_modes.py trx_drna_fasta 200000 8000 [/tmp/models]
    python tests/golden/make_golden_runs_modes.py trx_ir_drna_fasta 100000 8000 [/tmp/models]     (intron retention on; HTSeq /
                                                  pysam calls served by the stand-ins of oracle/ref_shim.py, fixture tests/golden/ir)

Besides the run_stats histograms the JSON holds ``by_chrom_reads`` / ``by_chrom_bases``: aligned reads and their emitted
bases per reference record of the read's FIRST segment (species-chromosome, or transcript).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import run_stats as rs  # noqa: E402

EVEN = "metagenome_ERR3152364_Even_v3.2.2"
DRNA = "human_NA12878_dRNA_Bham1_guppy"


def by_chrom(path, fastq):
    reads, bases = {}, {}
    for name, seq, _ in rs._records(path, fastq):
        first = name.split(";")[0] if ";" in name.split("_aligned_")[0] else name.split("_aligned_")[0]
        key = first.rsplit("_", 1)[0]
        reads[key] = reads.get(key, 0) + 1
        bases[key] = bases.get(key, 0) + len(seq)
    return reads, bases


def main():
    cfg, total, chunk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    models = sys.argv[4] if len(sys.argv) > 4 else "/tmp/models"
    work = tempfile.mkdtemp(prefix="golden_mode_", dir="/tmp")
    shim = os.path.join(ROOT, "oracle", "ref_shim.py")
    out = os.path.join(work, "sim")
    if cfg == "meta_even_fastq_chimeric":
        M = os.path.join(HERE, "meta")
        gl = os.path.join(work, "gl.tsv")
        with open(os.path.join(M, "genome_list.tsv")) as f, open(gl, "w") as o:
            for line in f:
                sp, path = line.rstrip("\n").split("\t")
                o.write("%s\t%s\n" % (sp, os.path.join(M, os.path.basename(path))))
        ab = os.path.join(work, "abun.tsv")
        with open(os.path.join(M, "abundance.tsv")) as f, open(ab, "w") as o:      # first sample's abundances, `chunk` reads
            f.readline()
            o.write("Size\t%d\n" % chunk)
            for line in f:
                p = line.rstrip("\n").split("\t")
                o.write("%s\t%s\n" % (p[0], p[1]))
        cmd = [sys.executable, shim, "metagenome", "-gl", gl, "-a", ab, "-dl", os.path.join(M, "dna_type.tsv"),
               "-c", os.path.join(models, EVEN, "training"), "-o", out, "--fastq", "--chimeric", "-t", "8"]
        fastq, prefix = True, out + "_sample0"
        shown = "simulator.py metagenome -gl genome_list.tsv -a <Size=%d, sample 0 of abundance.tsv> -dl dna_type.tsv " \
                "-c %s/training --fastq --chimeric -t 8" % (chunk, EVEN)
    elif cfg == "trx_drna_fasta":
        T = os.path.join(HERE, "trx")
        cmd = [sys.executable, shim, "transcriptome", "-rt", os.path.join(T, "transcripts.fa"), "-e", os.path.join(T, "expression.tsv"),
               "-c", os.path.join(models, DRNA, "training"), "-o", out, "-n", str(chunk), "--no_model_ir", "-b", "guppy",
               "--polya", os.path.join(T, "polya.txt"), "-t", "8"]
        fastq, prefix = False, out
        shown = "simulator.py transcriptome -rt trx/transcripts.fa -e trx/expression.tsv -c %s/training -n %d --no_model_ir " \
                "-b guppy --polya trx/polya.txt -t 8" % (DRNA, chunk)
    elif cfg == "trx_ir_drna_fasta":
        sys.path.insert(0, HERE)
        from make_golden_ir import IR, model_dir
        mp_ = model_dir(models)
        cmd = [sys.executable, shim, "transcriptome", "-rt", os.path.join(IR, "transcripts.fa"), "-rg", os.path.join(IR, "genome.fa"),
               "-e", os.path.join(IR, "expression.tsv"), "-c", mp_, "-o", out, "-n", str(chunk), "-b", "guppy",
               "--polya", os.path.join(IR, "polya.txt"), "-t", "8"]
        fastq, prefix = False, out
        shown = "simulator.py transcriptome -rt ir/transcripts.fa -rg ir/genome.fa -e ir/expression.tsv -c <%s + ir/IR_markov_model + " \
                "ir/annotation.gff3>/training -n %d -b guppy --polya ir/polya.txt -t 8" % (DRNA, chunk)
    else:
        raise SystemExit("unknown config " + cfg)
    out_json = os.path.join(HERE, "ref_stats_%s.json" % cfg)
    if os.path.exists(out_json):
        acc, meta = rs.load(out_json)
    else:
        acc, meta = rs.empty(), {"config": cfg, "cmd": shown, "chunk_reads": chunk, "chunks": []}
        acc["by_chrom_reads"], acc["by_chrom_bases"] = {}, {}
    done = acc["n_aligned"] + acc["n_unaligned"]
    while done < total:
        t0 = time.time()
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        t1 = time.time()
        s = rs.stats_from_prefix(prefix, fastq)
        r, b = by_chrom(prefix + "_aligned_reads." + ("fastq" if fastq else "fasta"), fastq)
        if cfg.startswith("trx_ir"):
            n_ir = n_iv = 0
            for name, _, _ in rs._records(prefix + "_aligned_reads." + ("fastq" if fastq else "fasta"), fastq):
                if "_RetainedIntron_" in name:
                    n_ir += 1
                    n_iv += name.split("_RetainedIntron_")[1].split("_")[0].count(";")
            acc["ir_reads"] = acc.get("ir_reads", 0) + n_ir
            acc["ir_intervals"] = acc.get("ir_intervals", 0) + n_iv
        for k in r:
            acc["by_chrom_reads"][k] = acc["by_chrom_reads"].get(k, 0) + r[k]
            acc["by_chrom_bases"][k] = acc["by_chrom_bases"].get(k, 0) + b[k]
        for fn in os.listdir(work):
            if fn.startswith("sim"):
                os.remove(os.path.join(work, fn))
        meta["chunks"].append({"n_aligned": int(s["n_aligned"]), "ref_bases": int(s["ref_bases"]),
                               "event_bases": {k: int(v) for k, v in s["event_bases"].items()}, "sim_seconds": round(t1 - t0, 1)})
        rs.merge(acc, s)
        rs.save(acc, out_json, meta)
        done = acc["n_aligned"] + acc["n_unaligned"]
        print("%s: %d/%d reads (chunk sim %.0fs, parse %.0fs)" % (cfg, done, total, t1 - t0, time.time() - t1), flush=True)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
