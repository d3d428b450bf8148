"""Deterministic synthetic references: i.i.d. uniform ACGT from numpy default_rng(seed)."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_chrom(rng, n):
    return _ACGT[rng.integers(0, 4, n, dtype=np.uint8)]


def write_fasta(path, records, width=80):
    with open(path, "w") as f:
        for name, arr in records:
            f.write(">" + name + "\n")
            s = arr.tobytes().decode()
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def ecoli5m(path=None):
    """BASELINE config 1: one chromosome ``chr1`` of 5,000,000 i.i.d. bases, default_rng(0)."""
    rng = np.random.default_rng(0)
    rec = [("chr1", synth_chrom(rng, 5_000_000))]
    if path:
        write_fasta(path, rec)
    return rec


HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
HG38_NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]


def config3_transcriptome(n_transcripts=200000, seed=2):
    """BASELINE config 3: ``n_transcripts`` records ``ENST%011d.1`` of length
    clip(lognormal(7.3, 0.6), 300, 20000), i.i.d. ACGT, and an expression profile with tpm ~ lognormal(2, 1.5);
    default_rng(2).  Returns (fasta names, lengths int64[n], bases uint8[sum], tpm float64[n])."""
    rng = np.random.default_rng(seed)
    lengths = np.clip(rng.lognormal(7.3, 0.6, n_transcripts), 300, 20000).astype(np.int64)
    tpm = rng.lognormal(2.0, 1.5, n_transcripts)
    bases = synth_chrom(rng, int(lengths.sum()))
    names = ["ENST%011d.1" % i for i in range(n_transcripts)]
    return names, lengths, bases, tpm


def write_config3_files(dirname, names, lengths, bases, tpm):
    """transcripts.fa + expression.tsv (``target_id est_counts tpm``, README.md of the reference) for the CLI / the oracle."""
    import os
    offs = np.concatenate([[0], np.cumsum(lengths)])
    fa = os.path.join(dirname, "transcripts.fa")
    with open(fa, "wb") as f:
        for i, nm in enumerate(names):
            f.write(b">" + nm.encode() + b"\n" + bases[offs[i]:offs[i + 1]].tobytes() + b"\n")
    ex = os.path.join(dirname, "expression.tsv")
    with open(ex, "w") as f:
        f.write("target_id\test_counts\ttpm\n")
        f.writelines("%s\t%.2f\t%.6f\n" % (nm, 10.0, t) for nm, t in zip(names, tpm))
    return fa, ex


def config4_metagenome(n_species=50, seed=3):
    """BASELINE config 4: ``n_species`` species of 1-3 chromosomes of 2-6 Mb, i.i.d. ACGT, all circular,
    Even abundance (100 / n_species each); default_rng(3).  Returns ordered [(species, [(chrom header, uint8 array)])]."""
    rng = np.random.default_rng(seed)
    genomes = []
    for s in range(n_species):
        n_chrom = int(rng.integers(1, 4))
        recs = [("contig%d" % c, synth_chrom(rng, int(rng.integers(2_000_000, 6_000_001)))) for c in range(n_chrom)]
        genomes.append(("Species_%02d" % s, recs))
    return genomes


def write_config4_files(dirname, genomes, n_reads=20000000):
    """genome list, dna type list (all circular) and abundance file (``Size<TAB>n``) for the CLI / the oracle."""
    import os
    gl, dl, ab = (os.path.join(dirname, x) for x in ("genome_list.tsv", "dna_type.tsv", "abundance.tsv"))
    with open(gl, "w") as fg, open(dl, "w") as fd, open(ab, "w") as fa:
        fa.write("Size\t%d\n" % n_reads)
        for sp, recs in genomes:
            path = os.path.join(dirname, sp + ".fa")
            with open(path, "wb") as f:
                for nm, arr in recs:
                    f.write(b">" + nm.encode() + b"\n" + arr.tobytes() + b"\n")
                    fd.write("%s\t%s\tcircular\n" % (sp, nm))
            fg.write("%s\t%s\n" % (sp, path))
            fa.write("%s\t%.6f\n" % (sp, 100.0 / len(genomes)))
    return gl, dl, ab
