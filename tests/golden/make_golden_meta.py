"""Golden vectors for METAGENOME mode from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_meta.py /tmp/models

Writes the fixture (tests/golden/meta/: four small species, genome list, abundance, dna-type list) and
tests/golden/vectors_meta.json with seeded outputs of assign_species, extract_read("metagenome") and whole
simulation_aligned_metagenome / simulation_unaligned loops (Even v3.2.2 model)."""
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

EVEN = "metagenome_ERR3152364_Even_v3.2.2"
META = os.path.join(HERE, "meta")


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def make_fixture():
    os.makedirs(META, exist_ok=True)
    rng = np.random.default_rng(21)
    species = [("Alpha bacter", [("chrom_1.1 alpha chromosome", 60000), ("plasmid_a", 9000)]),
               ("Beta coccus", [("NC_0001.2 beta complete genome", 45000)]),
               ("Gamma yeast", [("chrI", 30000), ("chrII", 22000), ("chrM", 6000)]),
               ("Delta phage", [("delta_genome", 14000)])]
    with open(os.path.join(META, "genome_list.tsv"), "w") as gl, open(os.path.join(META, "dna_type.tsv"), "w") as dl:
        for i, (sp, chroms) in enumerate(species):
            path = os.path.join(META, "sp%d.fa" % i)
            with open(path, "w") as f:
                for name, n in chroms:
                    s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes().decode()
                    f.write(">" + name + "\n")
                    for j in range(0, n, 60):
                        f.write(s[j:j + 60] + "\n")
                    linear = sp.startswith("Gamma") and not name.startswith("chrM")
                    dl.write("%s\t%s\t%s\n" % (sp, name, "linear" if linear else "circular"))
            gl.write("%s\t%s\n" % (sp, path))
    with open(os.path.join(META, "abundance.tsv"), "w") as f:
        f.write("Size\t300\t200\n")
        for (sp, _), a, b in zip(species, (40.0, 30.0, 20.0, 10.0), (5.0, 5.0, 60.0, 30.0)):
            f.write("%s\t%s\t%s\n" % (sp, a, b))


def main(models_dir):
    make_fixture()
    prefix = os.path.join(models_dir, EVEN, "training")
    sim = ref_shim.load_reference_module()
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    out = {}

    def profile(per=False):
        sys.stdout = devnull
        sim.read_profile(os.path.join(META, "genome_list.tsv"), [], prefix, per, "metagenome", None,
                         dna_type=os.path.join(META, "dna_type.tsv"), abun=os.path.join(META, "abundance.tsv"),
                         chimeric=True, homopolymer=False, fastq=True)
        sys.stdout = stdout

    profile()
    sim.dict_abun = sim.multi_dict_abun["sample0"]
    sim.dict_abun_inflated = {sp: sim.inflate_abun(sim.dict_abun, sp) for sp in sim.dict_abun}
    out["species"] = list(sim.seq_len.keys())
    out["chroms"] = {sp: {k: [sim.seq_len[sp][k], sim.dict_dna_type[sp][k]] for k in sim.seq_len[sp]} for sp in sim.seq_len}
    out["numbers"] = {"aligned": sim.number_aligned_l, "unaligned": sim.number_unaligned_l}
    out["inflated"] = sim.dict_abun_inflated
    # assign_species
    cases = []
    for s, n_reads in ((1, 40), (2, 200)):
        seed_all(800 + s)
        segs = np.random.geometric(1 / sim.segment_mean, n_reads)
        lens = [float(x) for x in np.random.uniform(200, 9000, int(segs.sum()))]
        cur = {sp: 0 for sp in sim.dict_abun}
        cur[out["species"][0]] = 5000 * s
        sp_list, len_list, seg_list = sim.assign_species(list(lens), segs, dict(cur))
        cases.append({"seed": 800 + s, "n_reads": n_reads, "segs": [int(x) for x in segs], "lens": lens, "current": cur,
                      "species": sp_list, "lengths": len_list, "seg_sorted": [int(x) for x in seg_list]})
    out["assign_species"] = cases
    ex = []
    for s, ln, sp in ((1, 500, "Alpha_bacter"), (2, 20000, "Alpha_bacter"), (3, 40000, "Gamma_yeast"), (4, 7000, None),
                      (5, 25000, "Delta_phage"), (6, 0, "Beta_coccus")):
        seed_all(820 + s)
        seq, name = sim.extract_read("metagenome", ln, sp)
        ex.append({"seed": 820 + s, "length": ln, "species": sp, "name": name, "md5": md5(seq)})
    out["extract_read"] = ex
    runs = []
    tmp = tempfile.mkdtemp(prefix="golden_meta_")
    for i, cf in enumerate((dict(fastq=True, per=False, chimeric=True, n=60), dict(fastq=False, per=False, chimeric=False, n=40),
                            dict(fastq=True, per=True, chimeric=False, n=25))):
        profile(cf["per"])
        sim.dict_abun = sim.multi_dict_abun["sample%d" % (i % 2)]
        sim.dict_abun_inflated = {sp: sim.inflate_abun(sim.dict_abun, sp) for sp in sim.dict_abun}
        sim.total_simulated = mp.Value("i", 0, lock=True)
        seed_all(840 + i)
        o_reads, o_err = os.path.join(tmp, "r%d" % i), os.path.join(tmp, "e%d" % i)
        max_l = max(sim.max_chrom.values())
        sys.stdout = devnull
        sim.simulation_aligned_metagenome(50, max_l, None, None, o_reads, o_err, None, cf["fastq"], cf["n"], cf["per"], cf["chimeric"])
        sys.stdout = stdout
        reads, err = open(o_reads).read(), open(o_err).read()
        runs.append({"cfg": cf, "seed": 840 + i, "sample": i % 2, "max_l": max_l, "reads_md5": md5(reads), "err_md5": md5(err),
                     "n_lines": reads.count("\n"), "first_header": reads.split("\n")[0]})
    profile()
    for i, fq in enumerate((False, True)):
        sim.total_simulated = mp.Value("i", 0, lock=True)
        seed_all(860 + i)
        o_reads = os.path.join(tmp, "u%d" % i)
        sys.stdout = devnull
        sim.simulation_unaligned("metagenome", 50, max(sim.max_chrom.values()), None, None, o_reads, fq, 15, False)
        sys.stdout = stdout
        reads = open(o_reads).read()
        runs.append({"cfg": dict(fastq=fq, n=15, unaligned=True), "seed": 860 + i, "reads_md5": md5(reads),
                     "n_lines": reads.count("\n"), "first_header": reads.split("\n")[0]})
    shutil.rmtree(tmp)
    out["runs"] = runs
    with open(os.path.join(HERE, "vectors_meta.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote vectors_meta.json", os.path.getsize(os.path.join(HERE, "vectors_meta.json")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/models")
