"""Generates the committed golden vectors by calling the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference and the extracted pre-trained models):

    python tests/golden/make_golden_vectors.py /tmp/models

It (1) compiles the two models used by the tests into nanosim_b200/data/*.npz (lossless copies of the
small text tables and of the KDE training samples), (2) writes tests/golden/mini_ref.fa, and
(3) imports the reference through oracle/ref_shim.py, seeds ``random`` and ``np.random`` and records
inputs/outputs of its per-read functions and of whole simulation_aligned_genome /
simulation_unaligned runs into tests/golden/vectors.json.

tests/test_oracle_golden.py replays the same seeds through oracle/nanosim_oracle.py and requires
bit-identical results.  That is what pins the oracle.
"""
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
from nanosim_b200.model import CompiledModel  # noqa: E402

GUPPY = "human_NA12878_DNA_FAB49712_guppy"
DORADO = "human_giab_hg002_sub1M_kitv14_dorado_v3.2.1"


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


def make_mini_ref(path):
    rng = np.random.default_rng(7)
    recs = []
    for name, n in (("chrA_1.2 some description", 30000), ("chrB", 20000), ("plasmid_x", 10000)):
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        # homopolymer runs, IUPAC codes and lower case so that every branch is exercised
        for _ in range(n // 400):
            p = int(rng.integers(0, n - 20))
            s[p:p + int(rng.integers(4, 14))] = b"ACGT"[int(rng.integers(0, 4))]
        for _ in range(n // 1500):
            p = int(rng.integers(0, n - 5))
            s[p:p + int(rng.integers(1, 4))] = ord("NRYKMSWBDHV"[int(rng.integers(0, 11))])
        txt = s.tobytes().decode()
        txt = txt[:200].lower() + txt[200:]
        recs.append((name, txt))
    with open(path, "w") as f:
        for name, txt in recs:
            f.write(">" + name + "\n")
            for i in range(0, len(txt), 70):
                f.write(txt[i:i + 70] + "\n")


def e_dict_to_list(d):
    return [[float(k), v[0], int(v[1])] for k, v in d.items()]


def main(models_dir):
    data_dir = os.path.join(ROOT, "nanosim_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    g_prefix = os.path.join(models_dir, GUPPY, "training")
    d_prefix = os.path.join(models_dir, DORADO, "training")
    # config-2 resolution: the guppy model has no quality table, so the augmented guppy
    # model carries dorado_v3.2.1's _base_qualities_model_parameters.tsv
    CompiledModel.from_prefix(g_prefix, extra_text={
        "base_qualities_model_parameters.tsv": d_prefix + "_base_qualities_model_parameters.tsv"}
    ).save(os.path.join(data_dir, "guppy_fab49712_plusq.npz"))
    CompiledModel.from_prefix(d_prefix).save(os.path.join(data_dir, "dorado_kitv14_v3.2.1.npz"))

    mini = os.path.join(HERE, "mini_ref.fa")
    make_mini_ref(mini)

    # the same augmented directory for the reference itself
    aug = tempfile.mkdtemp(prefix="guppy_aug_")
    for fn in os.listdir(os.path.dirname(g_prefix)):
        os.symlink(os.path.join(os.path.dirname(g_prefix), fn), os.path.join(aug, fn))
    os.symlink(d_prefix + "_base_qualities_model_parameters.tsv",
               os.path.join(aug, "training_base_qualities_model_parameters.tsv"))

    sim = ref_shim.load_reference_module()
    import mixed_model as mm
    import model_base_qualities as mbq

    out = {"cases": {}}
    devnull = open(os.devnull, "w")

    for tag, prefix, hp, chim in (("guppy", os.path.join(aug, "training"), False, True),
                                  ("dorado", d_prefix, True, True)):
        stdout = sys.stdout
        sys.stdout = devnull
        sim.read_profile(mini, [100000], prefix, False, "genome", None, dna_type="linear", chimeric=chim,
                         homopolymer=hp, fastq=True)
        sys.stdout = stdout
        c = {}
        # -- ECDF parsing fingerprints
        fp = {}
        for nm, d in (("first_match", sim.match_ht_list), ("match_markov", sim.match_markov_model)):
            bins = []
            for b, iv in d.items():
                items = list(iv.items())
                bins.append({"bin": list(b), "n": len(items),
                             "head": [[k[0], k[1], v[0], v[1]] for k, v in items[:4]],
                             "tail": [[k[0], k[1], v[0], v[1]] for k, v in items[-3:]],
                             "sum": float(sum(k[0] + k[1] + v[0] + v[1] for k, v in items))})
            fp[nm] = bins
        c["ecdf"] = fp
        # -- samplers
        seed_all(11)
        c["pois_geom"] = [int(mm.pois_geom(sim.error_par["mis"][0], sim.error_par["mis"][2],
                                           sim.error_par["mis"][3])) for _ in range(300)]
        seed_all(12)
        c["wei_geom_ins"] = [int(mm.wei_geom(*sim.error_par["ins"])) for _ in range(300)]
        seed_all(13)
        c["wei_geom_del"] = [int(mm.wei_geom(*sim.error_par["del"])) for _ in range(300)]
        q = {}
        for i, st in enumerate(("mis", "ins", "match", "ht", "unmapped")):
            seed_all(20 + i)
            p = sim.lognorm_base_qual[st]
            q[st] = [int(x) for x in mbq.predict_base_qualities(p["sd"], p["loc"], np.exp(p["mu"]), 64)]
        c["quals"] = q
        # -- error_list / unaligned_error_list
        el = []
        for s, m_ref, fq in [(1, 0, True), (2, 1, False), (3, 5, True), (4, 60, False), (5, 400, True),
                             (6, 3000, True), (7, 3000, False), (8, 12000, True)]:
            seed_all(100 + s)
            l_new, middle_ref, e_dict, e_count = sim.error_list(m_ref, sim.match_markov_model, sim.match_ht_list,
                                                                sim.error_par, sim.trans_error_pr, fq)
            el.append({"seed": 100 + s, "m_ref": m_ref, "fastq": fq, "l_new": int(l_new),
                       "middle_ref": int(middle_ref), "e_dict": e_dict_to_list(e_dict),
                       "e_count": {k: int(v) for k, v in e_count.items()}})
        c["error_list"] = el
        ul = []
        for s, m_ref in [(1, 0), (2, 1), (3, 17), (4, 250), (5, 2000)]:
            seed_all(200 + s)
            l_new, middle_ref, e_dict, e_count = sim.unaligned_error_list(m_ref, sim.error_par)
            ul.append({"seed": 200 + s, "m_ref": m_ref, "l_new": int(l_new), "middle_ref": int(middle_ref),
                       "e_dict": e_dict_to_list(e_dict)})
        c["unaligned_error_list"] = ul
        # -- case_convert / extract_read
        seed_all(300)
        frag = sim.seq_dict["chrA-1"][150:450] + "nryk" + sim.seq_dict["plasmid-x"][0:60]
        c["case_convert"] = {"seed": 300, "in": frag, "out": sim.case_convert(frag)}
        ex = []
        for s, dt, ln in [(1, "linear", 500), (2, "linear", 25000), (3, "linear", 1)]:
            seed_all(310 + s)
            seq, name = sim.extract_read(dt, ln)
            ex.append({"seed": 310 + s, "dna_type": dt, "length": ln, "name": name, "md5": md5(seq)})
        c["extract_read"] = ex
        # -- mutate_read (+ mutate_homo for the hp model)
        mr = []
        for s, ln, fq, k in [(1, 700, True, None), (2, 700, False, None), (3, 2500, True, 6 if hp else None),
                             (4, 1800, False, 5 if hp else None)]:
            seed_all(400 + s)
            l_new, middle_ref, e_dict, e_count = sim.error_list(ln, sim.match_markov_model, sim.match_ht_list,
                                                                sim.error_par, sim.trans_error_pr, fq)
            seq, name = sim.extract_read("linear", middle_ref)
            seq = sim.case_convert(seq)

            class _Log:
                def __init__(self):
                    self.rows = []

                def write(self, x):
                    self.rows.append(x)

            log = _Log()
            mutated, quals = sim.mutate_read(seq, name, log, e_dict, e_count, fq, k)
            rec = {"seed": 400 + s, "length": ln, "fastq": fq, "k": k, "mutated_md5": md5(mutated),
                   "mutated_len": len(mutated), "quals": [int(x) for x in quals], "log_md5": md5("".join(log.rows)),
                   "n_log": len(log.rows)}
            if k:
                m2, q2 = sim.mutate_homo(mutated, quals, k)
                rec["homo_md5"] = md5(m2)
                rec["homo_len"] = len(m2)
                rec["homo_quals"] = [int(x) for x in q2]
            mr.append(rec)
        c["mutate_read"] = mr
        # -- whole per-mode loops
        runs = []
        tmp = tempfile.mkdtemp(prefix="golden_runs_")
        cfgs = [("aligned_fasta", dict(fastq=False, per=False, chimeric=False, k=None, n=30)),
                ("aligned_fastq_chimeric", dict(fastq=True, per=False, chimeric=True, k=None, n=40)),
                ("perfect_fastq", dict(fastq=True, per=True, chimeric=False, k=None, n=20))]
        if hp:
            cfgs.append(("aligned_fastq_hp6_chimeric", dict(fastq=True, per=False, chimeric=True, k=6, n=40)))
        for i, (nm, cf) in enumerate(cfgs):
            # perfect mode reads kde_aligned from _aligned_reads.pkl: re-run read_profile accordingly
            sys.stdout = devnull
            sim.read_profile(mini, [100000], prefix, cf["per"], "genome", None, dna_type="linear", chimeric=chim,
                             homopolymer=hp, fastq=True)
            sys.stdout = stdout
            sim.total_simulated = mp.Value("i", 0, lock=True)
            seed_all(500 + i)
            o_reads, o_err = os.path.join(tmp, nm + ".reads"), os.path.join(tmp, nm + ".err")
            sys.stdout = devnull
            sim.simulation_aligned_genome("linear", 50, sim.max_chrom, None, None, o_reads, o_err, cf["k"],
                                          cf["fastq"], cf["n"], cf["per"], cf["chimeric"])
            sys.stdout = stdout
            reads = open(o_reads).read()
            err = open(o_err).read()
            runs.append({"name": nm, "seed": 500 + i, "cfg": cf, "reads_md5": md5(reads), "err_md5": md5(err),
                         "n_lines": reads.count("\n"), "first_header": reads.split("\n")[0]})
        sys.stdout = devnull
        sim.read_profile(mini, [100000], prefix, False, "genome", None, dna_type="linear", chimeric=chim,
                         homopolymer=hp, fastq=True)
        sys.stdout = stdout
        for i, fq in enumerate((False, True)):
            sim.total_simulated = mp.Value("i", 0, lock=True)
            seed_all(600 + i)
            o_reads = os.path.join(tmp, "unaligned%d.reads" % i)
            sys.stdout = devnull
            sim.simulation_unaligned("linear", 50, sim.max_chrom, None, None, o_reads, fq, 12, False)
            sys.stdout = stdout
            reads = open(o_reads).read()
            runs.append({"name": "unaligned_fastq" if fq else "unaligned_fasta", "seed": 600 + i,
                         "cfg": dict(fastq=fq, n=12), "reads_md5": md5(reads), "n_lines": reads.count("\n"),
                         "first_header": reads.split("\n")[0]})
        shutil.rmtree(tmp)
        c["runs"] = runs
        out["cases"][tag] = c

    # circular extraction uses a single-chromosome reference
    single = os.path.join(HERE, "mini_circular.fa")
    rng = np.random.default_rng(9)
    with open(single, "w") as f:
        f.write(">circ\n")
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5000)].tobytes().decode()
        for i in range(0, len(s), 60):
            f.write(s[i:i + 60] + "\n")
    sys.stdout = devnull
    sim.read_profile(single, [10], d_prefix, False, "genome", None, dna_type="circular", chimeric=False,
                     homopolymer=False, fastq=False)
    sys.stdout = sys.__stdout__
    circ = []
    for s_, ln in [(1, 800), (2, 4990), (3, 3)]:
        seed_all(700 + s_)
        seq, name = sim.extract_read("circular", ln)
        circ.append({"seed": 700 + s_, "length": ln, "name": name, "seq": seq if ln < 100 else md5(seq)})
    out["circular"] = circ

    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    shutil.rmtree(aug)
    print("wrote", os.path.join(HERE, "vectors.json"), os.path.getsize(os.path.join(HERE, "vectors.json")), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/models")
