"""Run-level golden histograms from the UNMODIFIED reference CLI (build container only).

    python tests/golden/make_golden_runs.py <config> <total_reads> <chunk_reads> [/tmp/models]

Runs ``oracle/ref_shim.py genome ...`` (i.e. /root/reference/src/simulator.py, unmodified) in chunks with
``-t 8``, extracts tests/run_stats.py statistics from each chunk's output files, deletes the files and
accumulates the histograms into tests/golden/ref_stats_<config>.json (resumable: re-running adds chunks
until <total_reads> is reached).  The reference cannot be seeded (simulator.py:1591-1592), so every chunk
is an independent sample; per-chunk totals are kept in the JSON meta so the reference's own run-to-run
noise can be estimated.
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
import synth  # noqa: E402

GUPPY = "human_NA12878_DNA_FAB49712_guppy"
DORADO = "human_giab_hg002_sub1M_kitv14_dorado_v3.2.1"

CONFIGS = {
    # name: (model, extra flags, fastq)
    "guppy_fasta": (GUPPY, [], False),
    "guppyq_fastq": (GUPPY + "+q", ["--fastq"], True),
    "dorado_fastq_chimeric": (DORADO, ["--fastq", "--chimeric"], True),
    "dorado_fastq_hp6_chimeric": (DORADO, ["--fastq", "--chimeric", "-hp", "-k", "6"], True),
    "guppy_perfect": (GUPPY, ["--perfect"], False),
    "guppy_medsd": (GUPPY, ["-med", "5000", "-sd", "1.05"], False),
    # eight independent single-threaded processes per chunk (independent numpy streams for the unaligned phase)
    "guppy_fasta_t1": (GUPPY, [], False),
    "guppyq_fastq_t1": (GUPPY + "+q", ["--fastq"], True),
}


def main():
    cfg, total, chunk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    models = sys.argv[4] if len(sys.argv) > 4 else "/tmp/models"
    model, flags, fastq = CONFIGS[cfg]
    work = tempfile.mkdtemp(prefix="golden_run_", dir="/tmp")
    ref = os.path.join(work, "ecoli5m.fa")
    synth.ecoli5m(ref)
    if model.endswith("+q"):
        base = os.path.join(models, model[:-2])
        aug = os.path.join(work, "aug")
        os.makedirs(aug)
        for fn in os.listdir(base):
            os.symlink(os.path.join(base, fn), os.path.join(aug, fn))
        os.symlink(os.path.join(models, DORADO, "training_base_qualities_model_parameters.tsv"),
                   os.path.join(aug, "training_base_qualities_model_parameters.tsv"))
        prefix = os.path.join(aug, "training")
    else:
        prefix = os.path.join(models, model, "training")
    out_json = os.path.join(HERE, "ref_stats_%s.json" % cfg)
    if os.path.exists(out_json):
        acc, meta = rs.load(out_json)
    else:
        acc, meta = rs.empty(), {"config": cfg, "model": model, "flags": flags, "reference": "synth.ecoli5m()",
                                 "cmd": "simulator.py genome -rg ecoli5m.fa -c <model>/training -n %d -t 8 %s" % (chunk, " ".join(flags)),
                                 "chunks": []}
    done = acc["n_aligned"] + acc["n_unaligned"]
    while done < total:
        out = os.path.join(work, "sim")
        t0 = time.time()
        if cfg.endswith("_t1"):
            procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "ref_shim.py"), "genome", "-rg", ref,
                                       "-c", prefix, "-n", str(chunk // 8), "-t", "1", "-o", out + "_p%d" % i] + flags,
                                      stdout=subprocess.DEVNULL) for i in range(8)]
            assert all(p.wait() == 0 for p in procs)
            t1 = time.time()
            s = rs.empty()
            for i in range(8):
                rs.merge(s, rs.stats_from_prefix(out + "_p%d" % i, fastq, with_errors=False))
        else:
            subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_shim.py"), "genome", "-rg", ref, "-c", prefix,
                            "-n", str(chunk), "-t", "8", "-o", out] + flags, check=True, stdout=subprocess.DEVNULL)
            t1 = time.time()
            s = rs.stats_from_prefix(out, fastq)
        for fn in os.listdir(work):
            if fn.startswith("sim_"):
                os.remove(os.path.join(work, fn))
        meta["chunks"].append({"n_aligned": int(s["n_aligned"]), "ref_bases": int(s["ref_bases"]),
                               "aligned_bases": int(s["aligned_bases"]),
                               "event_bases": {k: int(v) for k, v in s["event_bases"].items()},
                               "events": {k: int(v) for k, v in s["events"].items()},
                               "sim_seconds": round(t1 - t0, 1)})
        rs.merge(acc, s)
        rs.save(acc, out_json, meta)
        done = acc["n_aligned"] + acc["n_unaligned"]
        print("%s: %d/%d reads (chunk sim %.0fs, parse %.0fs)" % (cfg, done, total, t1 - t0, time.time() - t1), flush=True)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
