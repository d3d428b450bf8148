"""Pins the oracle: oracle/nanosim_oracle.py must reproduce, bit for bit, what the UNMODIFIED
reference functions returned under the same seeds (tests/golden/vectors.json, produced by
tests/golden/make_golden_vectors.py in the build container)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, oracle_model

import nanosim_oracle as no


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


@pytest.fixture(scope="module")
def vectors():
    with open(os.path.join(GOLDEN, "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def mini_ref():
    return no.OracleReference.from_fasta(os.path.join(GOLDEN, "mini_ref.fa"))


CASES = [("guppy", False), ("dorado", True)]


@pytest.mark.parametrize("tag,hp", CASES)
def test_ecdf_parse(tag, hp, vectors, compiled_models, tmp_path):
    m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True)
    for nm, table in (("first_match", m.first_match), ("match_markov", m.match_markov)):
        gold = vectors["cases"][tag]["ecdf"][nm]
        assert [list(b) for b in table.keys()] == [g["bin"] for g in gold]
        for g, (b, items) in zip(gold, table.items()):
            assert len(items) == g["n"]
            assert [list(x) for x in items[:4]] == g["head"]
            assert [list(x) for x in items[-3:]] == g["tail"]
            assert float(sum(sum(x) for x in items)) == g["sum"]


@pytest.mark.parametrize("tag,hp", CASES)
def test_samplers(tag, hp, vectors, compiled_models, tmp_path):
    m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True)
    c = vectors["cases"][tag]
    seed_all(11)
    p = m.error_par["mis"]
    assert [int(no.pois_geom(p[0], p[2], p[3])) for _ in range(300)] == c["pois_geom"]
    seed_all(12)
    assert [int(no.wei_geom(*m.error_par["ins"])) for _ in range(300)] == c["wei_geom_ins"]
    seed_all(13)
    assert [int(no.wei_geom(*m.error_par["del"])) for _ in range(300)] == c["wei_geom_del"]
    for i, st in enumerate(("mis", "ins", "match", "ht", "unmapped")):
        seed_all(20 + i)
        assert no.base_qualities(m.base_qual[st], 64) == c["quals"][st]


@pytest.mark.parametrize("tag,hp", CASES)
def test_error_lists(tag, hp, vectors, compiled_models, tmp_path):
    m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True)
    c = vectors["cases"][tag]
    for g in c["error_list"]:
        seed_all(g["seed"])
        l_new, middle_ref, e_dict, e_count = no.error_list(g["m_ref"], m, g["fastq"])
        assert (l_new, middle_ref) == (g["l_new"], g["middle_ref"])
        assert [[float(k), v[0], int(v[1])] for k, v in e_dict.items()] == g["e_dict"]
        assert {k: int(v) for k, v in e_count.items()} == g["e_count"]
    for g in c["unaligned_error_list"]:
        seed_all(g["seed"])
        l_new, middle_ref, e_dict, _ = no.unaligned_error_list(g["m_ref"], m)
        assert (l_new, middle_ref) == (g["l_new"], g["middle_ref"])
        assert [[float(k), v[0], int(v[1])] for k, v in e_dict.items()] == g["e_dict"]


@pytest.mark.parametrize("tag,hp", CASES)
def test_extract_case_mutate(tag, hp, vectors, compiled_models, tmp_path, mini_ref):
    m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True)
    c = vectors["cases"][tag]
    g = c["case_convert"]
    seed_all(g["seed"])
    assert no.case_convert(g["in"]) == g["out"]
    for g in c["extract_read"]:
        seed_all(g["seed"])
        seq, name = no.extract_read(mini_ref, g["dna_type"], g["length"])
        assert name == g["name"] and md5(seq) == g["md5"]
    for g in c["mutate_read"]:
        seed_all(g["seed"])
        l_new, middle_ref, e_dict, e_count = no.error_list(g["length"], m, g["fastq"])
        seq, name = no.extract_read(mini_ref, "linear", middle_ref)
        seq = no.case_convert(seq)
        log = []
        mutated, quals = no.mutate_read(seq, name, log, e_dict, e_count, g["fastq"], g["k"], m)
        assert len(mutated) == g["mutated_len"] and md5(mutated) == g["mutated_md5"]
        assert [int(x) for x in quals] == g["quals"]
        assert len(log) == g["n_log"] and md5("".join(r + "\n" for r in log)) == g["log_md5"]
        if g["k"]:
            m2, q2 = no.mutate_homo(mutated, quals, g["k"], m)
            assert len(m2) == g["homo_len"] and md5(m2) == g["homo_md5"]
            assert [int(x) for x in q2] == g["homo_quals"]


def test_circular_extract(vectors):
    ref = no.OracleReference.from_fasta(os.path.join(GOLDEN, "mini_circular.fa"))
    for g in vectors["circular"]:
        seed_all(g["seed"])
        seq, name = no.extract_read(ref, "circular", g["length"])
        assert name == g["name"]
        assert (seq if g["length"] < 100 else md5(seq)) == g["seq"]


@pytest.mark.parametrize("tag,hp", CASES)
def test_whole_loops(tag, hp, vectors, compiled_models, tmp_path, mini_ref):
    c = vectors["cases"][tag]
    for g in c["runs"]:
        cf = g["cfg"]
        if g["name"].startswith("unaligned"):
            m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True)
            sink = no.ReadSink()
            seed_all(g["seed"])
            no.simulation_unaligned(mini_ref, m, sink, "linear", 50, mini_ref.max_chrom, None, None, cf["fastq"], cf["n"])
            text = no.format_records(sink.records, cf["fastq"])
        else:
            m = oracle_model(compiled_models[tag], tmp_path, homopolymer=hp, chimeric=True, perfect=cf["per"])
            sink = no.ReadSink()
            seed_all(g["seed"])
            no.simulation_aligned_genome(mini_ref, m, sink, "linear", 50, mini_ref.max_chrom, None, None, cf["k"],
                                         cf["fastq"], cf["n"], cf["per"], cf["chimeric"])
            text = no.format_records(sink.records, cf["fastq"])
            assert md5("".join(r + "\n" for r in sink.error_rows)) == g["err_md5"], g["name"]
        assert text.split("\n")[0] == g["first_header"], g["name"]
        assert text.count("\n") == g["n_lines"], g["name"]
        assert md5(text) == g["reads_md5"], g["name"]
