"""Metagenome mode of the oracle, pinned bit-exactly against the unmodified reference (tests/golden/vectors_meta.json,
generator tests/golden/make_golden_meta.py)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, meta_fixture, oracle_model

import nanosim_oracle as no


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "vectors_meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def fx():
    return meta_fixture()


def test_reference_and_abundance_parsing(vec, fx, compiled_models, tmp_path):
    ref, numbers, multi = fx
    assert list(ref.seq_len.keys()) == vec["species"]
    for sp, d in vec["chroms"].items():
        assert {k: [ref.seq_len[sp][k], ref.dict_dna_type[sp][k]] for k in ref.seq_len[sp]} == d
    m = oracle_model(compiled_models["even"], tmp_path, chimeric=True, mode="metagenome")
    assert [m.split_counts(n)[0] for n in numbers] == vec["numbers"]["aligned"]
    infl = {sp: no.inflate_abun(multi["sample0"], sp, m.abun_inflation) for sp in multi["sample0"]}
    assert infl == vec["inflated"]


def test_assign_species_and_extract(vec, fx, compiled_models, tmp_path):
    ref, numbers, multi = fx
    m = oracle_model(compiled_models["even"], tmp_path, chimeric=True, mode="metagenome")
    abun = multi["sample0"]
    infl = {sp: no.inflate_abun(abun, sp, m.abun_inflation) for sp in abun}
    for g in vec["assign_species"]:
        seed_all(g["seed"])
        segs = np.random.geometric(1 / m.segment_mean, g["n_reads"])
        lens = [float(x) for x in np.random.uniform(200, 9000, int(segs.sum()))]
        assert [int(x) for x in segs] == g["segs"] and lens == g["lens"]
        sp_list, len_list, seg_list = no.assign_species(list(lens), segs, dict(g["current"]), abun, infl)
        assert sp_list == g["species"] and len_list == g["lengths"] and [int(x) for x in seg_list] == g["seg_sorted"]
    for g in vec["extract_read"]:
        seed_all(g["seed"])
        seq, name = no.extract_read_meta(ref, g["length"], g["species"])
        assert name == g["name"] and md5(seq) == g["md5"]


def test_whole_metagenome_loops(vec, fx, compiled_models, tmp_path):
    ref, numbers, multi = fx
    for g in vec["runs"]:
        cf = g["cfg"]
        sink = no.ReadSink()
        if cf.get("unaligned"):
            m = oracle_model(compiled_models["even"], tmp_path, chimeric=True, mode="metagenome")
            seed_all(g["seed"])
            no.simulation_unaligned_meta(ref, m, sink, 50, max(ref.max_chrom.values()), cf["fastq"], cf["n"])
        else:
            m = oracle_model(compiled_models["even"], tmp_path, chimeric=True, mode="metagenome", perfect=cf["per"])
            abun = multi["sample%d" % g["sample"]]
            infl = {sp: no.inflate_abun(abun, sp, m.abun_inflation) for sp in abun}
            seed_all(g["seed"])
            no.simulation_aligned_metagenome(ref, m, sink, abun, infl, 50, g["max_l"], None, cf["fastq"], cf["n"], cf["per"],
                                             cf["chimeric"])
            assert md5("".join(r + "\n" for r in sink.error_rows)) == g["err_md5"]
        text = no.format_records(sink.records, cf["fastq"])
        assert text.split("\n")[0] == g["first_header"]
        assert text.count("\n") == g["n_lines"] and md5(text) == g["reads_md5"]
