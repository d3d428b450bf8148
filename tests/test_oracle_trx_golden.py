"""Transcriptome mode (--no_model_ir) of the oracle, pinned bit-exactly against the unmodified reference
(tests/golden/vectors_trx.json, generator tests/golden/make_golden_trx.py)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, oracle_model

import nanosim_oracle as no

TRX = os.path.join(GOLDEN, "trx")


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "vectors_trx.json")) as f:
        return json.load(f)


def _ref(polya=True):
    return no.OracleTrxReference.from_files(os.path.join(TRX, "transcripts.fa"), os.path.join(TRX, "expression.tsv"),
                                            os.path.join(TRX, "polya.txt") if polya else None)


def test_expression_profile_and_helpers(vec, compiled_models, tmp_path):
    ref = _ref()
    g = vec["ecdf"]
    assert len(ref.ecdf_length_list) == g["n"]
    assert [[a, b] for a, b in ref.ecdf_length_list[:5]] == g["head"] and [[a, b] for a, b in ref.ecdf_length_list[-3:]] == g["tail"]
    assert ref.ecdf_weight_list[:5] == g["weights_head"] and float(sum(ref.ecdf_weight_list)) == g["weight_sum"]
    m = oracle_model(compiled_models["drna"], tmp_path)
    assert [m.split_counts(1000)[0]] == vec["numbers"]["aligned"]
    seed_all(900)
    sample = no.kde_lengths(m.kde_aligned_2d, 400, False, False)
    assert [[L, no.select_nearest_kde2d(sample, L)] for L, _ in vec["nearest"]] == vec["nearest"]
    for e in vec["extract"]:
        seed_all(e["seed"])
        if "key" in e:
            seq, pos, retain = no.extract_read_trx(ref, e["key"], e["length"], e["polya"])
            assert (pos, bool(retain), md5(seq)) == (e["pos"], e["retain"], e["md5"])
        else:
            seq, name = no.extract_read_transcriptome(ref, e["length"])
            assert (name, md5(seq)) == (e["name"], e["md5"])


def test_whole_transcriptome_loops(vec, compiled_models, tmp_path):
    for g in vec["runs"]:
        cf = g["cfg"]
        sink = no.ReadSink()
        if cf.get("unaligned"):
            ref = _ref()
            m = oracle_model(compiled_models["drna"], tmp_path)
            seed_all(g["seed"])
            no.simulation_unaligned_transcriptome(ref, m, sink, 50, g["max_l"], cf["fastq"], cf["n"])
        else:
            ref = _ref(cf["polya"])
            m = oracle_model(compiled_models["drna"], tmp_path, perfect=cf["per"])
            seed_all(g["seed"])
            no.simulation_aligned_transcriptome(ref, m, sink, None, cf["basecaller"], cf["n"], cf["polya"], cf["fastq"],
                                                cf["per"], cf["uracil"])
            assert md5("".join(r + "\n" for r in sink.error_rows)) == g["err_md5"]
        text = no.format_records(sink.records, cf["fastq"])
        assert text.split("\n")[0] == g["first_header"]
        assert text.count("\n") == g["n_lines"] and md5(text) == g["reads_md5"]
