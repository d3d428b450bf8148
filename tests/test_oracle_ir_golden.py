"""Intron-retention branch of transcriptome mode in the oracle, pinned bit-exactly against the unmodified reference's own
IR code (update_structure, extract_read_pos, the IR part of simulation_aligned_transcriptome) run through the HTSeq / pysam
stand-ins of oracle/ref_shim.py (tests/golden/vectors_ir.json, generator tests/golden/make_golden_ir.py)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, oracle_model

import nanosim_oracle as no

IR = os.path.join(GOLDEN, "ir")


def seed_all(s):
    random.seed(s)
    np.random.seed(s)


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "vectors_ir.json")) as f:
        return json.load(f)


def _ref(polya=True):
    ref = no.OracleTrxReference.from_files(os.path.join(IR, "transcripts.fa"), os.path.join(IR, "expression.tsv"),
                                           os.path.join(IR, "polya.txt") if polya else None)
    return ref.load_ir(os.path.join(IR, "genome.fa"), os.path.join(IR, "annotation.gff3"), os.path.join(IR, "IR_markov_model"))


def test_structure_model_and_per_read_functions(vec):
    ref = _ref()
    assert {k: [list(x) for x in v] for k, v in sorted(ref.structure.items())} == vec["structure"]
    assert {k: sorted([a, b, f] for (a, b), f in v) for k, v in ref.ir_model.items()} == vec["ir_model"]
    for g in vec["update_extract"]:
        seed_all(g["seed"])
        flag, st = no.update_structure(ref.structure[g["key"]], ref.ir_model)
        assert (bool(flag), [x[0] for x in st]) == (g["flag"], g["types"])
        if "extract" in g:
            e = g["extract"]
            ivs, retain, ir_list = no.extract_read_pos(e["length"], e["ref_len"], st, g["key"] in ref.trx_with_polya)
            assert ([list(iv) for iv in ivs], bool(retain), [list(x) for x in ir_list]) == (e["ivs"], e["retain"], e["ir_list"])


def test_whole_loops_with_intron_retention(vec, compiled_models, tmp_path):
    for g in vec["runs"]:
        cf = g["cfg"]
        ref = _ref(cf["polya"])
        m = oracle_model(compiled_models["drna"], tmp_path)
        sink = no.ReadSink()
        seed_all(g["seed"])
        no.simulation_aligned_transcriptome(ref, m, sink, None, cf["basecaller"], cf["n"], cf["polya"], cf["fastq"], False,
                                            cf["uracil"], model_ir=True)
        text = no.format_records(sink.records, cf["fastq"])
        heads = [l for l in text.split("\n") if l[:1] in "@>" and "_aligned_" in l]
        assert sum("_RetainedIntron_" in h for h in heads) == g["n_ir"]
        assert [h for h in heads if "_RetainedIntron_" in h][:6] == g["ir_headers"]
        assert md5("".join(r + "\n" for r in sink.error_rows)) == g["err_md5"]
        assert text.count("\n") == g["n_lines"] and md5(text) == g["reads_md5"]
