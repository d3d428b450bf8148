"""CPU tests of the host layer: the C-ABI library loads and exports everything the header declares, FASTA packing
matches the oracle's reader, CLI validation behaves like the reference's, rank sharding / merging works under gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

import nanosim_oracle as no


def test_library_exports_every_header_symbol():
    from nanosim_b200 import _lib

    header = open(os.path.join(ROOT, "include", "nanosim_b200.h")).read()
    declared = set(re.findall(r"\b(ns_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    # struct layouts the ctypes mirror relies on
    assert ctypes.sizeof(_lib.NsReadMeta) == 32 and ctypes.sizeof(_lib.NsPieceMeta) == 64


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nanosim_b200.engine import Engine, NanoSimError

    with pytest.raises(NanoSimError):
        Engine(device=0, seed=1)


def test_fasta_packing_matches_oracle_reader():
    from nanosim_b200.reference_fasta import PackedReference

    for fn in ("mini_ref.fa", "mini_circular.fa"):
        path = os.path.join(GOLDEN, fn)
        ref = PackedReference.from_fasta(path)
        recs = no.read_fasta(path)
        assert ref.names == [n for n, _ in recs]
        for i, (_, s) in enumerate(recs):
            a, b = int(ref.offsets[i]), int(ref.offsets[i + 1])
            assert ref.bases[a:b].tobytes().decode() == s
        assert ref.max_chrom == max(len(s) for _, s in recs)


@pytest.mark.parametrize("argv", [
    ["genome", "-rg", "x.fa", "-max", "10", "-min", "50"],
    ["genome", "-rg", "x.fa", "--perfect", "--chimeric"],
    ["genome", "-rg", "x.fa", "-med", "5000"],
    ["genome", "-rg", "x.fa", "-hp"],
    ["genome", "-rg", "x.fa", "-s", "1.5"],
    ["genome", "-rg", "x.fa", "-med", "5000", "-sd", "0.5", "--chimeric"],
    [],
])
def test_cli_validation_exits_like_the_reference(argv):
    from nanosim_b200 import simulator

    with pytest.raises(SystemExit) as e:
        simulator.main(argv)
    assert e.value.code == 1


def test_shard_matches_reference_worker_split():
    from nanosim_b200.simulator import _shard

    for n, w in ((1000, 8), (898, 3), (7, 8), (0, 2)):
        parts = [_shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert all(hi - lo == n // w for lo, hi in parts[:-1])      # simulator.py:1588, remainder to the last


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from nanosim_b200.simulator import _shard, merge_rank_files
rank, world, out = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), sys.argv[2]
dist.init_process_group("gloo")
lo, hi = _shard(101, rank, world)
with open(out + "_aligned_reads%d.fasta" % rank, "w") as f:
    for i in range(lo, hi):
        f.write(">r%d\nACGT\n" % i)
with open(out + "_error_profile%d" % rank, "w") as f:
    f.write("r%d\t0\tmis\t1\tA\tC\n" % lo)
with open(out + "_unaligned_reads%d.fasta" % rank, "w") as f:
    f.write(">u%d\nAC\n" % rank)
dist.barrier()
if rank == 0:
    merge_rank_files(out, False, False, world)
dist.barrier()
'''


def test_two_rank_gloo_merge(tmp_path):
    script = os.path.join(str(tmp_path), "w.py")
    with open(script, "w") as f:
        f.write(_GLOO_WORKER)
    out = os.path.join(str(tmp_path), "sim")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, script, ROOT, out], env=dict(env, RANK=str(r))) for r in range(2)]
    assert all(p.wait(timeout=120) == 0 for p in procs)
    names = [l[1:].strip() for l in open(out + "_aligned_reads.fasta") if l.startswith(">")]
    assert names == ["r%d" % i for i in range(101)]
    assert open(out + "_aligned_error_profile").read().startswith("Seq_name\t")
    assert open(out + "_unaligned_reads.fasta").read() == ">u0\nAC\n>u1\nAC\n"
    assert not os.path.exists(out + "_aligned_reads0.fasta")


def _synthetic_batch(rng, ref, n_reads, rewritten):
    """A hand-built fetched batch (no GPU): random edit scripts applied to `ref` in Python."""
    from nanosim_b200 import _lib as L
    from nanosim_b200.engine import Batch
    comp = {65: 84, 84: 65, 67: 71, 71: 67}
    reads = np.zeros(n_reads, dtype=L.READ_DTYPE)
    pieces, ops_all, seq_parts = [], [], []
    seq_off = 0
    clen = int(ref.lengths[0])
    for i in range(n_reads):
        n_seg = int(rng.integers(1, 3))
        fwd = []
        reads[i]["piece_first"] = len(pieces)
        reads[i]["n_pieces"] = 2 * n_seg - 1
        for k in range(2 * n_seg - 1):
            pc = np.zeros((), dtype=L.PIECE_DTYPE)
            pc["read_slot"] = i
            pc["out_rel"] = len(fwd)
            pc["chrom"] = 0
            if k & 1:                                    # a gap piece: not logged
                pc["kind"] = L.NS_PIECE_GAP
                fwd += list(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 7))
                pieces.append(pc)
                continue
            pc["kind"] = L.NS_PIECE_SEGMENT
            pos = int(rng.integers(0, clen))
            pc["pos"] = pos
            script, out, rf = [], [], 0
            if k == 0:
                h = int(rng.integers(0, 4))
                if h:
                    script.append((L.NS_OP_HT << 28) | h)
                    out += list(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), h))
            for _ in range(int(rng.integers(3, 30))):
                ty = int(rng.choice([L.NS_OP_COPY, L.NS_OP_MIS, L.NS_OP_INS, L.NS_OP_DEL]))
                ln = int(rng.integers(1, 6 if ty else 40))
                script.append((ty << 28) | ln)
                if ty == L.NS_OP_INS:
                    out += list(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln))
                    continue
                rb = [int(ref.bases[(pos + rf + t) % clen]) for t in range(ln)]         # circular wrap
                rf += ln
                if ty == L.NS_OP_COPY:
                    out += [(b & 0xDF) if (b & 0xDF) in b"ACGT" else 65 for b in rb]        # IUPAC codes are resolved on the device
                elif ty == L.NS_OP_MIS:
                    out += [int(rng.choice([c for c in b"ACGT" if c != (b & 0xDF)])) for b in rb]
            pc["ref_len"] = rf
            pc["ev_off"] = len(ops_all)
            pc["ev_n_ops"] = len(script)
            ops_all += script
            if rewritten and i % 2:
                pc["op_off"] = len(ops_all)              # a separate emitted script: bases come from the hp stream
                pc["n_ops"] = 1
                ops_all.append((L.NS_OP_HT << 28) | len(out))
            else:
                pc["op_off"], pc["n_ops"] = pc["ev_off"], pc["ev_n_ops"]
            pc["out_len"] = len(out)
            fwd += out
            pieces.append(pc)
        rev = bool(rng.integers(0, 2))
        raw = [comp[b] for b in reversed(fwd)] if rev else fwd
        reads[i]["reversed"] = rev
        reads[i]["seq_len"] = len(raw)
        reads[i]["seq_off"] = seq_off
        pad = (-len(raw)) % 16
        seq_parts.append(np.asarray(raw + [0] * pad, dtype=np.uint8))
        seq_off += len(raw) + pad
    info = type("I", (), {"n_reads": n_reads})()
    return Batch(info, np.concatenate(seq_parts), None, reads, np.asarray(pieces, dtype=L.PIECE_DTYPE),
                 np.asarray(ops_all, dtype=np.uint32), L.NS_KIND_ALIGNED, 1000)


@pytest.mark.parametrize("rewritten", [False, True])
def test_error_profile_formatter_matches_python_rows(rewritten):
    """ns_format_error_profile (host C++, threads) writes exactly the rows of records.error_profile_rows, including reverse
    reads, circular wrap, lower-case reference bases, gap pieces, and events whose bases the homopolymer pass fixed."""
    from nanosim_b200.records import error_profile_rows, format_error_profile
    from nanosim_b200.reference_fasta import PackedReference
    rng = np.random.default_rng(5)
    body = "".join(rng.choice(list("ACGTacgtN"), 3000))
    ref = PackedReference.from_records([("chrT test", body)])
    b = _synthetic_batch(rng, ref, 300, rewritten)
    names = ["chrT_%d_aligned_%d_F_0_1_0" % (i * 7, i) for i in range(300)]
    want = "".join(error_profile_rows(b, names, ref, seed=77)).encode()
    for nt in (1, 5):
        got = format_error_profile(b, names, ref, seed=77, n_threads=nt)
        assert got == want
    assert want.count(b"\n") > 1000
    # intron-retention layout: the second segment continues the first (one mutate_read call: positions keep counting,
    # rows right to left over both), minus-strand pieces show the complemented reference
    from nanosim_b200 import _lib as L
    two = np.flatnonzero(b.reads["n_pieces"] == 3)
    for i in two.tolist():
        p0 = int(b.reads["piece_first"][i])
        flags = L.NS_PIECE_GENOME | (L.NS_PIECE_REF_REV if i % 2 else 0)
        b.pieces["kind"][p0] |= flags
        b.pieces["kind"][p0 + 2] |= flags | L.NS_PIECE_CONT
    want_ir = "".join(error_profile_rows(b, names, ref, seed=77)).encode()
    assert want_ir != want and format_error_profile(b, names, ref, seed=77, n_threads=3) == want_ir


def test_name_formatter_matches_python_names():
    """ns_format_names writes the strings of records.read_names in every mode (chimeric genome, metagenome with gap
    lengths, perfect, transcriptome with polyA, unaligned)."""
    from nanosim_b200 import _lib as L
    from nanosim_b200.records import name_table, read_names
    from nanosim_b200.reference_fasta import PackedReference
    rng = np.random.default_rng(9)
    ref = PackedReference.from_records([("chrT test", "".join(rng.choice(list("ACGT"), 3000))), ("sp-two", "ACGT" * 50)])
    b = _synthetic_batch(rng, ref, 200, False)
    b.pieces["chrom"] = rng.integers(0, 2, len(b.pieces))
    b.pieces["polya_len"] = rng.integers(0, 9, len(b.pieces))
    b.reads["head"] = rng.integers(0, 500, len(b.reads))
    b.reads["tail"] = rng.integers(0, 500, len(b.reads))
    for kw in ({}, {"metagenome": True}, {"perfect": True}, {"transcriptome": True}, {"transcriptome": True, "perfect": True}):
        assert name_table(b, ref.names, 12345, **kw).tolist() == read_names(b, ref.names, 12345, **kw), kw
    two = np.flatnonzero(b.reads["n_pieces"] == 3)           # intron-retention layout of some reads
    for i in two.tolist():
        p0 = int(b.reads["piece_first"][i])
        b.pieces["kind"][p0] |= L.NS_PIECE_GENOME | (L.NS_PIECE_REF_REV if i % 2 else 0) | (L.NS_PIECE_RETAINED if i % 3 else 0)
        b.pieces["kind"][p0 + 2] |= L.NS_PIECE_GENOME | L.NS_PIECE_CONT | (L.NS_PIECE_REF_REV if i % 2 else 0) | (L.NS_PIECE_RETAINED if i % 5 == 0 else 0)
        b.pieces["ref_req"][p0] = b.pieces["ref_req"][p0 + 2] = i % 2
    got, want = name_table(b, ref.names, 99, transcriptome=True).tolist(), read_names(b, ref.names, 99, transcriptome=True)
    assert got == want and any("_RetainedIntron_" in x for x in want)
    b.kind = L.NS_KIND_UNALIGNED
    t = name_table(b, ref.names, 7)
    assert t.tolist() == read_names(b, ref.names, 7) and t[3] == read_names(b, ref.names, 7)[3] and len(t) == 200


def test_intron_retention_host_logic_matches_oracle(monkeypatch):
    """nanosim_b200.intron_retention (GFF3 structures, vectorised IR Markov chain, extract_read_pos, script cutting) against
    the pinned oracle functions fed the same uniforms."""
    import random
    from nanosim_b200 import _lib as L
    from nanosim_b200 import intron_retention as ir
    from nanosim_b200.reference_fasta import PackedReference
    D = os.path.join(GOLDEN, "ir")
    trx = PackedReference.from_fasta(os.path.join(D, "transcripts.fa"))
    oref = no.OracleTrxReference.from_files(os.path.join(D, "transcripts.fa"), os.path.join(D, "expression.tsv"), os.path.join(D, "polya.txt"))
    oref.load_ir(os.path.join(D, "genome.fa"), os.path.join(D, "annotation.gff3"), os.path.join(D, "IR_markov_model"))
    st = ir.TranscriptStructures.from_gff3(os.path.join(D, "annotation.gff3"), trx.names, oref.genome_names)
    p_no_ir = ir.read_ir_markov_model(os.path.join(D, "IR_markov_model"))
    assert p_no_ir.tolist() == [0.6, 0.7, 0.5]
    rng = np.random.default_rng(3)
    checked = 0
    for t, key in enumerate(trx.names):
        feats_o = oref.structure.get(key, [])
        a, b = int(st.first[t]), int(st.first[t + 1])
        assert b - a == len(feats_o) and int(st.n_introns[t]) == sum(1 for x in feats_o if x[0] == "intron")
        for j, f in enumerate(feats_o):                       # same features, genome record instead of the chromosome name
            g = int(st.chrom[a + j])
            assert (("exon", "intron")[st.ftype[a + j]], int(st.start[a + j]), int(st.end[a + j]), "-" if st.minus[a + j] else "+") == (f[0], f[2], f[3], f[5])
            assert g == (oref.genome_names.index("chr" + f[1]) if "chr" + f[1] in oref.genome_names else -1)
        for rep in range(40):
            n_int = int(st.n_introns[t])
            u = rng.random(n_int + 1)
            feed = iter(u[:n_int].tolist())
            monkeypatch.setattr(random, "random", lambda: next(feed))
            flag, st_new = no.update_structure(feats_o, oref.ir_model)
            ret = ir.draw_ir_states(p_no_ir, np.asarray([n_int]), u[None, :n_int]) if n_int else np.zeros((1, 0), dtype=bool)
            assert bool(flag) == bool(ret.any())
            assert [x[0] == "retained_intron" for x in st_new if x[0] != "exon"] == ret[0].tolist()
            if not flag:
                continue
            ref_len = int(trx.lengths[t])
            length = int(rng.integers(1, ref_len))
            monkeypatch.setattr(random, "randint", lambda lo, hi: min(int(u[n_int] * (hi + 1)), hi) if hi > 0 else 0)
            ivs_o, _, ir_o = no.extract_read_pos(length, ref_len, st_new, False)
            feats = list(zip(st.ftype[a:b].tolist(), st.chrom[a:b].tolist(), st.start[a:b].tolist(), st.end[a:b].tolist(), st.minus[a:b].tolist()))
            ivs = ir.extract_read_pos(length, ref_len, feats, ret[0].tolist(), float(u[n_int]))
            assert [(s, e) for _, s, e, _, _ in ivs] == [(s, e) for _, s, e, _ in ivs_o]
            assert [(s, e) for _, s, e, _, r in ivs if r] == [tuple(x) for x in ir_o]
            checked += 1
    assert checked > 100
    # script cutting: the pieces' ops concatenate to an equivalent script and consume exactly their intervals
    ops = np.asarray([(L.NS_OP_HT << 28) | 5, (L.NS_OP_COPY << 28) | 30, (L.NS_OP_INS << 28) | 2, (L.NS_OP_MIS << 28) | 3,
                      (L.NS_OP_COPY << 28) | 17, (L.NS_OP_DEL << 28) | 4, (L.NS_OP_COPY << 28) | 6, (L.NS_OP_LIT << 28) | (3 << 24) | 9,
                      (L.NS_OP_HT << 28) | 7], dtype=np.uint32)
    parts = ir.split_script(ops, [30, 41, 60])
    ref_used = [int(sum((o & 0x0fffffff) for o in p.tolist() if (o >> 28) in (0, 1, 3))) for p in parts]
    assert ref_used == [30, 11, 19]
    assert sum(ir._out_len(p) for p in parts) == ir._out_len(ops)
    assert parts[0][0] == ops[0] and parts[-1][-1] == ops[-1] and (parts[0][2] >> 28) == L.NS_OP_INS


class _FakeEngine:
    """Stands in for Engine in the pipeline tests: simulate() sleeps a job-dependent time, records what ran where."""

    def __init__(self, slot, log, fail_on=None):
        self.slot, self.log, self.fail_on, self.fastq, self.info = slot, log, fail_on, False, None

    def clone(self):
        raise AssertionError("not used")

    def simulate(self, kind, first, n):
        import time
        if self.fail_on is not None and first == self.fail_on:
            raise RuntimeError("boom at %d" % first)
        time.sleep(0.001 * (1 + (first * 7) % 5))
        self.log.append((self.slot, first))
        self.info = type("I", (), {"first": first, "n_reads": n, "seq_bytes": 0, "n_pieces": 0, "n_ops": 0})()
        return self.info

    def close(self):
        pass


def _fake_pipeline(depth, fail_on=None):
    from nanosim_b200.pipeline import BatchPipeline
    log = []
    pipe = BatchPipeline.__new__(BatchPipeline)
    pipe.engines = [_FakeEngine(s, log, fail_on) for s in range(depth)]
    pipe.depth, pipe.fetch, pipe.want_ops, pipe.want_pieces = depth, False, False, True
    pipe.bufs, pipe.hint = [None] * depth, {"seq": 0, "reads": 0, "pieces": 0, "ops": 0}
    return pipe, log


def test_pipeline_orders_results_and_assigns_contexts():
    """BatchPipeline: results are consumed in submission order whatever the completion order; static_assign pins job j to
    context j % depth; the after_simulate hook runs once per job on the job's context; a failing job surfaces."""
    jobs = [(0, i, 1) for i in range(40)]
    pipe, log = _fake_pipeline(3)
    seen, hooked = [], []
    infos = pipe.run(jobs, consume=lambda info, b, job: seen.append(job[1]), after_simulate=lambda e, info, job: hooked.append((e.slot, job[1])))
    assert seen == list(range(40)) and [i.first for i in infos] == list(range(40))
    assert sorted(f for _, f in log) == list(range(40)) and sorted(f for _, f in hooked) == list(range(40))
    assert dict((f, s) for s, f in log) == dict((f, s) for s, f in hooked)
    pipe, log = _fake_pipeline(3)
    pipe.run(jobs, static_assign=True)
    assert all(slot == first % 3 for slot, first in log)
    pipe, log = _fake_pipeline(2, fail_on=7)
    with pytest.raises(RuntimeError, match="boom at 7"):
        pipe.run(jobs)
    assert pipe.run([]) == []


def test_intron_retention_patch_is_consistent():
    """IntronRetention.plan_batch on a hand-built transcriptome batch (no GPU): every patched read keeps its length, its new
    pieces tile the extracted genomic intervals, the cut scripts consume exactly those intervals and produce the same number
    of bases, offsets point behind the batch, and the decision does not depend on which reads share a batch."""
    from nanosim_b200 import _lib as L
    from nanosim_b200 import intron_retention as ir
    from nanosim_b200.reference_fasta import PackedReference
    D = os.path.join(GOLDEN, "ir")
    trx = PackedReference.from_fasta(os.path.join(D, "transcripts.fa"))
    genome = PackedReference.from_fasta(os.path.join(D, "genome.fa"))
    ref = PackedReference.concat(trx, genome)
    assert ref.names[:len(trx.names)] == trx.names and ref.raw_names[len(trx.names):] == ["chr1", "chr2"]
    assert int(ref.offsets[len(trx.names)]) == trx.genome_len and ref.genome_len == trx.genome_len + genome.genome_len
    st = ir.TranscriptStructures.from_gff3(os.path.join(D, "annotation.gff3"), trx.names, genome.raw_names)
    irm = ir.IntronRetention(ir.read_ir_markov_model(os.path.join(D, "IR_markov_model")), st, trx.lengths, len(trx.names))
    rng = np.random.default_rng(11)
    n = 400
    reads = np.zeros(n, dtype=L.READ_DTYPE)
    pieces = np.zeros(n, dtype=L.PIECE_DTYPE)
    ops = []
    for i in range(n):
        t = int(rng.integers(0, len(trx.names)))
        tl = int(trx.lengths[t])
        script, rf, out = [(L.NS_OP_HT << 28) | 4], 0, 4
        want = int(rng.integers(20, tl - 5))
        while rf < want:
            ty = int(rng.choice([L.NS_OP_COPY, L.NS_OP_COPY, L.NS_OP_MIS, L.NS_OP_INS, L.NS_OP_DEL]))
            ln = int(rng.integers(1, 25 if ty == L.NS_OP_COPY else 4))
            if ty != L.NS_OP_INS:
                ln = min(ln, want - rf)
                rf += ln
            if ty != L.NS_OP_DEL:
                out += ln
            script.append((ty << 28) | ln)
        script += [(L.NS_OP_LIT << 28) | (3 << 24) | 6, (L.NS_OP_HT << 28) | 3]
        out += 9
        reads[i]["piece_first"], reads[i]["n_pieces"], reads[i]["seq_len"], reads[i]["head"], reads[i]["tail"] = i, 1, out, 4, 3
        pieces[i]["chrom"], pieces[i]["pos"], pieces[i]["ref_len"], pieces[i]["out_len"], pieces[i]["read_slot"] = t, 0, rf, out, i
        pieces[i]["op_off"] = pieces[i]["ev_off"] = len(ops)
        pieces[i]["n_ops"] = pieces[i]["ev_n_ops"] = len(script)
        pieces[i]["polya_len"] = 6
        ops += script
    ops = np.asarray(ops, dtype=np.uint32)
    patch = irm.plan_batch(reads, pieces, ops, 1000, 5, n, len(ops))
    slots, nr, npc, nops = patch
    assert 0.15 * n < len(slots) < 0.95 * n and len(nr) == len(slots)
    for k, i in enumerate(slots.tolist()):
        r = nr[k]
        assert int(r["seq_len"]) == int(reads[i]["seq_len"]) and int(r["piece_first"]) >= n
        own = npc[int(r["piece_first"]) - n: int(r["piece_first"]) - n + int(r["n_pieces"])]
        segs, gaps = own[::2], own[1::2]
        assert (gaps["kind"] == L.NS_PIECE_GAP).all() and (gaps["out_len"] == 0).all()
        assert (segs["kind"] & L.NS_PIECE_GENOME).all() and (segs["chrom"] >= len(trx.names)).all() and (segs["read_slot"] == i).all()
        assert int(segs["ref_len"].sum()) == int(pieces[i]["ref_len"]) and int(segs["out_len"].sum()) == int(reads[i]["seq_len"])
        assert (segs["kind"][1:] & L.NS_PIECE_CONT).all() and not int(segs["kind"][0]) & L.NS_PIECE_CONT
        rel = 0
        for q in segs:
            part = nops[int(q["op_off"]) - len(ops): int(q["op_off"]) - len(ops) + int(q["n_ops"])]
            ty = part >> 28
            assert int(sum(int(o & 0x0fffffff) for o, t_ in zip(part.tolist(), ty.tolist()) if t_ in (0, 1, 3))) == int(q["ref_len"])
            assert ir._out_len(part) == int(q["out_len"]) and int(q["out_rel"]) == rel
            rel += int(q["out_len"])
            g = int(q["chrom"]) - len(trx.names)
            assert 0 <= int(q["pos"]) and int(q["pos"]) + int(q["ref_len"]) <= int(genome.lengths[g])
    # the same reads in a different batch split decide the same way
    half = irm.plan_batch(reads[:200], pieces[:200], ops, 1000, 5, 200, len(ops))
    assert half[0].tolist() == [s for s in slots.tolist() if s < 200]
    a = npc[:len(half[2])]
    assert np.array_equal(a["pos"], half[2]["pos"]) and np.array_equal(a["ref_len"], half[2]["ref_len"])


def test_ir_expressed_set_needs_matching_exon_structure(tmp_path):
    """simulator.py:1094-1099: with intron retention on, a transcript that is missing from the GFF3, or whose exons do
    not add up to its FASTA length, is drawn again and again -- i.e. never simulated.  The host drops such transcripts
    from the expressed set; the oracle carries the same guard."""
    import random
    from nanosim_b200 import intron_retention as ir
    from nanosim_b200.reference_fasta import PackedReference, read_expression
    D = os.path.join(GOLDEN, "ir")
    fa = open(os.path.join(D, "transcripts.fa")).read()
    # a transcript the annotation does not know, and one whose sequence is one base longer than its exons
    recs = fa.split(">")[1:]
    name0 = recs[0].split("\n", 1)[0].split()[0]
    longer = ">" + recs[0].split("\n", 1)[0] + "\n" + "".join(recs[0].split("\n")[1:]) + "A\n"
    fa2 = longer + "".join(">" + r for r in recs[1:]) + ">ENST00000009999.1\n" + "ACGT" * 300 + "\n"
    tfa = os.path.join(str(tmp_path), "t.fa")
    open(tfa, "w").write(fa2)
    texp = os.path.join(str(tmp_path), "e.tsv")
    open(texp, "w").write(open(os.path.join(D, "expression.tsv")).read() + "ENST00000009999.1\t10.00\t50.0\n")
    trx = PackedReference.from_fasta(tfa)
    chrom, w = read_expression(texp, trx)
    genome = PackedReference.from_fasta(os.path.join(D, "genome.fa"))
    st = ir.TranscriptStructures.from_gff3(os.path.join(D, "annotation.gff3"), trx.names, genome.raw_names)
    c2, w2, dropped = ir.expressed_with_structure(chrom, w, st, trx.lengths)
    bad = {trx.names.index("ENST00000009999"), trx.names.index(name0.split(".")[0])}
    assert dropped == 2 and not (set(c2.tolist()) & bad) and set(c2.tolist()) | bad == set(chrom.tolist())
    assert np.array_equal(w2, w[~np.isin(chrom, list(bad))])
    # the oracle never emits them either (and does not raise KeyError on the unknown transcript)
    from conftest import oracle_model
    from nanosim_b200.model import CompiledModel
    cm = CompiledModel.load(os.path.join(ROOT, "nanosim_b200", "data", "drna_bham1_guppy_plusq.npz"))
    m = oracle_model(cm, tmp_path, fastq=False)
    oref = no.OracleTrxReference.from_files(tfa, texp, os.path.join(D, "polya.txt"))
    oref.load_ir(os.path.join(D, "genome.fa"), os.path.join(D, "annotation.gff3"), os.path.join(D, "IR_markov_model"))
    random.seed(5)
    np.random.seed(5)
    sink = no.ReadSink()
    no.simulation_aligned_transcriptome(oref, m, sink, None, "guppy", 60, False, False, False, False, True)
    used = {name.split("_")[0] for name, _, _ in sink.records}
    assert len(sink.records) == 60 and "ENST00000009999" not in used and name0.split(".")[0] not in used


def test_coverage_counts_the_transcriptome_only_and_abundance_rows_may_be_missing(tmp_path):
    """-x (calculate_read_number_from_coverage, simulator.py:2024-2068, called with ref_t at :2348): the reference size is the
    transcriptome's even when the genome is concatenated behind it for intron retention; the closed-form mean equals the
    reference's Monte-Carlo estimate.  A species without an abundance row is left out (zero), not an error."""
    from types import SimpleNamespace
    from nanosim_b200 import simulator
    from nanosim_b200.model import CompiledModel, DeviceTables
    from nanosim_b200.reference_fasta import MetaReference, PackedReference, read_abundance
    D = os.path.join(GOLDEN, "ir")
    trx = PackedReference.from_fasta(os.path.join(D, "transcripts.fa"))
    genome = PackedReference.from_fasta(os.path.join(D, "genome.fa"))
    cm = CompiledModel.load(os.path.join(ROOT, "nanosim_b200", "data", "drna_bham1_guppy_plusq.npz"))
    t = DeviceTables(cm, fastq=False)
    prof = SimpleNamespace(ref=PackedReference.concat(trx, genome), tables=t, coverage_ref_len=trx.genome_len)
    n = simulator.coverage_to_reads(prof, cm, 30.0)
    # Monte-Carlo estimate as the reference does it (KernelDensity.sample == data[randint] + N(0, bw)), 2M draws
    rng = np.random.default_rng(0)
    rate = t.aligned_ratio
    n_al = int(2000000 * rate / (rate + 1))
    al, bw_al = cm.kde["aligned_reads"]
    un, bw_un = cm.kde["unaligned_length"]
    draws = np.concatenate([al.reshape(-1)[rng.integers(0, al.size, n_al)] + rng.normal(0, bw_al, n_al),
                            un.reshape(-1)[rng.integers(0, un.size, 2000000 - n_al)] + rng.normal(0, bw_un, 2000000 - n_al)])
    want = trx.genome_len / draws.mean() * 30.0
    assert abs(n / want - 1) < 5e-3, (n, want)
    assert n < 0.5 * int(prof.ref.genome_len / draws.mean() * 30.0)          # the genome is not counted
    meta = os.path.join(GOLDEN, "meta")
    from conftest import meta_fixture
    meta_fixture()
    ref = MetaReference.from_genome_list(os.path.join(meta, "genome_list_local.tsv"), os.path.join(meta, "dna_type.tsv"))
    lines = open(os.path.join(meta, "abundance.tsv")).read().splitlines()
    ab = os.path.join(str(tmp_path), "abun.tsv")
    open(ab, "w").write("\n".join(lines[:-1]) + "\n")                        # the last species has no row
    numbers, samples = read_abundance(ab, ref.species)
    missing = MetaReference.species_key(lines[-1].split("\t")[0])
    assert samples[0][ref.species.index(missing)] == 0.0 and sum(samples[0]) > 0


def test_file_sinks_write_what_the_formatters_return(tmp_path):
    """ns_write_records / ns_write_error_profile (formatter threads pwrite() their own stretch of the file at the given
    offset) leave exactly the bytes ns_format_records / ns_format_error_profile return, behind whatever precedes them."""
    from nanosim_b200.records import format_error_profile, format_records, write_error_profile, write_records
    from nanosim_b200.reference_fasta import PackedReference
    rng = np.random.default_rng(15)
    ref = PackedReference.from_records([("chrT", "".join(rng.choice(list("ACGTacgtN"), 3000)))])
    b = _synthetic_batch(rng, ref, 400, False)
    b.qual = (33 + rng.integers(1, 94, len(b.seq))).astype(np.uint8)
    names = ["chrT_%d_aligned_%d_F_0_1_0" % (i * 7, i) for i in range(400)]
    for fastq in (False, True):
        want = format_records(b, names, fastq, n_threads=1)
        for nt in (1, 7):
            path = os.path.join(str(tmp_path), "r%d_%d" % (fastq, nt))
            fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            pos = os.pwrite(fd, b"HEADER\n", 0)
            pos += write_records(fd, pos, b, names, fastq, n_threads=nt)
            os.close(fd)
            assert open(path, "rb").read() == b"HEADER\n" + want and pos == 7 + len(want)
    want = format_error_profile(b, names, ref, seed=3, n_threads=1)
    for nt in (1, 6):
        path = os.path.join(str(tmp_path), "e%d" % nt)
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        n = write_error_profile(fd, 0, b, names, ref, seed=3, n_threads=nt)
        os.close(fd)
        assert n == len(want) > 1000 and open(path, "rb").read() == want


def test_native_fasta_reader_matches_python_reader(tmp_path):
    """ns_read_fasta (mmap, line-aligned chunks over threads) against the pure-Python reader: the repository's fixtures,
    a multi-MB file with CRLF line ends and ragged line lengths that the threads cut anywhere, and a FASTQ reference
    (readfq, simulator.py:709-740) whose quality lines start with '>' and '@'."""
    from nanosim_b200.reference_fasta import PackedReference
    rng = np.random.default_rng(21)
    paths = [os.path.join(GOLDEN, "mini_ref.fa"), os.path.join(GOLDEN, "mini_circular.fa"), os.path.join(GOLDEN, "trx", "transcripts.fa")]
    big = os.path.join(str(tmp_path), "big.fa")
    with open(big, "wb") as f:
        for r in range(40):
            f.write(b">rec_%d some description\r\n" % r)
            seq = np.frombuffer(b"ACGTNacgtRY", dtype=np.uint8)[rng.integers(0, 11, int(rng.integers(1000, 400000)))].tobytes()
            pos = 0
            while pos < len(seq):
                w = int(rng.integers(1, 120))
                f.write(seq[pos:pos + w] + (b"\r\n" if r % 2 else b"\n"))
                pos += w
            if r % 7 == 0:
                f.write(b"\n")
    paths.append(big)
    for path in paths:
        a = PackedReference._from_fasta_python(path)
        for nt in (1, 3, 16):
            b = PackedReference._from_fasta_native(path, nt)
            assert b is not None and a.names == b.names and a.raw_names == b.raw_names
            assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.bases, b.bases), (path, nt)
    fq = os.path.join(str(tmp_path), "ref.fq")
    recs = [("chrA", "ACGTACGTAC" * 30, ">@+!" * 75), ("chrB extra", "GGGTTTAAAC", "@>@>@>@>@>")]
    with open(fq, "w") as f:
        for name, seq, q in recs:
            f.write("@%s\n%s\n%s\n+\n%s\n%s\n" % (name, seq[:150], seq[150:], q[:200], q[200:]) if len(seq) > 150 else "@%s\n%s\n+\n%s\n" % (name, seq, q))
    b = PackedReference._from_fasta_native(fq, 4)
    assert b.names == ["chrA", "chrB"] and b.bases.tobytes().decode() == recs[0][1] + recs[1][1]
    assert b.lengths.tolist() == [300, 10]


def test_unpack_bases_matches_bit_definition(tmp_path):
    """ns_unpack_bases (the host half of ns_fetch's 2-bit transfer): base k sits in bits [2(k&3)+1 : 2(k&3)] of packed[k >> 2],
    A C T G = 0 1 2 3.  AVX2 path (when the CPU has it) and the table path (NANOSIM_B200_NO_AVX2=1, in a subprocess: the
    choice is made once per process), 1 and 5 threads, lengths around the 32-base step, a destination that is not 32-byte
    aligned, both alphabets."""
    import ctypes as C
    import subprocess
    import sys
    from nanosim_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(5)

    def check():
        for n in (0, 1, 31, 32, 33, 100, 4097, (1 << 20) + 13):
            for uracil in (0, 1):
                for threads in (1, 5):
                    for shift in (0, 8):
                        packed = rng.integers(0, 256, (n + 3) // 4 + 8, dtype=np.uint8)
                        k = np.arange(n)
                        want = np.frombuffer(b"ACUG" if uracil else b"ACTG", dtype=np.uint8)[(packed[k >> 2] >> (2 * (k & 3))) & 3]
                        buf = np.zeros(n + 64 + shift, dtype=np.uint8)
                        base = buf.ctypes.data
                        off = (-base) % 32 + shift                      # 32-byte aligned, or 8 past it
                        assert lib.ns_unpack_bases(packed.ctypes.data, base + off, n, uracil, threads) == 0
                        assert np.array_equal(buf[off:off + n], want), (n, uracil, threads, shift)
                        assert not buf[off + n:].any() and not buf[:off].any()

    check()
    if os.environ.get("NANOSIM_B200_NO_AVX2") is None:
        env = dict(os.environ, NANOSIM_B200_NO_AVX2="1", PYTHONPATH=ROOT)
        src = ("import os, sys\nsys.path[:0] = [%r, %r]\nimport pytest\n"
               "sys.exit(pytest.main(['-q', '-x', %r, '-k', 'unpack_bases', '-p', 'no:cacheprovider']))\n"
               % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "test_host_cpu.py")))
        r = subprocess.run([sys.executable, "-c", src], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
