"""Shared parity machinery for the GPU tests, __graft_entry__.smoke() and bench.py (test infrastructure).

  * check_edit_scripts : bit-exact.  Re-applies every piece's edit script to the reference on the host with
    mutate_read's semantics (/root/reference/src/simulator.py:1957-2015) and requires the device's bases to agree.
  * batch_stats        : the tests/run_stats.py histograms computed from a fetched device batch, so that device
    output, oracle output and the unmodified reference's output are compared with the same code.
  * compare_stats      : relative-rate and chi-square comparisons.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import run_stats as rs  # noqa: E402

DATA = os.path.join(ROOT, "nanosim_b200", "data")
MODELS = {"guppy": "guppy_fab49712_plusq.npz", "dorado": "dorado_kitv14_v3.2.1.npz", "even": "even_err3152364_v3.2.2.npz",
          "drna": "drna_bham1_guppy_plusq.npz"}

_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in (("A", "T"), ("C", "G")):
    _COMP[ord(_a)], _COMP[ord(_b)] = ord(_b), ord(_a)
_UPPER = np.arange(256, dtype=np.uint8)
_UPPER[ord("a"):ord("z") + 1] -= 32
_IS_ACGT = np.zeros(256, dtype=bool)
_IS_ACGT[[65, 67, 71, 84]] = True
_IUPAC = {"Y": "CT", "R": "AG", "W": "AT", "S": "GC", "K": "TG", "M": "CA", "D": "AGT", "V": "ACG", "H": "ACT",
          "B": "CGT", "N": "ATCG", "X": "ATCG"}
_MEMBER = np.zeros((256, 256), dtype=bool)     # _MEMBER[ref_code, base]
for _c in "ACGT":
    _MEMBER[ord(_c), ord(_c)] = True
for _k, _v in _IUPAC.items():
    for _c in _v:
        _MEMBER[ord(_k), ord(_c)] = True
_BIDX = np.full(256, -1, dtype=np.int64)
for _i, _c in enumerate("ACGT"):
    _BIDX[ord(_c)] = _i


def load_tables(model, **kw):
    from nanosim_b200.model import CompiledModel, DeviceTables

    cm = CompiledModel.load(os.path.join(DATA, MODELS[model]))
    return cm, DeviceTables(cm, **kw)


def make_engine(model, ref, fastq=True, chimeric=False, perfect=False, seed=1, min_len=50, max_len=None, circular=False,
                device=0, unaligned_scripts=False, kmer_bias=0, emit_exact=False, emit_whole=False):
    from nanosim_b200.engine import Engine

    cm, t = load_tables(model, fastq=fastq, chimeric=chimeric, perfect=perfect, homopolymer=bool(kmer_bias))
    eng = Engine(device=device, seed=seed)
    eng.set_reference(ref)
    eng.set_model(t, perfect=perfect)
    eng.configure(circular=circular, perfect=perfect, fastq=fastq, chimeric=chimeric, min_len=min_len,
                  max_len=min(max_len or ref.max_chrom, ref.max_chrom), unaligned_scripts=unaligned_scripts,
                  kmer_bias=kmer_bias, emit_exact=emit_exact, emit_whole=emit_whole)
    return eng, cm, t


def make_meta_engine(ref, abun, fastq=True, chimeric=True, perfect=False, seed=1, min_len=50, max_len=None, model="even"):
    """Metagenome-mode engine on a MetaReference with one sample's abundance vector (percent, species order)."""
    from nanosim_b200.engine import Engine

    cm, t = load_tables(model, fastq=fastq, chimeric=chimeric, perfect=perfect, mode="metagenome")
    eng = Engine(device=0, seed=seed)
    eng.set_reference(ref)
    eng.set_model(t, perfect=perfect)
    inflated = [1 - (1 - a) * t.abun_inflation for a in abun] if chimeric else None
    eng.set_abundance(abun, inflated)
    eng.configure(perfect=perfect, fastq=fastq, chimeric=chimeric, min_len=min_len,
                  max_len=min(max_len or ref.max_chrom, ref.max_chrom), metagenome=True)
    return eng, cm, t


def make_trx_engine(ref, expr_chrom, expr_weights, polya_flags=None, fastq=True, perfect=False, seed=1, polya_scale=0.0,
                    uracil=False, kde2d_sample=400, min_len=50, model="drna", unaligned_scripts=False, trx_records=0, max_len=None):
    """Transcriptome-mode engine (no intron retention)."""
    from nanosim_b200.engine import Engine
    from nanosim_b200.model import build_alias

    cm, t = load_tables(model, fastq=fastq, perfect=perfect)
    eng = Engine(device=0, seed=seed)
    eng.set_reference(ref)
    eng.set_model(t, perfect=perfect)
    pr, al = build_alias(expr_weights)
    eng.set_expression(pr, al, expr_chrom, polya_flags)
    eng.configure(perfect=perfect, fastq=fastq, min_len=min_len, max_len=max_len or ref.max_chrom, transcriptome=True, uracil=uracil,
                  polya_scale=polya_scale, kde2d_sample=kde2d_sample, unaligned_scripts=unaligned_scripts, trx_records=trx_records)
    return eng, cm, t


def _piece_layout(batch, pc, events=False):
    """events=True: the piece's error-event script (ev_off/ev_n_ops); else the script the emit kernel applied."""
    o, n = (int(pc["ev_off"]), int(pc["ev_n_ops"])) if events else (int(pc["op_off"]), int(pc["n_ops"]))
    ops = batch.ops[o:o + n]
    ty = (ops >> 28).astype(np.int64)
    ln = np.where(ty == 5, ops & 0x00ffffff, ops & 0x0fffffff).astype(np.int64)
    out_adv = np.where(ty == 3, 0, ln)
    ref_adv = np.where((ty == 2) | (ty == 4) | (ty == 5), 0, ln)
    out_start = np.concatenate([[0], np.cumsum(out_adv)[:-1]]) if len(ops) else np.zeros(0, dtype=np.int64)
    ref_start = np.concatenate([[0], np.cumsum(ref_adv)[:-1]]) if len(ops) else np.zeros(0, dtype=np.int64)
    return ty, ln, out_adv, ref_adv, out_start, ref_start


def check_edit_scripts(batch, ref, fastq, max_reads=None):
    """Bit-exact structural check of one fetched batch (needs ops + pieces).  Returns number of bases verified."""
    reads, pieces = batch.reads, batch.pieces
    ref_off = ref.offsets.astype(np.int64)
    verified = 0
    n = len(reads) if max_reads is None else min(len(reads), max_reads)
    for i in range(n):
        r = reads[i]
        Lr, so = int(r["seq_len"]), int(r["seq_off"])
        assert so % 16 == 0
        raw = batch.seq[so:so + Lr]
        fwd = _COMP[raw[::-1]] if r["reversed"] else raw
        assert _IS_ACGT[fwd].all(), "non-ACGT base in read %d" % i
        if fastq:
            q = batch.qual[so:so + Lr]
            assert q.min() >= 33 + 1 and q.max() <= 33 + 93, "quality out of [1,93] in read %d" % i
        p0, npc = int(r["piece_first"]), int(r["n_pieces"])
        cursor = 0
        for k in range(npc):
            pc = pieces[p0 + k]
            assert int(pc["read_slot"]) == i
            assert int(pc["out_rel"]) == cursor
            ty, ln, out_adv, ref_adv, out_start, ref_start = _piece_layout(batch, pc)
            assert int(out_adv.sum()) == int(pc["out_len"])
            assert int(ref_adv.sum()) == int(pc["ref_len"])
            assert (ln > 0).all()
            lit = ty == 5
            if lit.any():                               # literal bases of the homopolymer pass (A C T G = 0 1 2 3)
                o_ops = batch.ops[int(pc["op_off"]): int(pc["op_off"]) + int(pc["n_ops"])]
                want = np.frombuffer(b"ACTG", dtype=np.uint8)[((o_ops[lit] >> 26) & 3).astype(np.int64)]
                lo_idx = np.repeat(out_start[lit], ln[lit]) + (np.arange(int(ln[lit].sum())) - np.repeat(np.cumsum(ln[lit]) - ln[lit], ln[lit]))
                got = fwd[cursor:cursor + int(pc["out_len"])][lo_idx]
                assert (got == np.repeat(want, ln[lit])).all(), "literal base mismatch (read %d piece %d)" % (i, k)
            if k == 0 and int(r["head"]) > 0:
                assert ty[0] == 4 and ln[0] == int(r["head"])
            if k == npc - 1 and int(r["tail"]) > 0:
                assert ty[-1] == 4 and ln[-1] == int(r["tail"])
            cstart, clen = int(ref_off[pc["chrom"]]), int(ref_off[pc["chrom"] + 1] - ref_off[pc["chrom"]])
            if int(pc["ref_len"]) > 0:
                circ = getattr(ref, "chrom_circular", None)
                wraps_ok = clen == ref.genome_len if circ is None else bool(circ[pc["chrom"]])
                assert int(pc["pos"]) + int(pc["ref_len"]) <= clen or wraps_ok, "segment leaves its chromosome"
            seg = fwd[cursor:cursor + int(pc["out_len"])]
            use = (ty == 0) | (ty == 1)
            if use.any():
                per_ty = np.repeat(ty[use], ln[use])
                ridx = np.repeat(ref_start[use], ln[use]) + (np.arange(int(ln[use].sum())) - np.repeat(np.cumsum(ln[use]) - ln[use], ln[use]))
                oidx = np.repeat(out_start[use], ln[use]) + (np.arange(int(ln[use].sum())) - np.repeat(np.cumsum(ln[use]) - ln[use], ln[use]))
                if int(pc["kind"]) & 0x80000000:           # NS_PIECE_REF_REV: the genome read backwards, complemented
                    rb = _COMP[_UPPER[ref.bases[cstart + (int(pc["pos"]) + int(pc["ref_len"]) - 1 - ridx) % clen]]]
                else:
                    rb = _UPPER[ref.bases[cstart + (int(pc["pos"]) + ridx) % clen]]
                ob = seg[oidx]
                cp = per_ty == 0
                assert _MEMBER[rb[cp], ob[cp]].all(), "copied base differs from the reference (read %d piece %d)" % (i, k)
                ms = (per_ty == 1) & _IS_ACGT[rb]
                assert (rb[ms] != ob[ms]).all(), "substituted base equals the reference base (read %d)" % i
                verified += len(ridx)
            cursor += int(pc["out_len"])
        assert cursor == Lr
    return verified


def check_fast_unaligned(fast, scripted, ref, fastq):
    """fast: batch from the warp-per-read unaligned path; scripted: same ids/seed through plan/script/emit (with ops)."""
    from nanosim_b200.engine import Batch

    for k in ("seq_len", "reversed", "attempts", "seq_off"):
        assert np.array_equal(fast.reads[k], scripted.reads[k]), "unaligned fast path differs in reads.%s" % k
    for k in ("chrom", "pos", "ref_len", "out_len", "ref_req", "l_new"):
        assert np.array_equal(fast.pieces[k], scripted.pieces[k]), "unaligned fast path differs in pieces.%s" % k
    # both paths feed the same emit kernel with position-indexed randomness: identical bytes
    for i in range(len(fast.reads)):           # (the padding between the reads' 16-byte slots is never written)
        a, n = int(fast.reads["seq_off"][i]), int(fast.reads["seq_len"][i])
        assert np.array_equal(fast.seq[a:a + n], scripted.seq[a:a + n]), "unaligned fast path: bases of read %d differ" % i
        if fastq:
            assert np.array_equal(fast.qual[a:a + n], scripted.qual[a:a + n]), "unaligned fast path: qualities of read %d differ" % i
    hybrid = Batch(fast.info, fast.seq, fast.qual, scripted.reads, scripted.pieces, scripted.ops, fast.kind, fast.first_id)
    nb = check_edit_scripts(hybrid, ref, fastq)
    if fast.ops is not None:            # the fast path's own (less merged) scripts describe the same bases
        nb += check_edit_scripts(fast, ref, fastq)
    return nb


def batch_stats(batch, ref, fastq, s=None):
    """tests/run_stats.py statistics from a fetched device batch (needs ops for aligned batches)."""
    from nanosim_b200 import _lib as L

    s = s or rs.empty()
    reads, pieces = batch.reads, batch.pieces
    ref_off = ref.offsets.astype(np.int64)
    aligned = batch.kind == L.NS_KIND_ALIGNED
    qa = np.zeros(256, dtype=np.int64)
    qh = np.zeros(256, dtype=np.int64)
    comp = np.zeros(256, dtype=np.int64)
    for i in range(len(reads)):
        r = reads[i]
        Lr, so = int(r["seq_len"]), int(r["seq_off"])
        raw = batch.seq[so:so + Lr]
        if not aligned:
            s["n_unaligned"] += 1
            s["unaligned_bases"] += Lr
            s["strand_R_unaligned"] += int(r["reversed"])
            s["len_unaligned"][rs._bin(rs.LEN_EDGES, Lr)] += 1
            if fastq:
                qa += np.bincount(batch.qual[so:so + Lr], minlength=256)
            continue
        head, tail = int(r["head"]), int(r["tail"])
        p0, npc = int(r["piece_first"]), int(r["n_pieces"])
        segs = [pieces[p0 + k] for k in range(0, npc, 2)]
        s["n_aligned"] += 1
        s["aligned_bases"] += Lr
        s["ref_bases"] += sum(int(x["ref_len"]) for x in segs)
        s["head_bases"] += head
        s["tail_bases"] += tail
        s["strand_R_aligned"] += int(r["reversed"])
        s["n_chimeric"] += len(segs) > 1
        s["n_segments"] += len(segs)
        s["len_aligned"][rs._bin(rs.LEN_EDGES, Lr)] += 1
        for x in segs:
            s["len_middle_ref"][rs._bin(rs.LEN_EDGES, int(x["ref_len"]))] += 1
        s["len_head"][rs._bin(rs.HT_EDGES, head)] += 1
        s["len_tail"][rs._bin(rs.HT_EDGES, tail)] += 1
        comp += np.bincount(raw, minlength=256)
        s["hp_runs"] += rs.hp_run_hist(raw)
        if fastq:
            qq = batch.qual[so:so + Lr]
            lead, trail = (tail, head) if r["reversed"] else (head, tail)
            qh += np.bincount(qq[:lead], minlength=256) + np.bincount(qq[Lr - trail:] if trail else qq[:0], minlength=256)
            qa += np.bincount(qq[lead:Lr - trail], minlength=256)
        if batch.ops is None:
            continue
        fwd = _COMP[raw[::-1]] if r["reversed"] else raw
        for pc in segs:
            ty, ln, out_adv, ref_adv, out_start, ref_start = _piece_layout(batch, pc, events=True)
            rewritten = int(pc["ev_off"]) != int(pc["op_off"])      # -hp: output positions of events are not recoverable
            ev = np.nonzero((ty >= 1) & (ty <= 3))[0]
            if len(ev) == 0:
                continue
            s["events_per_read"][rs._bin(rs.EPR_EDGES, len(ev))] += 1
            cstart = int(ref_off[pc["chrom"]])
            clen = int(ref_off[pc["chrom"] + 1]) - cstart
            cur = 0
            first = True
            for j in ev:
                t, n, rp = int(ty[j]), int(ln[j]), int(ref_start[j])
                name = ("mis", "ins", "del")[t - 1]
                s["events"][name] += 1
                s["event_bases"][name] += n
                s["ev_len"][name][min(n, rs.EV_CAP)] += 1
                run = rp - cur
                (s["first_match"] if first else s["match_run"])[min(run, rs.RUN_CAP)] += 1
                first = False
                cur = rp + (n if t != 2 else 0)
                if rewritten:
                    continue
                if t == 1 and n == 1:
                    a = _BIDX[_UPPER[ref.bases[cstart + (int(pc["pos"]) + rp) % clen]]]
                    b = _BIDX[fwd[int(pc["out_rel"]) + int(out_start[j])]]
                    if a >= 0 and b >= 0:
                        s["mis_sub"][a, b] += 1
                elif t == 2:
                    o = int(pc["out_rel"]) + int(out_start[j])
                    bb = _BIDX[fwd[o:o + n]]
                    s["ins_base"] += np.bincount(bb[bb >= 0], minlength=4)
    if aligned:
        s["qual_middle"] += qa[33:33 + 94]
        s["qual_ht"] += qh[33:33 + 94]
        s["base_comp_aligned"] += comp[[65, 67, 71, 84]]
    else:
        s["qual_unaligned"] += qa[33:33 + 94]
    return s


def merge_op_stats(s, d, aligned_comp=False):
    """Fold Engine.op_stats() (device histograms of a whole batch) into a run_stats dict."""
    for k in ("mis", "ins", "del"):
        s["events"][k] += d["events"][k]
        s["event_bases"][k] += d["event_bases"][k]
        s["ev_len"][k] += d["ev_len"][k]
    s["match_run"] += d["match_run"]
    s["first_match"] += d["first_match"]
    # events per segment: exact device counts re-binned on run_stats' edges (the error profile only shows segments
    # with at least one event, so slot 0 stays empty on both sides)
    eps = d["events_per_segment"]
    nz = np.nonzero(eps)[0]
    np.add.at(s["events_per_read"], np.searchsorted(rs.EPR_EDGES, nz, side="right"), eps[nz])
    s["mis_sub"] += d["mis_sub"]
    s["ins_base"] += d["ins_base"]
    if aligned_comp:
        s["base_comp_aligned"] += d["base_comp"]
    return s


def meta_stats(batch, s=None):
    """Length / strand statistics from read + piece metadata only (no sequence or ops needed): fast for 1M reads."""
    from nanosim_b200 import _lib as L

    s = s or rs.empty()
    reads, pieces = batch.reads, batch.pieces
    Ls = reads["seq_len"].astype(np.int64)
    if batch.kind != L.NS_KIND_ALIGNED:
        s["n_unaligned"] += len(reads)
        s["unaligned_bases"] += int(Ls.sum())
        s["strand_R_unaligned"] += int(reads["reversed"].sum())
        s["len_unaligned"] += np.bincount(np.searchsorted(rs.LEN_EDGES, Ls, side="right"), minlength=len(rs.LEN_EDGES) + 1)
        return s
    seg = pieces[pieces["kind"] == L.NS_PIECE_SEGMENT]
    s["n_aligned"] += len(reads)
    s["aligned_bases"] += int(Ls.sum())
    s["ref_bases"] += int(seg["ref_len"].astype(np.int64).sum())
    s["head_bases"] += int(reads["head"].astype(np.int64).sum())
    s["tail_bases"] += int(reads["tail"].astype(np.int64).sum())
    s["strand_R_aligned"] += int(reads["reversed"].sum())
    s["n_chimeric"] += int((reads["n_pieces"] > 1).sum())
    s["n_segments"] += len(seg)
    s["len_aligned"] += np.bincount(np.searchsorted(rs.LEN_EDGES, Ls, side="right"), minlength=len(rs.LEN_EDGES) + 1)
    s["len_middle_ref"] += np.bincount(np.searchsorted(rs.LEN_EDGES, seg["ref_len"].astype(np.int64), side="right"),
                                       minlength=len(rs.LEN_EDGES) + 1)
    s["len_head"] += np.bincount(np.searchsorted(rs.HT_EDGES, reads["head"].astype(np.int64), side="right"),
                                 minlength=len(rs.HT_EDGES) + 1)
    s["len_tail"] += np.bincount(np.searchsorted(rs.HT_EDGES, reads["tail"].astype(np.int64), side="right"),
                                 minlength=len(rs.HT_EDGES) + 1)
    return s


# --------------------------------------------------------------------------------------
# comparisons
# --------------------------------------------------------------------------------------
def chi2_two_sample(a, b, min_count=20):
    """Two-sample chi-square on histograms (bins pooled until both have >= min_count).  Returns (stat, dof, p)."""
    from scipy.stats import chi2

    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    pa, pb = [], []
    ca = cb = 0.0
    for x, y in zip(a, b):
        ca += x
        cb += y
        if ca + cb >= 2 * min_count:
            pa.append(ca)
            pb.append(cb)
            ca = cb = 0.0
    if ca + cb > 0 and pa:
        pa[-1] += ca
        pb[-1] += cb
    pa, pb = np.asarray(pa), np.asarray(pb)
    if len(pa) < 2:
        return 0.0, 0, 1.0
    A, B = pa.sum(), pb.sum()
    stat = float((((np.sqrt(B / A) * pa - np.sqrt(A / B) * pb) ** 2) / (pa + pb)).sum())
    dof = len(pa) - 1
    return stat, dof, float(chi2.sf(stat, dof))


def rates(s):
    rb = max(s["ref_bases"], 1)
    return {k: s["event_bases"][k] / rb for k in ("mis", "ins", "del")}


def compare_stats(a, b, rate_tol, p_min, keys=None, label=""):
    """a, b: run_stats dicts.  Returns list of failure strings (empty == parity)."""
    fails = []
    ra, rb_ = rates(a), rates(b)
    for k in ra:
        if rb_[k] > 0 and abs(ra[k] / rb_[k] - 1) > rate_tol:
            fails.append("%s per-base %s rate %.6g vs %.6g (rel %.3g > %.3g)" % (label, k, ra[k], rb_[k], ra[k] / rb_[k] - 1, rate_tol))
    hist_keys = keys or ["len_aligned", "len_middle_ref", "len_head", "len_tail", "match_run", "first_match",
                         "events_per_read", "len_unaligned"]
    for k in hist_keys:
        if a[k].sum() == 0 or b[k].sum() == 0:
            # an unfilled histogram is a hole in the comparison, not a pass
            fails.append("%s histogram %s is empty on %s" % (label, k, "both sides" if a[k].sum() == b[k].sum() else
                                                             ("the first side" if a[k].sum() == 0 else "the second side")))
            continue
        st, dof, p = chi2_two_sample(a[k], b[k])
        if p < p_min:
            fails.append("%s histogram %s: chi2 %.1f dof %d p %.3g" % (label, k, st, dof, p))
    for k in ("mis", "ins", "del"):
        na, nb = a["ev_len"][k].sum(), b["ev_len"][k].sum()
        if na and nb:
            st, dof, p = chi2_two_sample(a["ev_len"][k], b["ev_len"][k])
            if p < p_min:
                fails.append("%s event-length %s: chi2 %.1f dof %d p %.3g" % (label, k, st, dof, p))
        elif na != nb:
            fails.append("%s event-length %s: %d events on one side, %d on the other" % (label, k, na, nb))
    return fails


def compare_base_choices(a, b, p_min, label=""):
    """Substituted / inserted base choices and read composition (run_stats mis_sub, ins_base, base_comp_aligned).
    mutate_read draws a substituted base uniformly from the three others (simulator.py:1968-1973) and an inserted base
    uniformly from ACGT (:1989-1991)."""
    fails = []
    off = ~np.eye(4, dtype=bool)
    for name, x, y in (("mis_sub", a["mis_sub"][off], b["mis_sub"][off]), ("ins_base", a["ins_base"], b["ins_base"]),
                       ("base_comp_aligned", a["base_comp_aligned"], b["base_comp_aligned"])):
        if x.sum() == 0 or y.sum() == 0:
            fails.append("%s %s is empty on one side" % (label, name))
            continue
        st, dof, p = chi2_two_sample(x, y)
        print(label, name, "chi2 %.1f dof %d p %.3g" % (st, dof, p))
        if p < p_min:
            fails.append("%s %s: chi2 %.1f dof %d p %.3g" % (label, name, st, dof, p))
    if np.trace(a["mis_sub"]) != 0:
        fails.append("%s a substituted base equals its reference base" % label)
    return fails


def golden(path):
    """Golden files are part of the repository: a missing one is a failure, never a skip."""
    assert os.path.exists(path), "golden file %s is missing (regenerate with tests/golden/make_golden_*.py)" % path
    return rs.load(path)


# --------------------------------------------------------------------------------------
# unaligned reads: unaligned_error_list + mutate_read (simulator.py:1784-1830, 1957-1995) as edit statistics
# --------------------------------------------------------------------------------------
def apply_edict_to_tags(e_dict, middle_ref):
    """The string splices of mutate_read (simulator.py:1960-1995: keys right to left, ceil(key), mis replaces, del cuts,
    ins inserts) on a list of provenance tags instead of characters: int k = reference base k, "I" = inserted base,
    ("M", t) = substituted base that replaced tag t."""
    import math

    tags = list(range(middle_ref))
    for key in sorted(e_dict.keys(), reverse=True):
        kind, length = e_dict[key]
        k = math.ceil(key)
        if kind == "mis":
            tags[k:k + length] = [("M", t) for t in tags[k:k + length]]
        elif kind == "del":
            del tags[k:k + length]
        else:
            tags[k:k] = ["I"] * length
    return tags


def tags_to_ops(tags, middle_ref):
    """Provenance tags -> [(type, length)] with the device's op types (COPY 0, MIS 1, INS 2, DEL 3), adjacent equal types
    merged.  A substituted inserted base is still an inserted (random) base."""
    ops = []

    def put(t, n):
        if n <= 0:
            return
        if ops and ops[-1][0] == t:
            ops[-1][1] += n
        else:
            ops.append([t, n])

    nxt = 0
    for t in tags:
        if t == "I" or (isinstance(t, tuple) and t[1] == "I"):
            put(2, 1)
            continue
        mis = isinstance(t, tuple)
        r = t[1] if mis else t
        put(3, r - nxt)
        put(1 if mis else 0, 1)
        nxt = r + 1
    put(3, middle_ref - nxt)
    return ops


def canonical_indels(ops):
    """Inserted and deleted bases that touch each other have no defined order (the read is the same whether the gap is
    written DEL INS DEL or INS DEL): every maximal run of INS / DEL ops becomes one INS followed by one DEL."""
    out, ins, dele = [], 0, 0
    for t, n in list(ops) + [(0, 0)]:
        if t == 2:
            ins += n
        elif t == 3:
            dele += n
        else:
            if ins:
                out.append((2, ins))
            if dele:
                out.append((3, dele))
            ins = dele = 0
            out.append((t, n))
    return out[:-1]


def script_stats(op_lists, canonical=False):
    """Event statistics of a list of op lists [(type, len), ...]: the run_stats event keys."""
    s = rs.empty()
    for ops in op_lists:
        if canonical:
            ops = canonical_indels(ops)
        merged = []
        for t, n in ops:
            if n == 0:
                continue
            if merged and merged[-1][0] == t:
                merged[-1][1] += n
            else:
                merged.append([t, n])
        run, first, n_ev, ref = 0, True, 0, 0
        for t, n in merged:
            if t == 0:
                run += n
                ref += n
                continue
            name = ("mis", "ins", "del")[t - 1]
            s["events"][name] += 1
            s["event_bases"][name] += n
            s["ev_len"][name][min(n, rs.EV_CAP)] += 1
            (s["first_match"] if first else s["match_run"])[min(run, rs.RUN_CAP)] += 1
            first, run = False, 0
            n_ev += 1
            if t != 2:
                ref += n
        s["ref_bases"] += ref
        if n_ev:
            s["events_per_read"][rs._bin(rs.EPR_EDGES, n_ev)] += 1
    return s


def device_piece_ops(batch, pc, events=True):
    ty, ln, _, _, _, _ = _piece_layout(batch, pc, events=events)
    return [(int(t), int(n)) for t, n in zip(ty, ln)]


# --------------------------------------------------------------------------------------
# oracle runs
# --------------------------------------------------------------------------------------
def oracle_stats(cm, ref_records, n_aligned, n_unaligned, fastq, chimeric=False, seed=1234, tmpdir=None, kmer_bias=None):
    """Runs the pure-Python oracle (oracle/nanosim_oracle.py) and returns its run_stats via the same text files the
    reference would write."""
    import random
    import tempfile

    import nanosim_oracle as no
    from conftest import oracle_model

    tmp = tmpdir or tempfile.mkdtemp(prefix="oracle_run_")
    m = oracle_model(cm, tmp, fastq=fastq, chimeric=chimeric, homopolymer=bool(kmer_bias))
    oref = no.OracleReference(ref_records)
    random.seed(seed)
    np.random.seed(seed)
    sink = no.ReadSink()
    no.simulation_aligned_genome(oref, m, sink, "linear", 50, oref.max_chrom, None, None, kmer_bias, fastq, n_aligned, False, chimeric)
    ext = ".fastq" if fastq else ".fasta"
    prefix = os.path.join(tmp, "oracle")
    with open(prefix + "_aligned_reads" + ext, "w") as f:
        f.write(no.format_records(sink.records, fastq))
    with open(prefix + "_aligned_error_profile", "w") as f:
        f.write("Seq_name\tSeq_pos\terror_type\terror_length\tref_base\tseq_base\n")
        f.writelines(r + "\n" for r in sink.error_rows)
    if n_unaligned:
        sink2 = no.ReadSink()
        sink2.next_index = sink.next_index
        no.simulation_unaligned(oref, m, sink2, "linear", 50, oref.max_chrom, None, None, fastq, n_unaligned)
        with open(prefix + "_unaligned_reads" + ext, "w") as f:
            f.write(no.format_records(sink2.records, fastq))
    return rs.stats_from_prefix(prefix, fastq)


def smoke_check(verbose=False):
    from nanosim_b200 import _lib as L
    from nanosim_b200.reference_fasta import PackedReference

    ref = PackedReference.from_fasta(os.path.join(HERE, "golden", "mini_ref.fa"))
    eng, cm, t = make_engine("guppy", ref, fastq=True, seed=7, unaligned_scripts=True)
    info = eng.simulate(L.NS_KIND_ALIGNED, 0, 3000)
    b = eng.fetch(want_ops=True)
    nb = check_edit_scripts(b, ref, True)
    s_gpu = batch_stats(b, ref, True)
    info_u = eng.simulate(L.NS_KIND_UNALIGNED, 0, 300)
    bu = eng.fetch(want_ops=True)
    nb += check_edit_scripts(bu, ref, True)
    # the warp-per-read fast path must give the same reads (lengths, strands, positions) and bases that satisfy the
    # scripted path's edit scripts
    eng.configure(fastq=True, min_len=50, max_len=ref.max_chrom, unaligned_scripts=False)
    eng.simulate(L.NS_KIND_UNALIGNED, 0, 300)
    bf = eng.fetch()
    nb += check_fast_unaligned(bf, bu, ref, True)
    import nanosim_oracle as no

    recs = no.read_fasta(os.path.join(HERE, "golden", "mini_ref.fa"))
    s_or = oracle_stats(cm, recs, 250, 0, True)
    rg, ro = rates(s_gpu), rates(s_or)
    if verbose:
        print("smoke: %d reads, %d bases, %d ops, kernels %.2f ms (plan %.2f, script %.2f, emit %.2f); %d bases verified bit-exactly"
              % (info.n_reads, info.total_bases, info.n_ops, info.ms_total, info.ms_plan, info.ms_script, info.ms_emit, nb))
        print("smoke: per-base rates gpu %s" % {k: round(v, 5) for k, v in rg.items()})
        print("smoke: per-base rates oracle %s" % {k: round(v, 5) for k, v in ro.items()})
    for k in rg:
        assert abs(rg[k] / ro[k] - 1) < 0.15, "rate %s: gpu %g oracle %g" % (k, rg[k], ro[k])
    eng.close()
    return True
