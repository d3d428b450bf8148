"""Run-level statistics extracted from NanoSim-format output files.

Test infrastructure shared by
  * tests/golden/make_golden_runs.py  (applied to outputs of the UNMODIFIED reference), and
  * the parity tests                   (applied to outputs of the CUDA path / the oracle).

Input is what ``simulator.py genome`` writes (README.md:685-701, simulator.py:1437-1443, 1633-1634,
2006-2008): ``<out>_aligned_reads.{fasta,fastq}``, ``<out>_aligned_error_profile`` and
``<out>_unaligned_reads.{fasta,fastq}``.  Output is a dict of integer histograms that can be summed
over chunks/runs and compared with chi-square / relative-rate tests.
"""
import json

import numpy as np

# histogram edges -----------------------------------------------------------------------
LEN_EDGES = np.unique(np.round(np.logspace(1.0, 6.0, 151)).astype(np.int64))      # read / region lengths
HT_EDGES = np.concatenate([[0, 1, 2, 3, 5, 8], np.unique(np.round(np.logspace(1.0, 5.5, 46)).astype(np.int64))])
EV_CAP = 64          # event lengths 1..63, 64+
RUN_CAP = 512        # match runs 0..511, 512+
HP_CAP = 48          # homopolymer run lengths 0..47, 48+
EPR_EDGES = np.concatenate([np.arange(0, 20), np.unique(np.round(np.logspace(np.log10(20), 5, 60)).astype(np.int64))])


def empty():
    return {
        "n_aligned": 0, "n_unaligned": 0,
        "aligned_bases": 0, "unaligned_bases": 0, "ref_bases": 0,
        "head_bases": 0, "tail_bases": 0,
        "strand_R_aligned": 0, "strand_R_unaligned": 0,
        "n_chimeric": 0, "n_segments": 0,
        "events": {"mis": 0, "ins": 0, "del": 0},
        "event_bases": {"mis": 0, "ins": 0, "del": 0},
        "len_aligned": np.zeros(len(LEN_EDGES) + 1, dtype=np.int64),
        "len_unaligned": np.zeros(len(LEN_EDGES) + 1, dtype=np.int64),
        "len_middle_ref": np.zeros(len(LEN_EDGES) + 1, dtype=np.int64),
        "len_head": np.zeros(len(HT_EDGES) + 1, dtype=np.int64),
        "len_tail": np.zeros(len(HT_EDGES) + 1, dtype=np.int64),
        "ev_len": {k: np.zeros(EV_CAP + 1, dtype=np.int64) for k in ("mis", "ins", "del")},
        "match_run": np.zeros(RUN_CAP + 1, dtype=np.int64),
        "first_match": np.zeros(RUN_CAP + 1, dtype=np.int64),
        "events_per_read": np.zeros(len(EPR_EDGES) + 1, dtype=np.int64),
        "mis_sub": np.zeros((4, 4), dtype=np.int64),     # ref base x read base for 1-base mismatches (ACGT order)
        "ins_base": np.zeros(4, dtype=np.int64),
        "qual_middle": np.zeros(94, dtype=np.int64),
        "qual_ht": np.zeros(94, dtype=np.int64),
        "qual_unaligned": np.zeros(94, dtype=np.int64),
        "base_comp_aligned": np.zeros(4, dtype=np.int64),
        "hp_runs": np.zeros(HP_CAP + 1, dtype=np.int64),      # homopolymer run lengths (>= 3) in aligned reads
    }


def hp_run_hist(arr):
    """Histogram of the lengths of maximal runs of equal bytes (only runs >= 3 are counted)."""
    if len(arr) == 0:
        return np.zeros(HP_CAP + 1, dtype=np.int64)
    change = np.flatnonzero(arr[1:] != arr[:-1])
    runs = np.diff(np.concatenate([[-1], change, [len(arr) - 1]]))
    runs = runs[runs >= 3]
    return np.bincount(np.minimum(runs, HP_CAP), minlength=HP_CAP + 1)


def merge(a, b):
    for k, v in b.items():
        if k not in a:
            a[k] = v
            continue
        if isinstance(v, dict):
            for kk in v:
                a[k][kk] = a[k][kk] + v[kk]
        else:
            a[k] = a[k] + v
    return a


def to_jsonable(s):
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, np.ndarray):
            return v.tolist()
        return int(v)
    return conv(s)


def from_jsonable(d):
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, list):
            return np.asarray(v, dtype=np.int64)
        return int(v)
    return conv(d)


def save(s, path, meta=None):
    with open(path, "w") as f:
        json.dump({"meta": meta or {}, "stats": to_jsonable(s)}, f, separators=(",", ":"))


def load(path):
    with open(path) as f:
        d = json.load(f)
    return from_jsonable(d["stats"]), d.get("meta", {})


def _bin(edges, x):
    return int(np.searchsorted(edges, x, side="right"))


_BIDX = {"A": 0, "C": 1, "G": 2, "T": 3}


def _records(path, fastq):
    with open(path) as f:
        while True:
            h = f.readline()
            if not h:
                return
            seq = f.readline().rstrip("\n")
            q = None
            if fastq:
                f.readline()
                q = f.readline().rstrip("\n")
            yield h[1:].rstrip("\n"), seq, q


def parse_aligned_name(name):
    """{chrom}_{pos}[;...]_aligned_{idx}[_chimeric]_{F|R}_{head}_{seg[;seg]}_{tail}  (simulator.py:1390-1402)
    or ..._perfect_{idx}_{F|R}_0_{len}_0 (:1332-1343)."""
    parts = name.rsplit("_", 4)
    tail = int(parts[4])
    segs = [int(x) for x in parts[3].split(";")]
    head = int(parts[2])
    strand = parts[1]
    return head, segs, tail, strand, "_chimeric" in parts[0]


def add_reads(s, path, fastq, aligned=True):
    qa = np.zeros(256, dtype=np.int64)
    qh = np.zeros(256, dtype=np.int64)
    comp = np.zeros(256, dtype=np.int64)
    for name, seq, q in _records(path, fastq):
        L = len(seq)
        if aligned:
            head, segs, tail, strand, chim = parse_aligned_name(name)
            s["n_aligned"] += 1
            s["aligned_bases"] += L
            s["ref_bases"] += sum(segs)
            s["head_bases"] += head
            s["tail_bases"] += tail
            s["strand_R_aligned"] += strand == "R"
            s["n_chimeric"] += chim
            s["n_segments"] += len(segs)
            s["len_aligned"][_bin(LEN_EDGES, L)] += 1
            for m in segs:
                s["len_middle_ref"][_bin(LEN_EDGES, m)] += 1
            s["len_head"][_bin(HT_EDGES, head)] += 1
            s["len_tail"][_bin(HT_EDGES, tail)] += 1
            arr8 = np.frombuffer(seq.encode(), dtype=np.uint8)
            comp += np.bincount(arr8, minlength=256)
            s["hp_runs"] += hp_run_hist(arr8)
            if q is not None:
                qq = np.frombuffer(q.encode(), dtype=np.uint8)
                lead, trail = (tail, head) if strand == "R" else (head, tail)
                qh += np.bincount(qq[:lead], minlength=256) + np.bincount(qq[L - trail:] if trail else qq[:0], minlength=256)
                qa += np.bincount(qq[lead:L - trail], minlength=256)
        else:
            parts = name.rsplit("_", 4)
            s["n_unaligned"] += 1
            s["unaligned_bases"] += L
            s["strand_R_unaligned"] += parts[1] == "R"
            s["len_unaligned"][_bin(LEN_EDGES, L)] += 1
            if q is not None:
                qa += np.bincount(np.frombuffer(q.encode(), dtype=np.uint8), minlength=256)
    if aligned:
        s["qual_middle"] += qa[33:33 + 94]
        s["qual_ht"] += qh[33:33 + 94]
        s["base_comp_aligned"] += comp[[65, 67, 71, 84]]
    else:
        s["qual_unaligned"] += qa[33:33 + 94]
    return s


def _flush_read(s, evs):
    """evs: (pos, type, len) rows of ONE segment of one read, in file order."""
    if not evs:
        return
    evs.reverse()         # rows are written right-to-left (simulator.py:1960)
    s["events_per_read"][_bin(EPR_EDGES, len(evs))] += 1
    cur = 0
    first = True
    for pos, typ, ln in evs:
        run = pos - cur
        tgt = s["first_match"] if first else s["match_run"]
        tgt[min(run, RUN_CAP)] += 1
        first = False
        cur = pos + (ln if typ != "ins" else 0)


def add_error_profile(s, path):
    """Rows: Seq_name Seq_pos error_type error_length ref_base seq_base (simulator.py:1634, 2006-2008).
    The reference writes one read's rows contiguously (segment by segment, right to left)."""
    prev = None
    evs = []
    with open(path) as f:
        first_line = f.readline()
        if not first_line.startswith("Seq_name"):
            f.seek(0)
        for line in f:
            name, pos, typ, ln, refb, seqb = line.rstrip("\n").split("\t")
            pos = int(pos)
            ln = int(ln)
            if name != prev or (evs and pos > evs[-1][0]):     # new read, or next segment of a chimeric read
                _flush_read(s, evs)
                evs = []
                prev = name
            evs.append((pos, typ, ln))
            s["events"][typ] += 1
            s["event_bases"][typ] += ln
            s["ev_len"][typ][min(ln, EV_CAP)] += 1
            if typ == "mis" and ln == 1:
                a, b = _BIDX.get(refb), _BIDX.get(seqb)
                if a is not None and b is not None:
                    s["mis_sub"][a, b] += 1
            elif typ == "ins":
                for ch in seqb:
                    b = _BIDX.get(ch)
                    if b is not None:
                        s["ins_base"][b] += 1
    _flush_read(s, evs)
    return s


def stats_from_prefix(prefix, fastq, with_errors=True):
    ext = ".fastq" if fastq else ".fasta"
    s = empty()
    add_reads(s, prefix + "_aligned_reads" + ext, fastq, aligned=True)
    import os
    if os.path.exists(prefix + "_unaligned_reads" + ext):
        add_reads(s, prefix + "_unaligned_reads" + ext, fastq, aligned=False)
    if with_errors and os.path.exists(prefix + "_aligned_error_profile"):
        add_error_profile(s, prefix + "_aligned_error_profile")
    return s
