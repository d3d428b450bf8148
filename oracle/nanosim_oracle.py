"""CPU oracle: a plain-Python restatement of NanoSim's per-read simulation path.

TEST INFRASTRUCTURE ONLY -- never imported by nanosim_b200/ (the product).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

Parity status: PINNED.  Every function below consumes the two RNG streams the reference uses
(stdlib ``random`` and numpy's legacy global ``np.random``) in the same order as the reference,
so under a fixed seed it reproduces the reference's outputs bit-for-bit.  tests/test_oracle_golden.py
checks that against vectors produced by calling the UNMODIFIED reference functions
(tests/golden/make_golden_vectors.py, run in the build container where /root/reference exists).

Each function cites the reference file:line (paths under /root/reference/src/) it restates.
The restatement keeps the reference's quirks on purpose (they shape the statistics the CUDA
path has to match); they are listed in DESIGN.md.
"""
import math
import random
import re

import numpy as np

ACGT_ORDER = ["A", "T", "C", "G"]          # simulator.py:49 (order matters for random.choice)

IUPAC = {                                   # simulator.py:744-746
    "Y": ["C", "T"], "R": ["A", "G"], "W": ["A", "T"], "S": ["G", "C"], "K": ["T", "G"],
    "M": ["C", "A"], "D": ["A", "G", "T"], "V": ["A", "C", "G"], "H": ["A", "C", "T"],
    "B": ["C", "G", "T"], "N": ["A", "T", "C", "G"], "X": ["A", "T", "C", "G"],
}
_COMP = {"A": "T", "T": "A", "C": "G", "G": "C"}  # simulator.py:1676


# --------------------------------------------------------------------------------------
# model files
# --------------------------------------------------------------------------------------
def parse_ecdf(lines):
    """simulator.py:194-231 (read_ecdf).  Returns {(bin_lo, bin_hi): [(cdf_lo, cdf_hi, len_lo, len_hi), ...]}
    with bins and intervals in the reference's dict insertion order."""
    it = iter(lines)
    head = next(it).strip().split()
    bins = []
    for tok in head[1:]:
        a, b = tok.split("-")
        bins.append((int(a), int(b)))
    table = {b: {} for b in bins}
    order = sorted(table.keys())
    lanes = len(bins)
    last_p = [0.0] * lanes
    last_len = [0.0] * lanes
    row_hi = 0.0
    for line in it:
        cols = line.strip().split("\t")
        lo, hi = (float(x) for x in cols[0].split("-"))
        row_hi = hi
        probs = [float(x) for x in cols[1:]]
        for i in range(lanes):
            if probs[i] == last_p[i]:
                continue
            if last_p[i] != 0:
                table[order[i]][(last_p[i], probs[i])] = (last_len[i], hi)
            else:
                table[order[i]][(last_p[i], probs[i])] = (max(last_len[i], hi - 10 * (hi - lo)), hi)
            last_len[i] = hi
            last_p[i] = probs[i]
    for b in order:
        last_key = sorted(table[b].keys())[-1]
        v = table[b][last_key]
        table[b][last_key] = (v[0], row_hi)
    return {b: [(k[0], k[1], v[0], v[1]) for k, v in table[b].items()] for b in table}


class OracleKDE:
    """sklearn KernelDensity.sample (gaussian) replayed on numpy's global RNG (third-party
    algorithm, scikit-learn 0.22.1 _kde.py; call sites simulator.py:235)."""

    def __init__(self, data, bandwidth):
        self.data = np.asarray(data, dtype=np.float64).reshape(len(data), -1)
        self.bandwidth = float(bandwidth)

    def sample(self, n):
        rng = np.random.mtrand._rand
        u = rng.uniform(0, 1, size=n)
        i = (u * self.data.shape[0]).astype(np.int64)
        return np.atleast_2d(rng.normal(self.data[i], self.bandwidth))


def kde_lengths(kde, n, log=False, flatten=True):
    """simulator.py:234-241 (get_length_kde)."""
    x = kde.sample(n)
    if log:
        x = np.power(10, x) - 1
    return x.flatten() if flatten else x


class OracleModel:
    """The module globals read_profile() fills (simulator.py:244-591), as one object."""

    def __init__(self):
        self.strandness_rate = 0.5
        self.error_par = {}            # {"mis": [lambda,k,prob,weight], "ins":..., "del":...}
        self.trans_error_pr = {}       # {state: [((lo,hi), name), ...]} insertion order mis, ins, del
        self.first_match = None        # parse_ecdf dict (one bin)
        self.match_markov = None       # parse_ecdf dict (14-15 bins)
        self.kde_aligned = None
        self.kde_ht = None
        self.kde_ht_ratio = None
        self.kde_unaligned = None
        self.kde_gap = None
        self.kde_aligned_2d = None
        self.segment_mean = None
        self.abun_inflation = None
        self.aligned_ratio = None      # "Aligned / Unaligned ratio" (None == "100%")
        self.base_qual = None          # {"mis": {"sd","loc","mu"}, ...}
        self.pw_hp_len = None
        self.lr_hp_len = None
        self.hp_mis_rate = None

    @staticmethod
    def load_text_tables(prefix, model=None, homopolymer=False, fastq=False, chimeric=False, mode="genome"):
        """Text model files only (simulator.py:470-542, 569-591).  KDEs are attached by the caller."""
        m = model or OracleModel()
        with open(prefix + "_strandness_rate") as f:            # :272-273
            m.strandness_rate = float(f.readline().split("\t")[1])
        with open(prefix + "_model_profile") as f:              # :475-484
            f.readline()
            for line in f:
                cols = line.strip().split("\t")
                if "mismatch" in line:
                    m.error_par["mis"] = [float(x) for x in cols[1:]]
                elif "insertion" in line:
                    m.error_par["ins"] = [float(x) for x in cols[1:]]
                else:
                    m.error_par["del"] = [float(x) for x in cols[1:]]
        with open(prefix + "_error_markov_model") as f:         # :487-495
            f.readline()
            for line in f:
                c = line.strip().split()
                p_mis, p_ins, p_del = float(c[1]), float(c[2]), float(c[3])
                m.trans_error_pr[c[0]] = [((0, p_mis), "mis"), ((p_mis, p_mis + p_ins), "ins"),
                                          ((1 - p_del, 1), "del")]
        with open(prefix + "_first_match.hist") as f:           # :497-498
            m.first_match = parse_ecdf(f.readlines())
        with open(prefix + "_match_markov_model") as f:         # :500-501
            m.match_markov = parse_ecdf(f.readlines())
        if homopolymer:                                         # :504-529
            with open(prefix + "_hp_lengths_model_parameters.tsv") as f:
                m.pw_hp_len, m.lr_hp_len = {}, {}
                m.hp_mis_rate = float(re.search(r"\d+\.?\d*", next(f))[0])
                names = next(f).strip().split("\t")
                for line in f:
                    cols = line.strip().split("\t")
                    m.pw_hp_len[cols[0]], m.lr_hp_len[cols[0]] = {}, {}
                    for i, nm in enumerate(names):
                        if i == 0:
                            continue
                        (m.lr_hp_len if nm in ("intercept", "slope") else m.pw_hp_len)[cols[0]][nm] = float(cols[i])
        with open(prefix + "_reads_alignment_rate") as f:       # :535-542
            rate = f.readline().strip().split("\t")[1]
            m.aligned_ratio = None if rate == "100%" else float(rate)
        if chimeric:                                            # :571-577
            with open(prefix + "_chimeric_info") as f:
                m.segment_mean = float(f.readline().split("\t")[1])
                if mode == "metagenome":
                    m.abun_inflation = float(f.readline().split("\t")[1])
        if fastq:                                               # :580-591
            with open(prefix + "_base_qualities_model_parameters.tsv") as f:
                next(f)
                m.base_qual = {}
                for line in f:
                    c = line.split("\t")
                    m.base_qual[c[0]] = {"sd": float(c[1]), "loc": float(c[2]), "mu": float(c[3])}
        return m

    def split_counts(self, number, perfect=False):
        """simulator.py:465-468, 538-542: number of aligned / unaligned reads."""
        if perfect or self.aligned_ratio is None:
            return number, 0
        r = self.aligned_ratio
        n_al = int(round(number * r / (r + 1)))
        return n_al, number - n_al


# --------------------------------------------------------------------------------------
# sampling primitives
# --------------------------------------------------------------------------------------
def pois_geom(lam, prob, weight):
    """mixed_model.py:41-49."""
    if np.random.random() < weight:
        return np.random.poisson(lam) + 1
    return np.random.geometric(prob)


def wei_geom(lam, k, prob, weight):
    """mixed_model.py:52-63."""
    if np.random.random() < weight:
        v = int(round(math.ceil(lam * np.random.weibull(k))))
    else:
        v = np.random.geometric(prob) - 1
    return 1 if v == 0 else v


def base_qualities(par, n):
    """model_base_qualities.py:9-20,120-130: truncated log-normal on [1, 93], floored.
    scipy's rv_discrete.rvs draws ``uniform(size=n)`` from numpy's global RNG, maps it through
    _ppf, adds loc, and casts to int64 (scipy 1.7.3 _distn_infrastructure.py rvs)."""
    from scipy.stats import lognorm

    sd, loc, scale = par["sd"], par["loc"], np.exp(par["mu"])
    u = np.random.mtrand._rand.uniform(size=n)
    fa = lognorm.cdf(1, sd, scale=scale)
    fb = lognorm.cdf(93, sd, scale=scale)
    y = lognorm.ppf(u * (fb - fa) + fa, sd, scale=scale)
    return (y + loc).astype(np.int64).tolist()


def hp_normal_params(length, pw, lr):
    """model_homopolymer_lengths.py:167-186, 204-209, 246-260 -> {"AT": (mu, sigma), "CG": ...}."""
    out = {}
    for cls in pw:
        p = pw[cls]
        mu = float(p["const"]) + float(p["alpha1"]) * length
        bps = [p[k] for k in p if "breakpoint" in k]
        betas = [p[k] for k in p if "beta" in k]
        for bp, beta in zip(bps, betas):
            mu += beta * np.maximum(length - bp, 0)
        out[cls] = (mu, lr[cls]["intercept"] + lr[cls]["slope"] * length)
    return out


# --------------------------------------------------------------------------------------
# per-read kernels
# --------------------------------------------------------------------------------------
def _pick_interval(items, p):
    """The reference's ``for k2, v2 in d.items(): if k2[0] < p <= k2[1]`` scan with interpolation."""
    for clo, chi, vlo, vhi in items:
        if clo < p <= chi:
            return int(np.floor((p - clo) / (chi - clo) * (vhi - vlo) + vlo))
    return None


def error_list(m_ref, model, fastq):
    """simulator.py:1833-1916."""
    l_new = m_ref
    pos = 0
    e_dict = {}
    middle_ref = m_ref
    prev_error = "start"
    e_count = {"mis": 0, "ins": 0, "match": 0}

    p = random.random()
    fm_items = model.first_match[list(model.first_match.keys())[0]]
    prev_match = None
    for clo, chi, vlo, vhi in fm_items:                 # no break in the reference (:1845-1849)
        if clo < p <= chi:
            prev_match = int(np.floor((p - clo) / (chi - clo) * (vhi - vlo) + vlo))
            if prev_match < 2:
                prev_match = 2
    pos += prev_match
    if fastq:
        e_count["match"] += middle_ref if prev_match > middle_ref else prev_match

    error = None
    step = None
    bins = list(model.match_markov.keys())
    while pos < middle_ref:
        p = random.random()
        for (lo, hi), name in model.trans_error_pr[prev_error]:
            if lo <= p < hi:
                error = name
                break
        par = model.error_par[error]
        if error == "mis":
            step = pois_geom(par[0], par[2], par[3])
        elif error == "ins":
            step = wei_geom(par[0], par[1], par[2], par[3])
            l_new += step
        else:
            step = wei_geom(par[0], par[1], par[2], par[3])
            l_new -= step

        if error != "ins":
            e_dict[pos] = [error, step]
            pos += step
            if pos >= middle_ref:
                l_new += pos - middle_ref
                middle_ref = pos
        else:
            e_dict[pos - 0.5] = [error, step]
        prev_error = error
        if fastq and error in ("mis", "ins"):
            e_count[error] += step

        k1 = bins[-1]
        for b in bins:
            if b[0] <= prev_match < b[1]:
                k1 = b
                break
        p = random.random()
        got = _pick_interval(model.match_markov[k1], p)
        if got is not None:                                # ECDF miss keeps the error length (:1895-1898)
            step = got
        if prev_match == 0 and step == 0:
            step = 1
        prev_match = step
        if fastq:
            e_count["match"] += step
        if pos + prev_match > middle_ref:
            l_new += pos + prev_match - middle_ref
            middle_ref = pos + prev_match
        pos += prev_match
        if prev_match == 0:
            prev_error += "0"
    return l_new, middle_ref, e_dict, e_count


_UNALIGNED_TYPES = [((0, 0.4), "match"), ((0.4, 0.7), "mis"), ((0.7, 0.85), "ins"), ((0.85, 1), "del")]


def unaligned_error_list(m_ref, model):
    """simulator.py:1784-1830."""
    l_new = m_ref
    e_dict = {}
    pos = 0
    middle_ref = m_ref
    last_is_ins = False
    e_count = {"match": 0, "mis": 0, "ins": 0}
    if m_ref == 0:
        return l_new, middle_ref, e_dict, e_count
    kind = None
    while pos < middle_ref:
        p = random.random()
        for (lo, hi), name in _UNALIGNED_TYPES:
            if lo <= p < hi:
                kind = name
                break
        if kind == "match":
            step = 1
        elif kind == "mis":
            par = model.error_par["mis"]
            step = pois_geom(par[0], par[2], par[3])
            e_dict[pos] = ["mis", step]
        elif kind == "ins":
            par = model.error_par["ins"]
            step = wei_geom(par[0], par[1], par[2], par[3])
            if last_is_ins:
                e_dict[pos + 0.1][1] += step
            else:
                e_dict[pos + 0.1] = ["ins", step]
                last_is_ins = True
            l_new += step
        else:
            par = model.error_par["del"]
            step = wei_geom(par[0], par[1], par[2], par[3])
            e_dict[pos] = ["del", step]
            l_new -= step
        if kind != "ins":
            pos += step
            last_is_ins = False
        if pos > middle_ref:
            l_new += pos - middle_ref
            middle_ref = pos
    return l_new, middle_ref, e_dict, e_count


def case_convert(seq):
    """simulator.py:743-755."""
    out = list(seq.upper())
    for i, c in enumerate(out):
        if c in IUPAC:
            out[i] = random.choice(IUPAC[c])
    return "".join(out)


def reverse_complement(seq):
    """simulator.py:1675-1680."""
    return "".join(_COMP.get(b, b) for b in reversed(seq))


def _hp_pattern(k):
    return "|".join(b + "{" + re.escape(str(k)) + ",}" for b in "ACGT")   # simulator.py:624-625, 1921-1922


def mutate_read(read, read_name, error_log, e_dict, e_count, fastq, k, model):
    """simulator.py:1919-2015.  ``error_log`` is a list receiving the TSV rows (or None)."""
    if k:
        hp_pos = [(m.start(), m.end()) for m in re.finditer(_hp_pattern(k), read)]
        kept = {}
        for start in e_dict.keys():
            kind, length = e_dict[start]
            end = start + length
            hit = False
            for hs, he in hp_pos:
                if not (he <= start or end <= hs):
                    hit = True
                    if fastq:
                        if kind != "ins":
                            e_count["match"] += length
                        if kind != "del":
                            e_count[kind] -= length
                    break
            if not hit:
                kept[start] = [kind, length]
    else:
        kept = e_dict

    if fastq:
        mis_q = base_qualities(model.base_qual["mis"], e_count["mis"])
        ins_q = base_qualities(model.base_qual["ins"], e_count["ins"])
        match_q = base_qualities(model.base_qual["match"], e_count["match"])

    quals = []
    prev = len(read)
    for key in sorted(kept.keys(), reverse=True):
        kind, length = kept[key]
        key = math.ceil(key)
        err_q = []
        if kind == "mis":
            ref_base = read[key: key + length]
            new_bases = ""
            for i in range(length):
                cand = list(ACGT_ORDER)
                cand.remove(read[key + i])
                new_bases += random.choice(cand)
                if fastq:
                    err_q.append(mis_q.pop())
            new_read = read[:key] + new_bases + read[key + length:]
            err_end = key + length
        elif kind == "del":
            new_bases = length * "-"
            ref_base = read[key: key + length]
            new_read = read[:key] + read[key + length:]
            err_end = key + length
        else:
            ref_base = length * "-"
            new_bases = ""
            for i in range(length):
                new_bases += random.choice(ACGT_ORDER)
                if fastq:
                    err_q.append(ins_q.pop())
            new_read = read[:key] + new_bases + read[key:]
            err_end = key
        if fastq:
            if err_end != prev:
                for _ in range(prev - err_end):
                    quals.append(match_q.pop())
            quals += err_q
        read = new_read
        prev = key
        if error_log is not None:
            error_log.append("\t".join([read_name, str(key), kind, str(length), ref_base, new_bases]))
    if fastq:
        while len(match_q) > 0:
            quals.append(match_q.pop())
    quals.reverse()
    return read, quals


def mutate_homo(seq, base_quals, k, model):
    """simulator.py:618-705."""
    runs = []
    hist = {}
    for m in re.finditer(_hp_pattern(k), seq):
        length = m.end() - m.start()
        base = m.group()[0]
        runs.append((base, m.start(), m.end()))
        hist.setdefault(length, {"A": 0, "T": 0, "C": 0, "G": 0})[base] += 1

    samples = {}
    for length in hist.keys():
        par = hp_normal_params(length, model.pw_hp_len, model.lr_hp_len)
        samples[length] = {}
        for base in ("A", "T", "C", "G"):
            if hist[length][base] > 0:
                mu, sigma = par["AT" if base in "AT" else "CG"]
                samples[length][base] = np.random.normal(mu, sigma, hist[length][base])
    for length in samples:
        for base in samples[length]:
            samples[length][base] = [0 if x < 0 else x for x in samples[length][base]]

    last = 0
    out = ""
    shift = 0
    for base, hs, he in runs:
        ref_len = he - hs
        size = int(round(samples[ref_len][base][-1]))
        samples[ref_len][base] = samples[ref_len][base][:-1]
        new_run = ""
        mis_pos = []
        for i in range(size):
            p = random.random()
            if 0 < p <= model.hp_mis_rate:
                while True:
                    nb = random.choice(list(ACGT_ORDER))
                    if nb != base:
                        break
                new_run += nb
                mis_pos.append(i)
            else:
                new_run += base
        out = out + seq[last:hs] + new_run
        if len(base_quals) != 0:
            diff = size - ref_len
            if diff < 0:
                for _ in range(-diff):
                    base_quals.pop(hs + shift)
            elif diff > 0:
                ins_q = base_qualities(model.base_qual["ins"], diff)
                base_quals = base_quals[:he + shift] + ins_q + base_quals[he + shift:]
            if len(mis_pos) != 0:
                mq = base_qualities(model.base_qual["mis"], 1)
                for i, q in zip(mis_pos, mq):
                    base_quals[hs + shift + i] = q
        shift += size - ref_len
        last = he
    return out + seq[last:], base_quals


class OracleReference:
    """seq_dict / seq_len / genome_len / max_chrom for genome mode (simulator.py:341-356)."""

    def __init__(self, seqs):
        self.seq_dict = dict(seqs)                      # name -> str, file order
        self.seq_len = {k: len(v) for k, v in self.seq_dict.items()}
        self.genome_len = sum(self.seq_len.values())
        self.max_chrom = max(self.seq_len.values()) if self.seq_len else 0

    @staticmethod
    def from_fasta(path):
        return OracleReference(read_fasta(path))


def read_fasta(path):
    """simulator.py:709-740 (readfq) + :344-347 name normalisation, FASTA records only."""
    out = []
    name, chunks = None, []
    with open(path) as f:
        for line in f:
            if line[:1] == ">":
                if name is not None:
                    out.append((name, "".join(chunks)))
                raw = line[1:-1].partition(" ")[0]
                name = "-".join(re.split(r"[_\s]\s*", raw)).split(".")[0]
                chunks = []
            else:
                chunks.append(line[:-1] if line.endswith("\n") else line)
    if name is not None:
        out.append((name, "".join(chunks)))
    return out


def extract_read(ref, dna_type, length):
    """simulator.py:1750-1781 (genome branches of extract_read) -> (sequence, "chrom_pos")."""
    if dna_type == "circular":
        pos = random.randint(0, ref.genome_len)
        chrom = list(ref.seq_dict.keys())[0]
        name = chrom + "_" + str(pos)
        s = ref.seq_dict[chrom]
        if length + pos <= ref.genome_len:
            return s[pos: pos + length], name
        return s[pos:] + s[0: length - ref.genome_len + pos], name
    while True:
        got = ""
        pos = random.randint(0, ref.genome_len)
        for key in ref.seq_len:
            if pos + length <= ref.seq_len[key]:
                got = ref.seq_dict[key][pos: pos + length]
                name = key + "_" + str(pos)
                break
            elif pos < ref.seq_len[key]:
                break
            else:
                pos -= ref.seq_len[key]
        if got != "":
            return got, name


# --------------------------------------------------------------------------------------
# per-mode read loops (genome)
# --------------------------------------------------------------------------------------
def lengths_and_ht_ratios(model, remaining):
    """simulator.py:1456-1479."""
    mult_r, mult_h = 1.3, 1.5
    rem, ratio = [], []
    it = 0
    while len(rem) < remaining or len(ratio) < remaining:
        if it > 50:
            break
        rem = [x for x in kde_lengths(model.kde_ht, int(remaining * mult_r), True) if x >= 0]
        ratio = [x for x in kde_lengths(model.kde_ht_ratio, int(remaining * mult_h)) if 0 <= x <= 1]
        mult_r *= 1.5
        mult_h *= 1.5
        it += 1
    return rem, ratio


def simulation_gap(ref, model, length, dna_type, fastq):
    """simulator.py:1552-1568."""
    if length == 0:
        return "", []
    _, middle_ref, e_dict, e_count = unaligned_error_list(length, model)
    gap, gap_name = extract_read(ref, dna_type, middle_ref)
    gap = case_convert(gap)
    mutated, _ = mutate_read(gap, gap_name, None, e_dict, e_count, False, False, model)
    quals = base_qualities(model.base_qual["unmapped"], len(mutated)) if fastq else []
    return mutated, quals


class ReadSink:
    """Collects what the reference writes to its output files (records + error-profile rows)."""

    def __init__(self):
        self.records = []      # (name, seq, quals or None)
        self.error_rows = []
        self.next_index = 0

    def take_index(self):
        i = self.next_index
        self.next_index += 1
        return i


def simulation_aligned_genome(ref, model, sink, dna_type, min_l, max_l, median_l, sd_l, kmer_bias, fastq,
                              num_simulate, per=False, chimeric=False):
    """simulator.py:1266-1454."""
    remaining = num_simulate
    if chimeric:
        num_segment = np.random.geometric(1 / model.segment_mean, num_simulate)
    else:
        num_segment = np.ones(num_simulate, dtype=int)
    rem_segments = num_segment
    rem_gaps = rem_segments - 1
    passed = 0
    while remaining > 0:
        if per:
            ref_lengths = kde_lengths(model.kde_aligned, sum(rem_segments)) if median_l is None else \
                np.random.lognormal(np.log(median_l), sd_l, rem_segments)
            ref_lengths = [x for x in ref_lengths if min_l <= x <= max_l]
        else:
            rem_lengths, ratio_list = lengths_and_ht_ratios(model, remaining)
            if median_l is None:
                ref_lengths = kde_lengths(model.kde_aligned, sum(rem_segments))
            else:
                totals = np.random.lognormal(np.log(median_l + sd_l ** 2 / 2), sd_l, remaining)
                n_cur = min(remaining, len(rem_lengths), len(ratio_list))
                ref_lengths = totals[:n_cur] - rem_lengths[:n_cur]
            ref_lengths = [x for x in ref_lengths if 0 < x <= max_l]
        gap_lengths = kde_lengths(model.kde_gap, sum(rem_gaps), True) if sum(rem_gaps) > 0 else []
        gap_lengths = [max(0, int(x)) for x in gap_lengths]

        seg_ptr = 0
        gap_ptr = 0
        for each in range(remaining):
            segments = rem_segments[each]
            if seg_ptr + segments > len(ref_lengths):
                break
            ref_len_list = [int(ref_lengths[seg_ptr + x]) for x in range(segments)]
            gap_len_list = [int(gap_lengths[gap_ptr + x]) for x in range(segments - 1)]
            is_reversed = random.random() > model.strandness_rate

            if per:
                seg_ptr += 1
                gap_ptr += 1
                index = sink.take_index()
                new_read, name, quals = "", "", []
                for each_ref in ref_len_list:
                    seg, seg_name = extract_read(ref, dna_type, each_ref)
                    new_read += seg
                    name += seg_name
                    if fastq:
                        quals.extend(base_qualities(model.base_qual["match"], each_ref))
                name = name + "_perfect_" + str(index)
                mutated = case_convert(new_read)
                head = tail = 0
                name += "_R" if is_reversed else "_F"
                name += "_0_" + str(sum(ref_len_list)) + "_0"
            else:
                gaps, gap_quals, seg_lens, seg_dicts, seg_counts = [], [], [], [], []
                remainder = int(rem_lengths[each])
                ratio = ratio_list[each]
                total = remainder
                for g in gap_len_list:
                    mg, gq = simulation_gap(ref, model, g, dna_type, fastq)
                    gaps.append(mg)
                    gap_quals.append(gq)
                for each_ref in ref_len_list:
                    middle, middle_ref, e_dict, e_count = error_list(each_ref, model, fastq)
                    total += middle
                    seg_lens.append(middle_ref)
                    seg_dicts.append(e_dict)
                    seg_counts.append(e_count)
                if total < min_l or total > max_l:
                    continue
                seg_ptr += segments
                gap_ptr += segments - 1
                index = sink.take_index()
                if remainder == 0:
                    head = tail = 0
                else:
                    head = int(round(remainder * ratio))
                    tail = remainder - head
                n_seg = len(seg_lens)
                segs, names = [None] * n_seg, [None] * n_seg
                for s in range(n_seg):
                    segs[s], names[s] = extract_read(ref, dna_type, seg_lens[s])
                name = ";".join(names) + "_aligned_" + str(index)
                if n_seg > 1:
                    name += "_chimeric"
                name += "_R" if is_reversed else "_F"
                name += "_" + str(head) + "_" + ";".join(str(x) for x in seg_lens) + "_" + str(tail)
                mutated, quals = "", []
                for s in range(n_seg):
                    seg = case_convert(segs[s])
                    sm, sq = mutate_read(seg, name, sink.error_rows, seg_dicts[s], seg_counts[s], fastq,
                                         kmer_bias, model)
                    if kmer_bias:
                        sm, sq = mutate_homo(sm, sq, kmer_bias, model)
                    mutated += sm
                    quals.extend(sq)
                    if s < len(gaps):
                        mutated += gaps[s]
                        quals.extend(gap_quals[s])
                if fastq:
                    ht = base_qualities(model.base_qual["ht"], head + tail)
                    quals = ht[:head] + quals + ht[head:]
            mutated = "".join(np.random.choice(ACGT_ORDER, head)) + mutated + \
                      "".join(np.random.choice(ACGT_ORDER, tail))
            if len(mutated) < min_l or len(mutated) > max_l:
                continue
            if is_reversed:
                mutated = reverse_complement(mutated)
                quals.reverse()
            sink.records.append((name, mutated, quals if fastq else None))
            passed += 1
        remaining = num_simulate - passed
        rem_segments = num_segment[passed:]
        rem_gaps = rem_segments - 1


def simulation_unaligned(ref, model, sink, dna_type, min_l, max_l, median_l, sd_l, fastq, num_simulate):
    """simulator.py:1482-1549 (without the --uracil typo branch at :1536)."""
    remaining = num_simulate
    passed = 0
    while remaining > 0:
        ref_l = kde_lengths(model.kde_unaligned, remaining) if median_l is None else \
            np.random.lognormal(np.log(median_l), sd_l, remaining)
        for j in range(len(ref_l)):
            m_ref = int(ref_l[j])
            _, middle_ref, e_dict, e_count = unaligned_error_list(m_ref, model)
            if middle_ref < min_l or middle_ref > max_l:
                continue
            index = sink.take_index()
            read, name = extract_read(ref, dna_type, middle_ref)
            name = name + "_unaligned_" + str(index)
            read = case_convert(read)
            mutated, _ = mutate_read(read, name, None, e_dict, e_count, False, False, model)
            if len(mutated) < min_l or len(mutated) > max_l:
                continue
            quals = base_qualities(model.base_qual["unmapped"], len(mutated)) if fastq else []
            p = random.random()
            if p > model.strandness_rate:
                mutated = reverse_complement(mutated)
                name += "_R"
                quals.reverse()
            else:
                name += "_F"
            sink.records.append((name + "_0_" + str(middle_ref) + "_0", mutated, quals if fastq else None))
            passed += 1
        remaining = num_simulate - passed


def format_records(records, fastq):
    """simulator.py:1437-1443 record layout."""
    out = []
    for name, seq, quals in records:
        out.append(("@" if fastq else ">") + name + "\n" + seq + "\n")
        if fastq:
            out.append("+\n" + "".join(chr(q + 33) for q in quals) + "\n")
    return "".join(out)


# --------------------------------------------------------------------------------------
# metagenome mode
# --------------------------------------------------------------------------------------
class OracleMetaReference:
    """seq_dict[species][chrom] / seq_len / dict_dna_type / max_chrom for metagenome mode (simulator.py:284-339)."""

    def __init__(self, genomes, dna_types=None):
        """genomes: ordered {species: [(chrom_key, sequence), ...]}; dna_types: {species: {chrom_key: type}}."""
        self.seq_dict, self.seq_len, self.dict_dna_type, self.max_chrom = {}, {}, {}, {}
        for sp, recs in genomes.items():
            self.seq_dict[sp], self.seq_len[sp], self.dict_dna_type[sp] = {}, {}, {}
            self.max_chrom[sp] = 0
            for key, seq in recs:
                self.seq_dict[sp][key] = seq
                self.seq_len[sp][key] = len(seq)
                self.dict_dna_type[sp][key] = "circular"          # local files default to circular (:323)
                self.max_chrom[sp] = max(self.max_chrom[sp], len(seq))
        for sp, d in (dna_types or {}).items():
            for key, ty in d.items():
                self.dict_dna_type[sp][key] = ty

    @staticmethod
    def from_genome_list(genome_list, dna_type_list=None):
        """genome list: species<TAB>fasta path (:259-266); dna type list: species<TAB>chrom header<TAB>type (:327-339)."""
        genomes = {}
        with open(genome_list) as f:
            for line in f.readlines():
                fields = line.split("\t")
                sp = "_".join(fields[0].split())
                genomes[sp] = read_fasta(fields[1].strip("\n"))
        types = {}
        if dna_type_list:
            with open(dna_type_list) as f:
                for line in f.readlines():
                    fields = line.split("\t")
                    sp = "_".join(fields[0].split())
                    key = "-".join(re.split(r"[_\s]\s*", fields[1].partition(" ")[0])).split(".")[0]
                    types.setdefault(sp, {})[key] = fields[2].strip("\n")
        return OracleMetaReference(genomes, types)


def read_abundance(path):
    """simulator.py:360-380 -> (number_list, {sampleN: {species: abundance}})."""
    with open(path) as f:
        header = f.readline()
        numbers = [int(x) for x in header.strip().split("\t")[1:]]
        samples = ["sample" + str(i) for i in range(len(numbers))]
        multi = {s: {} for s in samples}
        for line in f.readlines():
            fields = line.split("\t")
            sp = "_".join(fields[0].split())
            vals = [float(x) for x in fields[1:]]
            for i, s in enumerate(samples):
                multi[s][sp] = vals[i]
    return numbers, multi


def inflate_abun(dict_abun, species, abun_inflation):
    """simulator.py:2018-2022."""
    return 1 - (1 - dict_abun[species]) * abun_inflation


def extract_read_meta(ref, length, s=None):
    """simulator.py:1704-1749 (metagenome branch of extract_read)."""
    if not s:
        s = random.choice(list(ref.seq_len.keys()))
    key = random.choice(list(ref.seq_len[s].keys()))
    klen = ref.seq_len[s][key]
    if length > klen:
        longer, longer_target = [], []
        for ts in ref.seq_len:
            for tk in ref.seq_len[ts]:
                if length < ref.seq_len[ts][tk]:
                    (longer_target if ts == s else longer).append((ts, tk))
        assert len(longer) > 0 or len(longer_target) > 0
        s, key = random.choice(longer_target) if longer_target else random.choice(longer)
        klen = ref.seq_len[s][key]
    seq = ref.seq_dict[s][key]
    if ref.dict_dna_type[s][key] == "circular":
        pos = random.randint(0, klen)
        if length + pos > klen:
            out = seq[pos:] + seq[0: length - klen + pos]
        else:
            out = seq[pos: pos + length]
    else:
        pos = random.randint(0, klen - length)
        out = seq[pos: pos + length]
    return out, s + "-" + key + "_" + str(pos)


def assign_species(length_list, seg_list, current, dict_abun, dict_abun_inflated):
    """simulator.py:758-811."""
    seg_sorted = sorted(seg_list, reverse=True)
    segs_chimera = sum([x for x in seg_list if x > 1])
    lengths = length_list[:segs_chimera] + sorted(length_list[segs_chimera:], reverse=True)
    species_list = [""] * len(length_list)
    total_bases = sum(length_list) + sum(current.values())
    total_abun = sum(dict_abun.values())
    quota = {sp: total_bases * ab / total_abun - current[sp] for sp, ab in dict_abun.items()}
    ptr = 0
    pre = ""
    n = len(lengths)
    for seg in seg_sorted:
        if ptr + seg > n:
            break
        for e in range(seg):
            if e == 0:
                avail = [s for s, q in quota.items() if q - lengths[ptr] > 0]
                if len(avail) == 0:
                    avail = [s for s, q in quota.items() if q > 0]
                sp = random.choice(avail)
            else:
                avail = [s for s, q in quota.items() if q - lengths[ptr] > 0 and s != pre]
                p = random.uniform(0, 100)
                if p <= dict_abun_inflated[pre] and quota[pre] > 0:
                    sp = pre
                elif p > dict_abun_inflated[pre] and len(avail) > 0:
                    sp = random.choice(avail)
                else:
                    avail = [s for s, q in quota.items() if q - lengths[ptr] > 0]
                    if len(avail) == 0:
                        avail = [s for s, q in quota.items() if q > 0]
                    sp = random.choice(avail)
            species_list[ptr] = sp
            quota[sp] -= lengths[ptr]
            ptr += 1
            pre = sp
    return species_list[:ptr], lengths[:ptr], np.array(seg_sorted[:ptr])


def simulation_gap_meta(ref, model, length, fastq):
    """simulation_gap (:1552-1568) with dna_type == "metagenome"."""
    if length == 0:
        return "", []
    _, middle_ref, e_dict, e_count = unaligned_error_list(length, model)
    gap, gap_name = extract_read_meta(ref, middle_ref)
    gap = case_convert(gap)
    mutated, _ = mutate_read(gap, gap_name, None, e_dict, e_count, False, False, model)
    quals = base_qualities(model.base_qual["unmapped"], len(mutated)) if fastq else []
    return mutated, quals


def simulation_aligned_metagenome(ref, model, sink, dict_abun, dict_abun_inflated, min_l, max_l, kmer_bias, fastq,
                                  num_simulate, per=False, chimeric=False):
    """simulator.py:814-1040 (KDE lengths only; -med/-sd not restated)."""
    remaining = num_simulate
    if chimeric:
        num_segment = np.random.geometric(1 / model.segment_mean, num_simulate)
    else:
        num_segment = np.ones(num_simulate, dtype=int)
    rem_segments = num_segment
    rem_gaps = rem_segments - 1
    passed = 0
    current = {sp: 0 for sp in dict_abun.keys()}
    while remaining > 0:
        if per:
            ref_lengths = [x for x in kde_lengths(model.kde_aligned, sum(rem_segments)) if min_l <= x <= max_l]
            if len(ref_lengths) == 0:
                continue
        else:
            rem_lengths = [x for x in kde_lengths(model.kde_ht, int(remaining * 1.3), True) if x >= 0]
            ratio_list = [x for x in kde_lengths(model.kde_ht_ratio, int(remaining * 1.5)) if 0 <= x <= 1]
            ref_lengths = [x for x in kde_lengths(model.kde_aligned, sum(rem_segments)) if 0 < x <= max_l]
            if len(ref_lengths) == 0:
                continue
        gap_lengths = kde_lengths(model.kde_gap, sum(rem_gaps), True) if sum(rem_gaps) > 0 else []
        gap_lengths = [max(0, int(x)) for x in gap_lengths]
        species_pool, ref_lengths, rem_segments = assign_species(ref_lengths, rem_segments, current, dict_abun,
                                                                 dict_abun_inflated)
        is_reversed = random.random() > model.strandness_rate
        seg_ptr = gap_ptr = sp_ptr = 0
        for each in range(len(rem_segments)):
            segments = rem_segments[each]
            if (not per and each >= min(len(ratio_list), len(rem_lengths))) or seg_ptr + segments > len(ref_lengths):
                break
            ref_len_list = [int(round(ref_lengths[seg_ptr + x])) for x in range(segments)]
            gap_len_list = [int(round(gap_lengths[gap_ptr + x])) for x in range(segments - 1)]
            species_list = [species_pool[sp_ptr + x] for x in range(segments)]
            if per:
                seg_ptr += 1
                gap_ptr += 1
                sp_ptr += 1
                index = sink.take_index()
                new_read, name, quals = "", "", []
                for s in range(len(ref_len_list)):
                    seg, seg_name = extract_read_meta(ref, ref_len_list[s], species_list[s])
                    new_read += seg
                    name += seg_name
                    if fastq:
                        quals.extend(base_qualities(model.base_qual["match"], ref_len_list[s]))
                name = name + "_perfect_" + str(index)
                mutated = case_convert(new_read)
                if len(mutated) < min_l or len(mutated) > max_l:
                    continue
                head = tail = 0
                name += "_R" if is_reversed else "_F"
                name += "_0_" + str(sum(ref_len_list)) + "_0"
            else:
                gaps, gap_quals, seg_lens, seg_dicts, seg_counts = [], [], [], [], []
                remainder = int(round(rem_lengths[each]))
                ratio = ratio_list[each]
                total = remainder
                restart = False
                for each_ref in ref_len_list:
                    middle, middle_ref, e_dict, e_count = error_list(each_ref, model, fastq)
                    if total + middle_ref > max_l:
                        restart = True
                        break
                    total += middle_ref
                    seg_lens.append(middle_ref)
                    seg_dicts.append(e_dict)
                    seg_counts.append(e_count)
                if restart:
                    continue
                for g in gap_len_list:
                    mg, gq = simulation_gap_meta(ref, model, g, fastq)
                    gaps.append(mg)
                    gap_quals.append(gq)
                    if total + len(mg) > max_l:
                        restart = True
                        break
                    total += len(mg)
                if restart or total < min_l or total > max_l:
                    continue
                seg_ptr += segments
                gap_ptr += segments - 1
                sp_ptr += segments
                index = sink.take_index()
                if remainder == 0:
                    head = tail = 0
                else:
                    head = int(round(remainder * ratio))
                    tail = remainder - head
                n_seg = len(seg_lens)
                segs = [None] * n_seg
                comps = []
                for s in range(n_seg):
                    segs[s], seg_name = extract_read_meta(ref, seg_lens[s], species_list[s])
                    comps.append(seg_name)
                    if s < len(gaps):
                        comps.append("gap_" + str(len(gaps[s])))
                name = ";".join(comps) + "_aligned_" + str(index)
                if n_seg > 1:
                    name += "_chimeric"
                name += "_R" if is_reversed else "_F"
                name += "_" + str(head) + "_" + ";".join(str(x) for x in seg_lens) + "_" + str(tail)
                mutated, quals = "", []
                for s in range(n_seg):
                    seg = case_convert(segs[s])
                    sm, sq = mutate_read(seg, name, sink.error_rows, seg_dicts[s], seg_counts[s], fastq, kmer_bias, model)
                    if kmer_bias:
                        sm, sq = mutate_homo(sm, sq, kmer_bias, model)
                    mutated += sm
                    quals.extend(sq)
                    if s < len(gaps):
                        mutated += gaps[s]
                        quals.extend(gap_quals[s])
                    current[species_list[s]] += len(seg)
                if fastq:
                    ht = base_qualities(model.base_qual["ht"], head + tail)
                    quals = ht[:head] + quals + ht[head:]
            mutated = "".join(np.random.choice(ACGT_ORDER, head)) + mutated + "".join(np.random.choice(ACGT_ORDER, tail))
            if len(mutated) < min_l or len(mutated) > max_l:
                continue
            if is_reversed:
                mutated = reverse_complement(mutated)
                quals.reverse()
            sink.records.append((name, mutated, quals if fastq else None))
            passed += 1
        remaining = num_simulate - passed
        rem_segments = num_segment[passed:]
        rem_gaps = rem_segments - 1


def simulation_unaligned_meta(ref, model, sink, min_l, max_l, fastq, num_simulate):
    """simulation_unaligned (:1482-1549) with dna_type == "metagenome" (random species, :1705-1706)."""
    remaining = num_simulate
    passed = 0
    while remaining > 0:
        ref_l = kde_lengths(model.kde_unaligned, remaining)
        for j in range(len(ref_l)):
            m_ref = int(ref_l[j])
            _, middle_ref, e_dict, e_count = unaligned_error_list(m_ref, model)
            if middle_ref < min_l or middle_ref > max_l:
                continue
            index = sink.take_index()
            read, name = extract_read_meta(ref, middle_ref)
            name = name + "_unaligned_" + str(index)
            read = case_convert(read)
            mutated, _ = mutate_read(read, name, None, e_dict, e_count, False, False, model)
            if len(mutated) < min_l or len(mutated) > max_l:
                continue
            quals = base_qualities(model.base_qual["unmapped"], len(mutated)) if fastq else []
            p = random.random()
            if p > model.strandness_rate:
                mutated = reverse_complement(mutated)
                name += "_R"
                quals.reverse()
            else:
                name += "_F"
            sink.records.append((name + "_0_" + str(middle_ref) + "_0", mutated, quals if fastq else None))
            passed += 1
        remaining = num_simulate - passed


# --------------------------------------------------------------------------------------
# transcriptome mode (without intron retention: --no_model_ir)
# --------------------------------------------------------------------------------------
class OracleTrxReference:
    """seq_dict / seq_len for the reference transcriptome (simulator.py:341-349) + expression profile (:383-401)."""

    def __init__(self, seqs, dict_exp, polya_ids=None):
        self.seq_dict = dict(seqs)
        self.seq_len = {k: len(v) for k, v in self.seq_dict.items()}
        self.max_chrom = max(self.seq_len.values()) if self.seq_len else 0
        self.dict_exp = dict(dict_exp)
        self.ecdf_length_list, self.ecdf_weight_list = make_cdf(self.dict_exp, self.seq_len)
        self.trx_with_polya = {t: 0 for t in (polya_ids or [])}

    @staticmethod
    def from_files(fasta, exp_path, polya_path=None):
        dict_exp = {}
        with open(exp_path) as f:
            f.readline()
            for line in f:
                parts = line.split("\t")
                tid = parts[0].split(".")[0]
                tpm = float(parts[2])
                if tpm > 0:
                    dict_exp[tid] = tpm
        polya = None
        if polya_path:
            with open(polya_path) as f:
                polya = [line.strip().split(".")[0] for line in f.readlines()]
        return OracleTrxReference(read_fasta(fasta), dict_exp, polya)

    def load_ir(self, genome_fasta, gff3_path, ir_markov_path):
        """The model_ir part of read_profile (simulator.py:404-453): genome sequences (pysam.Fastafile: record name = header
        up to the first blank), the IR Markov model (two half-open intervals per state) and the exon/intron structure of
        every transcript from the GFF3 (HTSeq.GFF_Reader(end_included=True): start - 1, end; "chr" stripped)."""
        self.genome = {}
        name, parts = None, []
        with open(genome_fasta) as f:
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        self.genome[name] = "".join(parts)
                    name, parts = line[1:].split()[0], []
                else:
                    parts.append(line.strip())
        if name is not None:
            self.genome[name] = "".join(parts)
        self.genome_names = list(self.genome)
        self.ir_model = {}
        with open(ir_markov_path) as f:
            f.readline()
            for line in f:
                info = line.strip().split()
                if not info:
                    continue
                self.ir_model[info[0]] = [((0, float(info[1])), "no_IR"), ((float(info[1]), float(info[1]) + float(info[2])), "IR")]
        self.structure = {}
        with open(gff3_path) as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                c = line.rstrip("\n").split("\t", 8)
                if c[2] not in ("exon", "intron"):
                    continue
                attr, first = {}, None
                for tok in c[8].rstrip(";").split(";"):
                    tok = tok.strip()
                    if not tok:
                        continue
                    k, v = tok.split("=", 1) if "=" in tok else tok.split(None, 1)
                    v = v.strip().strip('"')
                    attr[k.strip()] = v
                    if first is None:
                        first = v
                if "transcript_id" in attr:
                    fid = attr["transcript_id"]
                elif "Parent" in attr:
                    info = first.split(":")
                    if len(info) == 1:
                        fid = info[0]
                    elif info[0] == "transcript":
                        fid = info[1]
                    else:
                        continue
                else:
                    continue
                fid = fid.split(".")[0]
                chrom = c[0].strip("chr") if "chr" in c[0] else c[0]
                start, end = int(c[3]) - 1, int(c[4])
                self.structure.setdefault(fid, []).append((c[2], chrom, start, end, end - start, c[6]))
        return self


def make_cdf(dict_exp, dict_len):
    """simulator.py:69-97."""
    total = 0
    vals = []
    for item in dict_exp:
        if item in dict_len:
            total += dict_exp[item]
    for item in dict_exp:
        if item in dict_len:
            vals.append((item, dict_exp[item] / float(total)))
    vals_sorted = sorted(vals, key=lambda x: x[1])
    cdf = np.cumsum([x[1] for x in vals_sorted])
    bounds = [0] + list(cdf)
    weights, lengths = [], []
    for i, t in enumerate(vals_sorted):
        weights.append(abs(bounds[i + 1] - bounds[i]))
        lengths.append((t[0], dict_len[t[0]]))
    return lengths, weights


def select_nearest_kde2d(sampled, ref_len_total):
    """simulator.py:108-111."""
    idx = np.abs(sampled[:, 0] - ref_len_total).argmin()
    return int(sampled[idx][1])


def update_structure(structure, ir_model):
    """simulator.py:114-146: one IR / no_IR state per intron from the two-state Markov chain."""
    count = sum(1 for item in structure if item[0] == "intron")
    states, flag_ir, prev = [], False, "start"
    for _ in range(count):
        p = random.random()
        for (lo, hi), flag in ir_model[prev]:
            if lo <= p < hi:
                if flag == "IR":
                    flag_ir = True
                states.append(flag)
                prev = flag
                break
    if not flag_ir:
        return False, structure
    out, j = list(structure), -1
    for i in range(len(out)):
        if out[i][0] == "intron":
            j += 1
            if states[j] == "IR":
                out[i] = ("retained_intron",) + out[i][1:]
    return True, out


def extract_read_pos(length, ref_len, structure, polya, buffer=10):
    """simulator.py:149-191 -> ([(chrom, start, end, strand)], retain_polya, [(start, end) of the retained introns hit])."""
    len_before = 0
    for item in structure:
        if item[0] == "exon":
            len_before += item[4]
        elif item[0] == "retained_intron":
            break
    start_pos = random.randint(0, min(ref_len - length, len_before))
    ivs, ir_list, end = [], [], None
    for item in structure:
        if length == 0:
            break
        if item[0] in ("exon", "retained_intron"):
            if start_pos < item[4]:
                start = start_pos + item[2]
                end = start + length if start + length <= item[3] else item[3]
                length -= end - start
                start_pos = 0
                ivs.append((item[1], start, end, item[5]))
                if item[0] == "retained_intron":
                    ir_list.append((start, end))
            else:
                start_pos -= item[4]
    retain = bool(polya and end + buffer >= structure[-1][3])
    return ivs, retain, ir_list


def extract_read_trx(ref, key, length, trx_has_polya, buffer=10):
    """simulator.py:1683-1691."""
    pos = random.randint(0, ref.seq_len[key] - length)
    seq = ref.seq_dict[key][pos: pos + length]
    retain = bool(trx_has_polya and pos + length + buffer >= ref.seq_len[key])
    return seq, pos, retain


def extract_read_transcriptome(ref, length):
    """simulator.py:1695-1703 (extract_read with dna_type == "transcriptome", used by unaligned reads)."""
    while True:
        key = random.choice(list(ref.seq_len.keys()))
        if length < ref.seq_len[key]:
            pos = random.randint(0, ref.seq_len[key] - length)
            return ref.seq_dict[key][pos: pos + length], key + "_" + str(pos)


def simulation_aligned_transcriptome(ref, model, sink, kmer_bias, basecaller, num_simulate, polya, fastq, per=False,
                                     uracil=False, model_ir=False):
    """simulator.py:1043-1263; model_ir needs ref.load_ir(...)."""
    flag_chrom = model_ir and any("chr" in item for item in ref.genome_names)
    import scipy.stats

    scale = 2.409858743694814 if basecaller == "albacore" else 4.168299657168961

    def polya_len_draw():
        return int(scipy.stats.expon.rvs(loc=2.0, scale=scale))

    remainder_l = kde_lengths(model.kde_ht, num_simulate, True)
    ratio_tmp = kde_lengths(model.kde_ht_ratio, num_simulate)
    ratio_l = [1 if x > 1 else x for x in ratio_tmp]
    ratio_l = [0 if x < 0 else x for x in ratio_l]
    simulated = 0
    sampled = kde_lengths(model.kde_aligned_2d, num_simulate, False, False)
    trx_sampled = set()
    while simulated < num_simulate:
        while True:
            ref_trx, ref_trx_len = random.choices(ref.ecdf_length_list, weights=ref.ecdf_weight_list, k=1)[0]
            if ref_trx in trx_sampled:
                sampled = kde_lengths(model.kde_aligned_2d, num_simulate, False, False)
                trx_sampled = set()
            if model_ir:
                # simulator.py:1094-1099: a transcript without GFF3 features, or whose exons do not add up to its length,
                # is drawn again (ref_len_from_structure :100-105)
                if ref_trx not in ref.structure:
                    continue
                if ref_trx_len != sum(item[-2] for item in ref.structure[ref_trx] if item[0] == "exon"):
                    continue
            ref_len_aligned = select_nearest_kde2d(sampled, ref_trx_len)
            if ref_len_aligned < ref_trx_len:
                break
        trx_sampled.add(ref_trx)
        has_polya = polya and ref_trx in ref.trx_with_polya
        is_reversed = random.random() > model.strandness_rate
        if per:
            index = sink.take_index()
            new_read, pos, retain = extract_read_trx(ref, ref_trx, ref_len_aligned, has_polya)
            name = ref_trx + "_" + str(pos) + "_perfect_" + str(index)
            mutated = case_convert(new_read)
            quals = base_qualities(model.base_qual["match"], ref_len_aligned) if fastq else []
            head = tail = 0
            name += "_R" if is_reversed else "_F"
            if retain:
                polya_len = polya_len_draw()
                if polya_len > 0:
                    mutated += "A" * polya_len
            else:
                polya_len = 0
            name += "_0_" + str(ref_len_aligned) + "_" + str(polya_len)
        else:
            middle, middle_ref, e_dict, e_count = error_list(ref_len_aligned, model, fastq)
            if middle_ref > ref_trx_len:
                continue
            index = sink.take_index()
            ir_list = []
            ir_flag = False
            if model_ir:
                ir_flag, structure_new = update_structure(ref.structure[ref_trx], ref.ir_model)
            if ir_flag:
                ivs, retain, ir_list = extract_read_pos(middle_ref, ref_trx_len, structure_new, has_polya)
                new_read, missing = "", False
                for chrom, start, end, strand in ivs:
                    if flag_chrom:
                        chrom = "chr" + chrom
                    if chrom not in ref.genome:
                        missing = True
                        break
                    new_read += ref.genome[chrom][start:end]
                if missing:
                    continue
                pos = ivs[0][1]
                if strand == "-":                   # keep the read in the direction of the reference transcript
                    new_read = reverse_complement(new_read)
            else:
                new_read, pos, retain = extract_read_trx(ref, ref_trx, middle_ref, has_polya)
            name = str(ref_trx) + "_" + str(pos) + "_aligned_" + str(index)
            if len(ir_list) > 0:
                name += "_RetainedIntron_"
                for ir_tuple in ir_list:
                    name += "-".join(str(x) for x in ir_tuple) + ";"
            name += "_R" if is_reversed else "_F"
            remainder = int(remainder_l[simulated])
            ratio = ratio_l[simulated]
            if remainder == 0:
                head = tail = 0
            else:
                head = int(round(remainder * ratio))
                tail = remainder - head
            polya_len = polya_len_draw() if retain else 0
            name += "_" + str(head) + "_" + str(middle_ref) + "_" + str(tail + polya_len)
            new_read = case_convert(new_read)
            mutated, quals = mutate_read(new_read, name, sink.error_rows, e_dict, e_count, fastq, kmer_bias, model)
            if kmer_bias:
                mutated, quals = mutate_homo(mutated, quals, kmer_bias, model)
            if polya_len > 0:
                mutated += "A" * polya_len
        if fastq:
            ht = base_qualities(model.base_qual["ht"], head + tail + polya_len)
            for _ in range(polya_len):
                quals.append(ht.pop())
            quals = ht[:head] + quals + ht[head:]
        mutated = "".join(np.random.choice(ACGT_ORDER, head)) + mutated + "".join(np.random.choice(ACGT_ORDER, tail))
        if is_reversed:
            mutated = reverse_complement(mutated)
            quals.reverse()
        if uracil:
            mutated = mutated.translate(str.maketrans("T", "U"))
        sink.records.append((name, mutated, quals if fastq else None))
        simulated += 1


def simulation_unaligned_transcriptome(ref, model, sink, min_l, max_l, fastq, num_simulate):
    """simulation_unaligned (:1482-1549) with dna_type == "transcriptome"."""
    remaining = num_simulate
    passed = 0
    while remaining > 0:
        ref_l = kde_lengths(model.kde_unaligned, remaining)
        for j in range(len(ref_l)):
            m_ref = int(ref_l[j])
            _, middle_ref, e_dict, e_count = unaligned_error_list(m_ref, model)
            if middle_ref < min_l or middle_ref > max_l:
                continue
            index = sink.take_index()
            read, name = extract_read_transcriptome(ref, middle_ref)
            name = name + "_unaligned_" + str(index)
            read = case_convert(read)
            mutated, _ = mutate_read(read, name, None, e_dict, e_count, False, False, model)
            if len(mutated) < min_l or len(mutated) > max_l:
                continue
            quals = base_qualities(model.base_qual["unmapped"], len(mutated)) if fastq else []
            p = random.random()
            if p > model.strandness_rate:
                mutated = reverse_complement(mutated)
                name += "_R"
                quals.reverse()
            else:
                name += "_F"
            sink.records.append((name + "_0_" + str(middle_ref) + "_0", mutated, quals if fastq else None))
            passed += 1
        remaining = num_simulate - passed
