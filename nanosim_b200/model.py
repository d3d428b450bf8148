"""Host-side model loader / compiler (the product's mirror of read_profile's model half).

Reads a NanoSim model directory in the reference's own on-disk format
(/root/reference/src/simulator.py:470-591: ``<prefix>_model_profile``, ``_error_markov_model``,
``_first_match.hist``, ``_match_markov_model``, ``_reads_alignment_rate``, ``_strandness_rate``,
``_chimeric_info``, ``_base_qualities_model_parameters.tsv``, ``_hp_lengths_model_parameters.tsv`` and the
joblib KernelDensity pickles) and turns it into the flat integer / float32 tables the CUDA path consumes
through the C-ABI (include/nanosim_b200.h, ``NsModel``).

Every discrete draw of the reference's hot path is a draw from a FIXED discrete distribution, so it is
tabulated exactly here (float64 on the host) and sampled on the device with Walker alias tables driven
by 32-bit Philox words:

  * first match length          simulator.py:1843-1850 (ECDF + linear interpolation + floor, min 2)
  * next match length | bin     simulator.py:1891-1898 (same, one table per previous-match-length bin)
  * mismatch run length         mixed_model.py:41-49   (w*(Poisson(lam)+1) + (1-w)*Geometric(p))
  * insertion / deletion length mixed_model.py:52-63   (w*ceil(lam*Weibull(k)) + (1-w)*(Geometric(p)-1), 0 -> 1)
  * base quality | state        model_base_qualities.py:9-20,120-130 (truncated log-normal, floored)

A compiled model (``CompiledModel``) is the raw small text files plus the KDE training samples, stored
losslessly in one ``.npz`` so that it can travel to machines without the reference checkout.
"""
import io
import math
import os
import re

import numpy as np

TEXT_FILES = [
    "model_profile", "error_markov_model", "first_match.hist", "match_markov_model",
    "reads_alignment_rate", "strandness_rate", "chimeric_info",
    "base_qualities_model_parameters.tsv", "hp_lengths_model_parameters.tsv", "error_rate.tsv",
]
KDE_FILES = ["aligned_region", "aligned_reads", "ht_length", "ht_ratio", "unaligned_length", "gap_length",
             "aligned_region_2d"]

QUAL_STATES = ["mis", "ins", "match", "ht", "unmapped"]   # device state ids 0..4
ERR_STATES = ["start", "mis", "ins", "del", "mis0", "ins0", "del0"]
Q_MIN, Q_MAX = 1, 93


# --------------------------------------------------------------------------------------
# KernelDensity pickles (scikit-learn 0.22/0.23) -> (training samples, bandwidth)
# --------------------------------------------------------------------------------------
class _Opaque:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.state = st


def _alloc(cls):
    return cls.__new__(cls)


def load_kde_pickle(path):
    """The shipped ``*.pkl`` files were written by scikit-learn 0.22/0.23 and reference private
    sklearn classes that no longer exist; only two fields are needed (the [N,d] float64 training
    matrix held by the KDTree, and the bandwidth), so every sklearn global is mapped to an opaque
    holder and those two fields are pulled out."""
    from joblib.numpy_pickle import NumpyUnpickler

    class _Unpickler(NumpyUnpickler):
        def find_class(self, module, name):
            if module.startswith("sklearn"):
                return _alloc if name == "newObj" else type(name, (_Opaque,), {})
            return super().find_class(module, name)

    with open(path, "rb") as f:
        obj = _Unpickler(path, f, ensure_native_byte_order=True).load()
    st = obj.state if hasattr(obj, "state") else obj.__dict__
    kernel = st.get("kernel", "gaussian")
    if kernel != "gaussian":
        raise ValueError("%s: only gaussian KernelDensity models are supported (got %r)" % (path, kernel))
    data = np.ascontiguousarray(np.asarray(st["tree_"].state[0], dtype=np.float64))
    return data, float(st["bandwidth"])


class CompiledModel:
    """Raw model text files + KDE samples, losslessly; round-trips through one .npz."""

    def __init__(self):
        self.text = {}     # name -> str
        self.kde = {}      # name -> (float64 [N,d], bandwidth)
        self.name = ""

    @staticmethod
    def from_prefix(prefix, extra_text=None):
        cm = CompiledModel()
        cm.name = os.path.basename(os.path.dirname(os.path.abspath(prefix + "_x")))
        for t in TEXT_FILES:
            p = prefix + "_" + t
            if os.path.exists(p):
                with open(p) as f:
                    cm.text[t] = f.read()
        for k in KDE_FILES:
            p = prefix + "_" + k + ".pkl"
            if os.path.exists(p):
                cm.kde[k] = load_kde_pickle(p)
        for t, p in (extra_text or {}).items():
            with open(p) as f:
                cm.text[t] = f.read()
        if "model_profile" not in cm.text:
            raise FileNotFoundError(prefix + "_model_profile")
        return cm

    def save(self, path):
        out = {"name": np.frombuffer(self.name.encode(), dtype=np.uint8)}
        for t, s in self.text.items():
            out["text/" + t] = np.frombuffer(s.encode(), dtype=np.uint8)
        for k, (data, bw) in self.kde.items():
            uniq, inv = np.unique(data, return_inverse=True)
            idt = np.uint16 if len(uniq) <= 65536 else np.uint32
            out["kde/%s/values" % k] = uniq
            out["kde/%s/index" % k] = inv.reshape(-1).astype(idt)
            out["kde/%s/shape" % k] = np.asarray(data.shape, dtype=np.int64)
            out["kde/%s/bw" % k] = np.asarray([bw], dtype=np.float64)
        np.savez_compressed(path, **out)

    @staticmethod
    def load(path):
        cm = CompiledModel()
        with np.load(path) as z:
            cm.name = bytes(z["name"]).decode()
            for key in z.files:
                if key.startswith("text/"):
                    cm.text[key[5:]] = bytes(z[key]).decode()
                elif key.endswith("/values"):
                    k = key.split("/")[1]
                    vals = z[key]
                    data = vals[z["kde/%s/index" % k].astype(np.int64)].reshape(tuple(z["kde/%s/shape" % k]))
                    cm.kde[k] = (np.ascontiguousarray(data), float(z["kde/%s/bw" % k][0]))
        return cm


def load_model(path_or_prefix):
    """Accepts a compiled ``.npz`` or a reference-format model prefix (``.../training``)."""
    if path_or_prefix.endswith(".npz"):
        return CompiledModel.load(path_or_prefix)
    return CompiledModel.from_prefix(path_or_prefix)


# --------------------------------------------------------------------------------------
# ECDF text -> exact probability mass functions
# --------------------------------------------------------------------------------------
def ecdf_intervals(text):
    """Restates how the reference turns an ECDF table into sampling intervals
    (simulator.py:194-231): per column, every row whose cumulative value differs from the
    previous distinct value opens an interval (cdf_prev, cdf_row] that maps linearly onto the
    length range (len_prev_hi, row_hi); the first non-zero row's range is widened downwards to
    max(0-ish, hi - 10*(hi-lo)); the last interval's upper length is the table's last row.
    Returns (bin_edges [(lo,hi)...], [array (n_i,4) of cdf_lo,cdf_hi,len_lo,len_hi per bin])."""
    lines = text.split("\n")
    head = lines[0].strip().split()
    bins = [tuple(int(v) for v in tok.split("-")) for tok in head[1:]]
    order = sorted(range(len(bins)), key=lambda i: bins[i])   # lane i of the file <-> i-th smallest bin
    rows_lo, rows_hi, probs = [], [], []
    for line in lines[1:]:
        if not line.strip():
            continue
        cols = line.strip().split("\t")
        a, b = cols[0].split("-")
        rows_lo.append(float(a))
        rows_hi.append(float(b))
        probs.append([float(x) for x in cols[1:]])
    rows_lo = np.asarray(rows_lo)
    rows_hi = np.asarray(rows_hi)
    probs = np.asarray(probs, dtype=np.float64).reshape(len(rows_hi), -1)
    out = [None] * len(bins)
    for lane in range(len(bins)):
        col = probs[:, lane]
        iv = []
        last_p, last_len = 0.0, 0.0
        for r in range(len(col)):
            p = col[r]
            if p == last_p:
                continue
            if last_p != 0:
                iv.append([last_p, p, last_len, rows_hi[r]])
            else:
                iv.append([last_p, p, max(last_len, rows_hi[r] - 10 * (rows_hi[r] - rows_lo[r])), rows_hi[r]])
            last_len = rows_hi[r]
            last_p = p
        # the reference extends the interval with the largest (cdf_lo, cdf_hi) key to the last row
        if iv:
            j = max(range(len(iv)), key=lambda t: (iv[t][0], iv[t][1]))
            iv[j][3] = rows_hi[-1]
        out[order[lane]] = np.asarray(iv, dtype=np.float64).reshape(-1, 4)
    sorted_bins = [bins[i] for i in order]
    return sorted_bins, out


def intervals_to_pmf(iv):
    """Exact pmf of ``floor((p-clo)/(chi-clo)*(vhi-vlo)+vlo)`` for p ~ U(0,1) restricted to the
    intervals (simulator.py:1847,1897).  Returns (pmf over 0..vmax, p_miss) where p_miss is the
    probability that no interval contains p (the reference then keeps a stale value)."""
    vmax = int(np.ceil(iv[:, 3].max()))
    pmf = np.zeros(vmax + 1, dtype=np.float64)
    covered = 0.0
    for clo, chi, vlo, vhi in iv:
        w = chi - clo
        if w <= 0:
            continue
        covered += w
        span = vhi - vlo
        if span <= 0:
            pmf[int(np.floor(vlo))] += w
            continue
        lo_i, hi_i = int(np.floor(vlo)), int(np.ceil(vhi))
        for s in range(lo_i, hi_i):
            a, b = max(vlo, s), min(vhi, s + 1)
            if b > a:
                pmf[s] += w * (b - a) / span
    # intervals are (lo, hi] with contiguous keys; what is not covered below 1 is a miss
    lo_all = iv[:, 0].min()
    p_miss = max(0.0, 1.0 - covered - max(0.0, lo_all)) + max(0.0, lo_all)
    total = pmf.sum() + p_miss
    return pmf / total, p_miss / total


def pois_geom_pmf(lam, prob, weight, tail=2.0 ** -44, kmax=4096):
    """pmf over 1..K of mixed_model.py:41-49."""
    pm = [0.0]
    x = 1
    pois = math.exp(-lam)
    geo = prob
    surv_p, surv_g = 1.0, 1.0
    while True:
        pm.append(weight * pois + (1 - weight) * geo)
        surv_p -= pois
        surv_g -= geo
        if (weight * max(surv_p, 0.0) + (1 - weight) * max(surv_g, 0.0)) < tail or x >= kmax:
            break
        pois = pois * lam / x
        geo = geo * (1 - prob)
        x += 1
    pm = np.asarray(pm)
    pm[-1] += max(0.0, 1.0 - pm.sum())
    return pm / pm.sum()


def wei_geom_pmf(lam, k, prob, weight, tail=2.0 ** -44, kmax=4096):
    """pmf over 1..K of mixed_model.py:52-63."""
    def wcdf(t):
        return 1.0 - math.exp(-((t / lam) ** k)) if t > 0 else 0.0

    pm = [0.0]
    x = 1
    while True:
        w_part = wcdf(x) - wcdf(x - 1)
        g_part = prob * (1 - prob) ** x
        if x == 1:
            g_part += prob
        pm.append(weight * w_part + (1 - weight) * g_part)
        surv = weight * (1.0 - wcdf(x)) + (1 - weight) * (1 - prob) ** (x + 1)
        if surv < tail or x >= kmax:
            break
        x += 1
    pm = np.asarray(pm)
    pm[-1] += max(0.0, 1.0 - pm.sum())
    return pm / pm.sum()


def quality_pmf(sd, loc, mu):
    """pmf over q = 0..93 of floor(truncated-lognormal on [1,93]) + loc
    (model_base_qualities.py:9-20, 120-130; scipy rv_discrete.rvs casts ppf+loc to int64)."""
    from scipy.stats import lognorm

    scale = math.exp(mu)
    grid = np.arange(Q_MIN, Q_MAX + 1, dtype=np.float64)
    cdf = lognorm.cdf(grid, sd, scale=scale)
    fa, fb = cdf[0], cdf[-1]
    mass = np.diff(cdf) / (fb - fa)             # q = 1..92
    pmf = np.zeros(Q_MAX + 1)
    iloc = int(loc)
    if float(iloc) != float(loc):
        raise ValueError("non-integer quality loc %r is not supported" % (loc,))
    for q, m in zip(range(Q_MIN, Q_MAX), mass):
        t = min(max(q + iloc, 0), Q_MAX)
        pmf[t] += m
    return pmf / pmf.sum()


# --------------------------------------------------------------------------------------
# Walker alias tables (uint32 fixed point)
# --------------------------------------------------------------------------------------
def build_alias(pmf):
    """Vose's alias method.  Returns (prob u32[n], alias u32[n]); a draw is
    ``j = (r*n)>>32; frac = (r*n)&0xffffffff; value = frac < prob[j] ? j : alias[j]``."""
    p = np.asarray(pmf, dtype=np.float64)
    n = len(p)
    scaled = p / p.sum() * n
    prob = np.ones(n, dtype=np.float64)
    alias = np.arange(n, dtype=np.int64)
    small = [i for i in range(n) if scaled[i] < 1.0]
    large = [i for i in range(n) if scaled[i] >= 1.0]
    scaled = scaled.copy()
    while small and large:
        s = small.pop()
        g = large.pop()
        prob[s] = scaled[s]
        alias[s] = g
        scaled[g] = scaled[g] - (1.0 - scaled[s])
        (small if scaled[g] < 1.0 else large).append(g)
    for i in small + large:
        prob[i] = 1.0
    q = np.minimum(np.floor(prob * 4294967296.0), 4294967295.0).astype(np.uint64)
    q[prob >= 1.0] = 4294967295
    return q.astype(np.uint32), alias.astype(np.uint32)


def alias_pmf(prob, alias):
    """The exact distribution an alias table realises with a 32-bit word (for tests)."""
    n = len(prob)
    acc = (prob.astype(np.float64) + 0.0) / 4294967296.0
    # slot j accepts when frac < prob[j]; prob==0xffffffff is treated as always-accept on device
    acc = np.where(prob == 4294967295, 1.0, acc)
    out = np.zeros(n)
    np.add.at(out, np.arange(n), acc / n)
    np.add.at(out, alias.astype(np.int64), (1.0 - acc) / n)
    return out


# --------------------------------------------------------------------------------------
# device tables
# --------------------------------------------------------------------------------------
T_FIRST, T_MIS, T_INS, T_DEL, T_MATCH0 = 0, 1, 2, 3, 4     # alias table ids (match bins follow)


class DeviceTables:
    """Flat arrays in the exact layout of include/nanosim_b200.h:NsModel."""

    def __init__(self, cm, fastq=False, homopolymer=False, chimeric=False, perfect=False,
                 strandness=None, mode="genome"):
        self.cm = cm
        t = cm.text
        # ---- scalars
        self.strandness = float(t["strandness_rate"].split("\t")[1]) if strandness is None else float(strandness)
        rate = t["reads_alignment_rate"].strip().split("\t")[1]
        self.aligned_ratio = None if rate == "100%" else float(rate)
        self.segment_mean = 1.0
        self.abun_inflation = 0.0
        if chimeric:
            lines = t["chimeric_info"].split("\n")
            self.segment_mean = float(lines[0].split("\t")[1])
            if mode == "metagenome":
                self.abun_inflation = float(lines[1].split("\t")[1])
        # ---- error model
        par = {}
        for line in t["model_profile"].split("\n")[1:]:
            if not line.strip():
                continue
            cols = line.strip().split("\t")
            key = "mis" if "mismatch" in line else ("ins" if "insertion" in line else "del")
            par[key] = [float(x) for x in cols[1:]]
        self.error_par = par
        tables = [None] * 4
        self.first_bins, first_iv = ecdf_intervals(t["first_match.hist"])
        pmf, p_miss = intervals_to_pmf(first_iv[0])
        pmf = pmf.copy()
        pmf[2] += pmf[0] + pmf[1]                       # simulator.py:1848-1849 floor of 2
        pmf[0] = pmf[1] = 0.0
        # a miss leaves prev_match unbound in the reference (UnboundLocalError); fold it into the mode
        pmf[int(np.argmax(pmf))] += p_miss
        tables[T_FIRST] = pmf
        tables[T_MIS] = pois_geom_pmf(par["mis"][0], par["mis"][2], par["mis"][3])
        tables[T_INS] = wei_geom_pmf(*par["ins"])
        tables[T_DEL] = wei_geom_pmf(*par["del"])
        self.match_bins, match_iv = ecdf_intervals(t["match_markov_model"])
        self.match_pmf = []
        for iv in match_iv:
            pmf, p_miss = intervals_to_pmf(iv)
            # extra last slot == "miss": the device then keeps the error length (simulator.py:1895-1898)
            self.match_pmf.append(np.concatenate([pmf, [p_miss]]))
        tables.extend(self.match_pmf)
        self.pmfs = tables
        desc, probs, aliases = [], [], []
        off = 0
        for pm in tables:
            pr, al = build_alias(pm)
            desc.append((off, len(pr)))
            probs.append(pr)
            aliases.append(al)
            off += len(pr)
        self.alias_desc = np.asarray(desc, dtype=np.uint32).reshape(-1, 2)
        self.alias_prob = np.concatenate(probs)
        self.alias_idx = np.concatenate(aliases)
        self.match_bin_lo = np.asarray([b[0] for b in self.match_bins], dtype=np.uint32)
        self.match_bin_hi = np.asarray([b[1] for b in self.match_bins], dtype=np.uint32)
        # ---- error-type Markov chain (simulator.py:486-495), thresholds on a 32-bit word
        trans = np.zeros((7, 3), dtype=np.uint32)
        rows = {}
        for line in t["error_markov_model"].split("\n")[1:]:
            c = line.strip().split()
            if len(c) >= 4:
                rows[c[0]] = (float(c[1]), float(c[2]), float(c[3]))
        self.trans_rows = rows
        for i, st in enumerate(ERR_STATES):
            pm, pi, pd = rows[st]
            trans[i, 0] = _u32(pm)                  # r <  t0            -> mis
            trans[i, 1] = _u32(pm + pi)             # t0 <= r < t1       -> ins
            trans[i, 2] = _u32(1.0 - pd)            # r >= t2            -> del ; between: stale value
        self.trans = trans
        # ---- KDEs (float32 on device; length KDEs hold integers < 2^24 exactly)
        self.kde = {}
        for k, (data, bw) in cm.kde.items():
            self.kde[k] = (np.ascontiguousarray(data.astype(np.float32)), np.float32(bw))
        # ---- 2-D KDE of (transcript length, aligned length), transcriptome mode: rows sorted by transcript length
        self.kde2d = None
        if "aligned_region_2d" in cm.kde:
            d2, bw2 = cm.kde["aligned_region_2d"]
            order = np.argsort(d2[:, 0], kind="stable")
            self.kde2d = (np.ascontiguousarray(d2[order, 0].astype(np.float32)),
                          np.ascontiguousarray(d2[order, 1].astype(np.float32)), np.float32(bw2))
        # ---- base qualities
        self.has_qual = "base_qualities_model_parameters.tsv" in t
        self.qual_cdf = np.zeros((5, Q_MAX + 1), dtype=np.uint32)
        self.qual_pmf = np.zeros((5, Q_MAX + 1))
        if self.has_qual:
            qp = {}
            for line in t["base_qualities_model_parameters.tsv"].split("\n")[1:]:
                c = line.split("\t")
                if len(c) >= 4:
                    qp[c[0]] = (float(c[1]), float(c[2]), float(c[3]))
            self.qual_par = qp
            for i, st in enumerate(QUAL_STATES):
                pm = quality_pmf(*qp[st])
                self.qual_pmf[i] = pm
                c = np.cumsum(pm)
                c[-1] = 1.0
                self.qual_cdf[i] = np.minimum(np.round(c * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
        elif fastq:
            raise FileNotFoundError("model has no _base_qualities_model_parameters.tsv (needed for --fastq)")
        # ---- homopolymer model (model_homopolymer_lengths.py:167-186,204-209,246-260)
        self.has_hp = "hp_lengths_model_parameters.tsv" in t
        self.hp = np.zeros((2, 6), dtype=np.float64)     # rows AT, CG: const, alpha1, beta1, breakpoint1, intercept, slope
        self.hp_mis_rate = 0.0
        if self.has_hp:
            lines = t["hp_lengths_model_parameters.tsv"].split("\n")
            self.hp_mis_rate = float(re.search(r"\d+\.?\d*", lines[0])[0])    # simulator.py:511
            names = lines[1].strip().split("\t")
            for line in lines[2:]:
                c = line.strip().split("\t")
                if len(c) < len(names):
                    continue
                d = dict(zip(names[1:], (float(x) for x in c[1:])))
                row = 0 if c[0] == "AT" else 1
                self.hp[row] = [d["const"], d["alpha1"], d["beta1"], d["breakpoint1"], d["intercept"], d["slope"]]
        elif homopolymer:
            raise FileNotFoundError("model has no _hp_lengths_model_parameters.tsv (needed for -hp)")
        # ---- expected ops/base, used to size the per-read op slots
        self.mean_ref_per_event, self.ref_per_event_cv = self._mean_ref_advance_per_event()

    def _mean_ref_advance_per_event(self, n_events=20000):
        """Reference bases consumed per error event in the stationary regime of the
        error/match renewal chain (simulator.py:1858-1914): mean and coefficient of variation, estimated by running
        the chain on the tabulated pmfs with a private generator.  Only used to size per-read op slots: the number of
        events in a segment of L bases has mean L / mean and variance ~ (L / mean) * cv^2 (renewal process)."""
        rng = np.random.default_rng(20240917)
        cdfs = [np.cumsum(pm) for pm in self.pmfs]
        state, prev_match, ref, ref2 = 0, 10, 0, 0
        u = rng.random((n_events, 3))
        for e in range(n_events):
            pm, pi, pd = self.trans_rows[ERR_STATES[state]]
            r = u[e, 0]
            kind = 1 if r < pm else (2 if r < pm + pi else 3)
            step = int(np.searchsorted(cdfs[kind], u[e, 1], side="right"))
            if kind != 2:
                ref += step
            b = len(self.match_bins) - 1
            for i, (lo, hi) in enumerate(self.match_bins):
                if lo <= prev_match < hi:
                    b = i
                    break
            m = int(np.searchsorted(cdfs[T_MATCH0 + b], u[e, 2], side="right"))
            if m >= len(self.match_pmf[b]) - 1:
                m = step
            if prev_match == 0 and m == 0:
                m = 1
            prev_match = m
            adv = m + (step if kind != 2 else 0)
            ref += m
            ref2 += adv * adv
            state = kind + (3 if m == 0 else 0)
        mean = ref / float(n_events)
        var = max(ref2 / float(n_events) - mean * mean, 0.0)
        return max(1.0, mean), max(1.0, math.sqrt(var) / max(mean, 1e-9))

    def split_counts(self, number, perfect=False):
        """simulator.py:465-468, 538-542."""
        if perfect or self.aligned_ratio is None:
            return number, 0
        r = self.aligned_ratio
        n_al = int(round(number * r / (r + 1)))
        return n_al, number - n_al


def _u32(x):
    return np.uint32(min(max(int(math.floor(x * 4294967296.0 + 0.5)), 0), 4294967295))
