"""Import/run shim for the UNMODIFIED reference (/root/reference/src/simulator.py).

TEST INFRASTRUCTURE ONLY.  Nothing under nanosim_b200/ may import this file.  It only
works in the build container (the GPU box has no /root/reference); its two jobs are

  * ``python oracle/ref_shim.py genome -rg ... -c ...``  runs the reference CLI unmodified
    (used by tests/golden/make_golden_runs.py to produce the run-level golden histograms),
  * ``load_reference_module()`` imports the reference as a module so that
    tests/golden/make_golden_vectors.py can call its per-read functions
    (error_list, mutate_read, ...) under a fixed seed and commit the input/output vectors.

What the shim does:
  1. stands in for ``HTSeq``, ``pysam``, ``piecewise_regression`` (absent here) -- imported at simulator.py:15-16,
     model_base_qualities.py:4, model_homopolymer_lengths.py:6.  Only the intron-retention branch uses them, through
     three calls: ``HTSeq.GFF_Reader(path, end_included=True)``, ``HTSeq.GenomicInterval(chrom, start, end, strand)``
     and ``pysam.Fastafile(path)`` (``.references``, ``.fetch(chrom, start, end)``).  The stand-ins below follow the
     documented behaviour of those calls (GFF coordinates 1-based inclusive -> 0-based half-open intervals; fetch =
     0-based half-open substring), so that the reference's own IR logic (update_structure, extract_read_pos,
     simulator.py:114-191, 1156-1183) runs unmodified;
  2. replaces ``joblib.load`` by a stub-unpickler for the scikit-learn 0.22/0.23 KernelDensity
     pickles, which do not load under the installed scikit-learn.  The returned object replays
     sklearn's ``KernelDensity.sample`` (gaussian kernel): ``i = floor(u*N)``,
     ``x = normal(data[i], bandwidth)`` on the global numpy RNG.
"""
import sys
import types
import warnings

import numpy as np

REFERENCE_SRC = "/root/reference/src"


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.state = st


def _new_obj(cls):
    return cls.__new__(cls)


class ShimKDE:
    """Replays sklearn.neighbors.KernelDensity.sample (gaussian) on numpy's global RNG."""

    def __init__(self, data, bandwidth):
        self.data = np.asarray(data, dtype=np.float64)
        self.bandwidth = float(bandwidth)

    def sample(self, n_samples=1, random_state=None):
        rng = np.random.mtrand._rand
        u = rng.uniform(0, 1, size=n_samples)
        i = (u * self.data.shape[0]).astype(np.int64)
        return np.atleast_2d(rng.normal(self.data[i], self.bandwidth))


def load_kde_pickle(path, *a, **k):
    """Stub-unpickle a KernelDensity pickle -> ShimKDE(training data [N,d] float64, bandwidth)."""
    from joblib.numpy_pickle import NumpyUnpickler

    class _U(NumpyUnpickler):
        def find_class(self, module, name):
            if module.startswith("sklearn"):
                return _new_obj if name == "newObj" else type(name, (_Stub,), {})
            return super().find_class(module, name)

    with open(path, "rb") as f:
        o = _U(path, f, ensure_native_byte_order=True).load()
    st = o.state if hasattr(o, "state") else o.__dict__
    return ShimKDE(np.asarray(st["tree_"].state[0]), st["bandwidth"])


class GenomicInterval:
    """HTSeq.GenomicInterval: 0-based, half-open [start, end) on ``chrom``/``strand``; ``length`` = end - start."""

    def __init__(self, chrom, start, end, strand="."):
        self.chrom, self.start, self.end, self.strand = chrom, int(start), int(end), strand

    @property
    def length(self):
        return self.end - self.start


class _GenomicFeature:
    def __init__(self, name, type_, iv, attr):
        self.name, self.type, self.iv, self.attr = name, type_, iv, attr


class GFF_Reader:
    """HTSeq.GFF_Reader: one GenomicFeature per non-comment line; GFF start/end are 1-based and inclusive, so
    iv.start = start - 1 and (end_included=True) iv.end = end; ``attr`` is the parsed 9th column and ``name`` the value
    of its first attribute (HTSeq.parse_GFF_attribute_string(..., extra_return_first_value=True))."""

    def __init__(self, filename, end_included=False):
        self.filename, self.end_included = filename, end_included

    def __iter__(self):
        with open(self.filename) as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                seqname, _source, feature, start, end, _score, strand, _frame, attr_str = line.rstrip("\n").split("\t", 8)
                attr, first = {}, None
                for tok in attr_str.rstrip(";").split(";"):
                    tok = tok.strip()
                    if not tok:
                        continue
                    if "=" in tok:
                        k, v = tok.split("=", 1)
                    else:
                        k, v = tok.split(None, 1)
                    v = v.strip().strip('"')
                    attr[k.strip()] = v
                    if first is None:
                        first = v
                iv = GenomicInterval(seqname, int(start) - 1, int(end) if self.end_included else int(end) - 1, strand)
                yield _GenomicFeature(first, feature, iv, attr)


class Fastafile:
    """pysam.Fastafile: ``references`` (record names up to the first whitespace), ``fetch(chrom, start, end)`` = the
    0-based half-open substring."""

    def __init__(self, path):
        self._seq, name, parts = {}, None, []
        with open(path) as f:
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        self._seq[name] = "".join(parts)
                    name, parts = line[1:].split()[0], []
                else:
                    parts.append(line.strip())
        if name is not None:
            self._seq[name] = "".join(parts)
        self.references = list(self._seq)

    def fetch(self, reference, start, end):
        return self._seq[reference][start:end]


def install():
    warnings.filterwarnings("ignore")
    for m in ("HTSeq", "pysam", "piecewise_regression"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["HTSeq"].GFF_Reader = GFF_Reader
    sys.modules["HTSeq"].GenomicInterval = GenomicInterval
    sys.modules["pysam"].Fastafile = Fastafile
    import joblib

    joblib.load = load_kde_pickle
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def load_reference_module():
    """Import the unmodified reference simulator as a module (``simulator``)."""
    install()
    import importlib

    return importlib.import_module("simulator")


def main():
    import runpy

    install()
    sys.argv = ["simulator.py"] + sys.argv[1:]
    runpy.run_path(REFERENCE_SRC + "/simulator.py", run_name="__main__")


if __name__ == "__main__":
    main()
