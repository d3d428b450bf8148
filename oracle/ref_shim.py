"""Import/run shim for the UNMODIFIED reference (/root/reference/src/simulator.py).

TEST INFRASTRUCTURE ONLY.  Nothing under nanosim_b200/ may import this file.  It only
works in the build container (the GPU box has no /root/reference); its two jobs are

  * ``python oracle/ref_shim.py genome -rg ... -c ...``  runs the reference CLI unmodified
    (used by tests/golden/make_golden_runs.py to produce the run-level golden histograms),
  * ``load_reference_module()`` imports the reference as a module so that
    tests/golden/make_golden_vectors.py can call its per-read functions
    (error_list, mutate_read, ...) under a fixed seed and commit the input/output vectors.

What the shim does:
  1. stubs ``HTSeq``, ``pysam``, ``piecewise_regression`` -- imported at simulator.py:15-16,
     model_base_qualities.py:4, model_homopolymer_lengths.py:6, used only by the
     intron-retention branch / the training side;
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


def install():
    warnings.filterwarnings("ignore")
    for m in ("HTSeq", "pysam", "piecewise_regression"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
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
