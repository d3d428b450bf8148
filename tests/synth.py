"""Deterministic synthetic references: i.i.d. uniform ACGT from numpy default_rng(seed)."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def synth_chrom(rng, n):
    return _ACGT[rng.integers(0, 4, n, dtype=np.uint8)]


def write_fasta(path, records, width=80):
    with open(path, "w") as f:
        for name, arr in records:
            f.write(">" + name + "\n")
            s = arr.tobytes().decode()
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def ecoli5m(path=None):
    """BASELINE config 1: one chromosome ``chr1`` of 5,000,000 i.i.d. bases, default_rng(0)."""
    rng = np.random.default_rng(0)
    rec = [("chr1", synth_chrom(rng, 5_000_000))]
    if path:
        write_fasta(path, rec)
    return rec
