"""Reference loading: the product-side mirror of read_profile's FASTA half
(/root/reference/src/simulator.py:341-349 with readfq :709-740).

The sequence is kept exactly as in the file (case and IUPAC codes included) as one uint8 per base, all
chromosomes concatenated in file order; ``case_convert`` (:743-755) happens per read on the device, as in the
reference.  Chromosome keys follow :344-347: header up to the first space, ``[_\\s]`` runs replaced by ``-``,
truncated at the first ``.``.
"""
import re

import numpy as np


def normalise_name(header):
    raw = header.partition(" ")[0]
    return "-".join(re.split(r"[_\s]\s*", raw)).split(".")[0]


class PackedReference:
    def __init__(self, names, bases, offsets):
        self.names = list(names)
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        assert len(self.offsets) == len(self.names) + 1
        assert int(self.offsets[-1]) == len(self.bases)

    @property
    def lengths(self):
        return np.diff(self.offsets.astype(np.int64))

    @property
    def genome_len(self):
        return int(self.offsets[-1])

    @property
    def max_chrom(self):
        return int(self.lengths.max()) if len(self.names) else 0

    @staticmethod
    def from_records(records):
        """records: iterable of (header, uint8 array or bytes or str)."""
        names, parts, offs = [], [], [0]
        seen = {}
        for header, seq in records:
            if isinstance(seq, str):
                seq = seq.encode()
            arr = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
            key = normalise_name(header)
            if key in seen:             # seq_dict[key] = seqS: a repeated key overwrites the earlier record (:346)
                i = seen[key]
                parts[i] = arr
            else:
                seen[key] = len(names)
                names.append(key)
                parts.append(arr)
        for p in parts:
            offs.append(offs[-1] + len(p))
        bases = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
        return PackedReference(names, bases, np.asarray(offs, dtype=np.uint64))

    @staticmethod
    def from_fasta(path):
        with open(path, "rb") as f:
            blob = f.read()
        recs = []
        # records start at '>' at the beginning of a line
        starts = [0] if blob[:1] == b">" else []
        pos = blob.find(b"\n>")
        while pos != -1:
            starts.append(pos + 1)
            pos = blob.find(b"\n>", pos + 1)
        for i, s in enumerate(starts):
            e = starts[i + 1] if i + 1 < len(starts) else len(blob)
            nl = blob.find(b"\n", s, e)
            if nl == -1:
                nl = e
            header = blob[s + 1:nl].decode()
            body = blob[nl + 1:e].replace(b"\n", b"").replace(b"\r", b"")
            recs.append((header, body))
        return PackedReference.from_records(recs)
