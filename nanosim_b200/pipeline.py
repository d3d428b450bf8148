"""Overlapped batch pipeline: the public way to run many batches.

``depth`` contexts (``ns_clone``: one copy of the reference and model in HBM, separate streams and batch buffers) are
driven by ``depth`` host threads (ctypes releases the GIL during library calls), so that while one batch is being copied
device->host into its pinned buffers the next batch's plan/emit kernels already run.  Results are handed to the consumer
strictly in submission order, which keeps output files identical to a sequential run.

This replaces the reference's ``for i in range(num_threads): mp.Process(...)`` fan-out
(/root/reference/src/simulator.py:1590-1622): same role (keep the machine busy), one GPU instead of N forks.
"""
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib as L
from .engine import Batch


class _HostBuffers:
    """Pinned (page-locked) host buffers for one in-flight batch; grown on demand."""

    def __init__(self, fastq):
        self.fastq = fastq
        self.cap = {"seq": 0, "qual": 0, "reads": 0, "pieces": 0, "ops": 0}
        self.t = {}

    def _alloc(self, nbytes):
        try:
            import torch
            if torch.cuda.is_available():
                return torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
        except Exception:
            pass
        return np.empty(int(nbytes), dtype=np.uint8)

    def ensure(self, key, nbytes, hint=0):
        if nbytes > self.cap[key]:
            want = int(max(nbytes, hint) * 1.25) + 4096
            self.t[key] = self._alloc(want)
            self.cap[key] = want
        t = self.t[key]
        return t.numpy() if hasattr(t, "numpy") and not isinstance(t, np.ndarray) else t

    def ptr(self, key):
        t = self.t[key]
        return t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data


class BatchPipeline:
    def __init__(self, engine, depth=2, fetch=True, want_ops=False, want_pieces=True):
        self.engines = [engine] + [engine.clone() for _ in range(max(1, depth) - 1)]
        self.depth = len(self.engines)
        self.fetch, self.want_ops, self.want_pieces = fetch, want_ops, want_pieces
        self.bufs = [_HostBuffers(engine.fastq) for _ in self.engines]
        self.pool = ThreadPoolExecutor(max_workers=self.depth)
        self.locks = [threading.Lock() for _ in self.engines]
        self.hint = {"seq": 0, "reads": 0, "pieces": 0, "ops": 0}     # largest batch seen by any slot (pinned allocs are slow)

    def close(self):
        self.pool.shutdown(wait=True)
        for e in self.engines[1:]:
            e.close()

    def _work(self, slot, job):
        kind, first, n = job
        eng, hb = self.engines[slot], self.bufs[slot]
        info = eng.simulate(kind, first, n)
        if not self.fetch:
            return info, None
        fastq = eng.fastq
        nb = {"seq": int(info.seq_bytes), "reads": int(info.n_reads) * L.READ_DTYPE.itemsize,
              "pieces": int(info.n_pieces) * L.PIECE_DTYPE.itemsize, "ops": int(info.n_ops) * 4}
        for k, v in nb.items():
            if v > self.hint[k]:
                self.hint[k] = v
        seq = hb.ensure("seq", nb["seq"], self.hint["seq"])[:nb["seq"]]
        qual = hb.ensure("qual", nb["seq"], self.hint["seq"])[:nb["seq"]] if fastq else None
        reads = hb.ensure("reads", nb["reads"], self.hint["reads"])[:nb["reads"]]
        pieces = ops = None
        if self.want_pieces:
            pieces = hb.ensure("pieces", nb["pieces"], self.hint["pieces"])[:nb["pieces"]]
        if self.want_ops and info.n_ops:
            ops = hb.ensure("ops", nb["ops"], self.hint["ops"])[:nb["ops"]]
        eng.fetch_into(hb.ptr("seq"), hb.ptr("qual") if fastq else None, hb.ptr("reads"),
                       hb.ptr("pieces") if pieces is not None else None, hb.ptr("ops") if ops is not None else None)
        b = Batch(info, seq, qual, reads.view(L.READ_DTYPE), pieces.view(L.PIECE_DTYPE) if pieces is not None else None,
                  ops.view(np.uint32) if ops is not None else np.zeros(0, dtype=np.uint32) if self.want_ops else None, kind, first)
        return info, b

    def run(self, jobs, consume=None):
        """jobs: iterable of (kind, first_read_id, n_reads).  consume(info, batch, job) runs on the calling thread in
        submission order; the batch's buffers are reused as soon as consume returns.  Returns the list of infos."""
        infos = []
        pending = deque()
        jobs = iter(jobs)
        i = 0

        def drain_one():
            fut, job = pending.popleft()
            info, b = fut.result()
            if consume is not None:
                consume(info, b, job)
            infos.append(info)

        for job in jobs:
            if len(pending) == self.depth:
                drain_one()
            pending.append((self.pool.submit(self._work, i % self.depth, job), job))
            i += 1
        while pending:
            drain_one()
        return infos
