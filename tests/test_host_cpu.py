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
