"""N>1 host logic on CPU: two `gloo` ranks shard one batch, each maps its shard, rank 0 merges in input order.

The per-rank aligner here is the CPU oracle (tests may use it; there is no GPU in the CPU suite). What is under test is
the sharding contract of DESIGN.md §6: contiguous ranges with GLOBAL read indices give records identical to a
single-process run, including the -S tie-breaks, and the ordered gather returns them in input order.
"""
import os
import socket
import sys

import numpy as np
import pytest

import helpers
from basal_b200 import capi, shard


def test_shard_range_partitions_the_input():
    for n in (0, 1, 7, 50000, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [shard.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_shard_batch_keeps_global_indices_and_ragged_offsets():
    seqs = ["ACGT" * 5, "A" * 33, "", "ACGTN" * 7, "C" * 16]
    b = capi.ReadBatch.from_strings(seqs, readset=1, first_index=100)
    parts = [shard.shard_batch(b, r, 2) for r in range(2)]
    assert [p.n for p in parts] == [3, 2] and [p.first_index for p in parts] == [100, 103]
    assert parts[1].offsets[0] == 0 and bytes(parts[1].bases) == (seqs[3] + seqs[4]).encode()
    assert all(p.readset == 1 for p in parts)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, pe, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, chrs, m1, m2 = helpers.small_case(2 if pe else 1, 0.0005 if pe else 0.02, limit=1001)
        params = helpers.flags_to_params(cfg, {"w": 3})
        cat, offs, lens = helpers.synth.reference_ascii(chrs)
        orc = helpers.oracle_context(params); orc.index_build(cat, offs, lens)       # every rank: its own replica of the index
        a = capi.ReadBatch.from_matrix(m1, readset=1 if pe else 0, first_index=17)
        if pe:
            b = capi.ReadBatch.from_matrix(m2, readset=2, first_index=17)
            ra, rb, rp = orc.align_pe(shard.shard_batch(a, rank, world), shard.shard_batch(b, rank, world))
            got = [shard.gather_records(x, dist) for x in (ra, rb, rp)]
        else:
            got = [shard.gather_records(orc.align_se(shard.shard_batch(a, rank, world)), dist)]
        if rank == 0:
            want = list(orc.align_pe(a, b)) if pe else [orc.align_se(a)]
            ok = all(g is not None and g.dtype == w.dtype and len(g) == len(w) and g.tobytes() == w.tobytes() for g, w in zip(got, want))
            multi = int(np.sum(want[0]["n_hits"] > 1))
            open(out_path, "w").write(f"{int(ok)} {len(want[0])} {multi}")
        else:
            assert all(g is None for g in got)
        orc.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pe", [False, True])
def test_two_gloo_ranks_reproduce_the_single_process_records(pe, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), pe, out), nprocs=2, join=True)
    ok, n, multi = open(out).read().split()
    assert ok == "1" and int(n) == 1001
    assert int(multi) > 0, "the case must contain multi-hit reads, otherwise the -S tie-break is not exercised"
