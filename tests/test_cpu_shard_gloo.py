"""CPU tier: the N>1 host path (contiguous sharding by ciphertext, key broadcast, output gather) on world_size 2 with
the gloo backend.  The per-shard compute is the oracle here (the GPU box runs the CUDA library in its place)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import _pkg
    _pkg.load_package()
    from learn_fhe_b200 import shard
    from oracle import orc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    P = orc.fhew_testing_param()
    shapes = dict(ksk_a=(P.n * P.ks_d, P.n_s), ksk_b=(P.n * P.ks_d,), brk=(P.n_s, 2 * P.rgsw_d, 2, P.n), ak=(P.w + 1, P.rlwe_d, 2, P.n))
    if rank == 0:
        K0 = orc.FhewKey(P, 0x5EED0001)
        ex = K0.export()
        arrays = [ex[k] for k in ("ksk_a", "ksk_b", "brk", "ak")] + [ex["ak_t"]]
    else:
        arrays = [np.zeros(shapes[k], dtype=np.uint64) for k in ("ksk_a", "ksk_b", "brk", "ak")] + [np.zeros(P.w + 1, dtype=np.int64)]
    arrays = shard.broadcast_host_arrays(dist, arrays, root=0)
    K = orc.FhewKey.from_arrays(P, *arrays)
    cts = orc.residues(77, 5 * (P.n + 1), P.big_q).reshape(5, P.n + 1)  # 5 ciphertexts over 2 ranks: ragged shards 3 + 2
    out = shard.sharded_map(dist, lambda rows: K.op([1, 1, 1, 0], rows, threads=2), cts)
    empty = shard.sharded_map(dist, lambda rows: rows, cts[:1])  # 1 item over 2 ranks: rank 1 has an empty shard
    if rank == 0:
        full = K.op([1, 1, 1, 0], cts, threads=4)
        q.put(bool((out == full).all()) and bool((empty == cts[:1]).all()))
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    import _pkg
    _pkg.load_package()
    from learn_fhe_b200.shard import shard_range
    for count in (0, 1, 5, 16384, 16385, 7):
        for world in (1, 2, 4, 8):
            spans = [shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharded_bootstrap():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
