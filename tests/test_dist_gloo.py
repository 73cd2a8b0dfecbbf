"""CPU, world_size 2, gloo: batch sharding and the variable-length match-list all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casmtr_b200 import dist as cdist


def test_shard_range_partitions_everything():
    for n in (1, 7, 16, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [cdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _matches(n, seed):
    g = torch.Generator().manual_seed(seed)
    return {'b_ids': torch.randint(0, 3, (n,), generator=g).sort()[0], 'i_ids': torch.randint(0, 999, (n,), generator=g),
            'j_ids': torch.randint(0, 999, (n,), generator=g), 'mconf': torch.rand(n, generator=g),
            'mkpts0': torch.rand(n, 2, generator=g), 'mkpts1': torch.rand(n, 2, generator=g)}


def test_pack_roundtrip():
    m = _matches(17, 0)
    r = cdist.unpack_matches(cdist.pack_matches(m, pair_offset=5))
    assert torch.equal(r['b_ids'], m['b_ids'] + 5) and torch.equal(r['j_ids'], m['j_ids'])
    assert torch.equal(r['mkpts1'], m['mkpts1']) and torch.equal(r['mconf'], m['mconf'])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_pairs = 5
    lo, hi = cdist.shard_range(n_pairs, rank, world)
    mine = _matches(0 if rank == 1 else 11, seed=rank)          # rank 1 contributes an EMPTY list
    out = cdist.gather_matches(mine, pair_offset=lo)
    capped = cdist.gather_matches(mine, pair_offset=lo, cap=16)          # single fixed-size all-gather variant
    assert all(torch.equal(out[k], capped[k]) for k in out)
    q.put((rank, lo, hi, {k: v.tolist() for k, v in out.items()}))       # plain lists: no shared-memory handles to outlive the worker
    dist.barrier()
    dist.destroy_process_group()


def test_gather_matches_world2():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, lo0, hi0, a), (_, lo1, hi1, b) = res
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 5)
    want = _matches(11, seed=0)
    for k in a:                                                  # both ranks hold the same global list
        assert a[k] == b[k]
    assert a['b_ids'] == want['b_ids'].tolist() and a['mkpts0'] == want['mkpts0'].tolist()
    assert len(a['b_ids']) == 11


def test_cpulist_parse():
    from casmtr_b200.dist import _parse_cpulist
    assert _parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist('') == set()
