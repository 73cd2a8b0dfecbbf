"""Two ranks, two GPUs, NCCL: the only exchange of the path -- pack kernel -> all_gather_into_tensor -> unpack -- checked for
CONTENT (only the count was asserted).  Each rank runs the hot path on its own shard of the pairs; every rank must
end up with the concatenation, in global pair order, of the two single-rank match lists, bit for bit, through both the
stream-ordered device path (casmtr_pack_matches_dev, count read on the device) and the host-synchronous variant.
Skipped on a single-GPU box (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ('b_ids', 'i_ids', 'j_ids', 'mconf', 'mkpts0', 'mkpts1')


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:      # noqa: BLE001  (the parent must hear about it instead of waiting for a rank that will never answer)
        import traceback
        q.put((rank, 'error', traceback.format_exc()[-2000:], None))


def _worker_body(rank, world, port, q):
    import torch.distributed as dist
    from casmtr_b200 import dist as cdist
    from casmtr_b200 import pipeline
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    pairs = 2 if rank == 0 else 1                               # uneven shards: 3 pairs over 2 ranks
    offset = 0 if rank == 0 else 2
    wl = pipeline.Workload(256, 256, pairs=pairs, config='4c', qt_layers=1)
    host = pipeline.make_host_inputs(wl, seed=100 + rank)
    hp = pipeline.HotPath(wl).to(dev)
    hp.load_level_weights(host)
    dev_in = pipeline.tree_map(lambda t: t.to(dev), host)
    mine = hp(dev_in)
    cap = 2 * wl.stages[0]['h'] * wl.stages[0]['w']             # the same static capacity on every rank
    # (a) host-synchronous gather of the trimmed list
    got_a = cdist.gather_matches({k: mine[k] for k in KEYS}, pair_offset=offset, cap=cap)
    # (b) stream-ordered: whole-step graph result (capacity-sized + device count) -> pack on device -> one all-gather
    gr = pipeline.GraphRunner(hp, dev_in)
    blocks, work = cdist.gather_matches_device(gr.step(), offset, cap, async_op=True)
    work.wait()
    got_b = cdist.unpack_gathered(blocks)
    q.put((rank, {k: mine[k].cpu().tolist() for k in KEYS}, {k: got_a[k].cpu().tolist() for k in KEYS},
           {k: got_b[k].cpu().tolist() for k in KEYS}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_gathered_match_list_equals_the_single_rank_lists():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=240) for _ in range(2)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()                                    # the exact processes this test started
    for r in res:
        assert r[1] != 'error', r[2]
    res = sorted(res, key=lambda t: t[0])
    (_, m0, a0, b0), (_, m1, a1, b1) = res
    assert len(m0['b_ids']) > 10 and len(m1['b_ids']) > 10
    want = {k: m0[k] + m1[k] for k in KEYS}
    want['b_ids'] = m0['b_ids'] + [b + 2 for b in m1['b_ids']]  # rank 1's pair ids become global
    for got in (a0, a1, b0, b1):                                 # both variants, both ranks
        for k in KEYS:
            assert got[k] == want[k], k
