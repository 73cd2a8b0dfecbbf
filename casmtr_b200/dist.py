"""Multi-GPU plumbing: image pairs are independent, so the hot path shards the batch contiguously
over ranks with no data-path collective; the only exchange is the final variable-length match
list.  The reference's analogue is a pickled gloo ``gather`` of result dicts
(src/utils/comm.py:180-220, src/lightning/lightning_cascade.py:388-396); here it is one all-gather
of the counts and one padded all-gather of a packed 44-byte-per-match record, on whatever backend
the process group uses (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

RECORD_BYTES = 3 * 8 + 5 * 4      # b_id, i_id, j_id (int64) + mconf, mkpts0[2], mkpts1[2] (fp32)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index, sysfs='/sys'):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated: first
    touch then places the staging buffers in memory local to the GPU's PCIe root, so the host->device copies of 8 ranks do
    not all cross the socket interconnect.  Returns {'node': n, 'cpus': count} or None when the topology is not exposed
    (containers without sysfs, single-node hosts): binding is an optimisation, never a requirement."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
        with open(f'{sysfs}/bus/pci/devices/{bdf}/numa_node') as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f'{sysfs}/devices/system/node/node{node}/cpulist') as fh:
            cpus = _parse_cpulist(fh.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {'node': node, 'cpus': len(cpus)}
    except (OSError, AttributeError, ValueError):
        return None


def shard_range(n_pairs, rank, world):
    """Contiguous split of n_pairs over `world` ranks: rank r gets [lo, hi)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_matches(m, pair_offset=0):
    """dict(b_ids,i_ids,j_ids,mconf,mkpts0,mkpts1) -> uint8 [M, 44]; b_ids become global pair ids."""
    ids = torch.stack([m['b_ids'] + pair_offset, m['i_ids'], m['j_ids']], dim=1).contiguous()
    fl = torch.cat([m['mconf'].reshape(-1, 1), m['mkpts0'].reshape(-1, 2), m['mkpts1'].reshape(-1, 2)], dim=1)
    fl = fl.to(torch.float32).contiguous()
    M = ids.shape[0]
    return torch.cat([ids.view(torch.uint8).reshape(M, 24), fl.view(torch.uint8).reshape(M, 20)], dim=1)


def unpack_matches(buf):
    M = buf.shape[0]
    ids = buf[:, :24].reshape(-1).clone().view(torch.int64).reshape(M, 3)
    fl = buf[:, 24:].reshape(-1).clone().view(torch.float32).reshape(M, 5)
    return {'b_ids': ids[:, 0], 'i_ids': ids[:, 1], 'j_ids': ids[:, 2], 'mconf': fl[:, 0],
            'mkpts0': fl[:, 1:3], 'mkpts1': fl[:, 3:5]}


def gather_matches_device(m, pair_offset, cap, group=None, async_op=False):
    """Stream-ordered variant for CUDA tensors: one pack kernel (libcasmtr_b200) + one fixed-size NCCL all-gather, no host
    synchronisation.  Returns the gathered blocks [world, cap+1, 44] uint8 on the device; unpack_gathered() turns them
    into the match dict when the host needs it.
    async_op: return (blocks, work) and do NOT make the compute stream wait for the collective -- the next pair's kernels
    do not depend on it, so it overlaps them; call work.wait() before reading the blocks."""
    from . import functional as F
    block = F.pack_matches(m, pair_offset, cap)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return (block.unsqueeze(0), None) if async_op else block.unsqueeze(0)
    allb = torch.empty(world * (cap + 1), RECORD_BYTES, dtype=torch.uint8, device=block.device)
    work = dist.all_gather_into_tensor(allb, block, group=group, async_op=async_op)
    allb = allb.reshape(world, cap + 1, RECORD_BYTES)
    return (allb, work) if async_op else allb


def unpack_gathered(allb):
    counts = allb[:, 0, :8].reshape(-1).clone().view(torch.int64).tolist()                # the one host read
    return unpack_matches(torch.cat([allb[r, 1:c + 1] for r, c in enumerate(counts)], dim=0))


def gather_matches(m, pair_offset=0, group=None, cap=None):
    """All ranks receive the concatenation (in rank order, i.e. global pair order) of every rank's
    match list.  Without an initialised process group this is the identity.
    cap: static per-rank capacity (>= any rank's match count).  With it the exchange is ONE fixed-size all-gather
    (row 0 of every rank's block carries its count) and one host read; without it the counts are exchanged first."""
    buf = pack_matches(m, pair_offset)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_matches(buf)
    world = dist.get_world_size(group)
    if cap is not None:
        n = buf.shape[0]
        if n > cap:
            raise RuntimeError(f'gather_matches: {n} matches exceed the static capacity {cap}')
        block = torch.zeros(cap + 1, RECORD_BYTES, dtype=torch.uint8, device=buf.device)
        block[0, :8] = torch.tensor([n], dtype=torch.int64, device=buf.device).view(torch.uint8)
        block[1:n + 1] = buf
        allb = torch.empty(world * (cap + 1), RECORD_BYTES, dtype=torch.uint8, device=buf.device)
        dist.all_gather_into_tensor(allb, block, group=group)
        allb = allb.reshape(world, cap + 1, RECORD_BYTES)
        counts = allb[:, 0, :8].reshape(-1).clone().view(torch.int64).tolist()               # the one host read
        return unpack_matches(torch.cat([allb[r, 1:c + 1] for r, c in enumerate(counts)], dim=0))
    count = torch.tensor([buf.shape[0]], dtype=torch.int64, device=buf.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros(cap, RECORD_BYTES, dtype=torch.uint8, device=buf.device)
    padded[:buf.shape[0]] = buf
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return unpack_matches(torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0))
