"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Read groups shard contiguously over the ranks (pairs are never split); the graph is replicated.  The abundance
pass has exactly two exchange steps:

1. after pass 1: all-reduce(SUM) of ``uniq_reads`` int64[P] -- the only quantity pass 2 needs from other shards
   (json2csv.py:1141,1194,1243); exact, independent of the sharding;
2. after pass 2: per-path scalars, histograms and counters are all-reduced; the two fp64 slot arrays are
   reduce-scattered onto a path-partitioned layout (rank k receives the slots of its contiguous path range),
   each rank reduces its own paths, and the small per-path table is all-gathered.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_groups, world, rank):
    """Contiguous range of read groups owned by ``rank``."""
    return n_groups * rank // world, n_groups * (rank + 1) // world


def path_partition(path_off_slots, world):
    """Split paths into ``world`` contiguous ranges of roughly equal slot count.  ``path_off_slots`` is the
    int array [P+1] of first slots (device layout).  Returns int64[world+1] path boundaries."""
    P = len(path_off_slots) - 1
    n_slots = int(path_off_slots[-1])
    targets = (np.arange(1, world, dtype=np.int64) * n_slots) // world
    cuts = np.searchsorted(np.asarray(path_off_slots[:-1], np.int64), targets, side="left")
    return np.concatenate(([0], np.minimum(cuts, P), [P])).astype(np.int64)


def all_reduce_sum(t, group):
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def reduce_scatter_ranges(t, bounds, group):
    """In place: after the call ``t[bounds[k]:bounds[k+1]]`` on rank k holds the sum over ranks of that range
    (other ranges are left with partial sums).  One reduce per destination, issued back to back."""
    world = dist.get_world_size(group)
    works = []
    for k in range(world):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        if hi > lo:
            works.append(dist.reduce(t[lo:hi], dst=dist.get_global_rank(group, k) if group is not None else k,
                                     op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    return t


class RangeReduceScatter:
    """Sum over ranks of several arrays that share one partition into contiguous ranges (``bounds``): rank k ends up
    with the sums of range k of every array -- ONE reduce_scatter_tensor per call instead of one reduce per destination
    and array.  The ranges are unequal (they end at path boundaries), the collective wants equal chunks: a gather
    index, built once, lays range k of array j at ``[k][j][0:len_k]`` of a staging buffer padded to the longest range
    (padding reads the LAST element of the source, which the caller keeps at zero).  ``arrays``: 1-D tensors of equal
    length and dtype, views of ONE flat tensor ``flat`` (the index addresses ``flat``); results are written back in place."""

    def __init__(self, flat, arrays, bounds, group):
        self.group, self.flat, self.arrays = group, flat, arrays
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        b = [int(x) for x in bounds]
        self.lo, self.hi = b[self.rank], b[self.rank + 1]
        self.chunk = max(1, max(b[k + 1] - b[k] for k in range(self.world)))
        base = [(a.data_ptr() - flat.data_ptr()) // flat.element_size() for a in arrays]
        zero = flat.numel() - 1                               # the spare element that stays zero
        idx = np.full((self.world, len(arrays), self.chunk), zero, np.int64)
        for k in range(self.world):
            n = b[k + 1] - b[k]
            for j, off in enumerate(base):
                idx[k, j, :n] = off + b[k] + np.arange(n, dtype=np.int64)
        self.idx = torch.from_numpy(idx.reshape(-1)).to(flat.device)
        self.send = torch.empty(self.idx.numel(), dtype=flat.dtype, device=flat.device)
        self.recv = torch.empty(len(arrays) * self.chunk, dtype=flat.dtype, device=flat.device)

    def __call__(self):
        torch.index_select(self.flat, 0, self.idx, out=self.send)
        dist.reduce_scatter_tensor(self.recv, self.send, op=dist.ReduceOp.SUM, group=self.group)
        n = self.hi - self.lo
        for j, a in enumerate(self.arrays):
            a[self.lo:self.hi].copy_(self.recv[j * self.chunk:j * self.chunk + n])


def all_gather_rows(local_rows, row_bounds, group):
    """local_rows: [n_k, C] on each rank (n_k = row_bounds[k+1]-row_bounds[k]); returns the full [N, C] on every rank."""
    world = dist.get_world_size(group)
    counts = [int(row_bounds[k + 1] - row_bounds[k]) for k in range(world)]
    width = local_rows.shape[1]
    pad = max(counts) if counts else 0
    send = torch.zeros((pad, width), dtype=local_rows.dtype, device=local_rows.device)
    send[:local_rows.shape[0]] = local_rows
    recv = torch.empty((world, pad, width), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return torch.cat([recv[k, :counts[k]] for k in range(world)], dim=0)


def bind_to_gpu_numa(device_index):
    """One process per GPU: restrict this process to the host cores NVML lists as local to its GPU, so that the pinned
    staging arenas it allocates afterwards (first touch) and the packer threads sit on the GPU's own NUMA node -- with
    eight ranks pulling their shards from one host, uploads that cross the socket interconnect share its bandwidth.
    Returns the cores bound to, or None when nothing was changed (no NVML, no affinity information, or the mask would
    not narrow the current one)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        now = os.sched_getaffinity(0)
        want = local & now
        if not want or want == now or len(want) < 2:
            return None
        os.sched_setaffinity(0, want)
        return sorted(want)
    except Exception:
        return None
