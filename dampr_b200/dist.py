"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference moves map output to the reducers through files and a multiprocessing Queue
(DefaultShuffler.shuffle, base.py:416-433; StageRunner.run, stagerunner.py:15-43).  Here every rank
partitions its (already map-side combined) records by owner = mix64(key) % world on the device
(dampr_kv_partition_by_owner writes destination-contiguous send buffers, so no pack kernel precedes
the collective) and ONE variable-size all-to-all moves the payload; a counts all-to-all precedes it.
Only the shuffle exchanges data: inputs are sharded by byte range, results stay on their owner.
"""
import numpy as np


def active():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def world():
    import torch.distributed as dist
    return dist.get_rank(), dist.get_world_size()


class _DeviceMemory(object):
    """Expose library-owned device memory to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def device_bytes(ptr, nbytes):
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_DeviceMemory(ptr, nbytes), device="cuda")


def exchange_counts(counts):
    """counts[d] = records this rank sends to rank d -> recv[s] = records rank s sends here."""
    import torch
    import torch.distributed as dist
    dev_ = "cuda" if dist.get_backend() == "nccl" else "cpu"
    send = torch.as_tensor(np.asarray(counts, dtype=np.int64)).to(dev_)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    return recv.cpu().numpy().astype(np.int64)


def all_to_all_bytes(send_tensor, send_counts, recv_tensor, recv_counts, item_bytes=16):
    """Variable-size all-to-all of item_bytes-sized records held in uint8 tensors."""
    import torch.distributed as dist
    dist.all_to_all_single(recv_tensor, send_tensor,
                           output_split_sizes=[int(c) * item_bytes for c in recv_counts],
                           input_split_sizes=[int(c) * item_bytes for c in send_counts])


def all_reduce_sum_int(values):
    import torch
    import torch.distributed as dist
    dev_ = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(np.asarray(values, dtype=np.int64)).to(dev_)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().tolist()


def all_gather_objects(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def shuffle_kv(ctx, kv):
    """Exchange a device kv so that every key lives on its owner rank. Returns (received kv, offsets): the
    records of source rank s are recv[offsets[s]:offsets[s+1]]. The split by owner is stable, so a kv that is
    key-sorted arrives as one SORTED RUN per source rank — the input of the k-way merge (csrc/merge.cu)."""
    import torch
    _rank, n = world()
    parts, counts = kv.partition_by_owner(n)
    ctx.sync()
    recv_counts = exchange_counts(counts)
    total = int(recv_counts.sum())
    out = ctx.kv(max(1, total))
    out.set_size(total)
    send_t = device_bytes(parts.devptr(), int(counts.sum()) * 16)
    recv_t = device_bytes(out.devptr(), total * 16)
    all_to_all_bytes(send_t, counts, recv_t, recv_counts)
    torch.cuda.synchronize()
    parts.free()
    offsets = np.concatenate(([0], np.cumsum(recv_counts))).astype(np.uint64)
    return out, offsets
