"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference moves map output to the reducers through files and a multiprocessing Queue
(DefaultShuffler.shuffle, base.py:416-433; StageRunner.run, stagerunner.py:15-43).  Here every rank
partitions its (already map-side combined) records by owner = mix64(key) % world on the device
(dampr_kv_partition_by_owner writes destination-contiguous send buffers, so no pack kernel precedes
the collective) and grouped NCCL send/recv move the payload; an all-gather of the counts (plus a small header:
line counts, flags) precedes it — all inside ONE C-ABI call (shuffle_kv -> dampr_kv_all_to_all, csrc/comm.cu).
torch.distributed is the launcher-facing bootstrap (rank / world, the NCCL unique id, object gathers of the rare
host-side strings); the helpers below (exchange_counts, all_to_all_bytes) remain for host-side tests.
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


def _comm(ctx):
    """The context's NCCL communicator of the C-ABI exchange (csrc/comm.cu), created on first use: rank 0 makes
    the unique id, torch.distributed (already the launcher's bootstrap) hands it to every rank."""
    if getattr(ctx, "comm", None) is None:
        import torch.distributed as dist
        rank, n = world()
        box = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_create(rank, n, box[0])
    return ctx.comm


def shuffle_kv(ctx, kv, header=None):
    """Exchange a device kv so that every key lives on its owner rank. Returns (received kv, offsets, headers):
    the records of source rank s are recv[offsets[s]:offsets[s+1]]. The split by owner is stable, so a kv that
    is key-sorted arrives as one SORTED RUN per source rank — the input of the k-way merge (csrc/merge.cu).
    `header`: a few int64 values of this rank that travel with the counts (one all-gather); headers[s] is rank
    s's row. The whole exchange is ONE C-ABI call (dampr_kv_all_to_all: partition by owner, counts + header
    all-gather, grouped NCCL send/recv between the kv buffers) with a single host synchronisation."""
    _comm(ctx)
    return ctx.kv_all_to_all(kv, header)
