"""
Data-parallel plumbing: one process per GPU (torchrun), images / tiles sharded by index, and the
only three exchanges the path has -- class histograms, confusion matrices, loss partials -- as
small all-reduces.  NCCL on GPUs; gloo in the CPU unit tests.

Integer sums are exactly associative, so histograms and confusion matrices are bit-identical for
any GPU count.  The reference has no distributed code (models/model.py:185-188 is commented out).
"""
import os

import numpy as np
import torch
import torch.distributed as td


def is_initialized():
    return td.is_available() and td.is_initialized()


def world_size():
    return td.get_world_size() if is_initialized() else 1


def rank():
    return td.get_rank() if is_initialized() else 0


def init_from_env(backend=None):
    """Join the torchrun rendezvous if WORLD_SIZE > 1.  Returns (rank, world_size, local_rank)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(lr)
            td.init_process_group(backend, device_id=torch.device("cuda", lr))
        else:
            td.init_process_group(backend)
    return rank(), world_size(), lr


def shard_indices(n_items, rank_=None, world=None):
    """Static round-robin: item i belongs to rank i mod world."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    return list(range(r, n_items, w))


def _comm_device():
    if is_initialized() and td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_reduce_(t):
    """In-place sum across ranks of a tensor already on the communication device (no host sync)."""
    if world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def all_reduce_i64(arr):
    t = torch.as_tensor(np.asarray(arr, dtype=np.int64)).to(_comm_device())
    return all_reduce_(t).cpu().numpy()


def all_reduce_f64(arr):
    t = torch.as_tensor(np.asarray(arr, dtype=np.float64)).to(_comm_device())
    return all_reduce_(t).cpu().numpy()


def barrier():
    if world_size() > 1:
        td.barrier()


def max_over_ranks(value):
    """Max of a python float across ranks (timing: the slowest rank defines the step)."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device())
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
