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


# ---- peer-addressable exchange workspace of the fused data-parallel loss ------------------------------------
_exchange = {"state": None}


def loss_exchange():
    """The symmetric-memory workspace pylc_multiloss_fwd_bwd_dp exchanges its partials through, created on first
    use (a COLLECTIVE: every rank must call it at the same point).  Returns a dict {ptrs_dev, rank, world,
    next_epoch()} or None when the group has no peer-addressable memory (gloo, one rank, torch without symmetric
    memory, PYLC_DP_FUSED=0): callers then use the two-launch route with an NCCL all-reduce in between.  All ranks
    agree on the outcome (one MIN all-reduce of the success flags)."""
    st = _exchange["state"]
    if st is not None:
        return st or None
    ok, st = 0, {}
    if world_size() > 1 and td.get_backend() == "nccl" and os.environ.get("PYLC_DP_FUSED", "1") != "0" and world_size() <= 16:
        try:
            import torch.distributed._symmetric_memory as symm
            dev = torch.device("cuda", torch.cuda.current_device())
            group = td.group.WORLD
            if hasattr(symm, "enable_symm_mem_for_group"):
                try:
                    symm.enable_symm_mem_for_group(group.group_name)
                except Exception:
                    pass
            buf = symm.empty(2048 // 8, dtype=torch.float64, device=dev)        # ops.DP_WS_BYTES
            buf.zero_()
            torch.cuda.synchronize()
            hdl = symm.rendezvous(buf, group)
            st = {"buf": buf, "hdl": hdl, "ptrs_dev": int(hdl.buffer_ptrs_dev), "rank": rank(), "world": world_size(), "epoch": 0}
            ok = 1
        except Exception as ex:       # noqa: BLE001 -- any failure means: no peer memory, use NCCL
            st = {"error": repr(ex)}
    if world_size() > 1:
        flag = torch.tensor([ok], dtype=torch.int32, device=_comm_device())
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        ok = int(flag.item())         # also orders every rank's zero-fill before anybody's first exchange
    if not ok:
        _exchange["state"] = False
        _exchange["error"] = st.get("error")
        return None

    def next_epoch():
        st["epoch"] += 1
        return st["epoch"]
    st["next_epoch"] = next_epoch
    _exchange["state"] = st
    return st


def max_over_ranks(value):
    """Max of a python float across ranks (timing: the slowest rank defines the step)."""
    if world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device())
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())
