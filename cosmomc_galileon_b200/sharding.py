"""Parameter-point sharding across ranks (parameter-batch mode).

Ranks own contiguous, disjoint blocks of a batch; the data path has no collective.  The one exchange step is the
gather of the final C_l (3 x (lmax-1) float64 per point), done with torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  Uneven blocks are padded to the largest block for the all-gather and trimmed afterwards.
"""
import numpy as np


def shard_range(n, rank, world):
    """[lo, hi) of the points rank owns: blocks differ by at most one point; earlier ranks get the extra ones."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_models(models, rank, world):
    lo, hi = shard_range(len(models), rank, world)
    return models[lo:hi]


_gather_state = {}


def gather_cls_to_root(local_cls, n_total, device=None, root=0):
    """Gather of the per-rank C_l blocks on ONE rank (the one that feeds the likelihoods): (n_total, 3, nL) in the
    original point order on `root`, None elsewhere.  The blocks travel device to device (NCCL over NVLink when `device`
    is a GPU); only the root copies the assembled batch back to the host.  The returned array is the root's (pinned,
    reused) result buffer: valid until the next call."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    nmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    shape = (nmax,) + tuple(local_cls.shape[1:])
    if local_cls.shape[0] == nmax and local_cls.dtype == np.float64 and local_cls.flags.c_contiguous:
        pad = local_cls  # equal shares (the usual case): no staging copy
    else:
        pad = np.zeros(shape, dtype=np.float64)
        pad[: local_cls.shape[0]] = local_cls
    # device buffers, the root's receive blocks and its pinned result are kept across calls (one batch shape per run)
    key = (shape, str(device), world, rank, root, n_total)
    st = _gather_state.get(key)
    if st is None:
        dev = device if device is not None else torch.device("cpu")
        st = {"mine": torch.empty(shape, dtype=torch.float64, device=dev)}
        if rank == root:
            st["bufs"] = [torch.empty(shape, dtype=torch.float64, device=dev) for _ in range(world)]
            st["out"] = torch.empty((n_total,) + shape[1:], dtype=torch.float64,
                                    pin_memory=(device is not None and torch.cuda.is_available()))
        _gather_state.clear()
        _gather_state[key] = st
    st["mine"].copy_(torch.from_numpy(pad))
    dist.gather(st["mine"], st.get("bufs"), dst=root)
    if rank != root:
        return None
    out = st["out"]
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi].copy_(st["bufs"][r][: hi - lo], non_blocking=True)
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)
    return out.numpy()


def gather_cls(local_cls, n_total, device=None):
    """All-gather of per-rank C_l blocks -> (n_total, 3, nL) on every rank, in the original point order."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    nmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    shape = (nmax,) + tuple(local_cls.shape[1:])
    pad = np.zeros(shape, dtype=np.float64)
    pad[: local_cls.shape[0]] = local_cls
    mine = torch.from_numpy(pad)
    if device is not None:
        mine = mine.to(device)
    out = torch.empty((world,) + shape, dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out.view(world * nmax, *shape[1:]), mine)
    out = out.cpu().numpy()
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r, : hi - lo])
    return np.concatenate(parts, axis=0)
