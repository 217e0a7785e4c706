"""Multi-GPU plumbing: spectra are independent, so the batch is sharded across ranks (one process per GPU) with no
collective in the solve and ONE gather of the packed results at the end."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank=None, world_size=None):
    """Contiguous block [start, stop) of rank's spectra; sizes differ by at most one."""
    if rank is None:
        rank, world_size = world()
    base, rem = divmod(n_total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_indices(n_total, rank=None, world_size=None, mode='strided'):
    """Global indices of rank's spectra.  'strided' (rank, rank + world_size, ...) interleaves the batch, which balances
    the load when iteration counts correlate with the position in the batch (SURVEY.md section 8e prefers it);
    'block' is shard_range.  Pass the indices to ``Inverter.fit(..., spectrum_ids=...)`` so that starts and random
    streams stay keyed by the global index, and to ``gather_results(..., indices=...)`` to restore the global order."""
    if rank is None:
        rank, world_size = world()
    if mode == 'block':
        a, b = shard_range(n_total, rank, world_size)
        return torch.arange(a, b, dtype=torch.int64)
    if mode != 'strided':
        raise ValueError(f"Invalid mode {mode}. Options are 'strided', 'block'")
    return torch.arange(rank, max(n_total, rank), world_size, dtype=torch.int64)  # (empty when rank >= n_total)


def gather_results(local, n_total=None, indices=None):
    """All-gather of per-spectrum result rows [n_local, P] -> [n_total, P] in global spectrum order (NCCL on GPUs, gloo
    on CPU tensors).  Shards may be ragged (sizes differ by one): rows are padded to the largest shard for the
    collective and trimmed afterwards."""
    rank, ws = world()
    if ws == 1:
        if indices is not None:
            out_ = torch.empty_like(local)
            out_[torch.as_tensor(indices, dtype=torch.int64, device=local.device)] = local
            return out_
        return local
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat((local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))))
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad.contiguous())
    res = torch.cat([o[:s] for o, s in zip(out, sizes)])
    if n_total is not None and res.shape[0] != n_total:
        raise RuntimeError(f'gathered {res.shape[0]} rows, expected {n_total}')
    if indices is not None:  # rows arrive rank by rank: put row i of rank r at its global index
        idx = torch.as_tensor(indices, dtype=torch.int64, device=local.device)
        ipad = idx if idx.shape[0] == mx else torch.cat((idx, idx.new_zeros(mx - idx.shape[0])))
        iall = [torch.empty_like(ipad) for _ in range(ws)]
        dist.all_gather(iall, ipad.contiguous())
        order = torch.cat([o[:s] for o, s in zip(iall, sizes)])
        out_ = torch.empty_like(res)
        out_[order] = res
        return out_
    return res
