"""Process-group helpers with the reference's names (src/dist_utils.py) plus the two fixed-shape
exchanges the B200 search path uses instead of the reference's 3 + 4*W var-size collectives
(src/index.py:127-143, src/dist_utils.py:46-113).
"""
import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def get_rank():  # src/dist_utils.py:125-130
    return dist.get_rank() if is_initialized() else 0


def get_world_size():  # src/dist_utils.py:137-142
    return dist.get_world_size() if is_initialized() else 1


def is_main():  # src/dist_utils.py:133-134
    return get_rank() == 0


def barrier():  # src/dist_utils.py:145-147
    if is_initialized():
        dist.barrier()


@torch.no_grad()
def get_varsize(x, dim=0):
    """Sizes of `x` along `dim` on every rank (src/dist_utils.py:102-113) as a python list."""
    if not is_initialized():
        return [x.size(dim)]
    size = torch.tensor([x.size(dim)], device=x.device, dtype=torch.int64)
    allsizes = [torch.zeros_like(size) for _ in range(get_world_size())]
    dist.all_gather(allsizes, size)
    return torch.cat(allsizes).cpu().tolist()


@torch.no_grad()
def all_gather_fixed(x):
    """all_gather of equally-shaped tensors -> [W, *x.shape] (one collective, no host sync)."""
    world = get_world_size()
    if world == 1:
        return x.unsqueeze(0)
    out = torch.empty((world,) + tuple(x.shape), dtype=x.dtype, device=x.device)
    try:
        dist.all_gather_into_tensor(out.view(-1), x.contiguous().view(-1))
    except (RuntimeError, AttributeError, NotImplementedError):
        dist.all_gather(list(out.unbind(0)), x.contiguous())
    return out


@torch.no_grad()
def varsize_all_gather(x, sizes=None):
    """Concatenation along dim 0 of per-rank tensors of different lengths (src/dist_utils.py:46-69):
    pad to the max length, ONE all_gather, drop the padding.  `sizes` may be passed if already known."""
    if not is_initialized():
        return x
    if sizes is None:
        sizes = get_varsize(x)
    max_size = max(sizes)
    if x.size(0) != max_size:
        pad = torch.zeros((max_size - x.size(0),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat((x, pad), dim=0)
    gathered = all_gather_fixed(x)
    return torch.cat([gathered[r, : sizes[r]] for r in range(len(sizes))], dim=0)
