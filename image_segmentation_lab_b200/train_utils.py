"""``parse_losses`` with the reference's signature and results (utils/train_utils.py:31-74), minus its host syncs.

The reference issues, per logged variable, one ``dist.all_reduce`` and one ``.item()`` (a device->host
synchronisation each, plus one more all_reduce for the length check). Here all variables of a step are stacked
into ONE device vector: one all_reduce (which carries the length check as an extra element) and one
device->host copy. ``lazy=True`` skips even that copy and returns the device vector for the caller to read later.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


def parse_losses(losses, lazy=False):
    """(loss, log_vars): ``loss`` = sum of every entry whose key contains 'loss' (stays in the autograd graph),
    ``log_vars`` = OrderedDict name -> python float, averaged over ranks when a process group is initialised,
    with the total under 'loss' — exactly the reference's layout. With ``lazy=True`` ``log_vars`` maps names to
    0-d views of one device tensor instead (no host synchronisation)."""
    log_vars = OrderedDict()
    for loss_name, loss_value in losses.items():
        if loss_name.startswith('_'):
            continue  # side channels such as '_stats'
        if isinstance(loss_value, torch.Tensor):
            log_vars[loss_name] = loss_value.mean()
        elif isinstance(loss_value, list):
            log_vars[loss_name] = sum(_loss.mean() for _loss in loss_value)
        else:
            raise TypeError(f'{loss_name} is not a tensor or list of tensors')
    loss = sum(_value for _key, _value in log_vars.items() if 'loss' in _key)
    log_vars['loss'] = loss

    names = list(log_vars.keys())
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    # torch.full is a device-side fill (a host tensor copied to the GPU here would break CUDA-graph capture of the step)
    vec = torch.stack([log_vars[k].detach().to(torch.float32).reshape(()) for k in names] +
                      [torch.full((), float(len(names)), dtype=torch.float32, device=loss.device)])
    if world > 1:
        dist.all_reduce(vec)
    if lazy:
        out = OrderedDict((k, vec[i] / world) for i, k in enumerate(names))
        return loss, out
    host = vec.tolist()   # the single device->host copy of the step
    if world > 1:
        # same guard as the reference: ranks must log the same variables, or collectives would hang later
        assert int(round(host[-1])) == len(names) * world, \
            'loss log variables are different across GPUs!\n' + \
            f'rank {dist.get_rank()} len(log_vars): {len(names)} keys: ' + ','.join(names) + '\n'
    return loss, OrderedDict((k, host[i] / world) for i, k in enumerate(names))
