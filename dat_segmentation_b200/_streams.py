"""Side stream for work that is off the critical path of the backward pass (weight / bias gradients:
nothing later in the backward reads them).  Such work is enqueued on one extra stream per device,
forked from the current stream when its inputs are ready, and joined once by an autograd-engine
callback at the end of the backward pass.  Under CUDA-graph capture the fork / join become graph
edges.  `DAT_B200_SERIAL_WGRAD=1` keeps everything on the current stream, and so does an initialised
torch.distributed process group unless `ALLOW_WITH_PROCESS_GROUP[0]` is set (see below).
"""
import os

import torch

_SIDE = {}
_PENDING = []
_JOIN_QUEUED = [False]


# A process group usually means DistributedDataParallel, whose gradient hooks hand a weight gradient to NCCL (on DDP's
# own stream, ordered only after the *current* stream) as soon as autograd has produced it - before the
# end-of-backward join of the side stream.  So with an initialised process group the side stream is off unless the
# caller opts in because it synchronises itself (bench.py packs and all-reduces the gradients after backward()).
ALLOW_WITH_PROCESS_GROUP = [False]


# Set by a caller whose gradient hooks order themselves after the side stream (bench.py's bucketed all-reduce waits
# for `existing_side_stream()` before it reads a gradient): such hooks do not force the early join.
HOOKS_SYNC_THEMSELVES = [False]


def _process_group_initialised():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def serial():
    if os.environ.get("DAT_B200_SERIAL_WGRAD"):
        return True
    return _process_group_initialised() and not ALLOW_WITH_PROCESS_GROUP[0]


def grads_consumed_at_end_only(*params):
    """True when nothing can read the gradients of `params` before the end-of-backward join: each parameter is a
    leaf without `.grad` yet (autograd then just stores the returned tensor; an existing `.grad` would make
    AccumulateGrad run `grad += dw` on the current stream while the side stream may still be writing dw -
    gradient accumulation over micro-batches, `zero_grad(set_to_none=False)`, shared weights) and without tensor
    hooks or post-accumulate-grad hooks (they would see dw before it is complete).  When this is False the caller
    joins the side stream into the current stream before returning the gradients."""
    for p in params:
        if p is None:
            continue
        if not isinstance(p, torch.Tensor) or not p.is_leaf or p.grad is not None:
            return False
        if not HOOKS_SYNC_THEMSELVES[0] and (getattr(p, "_backward_hooks", None)
                                             or getattr(p, "_post_accumulate_grad_hooks", None)):
            return False
    return True


def side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev)
    return _SIDE[key]


def existing_side_stream(dev):
    """The side stream of `dev` if one has been created (None otherwise)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    return _SIDE.get(key)


def _join_side_streams():
    _JOIN_QUEUED[0] = False
    for idx, side in _SIDE.items():
        torch.cuda.current_stream(idx).wait_stream(side)
    _PENDING.clear()           # tensors the side stream was reading may be reused from here on


def hold_until_join(*tensors):
    """Keep `tensors` (read by the side stream) alive until the end-of-backward join and make sure
    that join is queued."""
    _PENDING.append(tensors)
    if not _JOIN_QUEUED[0]:
        _JOIN_QUEUED[0] = True
        torch.autograd.Variable._execution_engine.queue_callback(_join_side_streams)
