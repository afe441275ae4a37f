"""Side stream for work that is off the critical path of the backward pass (weight / bias gradients:
nothing later in the backward reads them).  Such work is enqueued on one extra stream per device,
forked from the current stream when its inputs are ready, and joined once by an autograd-engine
callback at the end of the backward pass.  Under CUDA-graph capture the fork / join become graph
edges.  `DAT_B200_SERIAL_WGRAD=1` keeps everything on the current stream.
"""
import os

import torch

_SIDE = {}
_PENDING = []
_JOIN_QUEUED = [False]


def serial():
    return bool(os.environ.get("DAT_B200_SERIAL_WGRAD"))


def side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev)
    return _SIDE[key]


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
