"""Wire format of the world<->planner boundary: tensors as ``torch.save`` bytes
(same two functions as ``mppiisaac/utils/transport.py:5-14``)."""
import io

import torch


def torch_to_bytes(t: torch.Tensor) -> bytes:
    with io.BytesIO() as buf:
        torch.save(t, buf)
        return buf.getvalue()


def bytes_to_torch(b) -> torch.Tensor:
    if isinstance(b, torch.Tensor):      # in-process callers may skip the pickling round trip
        return b
    return torch.load(io.BytesIO(b), weights_only=True)
