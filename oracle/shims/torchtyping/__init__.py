"""Import shim (test infrastructure only): the reference annotates tensors with
torchtyping (src/models.py:11), which is not installed here. Only the names are
needed; no shape checking is performed."""
import torch


class _TensorTypeMeta(type):
    def __getitem__(cls, item):
        return torch.Tensor


class TensorType(metaclass=_TensorTypeMeta):
    pass


def patch_typeguard():
    return None
