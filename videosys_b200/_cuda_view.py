"""Wrap a raw device pointer (cudaMalloc'ed by the C-ABI) as a torch tensor without copying."""
import torch


class _Holder:
    def __init__(self, ptr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape),
            "typestr": typestr,
            "data": (int(ptr), False),
            "version": 3,
            "strides": None,
        }


def device_view(ptr: int, shape, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    if dtype == torch.bfloat16:
        # __cuda_array_interface__ has no bf16 typestr: view as int16 and reinterpret
        t = torch.as_tensor(_Holder(ptr, 0, "<i2", shape), device=device)
        return t.view(torch.bfloat16)
    if dtype == torch.int32:
        return torch.as_tensor(_Holder(ptr, 0, "<i4", shape), device=device)
    raise TypeError(dtype)
