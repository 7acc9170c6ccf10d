"""Helpers for the -m gpu tests: numpy uint64 <-> CUDA tensors (int64 storage; torch is only the allocator)."""
import numpy as np
import torch


def to_dev(a: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).cuda()
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).cuda()
    return torch.from_numpy(a).cuda()


def to_host(t: torch.Tensor) -> np.ndarray:
    torch.cuda.synchronize()
    a = t.cpu().numpy()
    if a.dtype == np.int64:
        return a.view(np.uint64)
    if a.dtype == np.int32:
        return a.view(np.uint32)
    return a
