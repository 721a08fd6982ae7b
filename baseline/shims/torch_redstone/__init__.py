"""Shim: the two torch_redstone symbols the reference uses (utils/raycaster.py:37, utils/shader_ops.py:389, 578-596)."""
import torch


def supercat(tensors, dim=0):
    """Concatenate after broadcasting every other dimension."""
    tensors = [t if isinstance(t, torch.Tensor) else torch.as_tensor(t) for t in tensors]
    nd = max(t.ndim for t in tensors)
    tensors = [t.reshape((1,) * (nd - t.ndim) + tuple(t.shape)) for t in tensors]
    d = dim % nd
    shape = [max(t.shape[k] for t in tensors) for k in range(nd)]
    out = []
    for t in tensors:
        s = list(shape)
        s[d] = t.shape[d]
        out.append(t.expand(*s))
    return torch.cat(out, dim=d)


def torch_to_numpy(t):
    return t.detach().cpu().numpy()
