"""Mirror of reference ``models/aasist_modules.py``: the four AASIST building blocks as
state-holding modules with the reference's class names, constructor signatures and parameter
names (``att_proj``, ``att_weight``, ``proj_with_att``, ``proj_without_att``, ``bn`` ...).

Their arithmetic is implemented by the fused kernels in ``csrc/aasist.cu`` (one CTA per graph
row, never materialising the (n,n,D) pairwise tensor; rank-by-count top-k) and is executed as part
of the enclosing model's forward; ``GraphPool`` is additionally callable on its own through the
``rtdf_graph_pool`` entry point.  Calling the other blocks stand-alone raises.
"""
from typing import Union

import torch
from torch import nn

__all__ = ["GraphAttentionLayer", "HtrgGraphAttentionLayer", "GraphPool", "Residual_block"]


def _fused_only(self, *args, **kwargs):
    raise RuntimeError(f"{type(self).__name__}.forward runs only inside the fused CUDA scoring path "
                       "(librtdf.so): call the enclosing XLSR_AASIST model")


def _new_param(*size):
    out = nn.Parameter(torch.FloatTensor(*size))
    nn.init.xavier_normal_(out)
    return out


class GraphAttentionLayer(nn.Module):
    """reference aasist_modules.py:17-110 (kernel: gat_rows_kernel)."""

    def __init__(self, in_dim, out_dim, **kwargs):
        super().__init__()
        self.att_proj = nn.Linear(in_dim, out_dim)
        self.att_weight = _new_param(out_dim, 1)
        self.proj_with_att = nn.Linear(in_dim, out_dim)
        self.proj_without_att = nn.Linear(in_dim, out_dim)
        self.bn = nn.BatchNorm1d(out_dim)
        self.input_drop = nn.Dropout(p=0.2)
        self.act = nn.SELU(inplace=True)
        self.temp = kwargs.get("temperature", 1.)

    forward = _fused_only


class HtrgGraphAttentionLayer(nn.Module):
    """reference aasist_modules.py:112-294 (kernels: type_proj_kernel + gat_rows_kernel with master row)."""

    def __init__(self, in_dim, out_dim, **kwargs):
        super().__init__()
        self.proj_type1 = nn.Linear(in_dim, in_dim)
        self.proj_type2 = nn.Linear(in_dim, in_dim)
        self.att_proj = nn.Linear(in_dim, out_dim)
        self.att_projM = nn.Linear(in_dim, out_dim)
        self.att_weight11 = _new_param(out_dim, 1)
        self.att_weight22 = _new_param(out_dim, 1)
        self.att_weight12 = _new_param(out_dim, 1)
        self.att_weightM = _new_param(out_dim, 1)
        self.proj_with_att = nn.Linear(in_dim, out_dim)
        self.proj_without_att = nn.Linear(in_dim, out_dim)
        self.proj_with_attM = nn.Linear(in_dim, out_dim)
        self.proj_without_attM = nn.Linear(in_dim, out_dim)
        self.bn = nn.BatchNorm1d(out_dim)
        self.input_drop = nn.Dropout(p=0.2)
        self.act = nn.SELU(inplace=True)
        self.temp = kwargs.get("temperature", 1.)

    forward = _fused_only


class GraphPool(nn.Module):
    """reference aasist_modules.py:296-338 (kernel: graph_pool_kernel)."""

    def __init__(self, k: float, in_dim: int, p: Union[float, int]):
        super().__init__()
        self.k = torch.tensor(k)
        self.sigmoid = nn.Sigmoid()
        self.proj = nn.Linear(in_dim, 1)
        self.drop = nn.Dropout(p=p) if p > 0 else nn.Identity()
        self.in_dim = in_dim

    def forward(self, h, return_idx=False):
        """h (B,n,D) CUDA fp32 -> (B,max(floor(n*k),1),D), nodes in descending score order."""
        from ._rt import native
        if self.training:
            raise RuntimeError("GraphPool: eval mode only (dropout is not implemented on the CUDA path)")
        if not h.is_cuda:
            raise RuntimeError("GraphPool: CUDA tensors only (no CPU path)")
        h = h.to(torch.float32).contiguous()
        B, n, D = h.shape
        keep = max(int((torch.as_tensor(n) * self.k).long()), 1)          # aasist_modules.py:329-330
        out = torch.empty(B, keep, D, dtype=torch.float32, device=h.device)
        idx = torch.empty(B, keep, dtype=torch.int32, device=h.device)
        lib = native.load()
        with torch.cuda.device(h.device):
            w = self.proj.weight.detach().to(torch.float32).contiguous()
            b = self.proj.bias.detach().to(torch.float32).contiguous()
            native.check(lib.rtdf_graph_pool(native.ptr(h), B, n, D, native.ptr(w), native.ptr(b), keep,
                                             native.ptr(out), native.ptr(idx),
                                             torch.cuda.current_stream(h.device).cuda_stream), "rtdf_graph_pool")
        return (out, idx) if return_idx else out


class Residual_block(nn.Module):
    """reference aasist_modules.py:340-397 (kernel: conv2d_kernel; bn1 exists but does not act, :376-383)."""

    def __init__(self, nb_filts, first=False):
        super().__init__()
        self.first = first
        self.bn1 = None
        self.conv_downsample = None
        if not self.first:
            self.bn1 = nn.BatchNorm2d(num_features=nb_filts[0])
        self.conv1 = nn.Conv2d(in_channels=nb_filts[0], out_channels=nb_filts[1], kernel_size=(2, 3),
                               padding=(1, 1), stride=1)
        self.selu = nn.SELU(inplace=True)
        self.bn2 = nn.BatchNorm2d(num_features=nb_filts[1])
        self.conv2 = nn.Conv2d(in_channels=nb_filts[1], out_channels=nb_filts[1], kernel_size=(2, 3),
                               padding=(0, 1), stride=1)
        if nb_filts[0] != nb_filts[1]:
            self.downsample = True
            self.conv_downsample = nn.Conv2d(in_channels=nb_filts[0], out_channels=nb_filts[1], padding=(0, 1),
                                             kernel_size=(1, 3), stride=1)
        else:
            self.downsample = False

    forward = _fused_only
