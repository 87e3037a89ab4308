"""Mirror of reference ``data/preprocess.py``: ``PreEmphasis`` on the CUDA path.

``forward`` keeps the reference's behaviour: identity when ``exp_config.is_pre_emphasis`` is
false (preprocess.py:19-20) and a trailing ``squeeze()`` that drops a batch dimension of 1
(preprocess.py:27).  The filter itself is the ``preemph_kernel`` of librtdf.so.
"""
import torch
import torch.nn as nn

try:
    from ..rtdf_runtime import native  # type: ignore
except (ImportError, ValueError):
    from rtdf_runtime import native  # type: ignore


class PreEmphasis(nn.Module):
    def __init__(self, device, sys_config, exp_config):
        super().__init__()
        self.exp_config = exp_config
        self.pre_emphasis_filter = torch.FloatTensor([[[-exp_config.pre_emphasis, 1.]]]).to(device)

    def forward(self, x):
        if not self.exp_config.is_pre_emphasis:
            return x
        if not x.is_cuda:
            raise RuntimeError("PreEmphasis: CUDA tensors only (no CPU path)")
        x = x.to(torch.float32).contiguous()
        if x.dim() != 2:
            raise ValueError("PreEmphasis expects (batch, samples)")
        y = torch.empty_like(x)
        lib = native.load()
        with torch.cuda.device(x.device):
            native.check(lib.rtdf_preemph(native.ptr(x), native.ptr(y), x.shape[0], x.shape[1],
                                          float(self.exp_config.pre_emphasis),
                                          torch.cuda.current_stream(x.device).cuda_stream), "rtdf_preemph")
        return y.squeeze()
