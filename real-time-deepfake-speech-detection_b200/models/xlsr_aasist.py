"""Mirror of reference ``models/xlsr_aasist.py``: ``XLSR_AASIST`` and ``My_XLSR_AASIST``.

Same constructor signature ``(device, ssl_cpkt_path=None, **kwargs)``, same parameters/buffers and
state-dict keys as the reference (xlsr_aasist.py:5-84, 180-245), so reference checkpoints load with
``load_state_dict``.  ``forward`` (eval mode, CUDA) runs the whole path -- XLS-R front-end, AASIST
back-end, read-out -- through librtdf.so; there is no PyTorch/CPU implementation behind it.
"""
import torch
import torch.nn as nn

from ._hooks import forward_with_hooks
from ._rt import engine_for
from .aasist_modules import *  # noqa: F401,F403
from .fe import *  # noqa: F401,F403
from .fe import My_XLSR_FE, XLSR_FE

__all__ = ["XLSR_AASIST", "My_XLSR_AASIST", "XLSR_FE", "My_XLSR_FE", "middle_indices",
           "GraphAttentionLayer", "HtrgGraphAttentionLayer", "GraphPool", "Residual_block"]


class _AasistBase(nn.Module):
    def _build_backend(self):
        filts = [128, [1, 32], [32, 32], [32, 64], [64, 64]]
        gat_dims = [64, 32]
        pool_ratios = [0.5, 0.5, 0.5, 0.5]
        temperatures = [2.0, 2.0, 100.0, 100.0]
        self.LL = nn.Linear(self.ssl_model.out_dim, 128)
        self.first_bn = nn.BatchNorm2d(num_features=1)
        self.first_bn1 = nn.BatchNorm2d(num_features=64)
        self.drop = nn.Dropout(0.5, inplace=True)
        self.drop_way = nn.Dropout(0.2, inplace=True)
        self.selu = nn.SELU(inplace=True)
        self.encoder = nn.Sequential(
            nn.Sequential(Residual_block(nb_filts=filts[1], first=True)),
            nn.Sequential(Residual_block(nb_filts=filts[2])),
            nn.Sequential(Residual_block(nb_filts=filts[3])),
            nn.Sequential(Residual_block(nb_filts=filts[4])),
            nn.Sequential(Residual_block(nb_filts=filts[4])),
            nn.Sequential(Residual_block(nb_filts=filts[4])))
        self.attention = nn.Sequential(
            nn.Conv2d(64, 128, kernel_size=(1, 1)), nn.SELU(inplace=True), nn.BatchNorm2d(128),
            nn.Conv2d(128, 64, kernel_size=(1, 1)))
        self.pos_S = nn.Parameter(torch.randn(1, 42, filts[-1][-1]))
        self.master1 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))
        self.master2 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))
        self.GAT_layer_S = GraphAttentionLayer(filts[-1][-1], gat_dims[0], temperature=temperatures[0])
        self.GAT_layer_T = GraphAttentionLayer(filts[-1][-1], gat_dims[0], temperature=temperatures[1])
        self.HtrgGAT_layer_ST11 = HtrgGraphAttentionLayer(gat_dims[0], gat_dims[1], temperature=temperatures[2])
        self.HtrgGAT_layer_ST12 = HtrgGraphAttentionLayer(gat_dims[1], gat_dims[1], temperature=temperatures[2])
        self.HtrgGAT_layer_ST21 = HtrgGraphAttentionLayer(gat_dims[0], gat_dims[1], temperature=temperatures[2])
        self.HtrgGAT_layer_ST22 = HtrgGraphAttentionLayer(gat_dims[1], gat_dims[1], temperature=temperatures[2])
        self.pool_S = GraphPool(pool_ratios[0], gat_dims[0], 0.3)
        self.pool_T = GraphPool(pool_ratios[1], gat_dims[0], 0.3)
        self.pool_hS1 = GraphPool(pool_ratios[2], gat_dims[1], 0.3)
        self.pool_hT1 = GraphPool(pool_ratios[2], gat_dims[1], 0.3)
        self.pool_hS2 = GraphPool(pool_ratios[2], gat_dims[1], 0.3)
        self.pool_hT2 = GraphPool(pool_ratios[2], gat_dims[1], 0.3)
        self.out_layer = nn.Linear(5 * gat_dims[1], 2)

    def _apply_freeze_kwargs(self, kwargs):
        partial_freeze_layers = kwargs.get('partial_freeze_layers', None)
        partial_freeze_init_layers = kwargs.get('partial_freeze_init_layers', [])
        if partial_freeze_layers:
            self.ssl_model.partial_freeze_layers(partial_freeze_layers.get('target_layers', []),
                                                 partial_freeze_layers.get('non_target_layers', []))
        if len(partial_freeze_init_layers) > 0:
            self.ssl_model.random_init_layers(partial_freeze_init_layers)

    def engine(self):
        return engine_for(self, "aasist", len(self.ssl_model.model.encoder.layers))

    def forward(self, x, return_taps=False):
        """x: (B,N) or (B,N,1) fp32 CUDA waveforms -> (B,2) logits  (reference forward, :86-177)."""
        x = x.squeeze(-1) if x.dim() == 3 else x
        if return_taps:
            return self.engine().forward(x, want_taps=True)
        return forward_with_hooks(self, self.engine(), x)


class XLSR_AASIST(_AasistBase):
    """reference models/xlsr_aasist.py:5-177."""

    def __init__(self, device, ssl_cpkt_path=None, **kwargs) -> None:
        super().__init__()
        self.ssl_model = XLSR_FE(device, kwargs.get('extractor_mode'), kwargs.get('conv_bias'))
        self._apply_freeze_kwargs(kwargs)
        self._build_backend()


class My_XLSR_AASIST(_AasistBase):
    """reference models/xlsr_aasist.py:180-339 (front-end truncated to ``num_layers``, :183)."""

    def __init__(self, device, ssl_cpkt_path=None, **kwargs) -> None:
        super().__init__()
        self.ssl_model = My_XLSR_FE(device, **kwargs)
        self._apply_freeze_kwargs(kwargs)
        self._build_backend()
