"""Mirror of reference ``models/fe.py``: SSL front-end wrappers ``XLSR_FE`` / ``My_XLSR_FE``.

Same class names, constructor arguments, attributes (``model``, ``out_dim``), methods and error
behaviour as the reference; ``extract_feat`` runs the hand-written sm_100a kernels through
librtdf.so instead of fairseq.  The reference loads a hard-coded ``xlsr2_300m.pt`` through fairseq
(fe.py:11-15, 63-66); here the checkpoint path comes from ``$RTDF_XLSR_CKPT`` (fairseq checkpoint or
plain state dict) and the model is randomly initialised when it is unset.
"""
import os

import torch
from torch import nn

from ._rt import engine_for
from .aasist_modules import *  # noqa: F401,F403  (reference fe.py:1)
from .wav2vec2_params import Wav2Vec2Model, extractor_config, read_checkpoint

__all__ = ["XLSR_FE", "My_XLSR_FE", "middle_indices"] + [
    "GraphAttentionLayer", "HtrgGraphAttentionLayer", "GraphPool", "Residual_block"]


def _build_ssl(device, extractor_mode=None, conv_bias=None):
    """The SSL model the reference gets from fairseq (fe.py:11-15).  With $RTDF_XLSR_CKPT the conv feature encoder's
    mode (per-frame LayerNorm + conv bias for XLS-R, GroupNorm after conv-0 for wav2vec2-base style checkpoints) follows
    the checkpoint; without it, ``extractor_mode`` / $RTDF_XLSR_EXTRACTOR_MODE (default "layer_norm") picks the layout of
    the randomly initialised model."""
    path = os.environ.get("RTDF_XLSR_CKPT")
    sd = read_checkpoint(path) if path else None
    if sd is not None:
        extractor_mode, conv_bias = extractor_config(sd)
    extractor_mode = extractor_mode or os.environ.get("RTDF_XLSR_EXTRACTOR_MODE", "layer_norm")
    if conv_bias is None:
        conv_bias = extractor_mode == "layer_norm"
    model = Wav2Vec2Model(extractor_mode=extractor_mode, conv_bias=conv_bias)
    if sd is not None:
        missing, _ = model.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"checkpoint {path} lacks XLS-R keys, e.g. {missing[:3]}")
    return model.to(device)


class XLSR_FE(nn.Module):
    """reference models/fe.py:8-40."""

    def __init__(self, device, extractor_mode=None, conv_bias=None):
        super().__init__()
        self.model = _build_ssl(device, extractor_mode, conv_bias)
        self.out_dim = 1024

    def extract_feat(self, input_data):
        input_tmp = input_data[:, :, 0] if input_data.ndim == 3 else input_data      # fe.py:18
        eng = engine_for(self, None, len(self.model.encoder.layers), key_prefix="ssl_model.")
        return eng.frontend(input_tmp)                                                # fe.py:19-21 ['x']

    def forward(self, input_data):
        return self.extract_feat(input_data)

    def partial_freeze_layers(self, target_layers: list, non_target_layers: list):   # fe.py:26-34
        for name, param in self.model.named_parameters():
            if any([layer in name for layer in target_layers]) and not any([layer in name for layer in non_target_layers]):
                param.requires_grad = False
        self.random_init_layers(non_target_layers)

    def random_init_layers(self, target_layers: list):                                # fe.py:36-40
        for name, param in self.model.named_parameters():
            if any([layer in name for layer in target_layers]) and param.dim() >= 2:
                torch.nn.init.xavier_uniform_(param)


def middle_indices(array_length, number_of_middle_elements):                          # fe.py:43-50
    start_index = (array_length - number_of_middle_elements) // 2
    return list(range(start_index, start_index + number_of_middle_elements))


class My_XLSR_FE(XLSR_FE):
    """reference models/fe.py:53-99: keeps a subset of the 24 transformer layers."""

    def __init__(self, device, **kwargs):
        num_layers = kwargs.get('num_layers', 24)
        order = kwargs.get('order', 'first')
        custom_order = kwargs.get('custom_order', None)
        if num_layers < 1 or num_layers > 24:
            raise ValueError("Number of layers must be at least 1 and at most 24.")
        super().__init__(device, kwargs.get('extractor_mode'), kwargs.get('conv_bias'))
        self.num_layers, self.order, self.custom_order = num_layers, order, custom_order
        layers = self.model.encoder.layers
        if order == 'last':
            self.model.encoder.layers = layers[-num_layers:]
        elif order == 'first':
            self.model.encoder.layers = layers[:num_layers]
        elif order == 'middle':
            self.model.encoder.layers = nn.ModuleList([layers[i] for i in middle_indices(24, num_layers)])
        else:
            if custom_order is None:
                raise ValueError("Custom order must be provided as a list of integers (0-23).")
            if type(custom_order) != list:
                raise ValueError("Custom order must be a list of integers.")
            self.model.encoder.layers = nn.ModuleList([layers[i] for i in custom_order])
