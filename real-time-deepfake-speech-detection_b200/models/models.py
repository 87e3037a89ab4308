"""Mirror of reference ``models/models.py``: ``SSLModel`` (models.py:13-46) plus the same four AASIST building
blocks the reference repeats there (models.py:49-432, identical to ``models/aasist_modules.py``).

``SSLModel(device, cp_path, out_dim)`` wraps the XLS-R front-end like ``XLSR_FE`` but takes the checkpoint path as
an argument (the reference hands it to fairseq, models.py:16-18); ``extract_feat`` runs through librtdf.so.
"""
import logging
import os

from torch import nn

from ._rt import engine_for
from .aasist_modules import GraphAttentionLayer, GraphPool, HtrgGraphAttentionLayer, Residual_block  # noqa: F401
from .wav2vec2_params import Wav2Vec2Model, load_pretrained

__all__ = ["SSLModel", "GraphAttentionLayer", "HtrgGraphAttentionLayer", "GraphPool", "Residual_block"]


class SSLModel(nn.Module):
    def __init__(self, device, cp_path, out_dim):
        super().__init__()
        self.model = Wav2Vec2Model()
        if cp_path and os.path.exists(cp_path):
            load_pretrained(self.model, cp_path)
        self.model = self.model.to(device)
        self.out_dim = out_dim
        self.freeze = False

    def extract_feat(self, input_data):
        input_tmp = input_data[:, :, 0] if input_data.ndim == 3 else input_data      # models.py:36
        eng = engine_for(self, None, len(self.model.encoder.layers), key_prefix="ssl_model.")
        return eng.frontend(input_tmp)

    def forward(self, input_data):
        return self.extract_feat(input_data)

    def frozen(self):                                                                 # models.py:42-46
        logging.info("Freezing the model")
        for param in self.model.parameters():
            param.requires_grad = False
        self.freeze = True

    def unfrozen(self):                                                               # models.py:48-52
        logging.info("Unfreezing the model")
        for param in self.model.parameters():
            param.requires_grad = True
        self.freeze = False
