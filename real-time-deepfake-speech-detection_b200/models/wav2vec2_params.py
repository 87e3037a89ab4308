"""Parameter container with fairseq ``Wav2Vec2Model`` names (XLS-R 300M configuration).

This module holds *state only* (SURVEY.md App. A.5 key contract: ``feature_extractor.conv_layers.
{i}.0.*``, ``...{i}.2.1.*``, ``layer_norm``, ``post_extract_proj``, ``encoder.pos_conv.0.{bias,
weight_g,weight_v}``, ``encoder.layers.{l}.*``, ``encoder.layer_norm``, ``mask_emb``).  The
arithmetic runs in librtdf.so; calling the container raises.  ``encoder.layers`` is a real,
re-assignable ``nn.ModuleList`` because the reference truncates the model through it
(models/fe.py:69-90) and copies teacher layers by index (main_kd.py:127-141).
"""
import math

import torch
import torch.nn as nn

CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2

_IGNORED_PREFIXES = ("quantizer.", "project_q.", "final_proj.", "target_glu.", "project_inp.")


def _no_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container: the arithmetic of this layer runs inside "
                       "the fused CUDA scoring path (librtdf.so); call the enclosing model instead")


class _Container(nn.Module):
    forward = _no_forward


class _PosConv(_Container):
    def __init__(self, dim, k, groups):
        super().__init__()
        v = torch.empty(dim, dim // groups, k)
        nn.init.normal_(v, mean=0.0, std=math.sqrt(4.0 / (k * dim)))
        self.weight_g = nn.Parameter(v.norm(dim=(0, 1), keepdim=True).clone())
        self.weight_v = nn.Parameter(v)
        self.bias = nn.Parameter(torch.zeros(dim))


class _SelfAttn(_Container):
    def __init__(self, dim):
        super().__init__()
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)


class TransformerSentenceEncoderLayer(_Container):
    def __init__(self, dim=1024, ffn=4096):
        super().__init__()
        self.self_attn = _SelfAttn(dim)
        self.self_attn_layer_norm = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, ffn)
        self.fc2 = nn.Linear(ffn, dim)
        self.final_layer_norm = nn.LayerNorm(dim)


class _FeatureExtractor(_Container):
    """fairseq ``ConvFeatureExtractionModel`` parameter layout.  mode "layer_norm" (XLS-R): conv bias and a LayerNorm(512)
    after every conv (keys ``conv_layers.{i}.2.1.*``); mode "default" (wav2vec2-base style): GroupNorm(512, 512) after
    conv-0 only (keys ``conv_layers.0.2.*``), ``conv_bias`` normally False."""

    def __init__(self, mode="layer_norm", conv_bias=True):
        super().__init__()
        if mode not in ("layer_norm", "default"):
            raise ValueError(f"extractor_mode must be 'layer_norm' or 'default', got {mode!r}")
        self.conv_layers = nn.ModuleList()
        in_d = 1
        for i, (dim, k, s) in enumerate(CONV_LAYERS):
            conv = nn.Conv1d(in_d, dim, k, stride=s, bias=conv_bias)
            nn.init.kaiming_normal_(conv.weight)
            # index 0: conv, 1: dropout, 2: norm, 3: GELU  (fairseq layout)
            if mode == "layer_norm":      # 2: (transpose, LayerNorm, transpose)
                block = nn.Sequential(conv, nn.Identity(), nn.Sequential(nn.Identity(), nn.LayerNorm(dim), nn.Identity()),
                                      nn.Identity())
            elif i == 0:                  # 2: GroupNorm(dim, dim)
                block = nn.Sequential(conv, nn.Identity(), nn.GroupNorm(dim, dim), nn.Identity())
            else:                         # conv, dropout, GELU
                block = nn.Sequential(conv, nn.Identity(), nn.Identity())
            self.conv_layers.append(block)
            in_d = dim


class _Encoder(_Container):
    def __init__(self, dim, ffn, layers):
        super().__init__()
        self.pos_conv = nn.Sequential(_PosConv(dim, 128, 16), nn.Identity(), nn.Identity())
        self.layers = nn.ModuleList([TransformerSentenceEncoderLayer(dim, ffn) for _ in range(layers)])
        self.layer_norm = nn.LayerNorm(dim)


class Wav2Vec2Model(_Container):
    def __init__(self, dim=1024, ffn=4096, layers=24, extractor_mode="layer_norm", conv_bias=True):
        super().__init__()
        self.feature_extractor = _FeatureExtractor(extractor_mode, conv_bias)
        self.layer_norm = nn.LayerNorm(512)
        self.post_extract_proj = nn.Linear(512, dim)
        self.mask_emb = nn.Parameter(torch.empty(dim).uniform_())
        self.encoder = _Encoder(dim, ffn, layers)
        for m in self.encoder.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=0.02)
                m.bias.data.zero_()
        # pre-training heads present in real checkpoints are not on the scoring path: drop them on load
        self._register_load_state_dict_pre_hook(self._drop_pretraining_heads)

    @staticmethod
    def _drop_pretraining_heads(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.startswith(prefix) and k[len(prefix):].startswith(_IGNORED_PREFIXES)]:
            del state_dict[k]


def read_checkpoint(path):
    """State dict of a fairseq checkpoint (``{'model': state_dict, ...}``) or of a plain state-dict file, without the
    pre-training heads."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt.get("model", ckpt) if isinstance(ckpt, dict) else ckpt
    return {k: v for k, v in sd.items() if not k.startswith(_IGNORED_PREFIXES)}


def extractor_config(sd):
    """(extractor_mode, conv_bias) a checkpoint was trained with, read off its keys."""
    mode = "layer_norm" if "feature_extractor.conv_layers.0.2.1.weight" in sd else "default"
    return mode, "feature_extractor.conv_layers.0.0.bias" in sd


def load_pretrained(model, path):
    """Load a fairseq XLS-R checkpoint (``{'model': state_dict, ...}``) or a plain state dict."""
    sd = read_checkpoint(path)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if missing:
        raise RuntimeError(f"checkpoint {path} lacks XLS-R keys, e.g. {missing[:3]}")
    return model
