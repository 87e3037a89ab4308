"""Oracle restatement of fairseq ``Wav2Vec2Model`` (XLS-R 300M configuration).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference reaches this arithmetic through
``fairseq.checkpoint_utils.load_model_ensemble_and_task`` (reference
``models/fe.py:11-15``) and calls
``self.model(x, mask=False, features_only=True)['x']`` (``models/fe.py:17-21``).
fairseq is an un-vendored, un-pinned dependency: **parity unpinned** by the
reference; the published algorithm (wav2vec 2.0, ``layer_norm_first=True``,
``extractor_mode="layer_norm"``, ``conv_bias=True``) is restated below with
fairseq's parameter names so that reference
state dicts load, and ``encoder.layers`` is the ``nn.ModuleList`` the forward
iterates, because ``My_XLSR_FE`` truncates the model by re-assigning it
(``models/fe.py:69-90``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2


class _TransposeLast(nn.Module):
    def forward(self, x):
        return x.transpose(-2, -1)


class _Fp32LayerNorm(nn.LayerNorm):
    def forward(self, x):
        out = F.layer_norm(x.float(), self.normalized_shape,
                           self.weight.float(), self.bias.float(), self.eps)
        return out.type_as(x)


class ConvFeatureExtractionModel(nn.Module):
    """7 x [Conv1d(no pad, bias) -> LayerNorm over channels -> GELU(erf)]."""

    def __init__(self, conv_layers=CONV_LAYERS, mode="layer_norm", conv_bias=True):
        super().__init__()
        assert mode in ("layer_norm", "default")
        self.mode = mode
        self.conv_layers = nn.ModuleList()
        in_d = 1
        for i, (dim, k, s) in enumerate(conv_layers):
            conv = nn.Conv1d(in_d, dim, k, stride=s, bias=conv_bias)
            nn.init.kaiming_normal_(conv.weight)
            if mode == "layer_norm":
                block = nn.Sequential(
                    conv, nn.Dropout(0.0),
                    nn.Sequential(_TransposeLast(), _Fp32LayerNorm(dim), _TransposeLast()),
                    nn.GELU())
            elif i == 0:  # wav2vec2-base style: GroupNorm(dim, dim) after conv-0 only
                block = nn.Sequential(conv, nn.Dropout(0.0), nn.GroupNorm(dim, dim), nn.GELU())
            else:
                block = nn.Sequential(conv, nn.Dropout(0.0), nn.GELU())
            self.conv_layers.append(block)
            in_d = dim

    def forward(self, x):
        x = x.unsqueeze(1)  # (B,1,N)
        for conv in self.conv_layers:
            x = conv(x)
        return x  # (B,512,T)


class _WeightNormConv1d(nn.Module):
    """Conv1d with old-style weight_norm(dim=2) parameters weight_g / weight_v."""

    def __init__(self, dim, k, groups):
        super().__init__()
        self.k, self.groups = k, groups
        v = torch.empty(dim, dim // groups, k)
        nn.init.normal_(v, mean=0.0, std=math.sqrt(4.0 / (k * dim)))
        self.weight_g = nn.Parameter(v.norm(dim=(0, 1), keepdim=True).clone())  # (1,1,k)
        self.weight_v = nn.Parameter(v)
        self.bias = nn.Parameter(torch.zeros(dim))

    def weight(self):
        v = self.weight_v
        return self.weight_g * v / v.norm(dim=(0, 1), keepdim=True)

    def forward(self, x):
        return F.conv1d(x, self.weight(), self.bias, padding=self.k // 2, groups=self.groups)


class _SamePad(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.remove = 1 if k % 2 == 0 else 0

    def forward(self, x):
        return x[:, :, :-self.remove] if self.remove else x


class _SelfAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads, self.head_dim = heads, dim // heads
        self.scaling = self.head_dim ** -0.5
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)

    def forward(self, x):  # (B,T,C)
        B, T, C = x.shape
        q = self.q_proj(x) * self.scaling
        k = self.k_proj(x)
        v = self.v_proj(x)
        q = q.view(B, T, self.heads, self.head_dim).transpose(1, 2)
        k = k.view(B, T, self.heads, self.head_dim).transpose(1, 2)
        v = v.view(B, T, self.heads, self.head_dim).transpose(1, 2)
        w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)).float(), dim=-1).type_as(q)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, T, C)
        return self.out_proj(a)


class TransformerSentenceEncoderLayer(nn.Module):
    """Pre-LN layer (``layer_norm_first=True``)."""

    def __init__(self, dim=1024, ffn=4096, heads=16):
        super().__init__()
        self.self_attn = _SelfAttention(dim, heads)
        self.self_attn_layer_norm = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, ffn)
        self.fc2 = nn.Linear(ffn, dim)
        self.final_layer_norm = nn.LayerNorm(dim)

    def forward(self, x):
        x = x + self.self_attn(self.self_attn_layer_norm(x))
        x = x + self.fc2(F.gelu(self.fc1(self.final_layer_norm(x))))
        return x


class TransformerEncoder(nn.Module):
    def __init__(self, dim=1024, ffn=4096, heads=16, layers=24, pos_k=128, pos_groups=16):
        super().__init__()
        self.pos_conv = nn.Sequential(_WeightNormConv1d(dim, pos_k, pos_groups), _SamePad(pos_k), nn.GELU())
        self.layers = nn.ModuleList([TransformerSentenceEncoderLayer(dim, ffn, heads) for _ in range(layers)])
        self.layer_norm = nn.LayerNorm(dim)

    def forward(self, x):  # (B,T,C)
        x = x + self.pos_conv(x.transpose(1, 2)).transpose(1, 2)
        for layer in self.layers:  # ModuleList is re-assignable (fe.py:69-90)
            x = layer(x)
        return self.layer_norm(x)


def _init_bert_params(module):
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()


class FairseqLikeWav2Vec2(nn.Module):
    """``model(source, mask=False, features_only=True) -> {'x': (B,T,1024)}``."""

    def __init__(self, dim=1024, ffn=4096, heads=16, layers=24, extractor_mode="layer_norm", conv_bias=True):
        super().__init__()
        self.feature_extractor = ConvFeatureExtractionModel(mode=extractor_mode, conv_bias=conv_bias)
        self.layer_norm = nn.LayerNorm(512)
        self.post_extract_proj = nn.Linear(512, dim)
        self.mask_emb = nn.Parameter(torch.empty(dim).uniform_())  # unused on this path
        self.encoder = TransformerEncoder(dim, ffn, heads, layers)
        self.encoder.apply(_init_bert_params)

    def forward(self, source, mask=False, features_only=True, **_):
        assert not mask and features_only
        f = self.feature_extractor(source).transpose(1, 2)  # (B,T,512)
        f = self.layer_norm(f)
        x = self.post_extract_proj(f)
        x = self.encoder(x)
        return {"x": x, "padding_mask": None, "features": f}


def conv_out_lengths(n):
    """L_i after each conv layer for an n-sample input."""
    out = []
    for _, k, s in CONV_LAYERS:
        n = (n - k) // s + 1
        out.append(n)
    return out
