"""Oracle restatement of the reference's model classes and pre-emphasis (eval, fp32, CPU).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Follows ``/root/reference/models/fe.py``, ``models/xlsr_aasist.py``,
``models/conformer_baseline.py`` and ``data/preprocess.py``.  State-dict keys
equal the reference's.
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from .aasist_ref import AasistBackend
from .conformer_block_ref import ConformerBlock
from .wav2vec2_ref import FairseqLikeWav2Vec2


def middle_indices(array_length, number_of_middle_elements):
    """fe.py:43-50."""
    start = (array_length - number_of_middle_elements) // 2
    return list(range(start, start + number_of_middle_elements))


class XLSR_FE(nn.Module):
    """fe.py:8-24 -- wraps the fairseq model; extract_feat returns ['x']."""

    def __init__(self, device="cpu", extractor_mode=None, conv_bias=None, **_):
        super().__init__()
        # what fairseq would build from the checkpoint's config: XLS-R is extractor_mode="layer_norm", conv_bias=True;
        # "default" is the wav2vec2-base style group-norm feature encoder
        mode = extractor_mode or "layer_norm"
        self.model = FairseqLikeWav2Vec2(extractor_mode=mode,
                                         conv_bias=(mode == "layer_norm") if conv_bias is None else conv_bias).to(device)
        self.out_dim = 1024

    def extract_feat(self, x):
        x = x[:, :, 0] if x.ndim == 3 else x                                # fe.py:18
        return self.model(x, mask=False, features_only=True)["x"]           # fe.py:19-21

    forward = extract_feat


class My_XLSR_FE(XLSR_FE):
    """fe.py:53-99 -- keeps a first/last/middle/custom subset of the 24 layers."""

    def __init__(self, device="cpu", **kwargs):
        num_layers = kwargs.get("num_layers", 24)
        order = kwargs.get("order", "first")
        custom_order = kwargs.get("custom_order", None)
        if num_layers < 1 or num_layers > 24:                               # fe.py:60-62
            raise ValueError("Number of layers must be at least 1 and at most 24.")
        super().__init__(device, kwargs.get("extractor_mode"), kwargs.get("conv_bias"))
        layers = self.model.encoder.layers
        if order == "last":                                                 # fe.py:69-71
            self.model.encoder.layers = layers[-num_layers:]
        elif order == "first":                                              # fe.py:72-74
            self.model.encoder.layers = layers[:num_layers]
        elif order == "middle":                                             # fe.py:75-79
            self.model.encoder.layers = nn.ModuleList([layers[i] for i in middle_indices(24, num_layers)])
        else:                                                               # fe.py:80-90
            if custom_order is None:
                raise ValueError("Custom order must be provided as a list of integers (0-23).")
            if type(custom_order) != list:
                raise ValueError("Custom order must be a list of integers.")
            self.model.encoder.layers = nn.ModuleList([layers[i] for i in custom_order])


class XLSR_AASIST(AasistBackend):
    """xlsr_aasist.py:5-177."""

    fe_cls = XLSR_FE

    def __init__(self, device="cpu", ssl_cpkt_path=None, **kwargs):
        super().__init__()
        self.ssl_model = self.fe_cls(device, **kwargs)

    def forward(self, x, taps=None):
        feats = self.ssl_model.extract_feat(x.squeeze(-1))                  # xlsr_aasist.py:88
        if taps is not None:
            taps["feats"] = feats
        return self.backend(feats, taps)


class My_XLSR_AASIST(XLSR_AASIST):
    """xlsr_aasist.py:180-339 (identical except for the truncated front-end, :183)."""

    fe_cls = My_XLSR_FE


class MyConformer(nn.Module):
    """conformer_baseline.py:8-29."""

    def __init__(self, emb_size=128, heads=4, ffmult=4, exp_fac=2, kernel_size=16, n_encoders=1):
        super().__init__()
        block = ConformerBlock(dim=emb_size, dim_head=int(emb_size / heads), heads=heads, ff_mult=ffmult,
                               conv_expansion_factor=exp_fac, conv_kernel_size=kernel_size)
        self.encoder_blocks = nn.ModuleList([copy.deepcopy(block) for _ in range(n_encoders)])  # _get_clones
        self.class_token = nn.Parameter(torch.rand(1, emb_size))
        self.fc5 = nn.Linear(emb_size, 2)

    def forward(self, x):
        tok = self.class_token.unsqueeze(0).expand(x.shape[0], -1, -1)
        x = torch.cat([tok, x], dim=1)                                      # :23-24
        for blk in self.encoder_blocks:
            x = blk(x)                                                      # :25-26
        emb = x[:, 0, :]
        return self.fc5(emb), emb                                           # :27-29


class ConformerModel(nn.Module):
    """conformer_baseline.py:31-64 (``Model``; main.py imports it as ``ConformerModel``)."""

    fe_cls = XLSR_FE

    def __init__(self, device="cpu", ssl_cpkt_path=None, **kwargs):
        super().__init__()
        emb = kwargs.get("emb_size", 144)
        self.ssl_model = self.fe_cls(device, **kwargs)
        self.LL = nn.Linear(1024, emb)
        self.first_bn = nn.BatchNorm2d(1)
        self.conformer = MyConformer(emb_size=emb, n_encoders=kwargs.get("n_encoders", 4),
                                     heads=kwargs.get("heads", 4), kernel_size=kwargs.get("kernel_size", 31))

    def forward(self, x, taps=None):
        feats = self.ssl_model.extract_feat(x.squeeze(-1))                  # :56
        if taps is not None:
            taps["feats"] = feats
        x = self.LL(feats).unsqueeze(1)                                     # :58-59
        x = F.selu(self.first_bn(x)).squeeze(1)                             # :60-62
        return self.conformer(x)[0]                                         # :63-64


Model = ConformerModel


class MyModel(ConformerModel):
    """conformer_baseline.py:66-99.  As shipped, forward passes an extra argument to
    ``MyConformer.forward`` (:98) and raises TypeError; ``fixed_call=True`` gives the
    intended one-argument call."""

    fe_cls = My_XLSR_FE

    def __init__(self, device="cpu", ssl_cpkt_path=None, fixed_call=False, **kwargs):
        super().__init__(device, ssl_cpkt_path, **kwargs)
        self.fixed_call = fixed_call

    def forward(self, x, taps=None):
        if not self.fixed_call:
            raise TypeError("MyConformer.forward() takes 2 positional arguments but 3 were given")
        return super().forward(x, taps)


def pre_emphasis(x, coef=0.97):
    """data/preprocess.py:16-29: y[t] = x[t] - coef * x[t-1] with reflect padding on the left
    (x[-1] := x[1]); the trailing ``squeeze()`` (:27) drops a batch dim of 1."""
    xp = F.pad(x.unsqueeze(1), (1, 0), mode="reflect")
    w = torch.tensor([[[-coef, 1.0]]], dtype=x.dtype, device=x.device)
    return F.conv1d(xp, w).squeeze()


def synth_waveforms(batch, n, seed=2021):
    """Synthetic 16 kHz utterances: 0.1*randn + 0.05*sin(2*pi*220 t), clamped."""
    out = torch.empty(batch, n)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + i)
        out[i] = (0.1 * torch.randn(n, generator=g) + 0.05 * torch.sin(2 * torch.pi * 220.0 * t)).clamp(-1, 1)
    return out


def build(kind, seed=1024, perturb=True, **kwargs):
    """Seeded random-init model of the given kind with perturbed norm statistics."""
    from .aasist_ref import perturb_norm_stats
    torch.manual_seed(seed)
    cls = {"XLSR_AASIST": XLSR_AASIST, "My_XLSR_AASIST": My_XLSR_AASIST,
           "ConformerModel": ConformerModel, "MyModel": MyModel}[kind]
    m = cls("cpu", None, **kwargs).eval()
    if perturb:
        perturb_norm_stats(m, seed=seed + 1)
    return m
