"""Mirror of reference ``models/conformer_baseline.py``: ``MyConformer``, ``Model``, ``MyModel``.

State-dict keys follow the reference (which nests lucidrains' ``conformer.ConformerBlock``:
``conformer.encoder_blocks.{b}.ff1.fn.norm`` ... SURVEY.md App. A.5); ``forward`` runs the fused
CUDA path.  ``MyModel.forward`` reproduces the reference's TypeError (conformer_baseline.py:98 passes
an extra positional argument) unless the instance was built with ``fixed_call=True``.
"""
import copy

import torch
import torch.nn as nn

from ._hooks import forward_with_hooks
from ._rt import engine_for
from .fe import *  # noqa: F401,F403
from .fe import My_XLSR_FE, XLSR_FE


def _container_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} holds parameters only; its arithmetic runs inside librtdf.so")


class _C(nn.Module):
    forward = _container_forward


class _FF(_C):          # lucidrains FeedForward: net = [Linear, Swish, Dropout, Linear, Dropout]
    def __init__(self, dim, mult):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult), nn.Identity(), nn.Identity(),
                                 nn.Linear(dim * mult, dim), nn.Identity())


class _PreNorm(_C):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)


class _Scale(_C):
    def __init__(self, scale, fn):
        super().__init__()
        self.fn = fn
        self.scale = scale


class _Attention(_C):
    def __init__(self, dim, heads, dim_head, max_pos_emb=512):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale, self.max_pos_emb = heads, dim_head ** -0.5, max_pos_emb
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim)
        self.rel_pos_emb = nn.Embedding(2 * max_pos_emb + 1, dim_head)


class _DepthWise(_C):
    def __init__(self, chan, k):
        super().__init__()
        self.conv = nn.Conv1d(chan, chan, k, groups=chan)


class _ConvModule(_C):
    def __init__(self, dim, expansion_factor, kernel_size):
        super().__init__()
        inner = dim * expansion_factor
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Identity(), nn.Conv1d(dim, inner * 2, 1), nn.Identity(),
                                 _DepthWise(inner, kernel_size), nn.BatchNorm1d(inner), nn.Identity(),
                                 nn.Conv1d(inner, dim, 1), nn.Identity(), nn.Identity())


class ConformerBlock(_C):
    """Parameter layout of lucidrains ``conformer.ConformerBlock`` (third-party)."""

    def __init__(self, *, dim, dim_head=64, heads=8, ff_mult=4, conv_expansion_factor=2, conv_kernel_size=31,
                 attn_dropout=0., ff_dropout=0., conv_dropout=0., conv_causal=False):
        super().__init__()
        if conv_causal:
            raise NotImplementedError("causal Conformer convolution is not on the reference path")
        if conv_expansion_factor != 2 or ff_mult != 4:
            raise NotImplementedError("only ff_mult=4, conv_expansion_factor=2 (reference defaults) are implemented")
        self.ff1 = _Scale(0.5, _PreNorm(dim, _FF(dim, ff_mult)))
        self.attn = _PreNorm(dim, _Attention(dim, heads, dim_head))
        self.conv = _ConvModule(dim, conv_expansion_factor, conv_kernel_size)
        self.ff2 = _Scale(0.5, _PreNorm(dim, _FF(dim, ff_mult)))
        self.post_norm = nn.LayerNorm(dim)


class MyConformer(_C):
    """reference conformer_baseline.py:8-29."""

    def __init__(self, emb_size=128, heads=4, ffmult=4, exp_fac=2, kernel_size=16, n_encoders=1):
        super().__init__()
        self.dim_head = int(emb_size / heads)
        self.dim = emb_size
        self.heads = heads
        self.kernel_size = kernel_size
        self.n_encoders = n_encoders
        block = ConformerBlock(dim=emb_size, dim_head=self.dim_head, heads=heads, ff_mult=ffmult,
                               conv_expansion_factor=exp_fac, conv_kernel_size=kernel_size)
        self.encoder_blocks = nn.ModuleList([copy.deepcopy(block) for _ in range(n_encoders)])  # _get_clones
        self.class_token = nn.Parameter(torch.rand(1, emb_size))
        self.fc5 = nn.Linear(emb_size, 2)


class _ConformerBase(nn.Module):
    def _build(self, kwargs):
        self._cfg = dict(emb_size=kwargs.get('emb_size', 144), heads=kwargs.get('heads', 4),
                         kernel_size=kwargs.get('kernel_size', 31), n_encoders=kwargs.get('n_encoders', 4))
        self.LL = nn.Linear(1024, self._cfg['emb_size'])
        self.first_bn = nn.BatchNorm2d(num_features=1)
        self.selu = nn.SELU(inplace=True)
        self.conformer = MyConformer(emb_size=self._cfg['emb_size'], n_encoders=self._cfg['n_encoders'],
                                     heads=self._cfg['heads'], kernel_size=self._cfg['kernel_size'])

    def engine(self):
        return engine_for(self, "conformer", len(self.ssl_model.model.encoder.layers), conformer=self._cfg)

    def _score(self, x):
        x = x.squeeze(-1) if x.dim() == 3 else x
        return forward_with_hooks(self, self.engine(), x)


class Model(_ConformerBase):
    """reference conformer_baseline.py:31-64 (imported as ``ConformerModel`` by main.py:21)."""

    def __init__(self, device, ssl_cpkt_path, **kwargs):
        super().__init__()
        self.device = device
        self.ssl_model = XLSR_FE(device)
        self._build(kwargs)

    def forward(self, x):
        return self._score(x)


class MyModel(_ConformerBase):
    """reference conformer_baseline.py:66-99."""

    def __init__(self, device, ssl_cpkt_path, fixed_call=False, **kwargs):
        super().__init__()
        self.device = device
        self.fixed_call = fixed_call
        self.ssl_model = My_XLSR_FE(device, **kwargs)
        self._build(kwargs)

    def forward(self, x):
        if not self.fixed_call:  # reference conformer_baseline.py:98: self.conformer(x, self.device)
            raise TypeError("MyConformer.forward() takes 2 positional arguments but 3 were given")
        return self._score(x)
