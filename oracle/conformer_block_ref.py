"""Oracle restatement of lucidrains ``conformer.ConformerBlock`` (eval mode).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference imports it at ``models/conformer_baseline.py:3`` and instantiates
it at ``models/conformer_baseline.py:16-18``.  The PyPI package ``conformer``
is un-vendored and un-pinned (**parity unpinned**); its published algorithm is
restated here with the package's module nesting so that
state-dict keys match (App. A.5): ``ff1.fn.norm``, ``ff1.fn.fn.net.{0,3}``,
``attn.norm``, ``attn.fn.{to_q,to_kv,to_out,rel_pos_emb}``,
``conv.net.{0,2,4.conv,5,7}``, ``post_norm``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Swish(nn.Module):
    def forward(self, x):
        return x * x.sigmoid()


class GLU(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        out, gate = x.chunk(2, dim=self.dim)
        return out * gate.sigmoid()


class DepthWiseConv1d(nn.Module):
    def __init__(self, chan_in, chan_out, kernel_size, padding):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv1d(chan_in, chan_out, kernel_size, groups=chan_in)

    def forward(self, x):
        return self.conv(F.pad(x, self.padding))


class Scale(nn.Module):
    def __init__(self, scale, fn):
        super().__init__()
        self.fn, self.scale = fn, scale

    def forward(self, x, **kw):
        return self.fn(x, **kw) * self.scale


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)

    def forward(self, x, **kw):
        return self.fn(self.norm(x), **kw)


class _Transpose12(nn.Module):  # stands in for einops Rearrange('b n c -> b c n') (no parameters)
    def forward(self, x):
        return x.transpose(1, 2)


class Attention(nn.Module):
    """MHSA with Shaw relative-position bias; q and kv projections have no bias."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, max_pos_emb=512):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale, self.max_pos_emb = heads, dim_head ** -0.5, max_pos_emb
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim)
        self.rel_pos_emb = nn.Embedding(2 * max_pos_emb + 1, dim_head)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, mask=None):
        B, n, _ = x.shape
        h = self.heads
        q = self.to_q(x)
        k, v = self.to_kv(x).chunk(2, dim=-1)
        q, k, v = (t.view(B, n, h, -1).transpose(1, 2) for t in (q, k, v))  # b h n d
        dots = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        seq = torch.arange(n, device=x.device)
        dist = (seq[:, None] - seq[None, :]).clamp(-self.max_pos_emb, self.max_pos_emb) + self.max_pos_emb
        rel = self.rel_pos_emb(dist).to(q)  # (n, n, d)
        dots = dots + torch.einsum("bhnd,nrd->bhnr", q, rel) * self.scale
        attn = dots.softmax(dim=-1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(B, n, -1)
        return self.dropout(self.to_out(out))


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult), Swish(), nn.Dropout(dropout),
                                 nn.Linear(dim * mult, dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class ConformerConvModule(nn.Module):
    def __init__(self, dim, expansion_factor=2, kernel_size=31, dropout=0.0):
        super().__init__()
        inner = dim * expansion_factor
        pad = kernel_size // 2
        padding = (pad, pad - (kernel_size + 1) % 2)
        self.net = nn.Sequential(
            nn.LayerNorm(dim), _Transpose12(), nn.Conv1d(dim, inner * 2, 1), GLU(dim=1),
            DepthWiseConv1d(inner, inner, kernel_size, padding), nn.BatchNorm1d(inner), Swish(),
            nn.Conv1d(inner, dim, 1), _Transpose12(), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class ConformerBlock(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8, ff_mult=4, conv_expansion_factor=2,
                 conv_kernel_size=31, attn_dropout=0.0, ff_dropout=0.0, conv_dropout=0.0):
        super().__init__()
        self.ff1 = Scale(0.5, PreNorm(dim, FeedForward(dim, ff_mult, ff_dropout)))
        self.attn = PreNorm(dim, Attention(dim, heads, dim_head, attn_dropout))
        self.conv = ConformerConvModule(dim, conv_expansion_factor, conv_kernel_size, conv_dropout)
        self.ff2 = Scale(0.5, PreNorm(dim, FeedForward(dim, ff_mult, ff_dropout)))
        self.post_norm = nn.LayerNorm(dim)

    def forward(self, x, mask=None):
        x = self.ff1(x) + x
        x = self.attn(x, mask=mask) + x
        x = self.conv(x) + x
        x = self.ff2(x) + x
        return self.post_norm(x)
