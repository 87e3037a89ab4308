"""Oracle restatement of the reference's AASIST back-end (eval mode, fp32).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Follows ``/root/reference/models/aasist_modules.py`` and
``/root/reference/models/xlsr_aasist.py`` (line numbers cited per function).
Parameter names equal the reference's so its state dicts load unchanged
(SURVEY.md App. A.5).  Pinned against the reference's own files, executed
unmodified, by ``oracle/check_against_reference.py``.

Reference quirks reproduced on purpose:
  * ``out_S1 = out_S1 + 1`` (xlsr_aasist.py:138);
  * ``Residual_block`` computes bn1+selu and then feeds the raw input to conv1
    (aasist_modules.py:376-383) -- bn1 parameters exist but do not act;
  * ``master1/master2.expand`` results are unused (xlsr_aasist.py:125-126);
  * pooled nodes come out in descending-score order (aasist_modules.py:332).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _xavier_col(n):
    p = nn.Parameter(torch.empty(n, 1))
    nn.init.xavier_normal_(p)
    return p


def _bn_rows(bn, x):
    """BatchNorm1d over the last dim of (B,n,D) (aasist_modules.py:98-104, 278-284)."""
    return bn(x.reshape(-1, x.shape[-1])).view_as(x)


class GraphAttentionLayer(nn.Module):
    """aasist_modules.py:17-110."""

    def __init__(self, in_dim, out_dim, temperature=1.0):
        super().__init__()
        self.att_proj = nn.Linear(in_dim, out_dim)
        self.att_weight = _xavier_col(out_dim)
        self.proj_with_att = nn.Linear(in_dim, out_dim)
        self.proj_without_att = nn.Linear(in_dim, out_dim)
        self.bn = nn.BatchNorm1d(out_dim)
        self.temp = temperature

    def forward(self, x):  # (B,n,D)
        pair = x[:, :, None, :] * x[:, None, :, :]                      # :57-69  x_i * x_j
        e = torch.tanh(self.att_proj(pair)) @ self.att_weight           # :76-80  (B,n,n,1)
        a = torch.softmax(e / self.temp, dim=-2).squeeze(-1)            # :83-85  softmax over j
        y = self.proj_with_att(a @ x) + self.proj_without_att(x)        # :89-93
        return F.selu(_bn_rows(self.bn, y))                             # :53-55


class HtrgGraphAttentionLayer(nn.Module):
    """aasist_modules.py:112-294."""

    def __init__(self, in_dim, out_dim, temperature=1.0):
        super().__init__()
        self.proj_type1 = nn.Linear(in_dim, in_dim)
        self.proj_type2 = nn.Linear(in_dim, in_dim)
        self.att_proj = nn.Linear(in_dim, out_dim)
        self.att_projM = nn.Linear(in_dim, out_dim)
        self.att_weight11 = _xavier_col(out_dim)
        self.att_weight22 = _xavier_col(out_dim)
        self.att_weight12 = _xavier_col(out_dim)
        self.att_weightM = _xavier_col(out_dim)
        self.proj_with_att = nn.Linear(in_dim, out_dim)
        self.proj_without_att = nn.Linear(in_dim, out_dim)
        self.proj_with_attM = nn.Linear(in_dim, out_dim)
        self.proj_without_attM = nn.Linear(in_dim, out_dim)
        self.bn = nn.BatchNorm1d(out_dim)
        self.temp = temperature

    def forward(self, x1, x2, master=None):
        n1, n2 = x1.shape[1], x2.shape[1]
        x = torch.cat([self.proj_type1(x1), self.proj_type2(x2)], dim=1)          # :159-164
        if master is None:
            master = x.mean(dim=1, keepdim=True)                                  # :167-168
        # pairwise attention with quadrant-specific weight vectors (:239-268)
        h = torch.tanh(self.att_proj(x[:, :, None, :] * x[:, None, :, :]))        # (B,n,n,Do)
        first = torch.arange(n1 + n2, device=x.device) < n1
        same1 = first[:, None] & first[None, :]
        same2 = (~first[:, None]) & (~first[None, :])
        w = torch.where(same1[..., None], self.att_weight11[:, 0],
                        torch.where(same2[..., None], self.att_weight22[:, 0], self.att_weight12[:, 0]))
        e = (h * w).sum(-1)                                                        # (B,n,n)
        a = torch.softmax(e / self.temp, dim=-1)
        # master node (:201-237, 275-281); uses the pre-update x
        em = torch.tanh(self.att_projM(x * master)) @ self.att_weightM            # (B,n,1)
        am = torch.softmax(em / self.temp, dim=-2)                                 # over nodes
        master = self.proj_with_attM(am.transpose(1, 2) @ x) + self.proj_without_attM(master)
        y = self.proj_with_att(a @ x) + self.proj_without_att(x)                  # :270-274
        y = F.selu(_bn_rows(self.bn, y))
        return y[:, :n1], y[:, n1:], master


class GraphPool(nn.Module):
    """aasist_modules.py:296-338: keep the top floor(n*k) nodes, descending score order."""

    def __init__(self, k, in_dim, p=0.0):
        super().__init__()
        self.k = k
        self.proj = nn.Linear(in_dim, 1)

    def forward(self, h, return_idx=False):
        s = torch.sigmoid(self.proj(h))                                 # :307-308
        keep = max(int((torch.as_tensor(h.shape[1]) * torch.tensor(self.k)).long()), 1)  # :329-330
        idx = torch.topk(s, keep, dim=1).indices                        # :332 (sorted desc)
        out = torch.gather(h * s, 1, idx.expand(-1, -1, h.shape[2]))    # :333-336
        return (out, idx.squeeze(-1)) if return_idx else out


class Residual_block(nn.Module):
    """aasist_modules.py:340-397."""

    def __init__(self, nb_filts, first=False):
        super().__init__()
        if not first:
            self.bn1 = nn.BatchNorm2d(nb_filts[0])  # parameters exist; output discarded (:376-383)
        self.conv1 = nn.Conv2d(nb_filts[0], nb_filts[1], (2, 3), padding=(1, 1))
        self.bn2 = nn.BatchNorm2d(nb_filts[1])
        self.conv2 = nn.Conv2d(nb_filts[1], nb_filts[1], (2, 3), padding=(0, 1))
        if nb_filts[0] != nb_filts[1]:
            self.conv_downsample = nn.Conv2d(nb_filts[0], nb_filts[1], (1, 3), padding=(0, 1))
        else:
            self.conv_downsample = None

    def forward(self, x):
        out = self.conv2(F.selu(self.bn2(self.conv1(x))))
        idt = x if self.conv_downsample is None else self.conv_downsample(x)
        return out + idt


class AasistBackend(nn.Module):
    """Everything of ``XLSR_AASIST`` after the SSL front-end (xlsr_aasist.py:24-84, 89-177)."""

    def __init__(self):
        super().__init__()
        filts = [128, [1, 32], [32, 32], [32, 64], [64, 64]]
        gat_dims = [64, 32]
        self.LL = nn.Linear(1024, 128)
        self.first_bn = nn.BatchNorm2d(1)
        self.first_bn1 = nn.BatchNorm2d(64)
        self.encoder = nn.Sequential(
            nn.Sequential(Residual_block(filts[1], first=True)),
            nn.Sequential(Residual_block(filts[2])),
            nn.Sequential(Residual_block(filts[3])),
            nn.Sequential(Residual_block(filts[4])),
            nn.Sequential(Residual_block(filts[4])),
            nn.Sequential(Residual_block(filts[4])))
        self.attention = nn.Sequential(
            nn.Conv2d(64, 128, (1, 1)), nn.SELU(), nn.BatchNorm2d(128), nn.Conv2d(128, 64, (1, 1)))
        self.pos_S = nn.Parameter(torch.randn(1, 42, 64))
        self.master1 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))
        self.master2 = nn.Parameter(torch.randn(1, 1, gat_dims[0]))
        self.GAT_layer_S = GraphAttentionLayer(64, gat_dims[0], temperature=2.0)
        self.GAT_layer_T = GraphAttentionLayer(64, gat_dims[0], temperature=2.0)
        for name, (i, o) in {"11": (0, 1), "12": (1, 1), "21": (0, 1), "22": (1, 1)}.items():
            setattr(self, "HtrgGAT_layer_ST" + name,
                    HtrgGraphAttentionLayer(gat_dims[i], gat_dims[o], temperature=100.0))
        self.pool_S = GraphPool(0.5, gat_dims[0])
        self.pool_T = GraphPool(0.5, gat_dims[0])
        for name in ("hS1", "hT1", "hS2", "hT2"):
            setattr(self, "pool_" + name, GraphPool(0.5, gat_dims[1]))
        self.out_layer = nn.Linear(5 * gat_dims[1], 2)

    def backend(self, feats, taps=None):
        """feats: (B,T,1024) SSL features -> (B,2) logits."""
        x = self.LL(feats).transpose(1, 2).unsqueeze(1)                  # :89-93
        x = F.selu(self.first_bn(F.max_pool2d(x, (3, 3))))                # :94-96
        x = F.selu(self.first_bn1(self.encoder(x)))                       # :99-101
        w = self.attention(x)                                             # :103
        e_S = (x * torch.softmax(w, dim=-1)).sum(-1).transpose(1, 2) + self.pos_S   # :106-108
        e_T = (x * torch.softmax(w, dim=-2)).sum(-2).transpose(1, 2)                # :115-118
        gat_S = self.GAT_layer_S(e_S)
        gat_T = self.GAT_layer_T(e_T)
        out_S, idx_S = self.pool_S(gat_S, return_idx=True)                # :111-112
        out_T, idx_T = self.pool_T(gat_T, return_idx=True)                # :121-122
        if taps is not None:
            taps.update(e_S=e_S, e_T=e_T, gat_S=gat_S, gat_T=gat_T, idx_S=idx_S, idx_T=idx_T,
                        out_S=out_S, out_T=out_T)

        def branch(l1, l2, pS, pT, master, quirk):
            T1, S1, m = l1(out_T, out_S, master=master)                   # :129-130 / :143-144
            S1, T1 = pS(S1), pT(T1)
            Ta, Sa, ma = l2(T1, S1, master=m)                             # :135-136 / :148-149
            S1 = S1 + 1 if quirk else S1 + Sa                             # :138 (quirk) / :150
            return T1 + Ta, S1, m + ma

        T1, S1, m1 = branch(self.HtrgGAT_layer_ST11, self.HtrgGAT_layer_ST12,
                            self.pool_hS1, self.pool_hT1, self.master1, True)
        T2, S2, m2 = branch(self.HtrgGAT_layer_ST21, self.HtrgGAT_layer_ST22,
                            self.pool_hS2, self.pool_hT2, self.master2, False)
        T, S, m = torch.max(T1, T2), torch.max(S1, S2), torch.max(m1, m2)  # :160-162
        hidden = torch.cat([T.abs().max(dim=1).values, T.mean(dim=1),
                            S.abs().max(dim=1).values, S.mean(dim=1), m.squeeze(1)], dim=1)  # :165-172
        return self.out_layer(hidden)                                      # :175


def perturb_norm_stats(model, seed=7):
    """Give every BN non-trivial running stats / affine and every LN a +-5% affine so that
    folding bugs are visible.  Biases of Linear/Conv that init to zero
    are also randomised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            elif isinstance(m, (nn.LayerNorm, nn.GroupNorm)):
                m.weight.copy_(1.0 + (torch.rand(m.weight.shape, generator=g) - 0.5) * 0.1)
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) - 0.5) * 0.1)
            elif isinstance(m, (nn.Linear, nn.Conv1d, nn.Conv2d)) and m.bias is not None:
                if float(m.bias.abs().max()) == 0.0:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
        for n, p in model.named_parameters():
            if n.endswith("pos_conv.0.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model
