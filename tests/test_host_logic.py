"""CPU: host-side mirror of the reference interface (state-dict contract, layer selection, error
behaviour) and the sharding / gather logic (world_size 2 over gloo)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from tests.util import pkg


def test_state_dict_contract_matches_oracle_and_reference():
    from oracle import models_ref as O
    xa = pkg("models.xlsr_aasist")
    cb = pkg("models.conformer_baseline")
    ora = O.build("My_XLSR_AASIST", perturb=False, num_layers=2)
    prod = xa.My_XLSR_AASIST("cpu", None, num_layers=2)
    assert set(prod.state_dict()) == set(ora.state_dict())
    for k, v in ora.state_dict().items():
        assert prod.state_dict()[k].shape == v.shape, k
    prod.load_state_dict(ora.state_dict(), strict=True)
    orc = O.build("MyModel", perturb=False, num_layers=1)
    prc = cb.MyModel("cpu", None, num_layers=1)
    assert set(prc.state_dict()) == set(orc.state_dict())
    # key spot checks from SURVEY.md App. A.5
    keys = set(prod.state_dict())
    for k in ("ssl_model.model.feature_extractor.conv_layers.0.0.weight",
              "ssl_model.model.feature_extractor.conv_layers.6.2.1.bias",
              "ssl_model.model.encoder.pos_conv.0.weight_g", "ssl_model.model.encoder.layers.1.self_attn.q_proj.weight",
              "ssl_model.model.mask_emb", "encoder.1.0.bn1.weight", "encoder.2.0.conv_downsample.bias",
              "HtrgGAT_layer_ST12.att_weightM", "pool_hT2.proj.weight", "pos_S", "master2", "out_layer.bias"):
        assert k in keys, k
    assert "conformer.encoder_blocks.0.attn.fn.rel_pos_emb.weight" in set(prc.state_dict())
    assert prc.state_dict()["conformer.encoder_blocks.0.conv.net.4.conv.weight"].shape == (288, 1, 31)


def test_layer_selection_and_errors():
    fe = pkg("models.fe")
    assert fe.middle_indices(24, 6) == list(range(9, 15))
    m = fe.My_XLSR_FE("cpu", num_layers=3, order="last")
    assert len(m.model.encoder.layers) == 3
    full = fe.XLSR_FE("cpu")
    assert len(full.model.encoder.layers) == 24 and full.out_dim == 1024
    # the kept modules are the *same* objects in the requested order (fe.py:69-90)
    c = fe.My_XLSR_FE("cpu", num_layers=2, order="custom", custom_order=[5, 1])
    assert len(c.model.encoder.layers) == 2
    with pytest.raises(ValueError):
        fe.My_XLSR_FE("cpu", num_layers=0)
    with pytest.raises(ValueError):
        fe.My_XLSR_FE("cpu", num_layers=25)
    with pytest.raises(ValueError):
        fe.My_XLSR_FE("cpu", num_layers=2, order="custom")
    with pytest.raises(ValueError):
        fe.My_XLSR_FE("cpu", num_layers=2, order="custom", custom_order=(1, 2))


def test_no_cpu_fallback_and_reference_typeerror():
    xa = pkg("models.xlsr_aasist")
    cb = pkg("models.conformer_baseline")
    m = xa.My_XLSR_AASIST("cpu", None, num_layers=1).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 16000))
    with pytest.raises(RuntimeError, match="eval"):
        m.train()(torch.zeros(1, 16000))
    with pytest.raises(RuntimeError):
        m.GAT_layer_S(torch.zeros(1, 4, 64))           # sub-blocks run only inside the fused path
    s = cb.MyModel("cpu", None, num_layers=1).eval()
    with pytest.raises(TypeError):                      # conformer_baseline.py:98 as shipped
        s(torch.zeros(1, 16000))


def test_pretraining_heads_in_checkpoints_are_tolerated():
    w2v = pkg("models.wav2vec2_params")
    m = w2v.Wav2Vec2Model(layers=1)
    sd = dict(m.state_dict())
    sd["quantizer.vars"] = torch.zeros(3)
    sd["final_proj.weight"] = torch.zeros(2, 2)
    m.load_state_dict(sd, strict=True)


def test_shard_ranges_cover_everything_once():
    sc = pkg("scoring")
    for n, w in ((180000, 8), (10, 4), (7, 8), (0, 2), (64, 1), (65, 2)):
        seen = []
        per = None
        for r in range(w):
            lo, hi, per = sc.shard_range(n, r, w)
            assert 0 <= lo <= hi <= n and hi - lo <= per
            seen += list(range(lo, hi))
        assert seen == list(range(n))
    assert sc.batch_ranges(3, 10, 4) == [(3, 7), (7, 10)]
    with pytest.raises(ValueError):
        sc.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_items, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = pkg("scoring")
    lo, hi, per = sc.shard_range(n_items, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5      # "score" of utterance i is i/2
    out = sc.gather_scores(local, n_items, per)
    ret[rank] = out.tolist()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [11, 8])
def test_two_rank_gloo_gather_restores_global_order(n_items):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_gather_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    expect = [i * 0.5 for i in range(n_items)]
    assert ret[0] == expect and ret[1] == expect


def test_load_pretrained_fairseq_style_checkpoint(tmp_path, monkeypatch):
    """The reference loads ``xlsr2_300m.pt`` through fairseq (models/fe.py:11-15): a pickled ``{'model': state_dict, ...}``
    that also carries the pre-training heads.  Here the path comes from $RTDF_XLSR_CKPT; heads are dropped, every XLS-R
    key must be present, a plain state-dict file works too."""
    import torch
    w2v = pkg("models.wav2vec2_params")
    fe = pkg("models.fe")
    skeleton = w2v.Wav2Vec2Model()
    keys = list(skeleton.state_dict().keys())
    # every tensor is a stride-0 expansion of one element, so the 1.2 GB checkpoint is a few hundred KB on disk
    sd = {k: torch.full((1,), (i + 1) / 1024.0).expand(skeleton.state_dict()[k].shape) for i, k in enumerate(keys)}
    heads = {"quantizer.vars": torch.zeros(1, 640, 384), "quantizer.weight_proj.weight": torch.zeros(640, 512),
             "project_q.weight": torch.zeros(768, 768), "final_proj.bias": torch.zeros(768)}
    path = str(tmp_path / "xlsr2_300m.pt")
    torch.save({"model": {**sd, **heads}, "cfg": {"model": {"_name": "wav2vec2"}}, "args": None}, path)
    monkeypatch.setenv("RTDF_XLSR_CKPT", path)
    m = fe.XLSR_FE("cpu")
    got = m.model.state_dict()
    assert list(got.keys()) == keys                       # no pre-training head leaked into the module
    for i, k in enumerate(keys):
        assert bool((got[k] == (i + 1) / 1024.0).all()), k
    # truncated student (fe.py:53-90) from the same checkpoint: layers selected after the load
    s = fe.My_XLSR_FE("cpu", num_layers=2, order="last")
    want = (keys.index("encoder.layers.23.fc1.bias") + 1) / 1024.0
    assert bool((s.model.encoder.layers[1].fc1.bias == want).all())
    # plain state dict (no 'model' wrapper)
    plain = str(tmp_path / "plain.pt")
    torch.save(sd, plain)
    w2v.load_pretrained(skeleton, plain)
    assert bool((skeleton.post_extract_proj.bias == (keys.index("post_extract_proj.bias") + 1) / 1024.0).all())
    # a checkpoint that lacks XLS-R keys fails loudly
    broken = str(tmp_path / "broken.pt")
    torch.save({"model": {k: v for k, v in sd.items() if "layers.3." not in k}}, broken)
    with pytest.raises(RuntimeError, match="lacks XLS-R keys"):
        w2v.load_pretrained(skeleton, broken)


def test_shard_and_batch_ranges_properties():
    """Property test (hypothesis): for any utterance count, world size and batch size the shards partition [0, n) in order,
    every shard but possibly trailing ones has ceil(n / W) items, and the batches of a shard tile it exactly with one ragged
    batch at most (drop_last=False, reference main.py:200)."""
    from hypothesis import given, settings, strategies as st
    sc = pkg("scoring")

    @settings(max_examples=300, deadline=None)
    @given(n=st.integers(0, 200_000), world=st.integers(1, 8), batch=st.integers(1, 257))
    def prop(n, world, batch):
        pos = 0
        per_ref = -(-n // world) if n else 0
        for r in range(world):
            lo, hi, per = sc.shard_range(n, r, world)
            assert per == per_ref and lo == pos and lo <= hi <= n and hi - lo <= per
            pos = hi
            ranges = sc.batch_ranges(lo, hi, batch)
            assert [a for a, _ in ranges] == list(range(lo, hi, batch))
            assert all(0 < b - a <= batch for a, b in ranges)
            assert sum(b - a for a, b in ranges) == hi - lo
            assert sum(1 for a, b in ranges if b - a < batch) <= 1
        assert pos == n

    prop()
    with pytest.raises(ValueError):
        sc.shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        sc.batch_ranges(0, 10, 0)


def test_extractor_config_is_read_from_checkpoint_keys():
    """XLS-R checkpoints carry per-conv LayerNorms and conv biases (extractor_mode=layer_norm); wav2vec2-base style ones a
    GroupNorm after conv-0 and no conv bias (default).  The mirror classes build the matching parameter layout."""
    w2v = pkg("models.wav2vec2_params")
    fe = pkg("models.fe")
    ln = w2v.Wav2Vec2Model(layers=1)
    gn = w2v.Wav2Vec2Model(layers=1, extractor_mode="default", conv_bias=False)
    assert w2v.extractor_config(ln.state_dict()) == ("layer_norm", True)
    assert w2v.extractor_config(gn.state_dict()) == ("default", False)
    assert "feature_extractor.conv_layers.0.2.weight" in gn.state_dict()           # GroupNorm affine
    assert "feature_extractor.conv_layers.1.2.1.weight" not in gn.state_dict()     # no LayerNorm after conv-1
    assert "feature_extractor.conv_layers.0.0.bias" not in gn.state_dict()
    with pytest.raises(ValueError):
        w2v.Wav2Vec2Model(layers=1, extractor_mode="group")
    m = fe.My_XLSR_FE("cpu", num_layers=1, extractor_mode="default")
    assert w2v.extractor_config(m.model.state_dict()) == ("default", False)
