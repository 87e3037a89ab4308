"""Cross-check ``oracle/wav2vec2_ref.py`` against two independent in-image implementations
of XLS-R 300M: ``torchaudio.models.wav2vec2_xlsr_300m`` and (optionally) HuggingFace
``Wav2Vec2Model(do_stable_layer_norm=True, feat_extract_norm="layer", conv_bias=True)``.

TEST INFRASTRUCTURE ONLY.  fairseq itself is absent (un-pinned third-party dependency of
the reference, ``models/fe.py:5``): parity of the XLS-R arithmetic is therefore *unpinned by
the reference*; these witnesses are the strongest check available offline.

Usage:  python -m oracle.check_against_witnesses [--hf]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fairseq_to_torchaudio(sd):
    """Key map following torchaudio/models/wav2vec2/utils/import_fairseq.py:51-115."""
    out = {}
    for k, v in sd.items():
        if k == "mask_emb":
            continue
        if k.startswith("feature_extractor.conv_layers."):
            p = k.split(".")
            i = p[2]
            if p[3] == "0":
                out[f"feature_extractor.conv_layers.{i}.conv.{p[4]}"] = v
            else:  # "2.1.weight"
                out[f"feature_extractor.conv_layers.{i}.layer_norm.{p[5]}"] = v
        elif k.startswith("layer_norm."):
            out["encoder.feature_projection." + k] = v
        elif k.startswith("post_extract_proj."):
            out["encoder.feature_projection.projection." + k.split(".")[1]] = v
        elif k.startswith("encoder.pos_conv.0."):
            name = k.split(".")[-1]
            name = {"weight_g": "parametrizations.weight.original0",
                    "weight_v": "parametrizations.weight.original1", "bias": "bias"}[name]
            out["encoder.transformer.pos_conv_embed.conv." + name] = v
        elif k.startswith("encoder.layer_norm."):
            out["encoder.transformer.layer_norm." + k.split(".")[-1]] = v
        elif k.startswith("encoder.layers."):
            p = k.split(".")
            l, rest = p[2], p[3:]
            if rest[0] == "self_attn":
                out[f"encoder.transformer.layers.{l}.attention.{rest[1]}.{rest[2]}"] = v
            elif rest[0] == "self_attn_layer_norm":
                out[f"encoder.transformer.layers.{l}.layer_norm.{rest[1]}"] = v
            elif rest[0] == "fc1":
                out[f"encoder.transformer.layers.{l}.feed_forward.intermediate_dense.{rest[1]}"] = v
            elif rest[0] == "fc2":
                out[f"encoder.transformer.layers.{l}.feed_forward.output_dense.{rest[1]}"] = v
            elif rest[0] == "final_layer_norm":
                out[f"encoder.transformer.layers.{l}.final_layer_norm.{rest[1]}"] = v
        else:
            raise KeyError(k)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hf", action="store_true")
    args = ap.parse_args()
    import torchaudio
    from oracle.aasist_ref import perturb_norm_stats
    from oracle.models_ref import synth_waveforms
    from oracle.wav2vec2_ref import FairseqLikeWav2Vec2

    torch.manual_seed(0)
    ours = perturb_norm_stats(FairseqLikeWav2Vec2().eval())
    ta = torchaudio.models.wav2vec2_xlsr_300m().eval()
    res = ta.load_state_dict(fairseq_to_torchaudio(ours.state_dict()), strict=True)
    x = synth_waveforms(2, 64000)
    with torch.no_grad():
        a = ours(x)["x"]
        b = ta.encoder(ta.feature_extractor(x, None)[0], None)
    d = float((a - b).abs().max())
    print(f"oracle vs torchaudio.wav2vec2_xlsr_300m: shape {tuple(a.shape)} max|diff| = {d:.3e} (|x| mean {float(a.abs().mean()):.3f})")
    assert d < 5e-5
    # extractor_mode="default" (wav2vec2-base style: GroupNorm(512, 512) after conv-0 only, no conv bias): the conv stack of
    # the oracle against torchaudio's "group_norm" feature extractor (TA:components.py:564-574)
    from torchaudio.models.wav2vec2 import components as tac
    from oracle.wav2vec2_ref import CONV_LAYERS, ConvFeatureExtractionModel
    torch.manual_seed(1)
    fe = ConvFeatureExtractionModel(mode="default", conv_bias=False).eval()
    with torch.no_grad():
        fe.conv_layers[0][2].weight.uniform_(0.8, 1.2)
        fe.conv_layers[0][2].bias.normal_(0, 0.1)
    ta_fe = tac._get_feature_extractor("group_norm", list(CONV_LAYERS), bias=False).eval()
    sd = {}
    for k, v in fe.state_dict().items():
        i, j, name = k.split(".")[1], k.split(".")[2], k.split(".")[-1]
        sd[f"conv_layers.{i}.conv.{name}" if j == "0" else f"conv_layers.{i}.layer_norm.{name}"] = v
    ta_fe.load_state_dict(sd, strict=True)
    with torch.no_grad():
        a = fe(x).transpose(1, 2)
        b = ta_fe(x, None)[0]
    d = float((a - b).abs().max())
    print(f"oracle conv stack (group-norm mode) vs torchaudio group_norm feature extractor: shape {tuple(a.shape)} "
          f"max|diff| = {d:.3e}")
    assert d < 2e-5
    if args.hf:
        from torchaudio.models.wav2vec2.utils import import_huggingface_model
        from transformers import Wav2Vec2Config, Wav2Vec2Model
        cfg = Wav2Vec2Config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                             feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
                             num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
                             hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, layerdrop=0.0,
                             activation_dropout=0.0, final_dropout=0.0)
        hf = Wav2Vec2Model(cfg).eval()
        ta2 = import_huggingface_model(hf).eval()
        with torch.no_grad():
            h = hf(x).last_hidden_state
            t = ta2.encoder(ta2.feature_extractor(x, None)[0], None)
        print(f"HF Wav2Vec2Model vs torchaudio import: max|diff| = {float((h - t).abs().max()):.3e}")
    print("OK")


if __name__ == "__main__":
    main()
