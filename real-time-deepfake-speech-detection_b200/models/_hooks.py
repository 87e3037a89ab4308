"""Forward-hook support for the fused scoring forward.

The reference distils through ``torchdistill.ForwardHookManager`` hooks registered on sub-modules such as
``ssl_model.model.encoder.layers.N`` (reference trainer.py:176-195, 248-257; main_kd.py:88-141).  Here those
sub-modules are parameter containers -- their arithmetic runs inside one fused CUDA forward -- so the hooks
registered on them are fired by hand, after the forward, with the tensors the fused path tapped:

  * ``ssl_model.model.encoder.layers[i]``: input ``(x_in,)``, output ``(x_out, (None, None))`` in fairseq's
    time-major ``(T, B, C)`` layout (``TransformerSentenceEncoderLayer.forward`` returns ``x, (attn, layer_result)``);
  * ``ssl_model``: input ``(waveform,)``, output ``(B, T, 1024)`` features (models/fe.py:17-21).

Hooks on the model itself fire through ``nn.Module.__call__`` as usual.
"""


def _hooks(module):
    return list(module._forward_hooks.values())


def forward_with_hooks(model, engine, x):
    """engine.forward(x) and fire the forward hooks registered on the tapped sub-modules; returns the logits."""
    ssl = model.ssl_model
    layers = ssl.model.encoder.layers
    hooked = [i for i, layer in enumerate(layers) if layer._forward_hooks]
    ssl_hooked = bool(ssl._forward_hooks)
    if not hooked and not ssl_hooked:
        return engine.forward(x)
    logits, taps = engine.forward(x, want_taps=True, layer_taps=bool(hooked))
    for i in hooked:
        x_in = taps["layers"][i].transpose(0, 1)
        out = (taps["layers"][i + 1].transpose(0, 1), (None, None))
        for hook in _hooks(layers[i]):
            hook(layers[i], (x_in,), out)
    if ssl_hooked:
        for hook in _hooks(ssl):
            hook(ssl, (x,), taps["feats"])
    return logits
