"""Engine: owns one ``rtdf_ctx`` (packed weights on one GPU) and a workspace, and runs the
scoring forward through the C-ABI on the caller's current CUDA stream.

PyTorch is used for device memory and streams only.  Nothing here computes on the CPU and
nothing falls back to PyTorch ops: a missing library, a non-CUDA tensor or a failed call raise.
"""
import ctypes
import os

import torch

from . import native


def _precision_from_env(default="bf16"):
    p = os.environ.get("RTDF_PRECISION", default).lower()
    if p not in ("bf16", "fp32"):
        raise ValueError(f"RTDF_PRECISION must be 'bf16' or 'fp32', got {p!r}")
    return p


class Engine:
    """One packed model instance on one device.

    state_dict keys follow the reference; a leading ``module.`` is
    stripped (reference utils.py:13-43).
    """

    def __init__(self, state_dict, device, backend, n_layers, precision=None, conformer=None,
                 attention_impl=None, aasist_conv_impl=None, use_graph=None, gat_impl=None):
        self.lib = native.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the rtdf engine runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision or _precision_from_env()
        conformer = conformer or {}
        if attention_impl is None:
            attention_impl = int(os.environ.get("RTDF_ATTENTION_IMPL", "0"))
        if aasist_conv_impl is None:
            aasist_conv_impl = int(os.environ.get("RTDF_AASIST_CONV_IMPL", "0"))
        if gat_impl is None:
            gat_impl = int(os.environ.get("RTDF_GAT_IMPL", "0"))
        desc = native.ModelDesc(
            backend={"aasist": native.BACKEND_AASIST, "conformer": native.BACKEND_CONFORMER,
                     None: native.BACKEND_NONE, "none": native.BACKEND_NONE}[backend],
            n_layers=int(n_layers),
            precision=native.PREC_BF16 if self.precision == "bf16" else native.PREC_FP32,
            conf_emb=int(conformer.get("emb_size", 144)), conf_heads=int(conformer.get("heads", 4)),
            conf_kernel=int(conformer.get("kernel_size", 31)), conf_blocks=int(conformer.get("n_encoders", 4)),
            attention_impl=attention_impl, aasist_conv_impl=aasist_conv_impl,
            gat_impl=gat_impl)
        self.backend_kind = backend
        self.n_layers = int(n_layers)
        self._ctx = ctypes.c_void_p()
        self._ws = None
        # CUDA graphs: the forward allocates nothing and never synchronises, so one captured graph per input
        # shape replaces ~215 launches by one (RTDF_CUDA_GRAPH=0 disables).
        self.use_graph = bool(int(os.environ.get("RTDF_CUDA_GRAPH", "1"))) if use_graph is None else bool(use_graph)
        self._graphs = {}
        # "auto": streaming chunks (<= 512 frames in flight) use the low-latency weight-streaming kernels;
        # "throughput": batch-composition-invariant large-batch kernels only (include/rtdf.h rtdf_regime).
        self.regime = os.environ.get("RTDF_REGIME", "auto")
        self._regime_set = None
        with torch.cuda.device(self.device):
            native.check(self.lib.rtdf_create(ctypes.byref(self._ctx), self.device.index, ctypes.byref(desc)),
                         "rtdf_create")
            try:
                for key, t in state_dict.items():
                    if not torch.is_tensor(t) or not t.is_floating_point():
                        continue  # e.g. BatchNorm num_batches_tracked
                    t = t.detach().to(dtype=torch.float32).contiguous()
                    shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                    native.check(self.lib.rtdf_load_weight(self._ctx, key.encode(), native.ptr(t), shape, t.dim()),
                                 f"rtdf_load_weight({key})")
                native.check(self.lib.rtdf_finalize(self._ctx), "rtdf_finalize")
            except Exception:
                self.close()
                raise

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.rtdf_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()
        self._ws = None
        self._graphs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_frames(self, n_samples):
        return int(self.lib.rtdf_num_frames(int(n_samples)))

    def _workspace(self, B, N):
        need = ctypes.c_size_t()
        native.check(self.lib.rtdf_workspace_bytes(self._ctx, B, N, ctypes.byref(need)), "rtdf_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._graphs.clear()  # captured graphs hold pointers into the old workspace
            self._ws = None  # release before growing
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def _apply_regime(self, regime):
        regime = regime or self.regime
        if regime not in native.REGIMES:
            raise ValueError(f"rtdf: unknown regime {regime!r} (expected one of {sorted(native.REGIMES)})")
        if regime != self._regime_set:
            native.check(self.lib.rtdf_set_regime(self._ctx, native.REGIMES[regime]), "rtdf_set_regime")
            self._regime_set = regime
        return regime

    def _launch_forward(self, wav, B, N, preemph, coef, logits, ws):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        torch.cuda.nvtx.range_push(f"rtdf_forward B={B} N={N} {self.precision}")     # visible in nsys / ncu --nvtx timelines
        try:
            native.check(self.lib.rtdf_forward(self._ctx, native.ptr(wav), B, N, int(bool(preemph)), float(coef),
                                               native.ptr(logits), native.ptr(ws), ws.numel(), None,
                                               ctypes.c_void_p(stream)), "rtdf_forward")
        finally:
            torch.cuda.nvtx.range_pop()

    def _graph_entry(self, B, N, preemph, coef, regime, first_input=None):
        """(graph, static_in, static_out) for one input shape; captured on first use."""
        key = (B, N, bool(preemph), float(coef), regime)
        ws = self._workspace(B, N)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.zeros(B, N, dtype=torch.float32, device=self.device)
            static_out = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            if first_input is not None:
                static_in.copy_(first_input)
            self._launch_forward(static_in, B, N, preemph, coef, static_out, ws)   # eager warm-up (also validates)
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._launch_forward(static_in, B, N, preemph, coef, static_out, ws)
            if len(self._graphs) >= 16:
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._graphs[key] = (graph, static_in, static_out)
        return entry

    def _graph_forward(self, wav, B, N, preemph, coef, regime):
        graph, static_in, static_out = self._graph_entry(B, N, preemph, coef, regime, first_input=wav)
        if wav.data_ptr() != static_in.data_ptr():      # zero-copy when the caller filled static_input() directly
            static_in.copy_(wav)
        graph.replay()
        return static_out.clone()

    def static_input(self, B, N, preemph=False, coef=0.97, regime=None):
        """The (B,N) fp32 device buffer the captured graph of this shape reads.  A streaming caller copies its chunk
        straight into it (e.g. ``buf.copy_(pinned_host, non_blocking=True)``) and passes it to ``forward`` /
        ``forward_static``: no device-to-device staging copy on the latency path."""
        with torch.cuda.device(self.device):
            regime = self._apply_regime(regime)
            return self._graph_entry(int(B), int(N), preemph, coef, regime)[1]

    def forward_static(self, B, N, preemph=False, coef=0.97, regime=None):
        """Replay the captured graph on whatever ``static_input(B, N)`` holds; returns the graph's own (B,2) output buffer
        (valid until the next replay of this shape).  Two launches fewer than ``forward`` per call."""
        with torch.cuda.device(self.device):
            regime = self._apply_regime(regime)
            graph, _, static_out = self._graph_entry(int(B), int(N), preemph, coef, regime)
            graph.replay()
        return static_out

    def _check_wav(self, wav):
        if not torch.is_tensor(wav) or not wav.is_cuda:
            raise RuntimeError("rtdf: input must be a CUDA tensor (the scoring path has no CPU implementation)")
        if wav.device != self.device:
            raise RuntimeError(f"rtdf: input on {wav.device}, engine on {self.device}")
        if wav.dim() == 3:            # (B,N,1)  -- reference models/fe.py:18
            wav = wav[:, :, 0]
        if wav.dim() != 2:
            raise ValueError(f"rtdf: expected (B,N) waveforms, got shape {tuple(wav.shape)}")
        return wav.to(torch.float32).contiguous()

    def forward(self, wav, preemph=False, coef=0.97, want_taps=False, layer_taps=False, regime=None):
        """(B,N) fp32 CUDA waveforms -> (B,2) fp32 logits [, taps dict].

        regime: None (the engine's default, ``self.regime``), "auto" or "throughput" (batch-composition-invariant
        kernels only; what the sharded scoring sweep uses).

        layer_taps (with want_taps): also return taps['layers'], (n_layers+1, B, T, 1024) fp32 -- the residual
        stream entering layer 0 and leaving each transformer layer (the I/O of ``encoder.layers.N`` that the
        reference's KD forward hooks read, trainer.py:176-195)."""
        wav = self._check_wav(wav)
        B, N = wav.shape
        if B == 0:
            out = torch.empty(0, 2, dtype=torch.float32, device=self.device)
            return (out, {}) if want_taps else out
        with torch.cuda.device(self.device):
            regime = self._apply_regime(regime)
            if self.use_graph and not want_taps and not torch.cuda.is_current_stream_capturing():
                return self._graph_forward(wav, B, N, preemph, coef, regime)
            ws = self._workspace(B, N)
            logits = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            taps_struct, taps = None, {}
            if want_taps:
                T = self.num_frames(N)
                taps["feats"] = torch.empty(B, T, 1024, dtype=torch.float32, device=self.device)
                taps_struct = native.Taps(feats=taps["feats"].data_ptr())
                if layer_taps:
                    taps["layers"] = torch.empty(self.n_layers + 1, B, T, 1024, dtype=torch.float32, device=self.device)
                    taps_struct.layers = taps["layers"].data_ptr()
                if self.backend_kind == "aasist":
                    Tp = T // 3
                    taps["hidden"] = torch.empty(B, 160, dtype=torch.float32, device=self.device)
                    taps["idx_S"] = torch.empty(B, 21, dtype=torch.int32, device=self.device)
                    taps["idx_T"] = torch.empty(B, max(Tp // 2, 1), dtype=torch.int32, device=self.device)
                    taps_struct.hidden = taps["hidden"].data_ptr()
                    taps_struct.idx_S = taps["idx_S"].data_ptr()
                    taps_struct.idx_T = taps["idx_T"].data_ptr()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            native.check(self.lib.rtdf_forward(self._ctx, native.ptr(wav), B, N, int(bool(preemph)), float(coef),
                                               native.ptr(logits), native.ptr(ws), ws.numel(),
                                               ctypes.byref(taps_struct) if taps_struct is not None else None,
                                               ctypes.c_void_p(stream)), "rtdf_forward")
        return (logits, taps) if want_taps else logits

    def frontend(self, wav, preemph=False, coef=0.97, regime=None):
        """XLSR_FE.extract_feat: (B,N) -> (B,T,1024) fp32."""
        wav = self._check_wav(wav)
        self._apply_regime(regime)
        B, N = wav.shape
        T = self.num_frames(N)
        if T < 1:
            raise ValueError(f"rtdf: {N} samples are too few for the conv feature encoder")
        with torch.cuda.device(self.device):
            ws = self._workspace(B, N)
            feats = torch.empty(B, T, 1024, dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            native.check(self.lib.rtdf_frontend(self._ctx, native.ptr(wav), B, N, int(bool(preemph)), float(coef),
                                                native.ptr(feats), native.ptr(ws), ws.numel(), ctypes.c_void_p(stream)),
                         "rtdf_frontend")
        return feats

    def backend(self, feats, want_taps=False):
        """(B,T,1024) fp32 features -> (B,2) logits."""
        if not feats.is_cuda or feats.dim() != 3 or feats.shape[2] != 1024:
            raise ValueError("rtdf: backend expects (B,T,1024) CUDA features")
        feats = feats.to(torch.float32).contiguous()
        B, T, _ = feats.shape
        self._apply_regime(None)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, max(400, T * 320 + 80))
            logits = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            taps_struct, taps = None, {}
            if want_taps and self.backend_kind == "aasist":
                taps["hidden"] = torch.empty(B, 160, dtype=torch.float32, device=self.device)
                taps["idx_S"] = torch.empty(B, 21, dtype=torch.int32, device=self.device)
                taps["idx_T"] = torch.empty(B, max((T // 3) // 2, 1), dtype=torch.int32, device=self.device)
                taps_struct = native.Taps(hidden=taps["hidden"].data_ptr(), idx_S=taps["idx_S"].data_ptr(),
                                          idx_T=taps["idx_T"].data_ptr())
            stream = torch.cuda.current_stream(self.device).cuda_stream
            native.check(self.lib.rtdf_backend(self._ctx, native.ptr(feats), B, T, native.ptr(logits), native.ptr(ws),
                                               ws.numel(), ctypes.byref(taps_struct) if taps_struct is not None else None,
                                               ctypes.c_void_p(stream)), "rtdf_backend")
        return (logits, taps) if want_taps else logits


def _fingerprint(module):
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def invalidate(module):
    """Drop the engine cached on ``module`` so the next forward re-packs the weights.

    ``engine_for`` notices re-assigned parameters, ``load_state_dict`` and in-place ops through the tensors' version
    counters; writes that bypass them -- ``param.data.normal_()``, ``param.data[...] = ...``, raw pointer writes -- are
    invisible to it: call ``invalidate(model)`` after such a write."""
    cached = module.__dict__.pop("_rtdf_engine", None)
    if cached is not None:
        cached[1].close()


def engine_for(module, backend, n_layers, conformer=None, key_prefix=""):
    """Engine cached on ``module``; rebuilt when any parameter/buffer changed (version counters),
    when the module moved to another device or when the requested precision changed.  ``module.rtdf_frozen = True``
    skips the per-call scan of the ~450 tensors (weights declared final; see ``invalidate``)."""
    if module.training:
        raise RuntimeError("rtdf accelerates the eval-mode scoring forward only: call model.eval() first "
                           "(the reference's scoring callers do: main.py:202, trainer.py:86)")
    cached = module.__dict__.get("_rtdf_engine")
    if cached is not None and getattr(module, "rtdf_frozen", False):
        return cached[1]
    try:
        dev = next(module.parameters()).device
    except StopIteration:
        raise RuntimeError("model has no parameters")
    if dev.type != "cuda":
        raise RuntimeError("rtdf has no CPU path: move the model to a CUDA device with .to(device)")
    precision = getattr(module, "rtdf_precision", None) or _precision_from_env()
    fp = (_fingerprint(module), str(dev), precision, backend, n_layers)
    if cached is not None and cached[0] == fp:
        return cached[1]
    if cached is not None:
        cached[1].close()
    state = {key_prefix + k: v for k, v in module.state_dict().items()}
    eng = Engine(state, dev, backend, n_layers, precision=precision, conformer=conformer)
    module.__dict__["_rtdf_engine"] = (fp, eng)
    return eng
