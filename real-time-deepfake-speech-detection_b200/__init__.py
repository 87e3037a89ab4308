"""B200-native (sm_100a) batched evaluation-scoring forward for
hungdinhxuan/real-time-deepfake-speech-detection.

Layout
  csrc/          hand-written CUDA kernels + the C-ABI (``include/rtdf.h``) -> ``librtdf.so``
  rtdf_runtime/  ctypes binding and engine (device memory / streams via PyTorch)
  models/        mirror of the reference's ``models`` package: same module paths, class names,
                 constructor signatures and state-dict keys (drop-in boundary)
  data/          mirror of ``data/preprocess.py`` (PreEmphasis)
  scoring.py     data-parallel sharded scoring driver with a single NCCL gather of scores

The directory name is not a Python identifier; import it with
``importlib.import_module("real-time-deepfake-speech-detection_b200")`` or put this directory on
``sys.path`` so that ``import models`` resolves here instead of the reference.
"""
