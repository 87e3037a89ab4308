"""Package the reference's own Python files as a build artefact for the checker.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it means zipping its modules from where
they lie under ``/root/reference`` into the git-ignored ``oracle/_ref/reference_py.zip`` (importable through
``zipimport``).  Nothing of the reference is committed; the archive exists only in a working tree where
``__graft_entry__.build()`` ran with ``/root/reference`` present, and travels to the GPU box like a built ``.so``.

Who may use the archive (and only as checker / timed baseline, never on the product path):
  * ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg: the reference's own model classes
    (``models/xlsr_aasist.py`` ...) over import shims for the two un-vendored third-party packages
    (``fairseq`` -> ``oracle/wav2vec2_ref.py``, ``conformer`` -> ``oracle/conformer_block_ref.py``);
  * ``tests/test_dropin_*.py``: the reference's own callers ``produce_evaluation_file`` (main.py:199-221) and
    ``Trainer._test`` (trainer.py:85-132), run unmodified against this package's ``models``.

Usage:  python -m oracle.build_ref            (no-op with a message when /root/reference is absent)
"""
import os
import sys
import types
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RTDF_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(OUT_DIR, "reference_py.zip")
SUBDIRS = ("", "models", "data")


def build(verbose=True):
    """Zip <REF>/*.py, models/*.py, data/*.py -> oracle/_ref/reference_py.zip.  Returns the path or None."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/build_ref: {REF} not present; keeping {'existing' if os.path.exists(ARCHIVE) else 'no'} archive")
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    names = []
    for sub in SUBDIRS:
        d = os.path.join(REF, sub)
        for f in sorted(os.listdir(d)):
            if f.endswith(".py"):
                names.append(os.path.join(sub, f) if sub else f)
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for sub in SUBDIRS:
            if sub:   # explicit directory entries: zipimport only treats `data/` (no __init__.py in the reference) as a
                      # namespace-package portion when the archive lists the directory itself
                z.writestr(zipfile.ZipInfo(sub + "/", date_time=(2020, 1, 1, 0, 0, 0)), b"")
        for n in names:
            info = zipfile.ZipInfo(n, date_time=(2020, 1, 1, 0, 0, 0))   # reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            with open(os.path.join(REF, n), "rb") as fh:
                z.writestr(info, fh.read())
    os.replace(tmp, ARCHIVE)
    if verbose:
        print(f"oracle/build_ref: {len(names)} reference modules -> {ARCHIVE}")
    return ARCHIVE


def reference_root():
    """Path to put on sys.path to import the reference's modules: the source tree when it exists (build container),
    else the archive (GPU box), else None."""
    if os.path.isdir(REF):
        return REF
    return ARCHIVE if os.path.exists(ARCHIVE) else None


def install_third_party_stubs():
    """Empty stand-ins for packages this image lacks and the scoring callers never touch
    (librosa, wandb, audiomentations, torch_audiomentations)."""
    for name in ("librosa", "wandb", "audiomentations"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "torch_audiomentations" not in sys.modules:
        ta = types.ModuleType("torch_audiomentations")
        for n in ("Compose", "AddColoredNoise", "HighPassFilter", "LowPassFilter", "Gain"):
            setattr(ta, n, object)
        sys.modules["torch_audiomentations"] = ta


def install_model_shims():
    """``fairseq`` / ``conformer`` import shims backed by the oracle's restatements of those two un-vendored
    packages (same as oracle/check_against_reference.py), so the reference's model files import unmodified."""
    from oracle.conformer_block_ref import ConformerBlock
    from oracle.wav2vec2_ref import FairseqLikeWav2Vec2

    fairseq = types.ModuleType("fairseq")
    cu = types.ModuleType("fairseq.checkpoint_utils")
    cu.load_model_ensemble_and_task = lambda paths, *a, **k: ([FairseqLikeWav2Vec2()], None, None)
    fairseq.checkpoint_utils = cu
    sys.modules["fairseq"] = fairseq
    sys.modules["fairseq.checkpoint_utils"] = cu
    conformer = types.ModuleType("conformer")
    conformer.ConformerBlock = ConformerBlock
    sys.modules["conformer"] = conformer


def import_reference_models():
    """(models.xlsr_aasist, models.conformer_baseline) of the reference itself, or None when neither the source
    tree nor the archive is available.  The caller's process must not have a top-level ``models`` package yet."""
    root = reference_root()
    if root is None:
        return None
    import importlib
    install_model_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    xa = importlib.import_module("models.xlsr_aasist")
    cb = importlib.import_module("models.conformer_baseline")
    if not (xa.__file__ or "").startswith(root):
        raise RuntimeError(f"top-level 'models' resolved to {xa.__file__}, not the reference at {root}")
    return xa, cb


if __name__ == "__main__":
    build()
