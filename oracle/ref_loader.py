"""Import the reference's own hot-path modules from ``/root/reference`` (read-only).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only where
``/root/reference`` exists (the build container); the GPU box never calls it.

Recipe:
  * stub ``matplotlib``, ``matplotlib.pyplot``, ``unidecode``, ``inflect`` with real
    module objects (``utils/tools.py:11``, ``text/cleaners.py:18``, ``text/numbers.py:3``);
  * pre-register an empty package ``model`` so ``model/__init__.py`` (which pulls
    in loss / pitch tools / tensorflow) never runs;
  * restore ``CUDA_VISIBLE_DEVICES``, which ``model/blocks.py:2`` overwrites.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("MIXGAN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "diffusion.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = None


def _missing_attr(mod, attr):
    if attr.startswith("__"):
        raise AttributeError(attr)

    def _absent(*a, **k):
        raise RuntimeError(f"{mod}.{attr} is a stub: {mod} is not installed in this container")
    return _absent


def load():
    """Return the reference's ``model.diffusion`` module (with ``tqdm`` silenced)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.__path__ = []
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("unidecode", unidecode=lambda s: s)
    _stub("inflect", engine=lambda: SimpleNamespace())
    if "model" not in sys.modules:
        pkg = types.ModuleType("model")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "model")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("model", loader=None, is_package=True)
        sys.modules["model"] = pkg
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mod = importlib.import_module("model.diffusion")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        if saved is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved
    mod.tqdm = lambda it, **k: it
    _loaded = mod
    return mod


def load_module(name: str):
    """Import another module of the reference tree (e.g. ``model.linguistic_encoder``, ``model.mixgantts``,
    ``transformer.Models``, ``hifigan.models``) under the same stubs; ``CUDA_VISIBLE_DEVICES`` is restored."""
    load()
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        # Third-party packages that only the reference's preprocessing / plotting code needs (librosa, pycwt, parselmouth,
        # g2p_en ...) are absent here: stub each one the import trips over and retry.  Nothing on the paths the goldens
        # exercise touches them.
        for _ in range(32):
            try:
                return importlib.import_module(name)
            except ModuleNotFoundError as e:
                missing = e.name or ""
                if not missing or os.path.exists(os.path.join(REFERENCE_ROOT, *missing.split("."))) \
                        or os.path.exists(os.path.join(REFERENCE_ROOT, *missing.split(".")) + ".py"):
                    raise
                parts = missing.split(".")
                for i in range(1, len(parts) + 1):
                    m = _stub(".".join(parts[:i]))
                    m.__path__ = []
                    m.__getattr__ = (lambda mod: (lambda attr: _missing_attr(mod, attr)))(".".join(parts[:i]))
        raise RuntimeError(f"could not import {name} from the reference")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        if saved is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved


def build_reference_diffusion(args, preprocess_config, model_config, train_config, weights: dict):
    """Construct the reference ``GaussianDiffusion`` and load ``weights`` into its Denoiser."""
    import torch
    mod = load()
    gd = mod.GaussianDiffusion(args, preprocess_config, model_config, train_config)
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    gd.denoise_fn.load_state_dict(sd, strict=True)
    gd.eval()
    return gd


class injected_noise:
    """Context manager: make the reference's ``noise_like`` / ``randn_like`` return
    the given tensors, in call order."""

    def __init__(self, noise_like_seq=(), randn_like_seq=(), randint_seq=()):
        self.nl = list(noise_like_seq)
        self.rl = list(randn_like_seq)
        self.ri = list(randint_seq)

    def __enter__(self):
        import torch
        self.mod = load()
        self._nl, self._rl, self._ri = self.mod.noise_like, torch.randn_like, torch.randint
        nl, rl, ri = iter(self.nl), iter(self.rl), iter(self.ri)
        self.mod.noise_like = lambda shape, device, repeat=False: next(nl)
        self.mod.torch.randn_like = lambda x, **k: next(rl)
        if self.ri:
            self.mod.torch.randint = lambda *a, **k: next(ri)
        return self

    def __exit__(self, *exc):
        import torch
        self.mod.noise_like = self._nl
        torch.randn_like = self._rl
        torch.randint = self._ri
        return False
