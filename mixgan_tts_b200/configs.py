"""Hot-path hyper-parameters of the two shipped reference configurations.

Only the keys the diffusion decoder reads are restated here (values from
``config/LJSpeech/model.yaml:4,27-38,58``, ``config/AISHELL3/model.yaml:33,60``,
``config/*/preprocess.yaml:25`` and ``config/*/train.yaml:23``), so that the
package, its tests and the bench run on a box where ``/root/reference`` does
not exist.  The dictionaries have the same nesting as the reference's YAML so
they can be passed straight to ``Denoiser(preprocess_config, model_config)``
and ``GaussianDiffusion(args, preprocess_config, model_config, train_config)``.
"""
from __future__ import annotations

import copy
import json
import os
import tempfile
from types import SimpleNamespace

N_MEL = 80
SPEC_MIN = -11.5129  # ln(1e-5): the log-compression floor, audio/audio_processing.py:85-92
SPEC_MAX = 2.0

_DENOISER = {
    "denoiser_hidden": 512,      # read by nobody on the hot path
    "denoiser_dropout": 0.2,     # accepted and ignored (model/blocks.py:1136)
    "residual_layers": 20,
    "residual_channels": 256,
    "noise_schedule_naive": "vpsde",
    "timesteps": 4,
    "shallow_timesteps": 1,
    "min_beta": 0.1,
    "max_beta": 40,
    "s": 0.008,
    "keep_bins": 80,
}


def model_config(dataset: str = "LJSpeech", multi_speaker: bool | None = None) -> dict:
    den = dict(_DENOISER)
    if dataset == "AISHELL3":
        den["timesteps"] = 1
    elif dataset != "LJSpeech":
        raise ValueError(f"unknown dataset {dataset!r}")
    cfg = {
        # config/*/model.yaml:1-13 (the decoder keys are read by the aux decoder)
        "transformer": {"encoder_hidden": 256, "decoder_layer": 6, "decoder_head": 2, "decoder_hidden": 256,
                        "conv_filter_size": 1024, "conv_kernel_size": 9, "decoder_dropout": 0.2},
        "denoiser": den,
        # config/*/model.yaml:40-46 (JCU discriminator, the training config)
        "discriminator": {"n_layer": 3, "n_uncond_layer": 2, "n_cond_layer": 2, "n_channels": [64, 128, 512, 128, 1],
                          "kernel_sizes": [3, 5, 5, 5, 3], "strides": [1, 2, 2, 1, 1]},
        "multi_speaker": False if multi_speaker is None else bool(multi_speaker),
        "max_seq_len": 1000 if dataset == "LJSpeech" else 1500,
    }
    return cfg


def preprocess_config(stats_dir: str) -> dict:
    return {
        "path": {"preprocessed_path": stats_dir},
        "preprocessing": {
            "mel": {"n_mel_channels": N_MEL},
            "audio": {"sampling_rate": 22050},
            "stft": {"hop_length": 256},
        },
    }


def train_config() -> dict:
    return {"loss": {"noise_loss": "l1"}}


def write_stats(stats_dir: str | None = None, spec_min: float = SPEC_MIN,
                spec_max: float = SPEC_MAX, n_mel: int = N_MEL) -> str:
    """Write a synthetic ``stats.json`` (schema: preprocessor/preprocessor.py:193-212)."""
    if stats_dir is None:
        stats_dir = tempfile.mkdtemp(prefix="mixgan_stats_")
    os.makedirs(stats_dir, exist_ok=True)
    with open(os.path.join(stats_dir, "stats.json"), "w") as f:
        json.dump({"spec_min": [spec_min] * n_mel, "spec_max": [spec_max] * n_mel}, f)
    return stats_dir


def make_configs(dataset: str = "LJSpeech", model: str = "naive",
                 multi_speaker: bool | None = None, stats_dir: str | None = None,
                 residual_layers: int | None = None, timesteps: int | None = None):
    """Return ``(args, preprocess_config, model_config, train_config)``."""
    mc = copy.deepcopy(model_config(dataset, multi_speaker))
    if residual_layers is not None:
        mc["denoiser"]["residual_layers"] = int(residual_layers)
    if timesteps is not None:
        mc["denoiser"]["timesteps" if model == "naive" else "shallow_timesteps"] = int(timesteps)
    stats_dir = write_stats(stats_dir)
    return SimpleNamespace(model=model), preprocess_config(stats_dir), mc, train_config()
