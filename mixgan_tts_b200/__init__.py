"""B200-native reverse-diffusion decoder of MixGAN-TTS (drop-in ``Denoiser`` / ``GaussianDiffusion``).

Only the hot path lives here: the sm_100a CUDA library (``csrc/`` -> ``libmixgan_b200.so``, C ABI in
``include/mixgan_b200.h``) and the torch modules that mirror the reference's interface for it.
"""
from .modules import Denoiser  # noqa: F401
from .diffusion import GaussianDiffusion  # noqa: F401
from .discriminator import JCUDiscriminator  # noqa: F401
from .length_regulator import LengthRegulator  # noqa: F401
from .aux_decoder import AuxDecoder  # noqa: F401
from .vocoder import Generator  # noqa: F401

__all__ = ["Denoiser", "GaussianDiffusion", "JCUDiscriminator", "LengthRegulator", "AuxDecoder", "Generator"]
