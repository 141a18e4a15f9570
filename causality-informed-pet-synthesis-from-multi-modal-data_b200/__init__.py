"""petsyn_b200 -- B200-native (sm_100a) drop-in replacements for the 3D T1->PET generator hot path of
jessyblues/Causality-Informed-PET-Synthesis-from-Multi-modal-Data.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute goes through the
C ABI of ``lib/libpetsyn.so`` (``include/petsyn.h``).  No Triton, no backend dispatch, no CPU fallback.
"""
from . import _cabi  # noqa: F401  (fails loudly when the CUDA library has not been built)
from . import ops  # noqa: F401
from .unet_model import UnetGenerator3d, UnetSkipConnectionBlock3d  # noqa: F401
from .atten_unet_model import AttenUNet, DiffusionModelEncoder  # noqa: F401
from .data import PairVolumeLoader, SyntheticPairSource, volume_prepare  # noqa: F401
from .causal_model import Decoder, DiffusionModelDecoder, kl_divergence, reparameterize  # noqa: F401
from .bmgan_model import PatchDiscriminator, ResNet_encoder, dense_unet_generator, patch_discriminator  # noqa: F401

__all__ = ["AttenUNet", "DiffusionModelEncoder", "UnetGenerator3d", "UnetSkipConnectionBlock3d", "dense_unet_generator", "patch_discriminator", "PatchDiscriminator", "Decoder", "DiffusionModelDecoder", "kl_divergence", "reparameterize", "ResNet_encoder", "ops",
           "PairVolumeLoader", "SyntheticPairSource", "volume_prepare"]
