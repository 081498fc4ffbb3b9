"""vibravox_b200: B200-native EBEN bandwidth-extension training step (drop-in for that one path
of jhauret/vibravox).  Importing the package does not load the CUDA library; the first op does,
and raises if libvbx_b200.so is missing (there is no CPU fallback)."""

__all__ = ["build_model"]


def build_model(m: int = 4, n: int = 32, p: int = 2, q: int = 4, min_channels: int = 24, seed: int = 42,
                lr: float = 3e-4, betas=(0.5, 0.9), dynamic_loss_balancing="ema", beta_ema: float = 0.9,
                device="cuda", sample_rate: int = 16000):
    """The reference's default EBEN training configuration (configs/lightning_module/eben.yaml and
    the YAML it composes) assembled from the drop-in modules, on `device`."""
    from functools import partial

    import torch

    from .lightning_modules.eben import EBENLightningModule
    from .optim import FlatAdam
    from .torch_modules.dnn.eben_discriminator import DiscriminatorEBENMultiScales
    from .torch_modules.dnn.eben_generator import EBENGenerator
    from .torch_modules.losses.feature_loss import FeatureLossForDiscriminatorMelganMultiScales
    from .torch_modules.losses.hinge_loss import HingeLossForDiscriminatorMelganMultiScales
    from .torch_modules.losses.mrstft_loss import MultiResolutionSTFTLoss

    torch.manual_seed(seed)
    generator = EBENGenerator(m=m, n=n, p=p)
    discriminator = DiscriminatorEBENMultiScales(q=q, min_channels=min_channels)
    stft = MultiResolutionSTFTLoss(fft_sizes=(512, 1024, 2048), hop_sizes=(50, 120, 240),
                                   win_lengths=(240, 600, 1200), sample_rate=sample_rate,
                                   perceptual_weighting=True)
    generator, discriminator, stft = generator.to(device), discriminator.to(device), stft.to(device)
    adam = partial(FlatAdam, lr=lr, betas=betas)
    return EBENLightningModule(sample_rate=sample_rate, generator=generator, discriminator=discriminator,
                               generator_optimizer=adam, discriminator_optimizer=adam,
                               reconstructive_loss_freq_fn=stft,
                               feature_matching_loss_fn=FeatureLossForDiscriminatorMelganMultiScales(),
                               adversarial_loss_fn=HingeLossForDiscriminatorMelganMultiScales(),
                               dynamic_loss_balancing=dynamic_loss_balancing, beta_ema=beta_ema,
                               update_discriminator_ratio=1.0, description="vibravox_b200")
