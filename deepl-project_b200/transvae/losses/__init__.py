from .vae_loss import TransVAELoss

__all__ = ["TransVAELoss"]
