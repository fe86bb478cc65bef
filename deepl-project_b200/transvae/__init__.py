"""TransVAE, B200-native: drop-in for the reference package ``transvae`` (transvae/__init__.py:5-9).

Same public names -- ``TransVAE``, ``create_transvae``, ``TransVAELoss`` -- same constructor arguments,
``state_dict`` keys and return values; every FLOP runs in hand-written sm_100a kernels from
``libtransvae_sm100.so`` (include/transvae_sm100.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
__all__ = ["TransVAE", "create_transvae", "TransVAELoss"]


def __getattr__(name):
    if name in ("TransVAE", "create_transvae"):
        from .models.transvae import TransVAE, create_transvae
        return {"TransVAE": TransVAE, "create_transvae": create_transvae}[name]
    if name == "TransVAELoss":
        from .losses.vae_loss import TransVAELoss
        return TransVAELoss
    raise AttributeError(name)
