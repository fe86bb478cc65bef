from .decoder import TransVAEDecoder
from .encoder import TransVAEEncoder
from .transvae import TransVAE, create_transvae

__all__ = ["TransVAE", "create_transvae", "TransVAEEncoder", "TransVAEDecoder"]
