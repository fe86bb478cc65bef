"""Stand-in for the `lpips` package (not installed, no network): the reference imports it at module import time
(transvae/losses/vae_loss.py:8) although the L1 + KL terms this repository covers never call it.  Constructing the
network is allowed (TransVAELoss.__init__ does it unconditionally); calling it is not."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, net="vgg", **kwargs):
        super().__init__()
        self.net = net

    def forward(self, *a, **k):
        raise RuntimeError("lpips stub: the LPIPS term needs VGG weights that are unavailable offline")
