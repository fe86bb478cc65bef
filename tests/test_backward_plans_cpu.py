"""Host logic of the backward pass: transposed tap tables (dgrad) and the weight-gradient addressing reproduce
torch.autograd through the reference ops (CPU, fp32, addressing emulator)."""
import torch
import torch.nn.functional as F

from emu import emulate, emulate_wgrad
from transvae import _taps as T


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(a, b, tol=5e-5):
    err = float((a - b).abs().max() / b.abs().max())
    assert err < tol, err


def grads(fn, inputs, dout):
    inputs = [t.clone().requires_grad_(True) for t in inputs]
    fn(*inputs).backward(dout)
    return [t.grad for t in inputs]


def test_conv3x3_dgrad_and_wgrad():
    C, N = 64, 128
    x, w, dz = rnd(2, C, 6, 10), rnd(N, C, 3, 3, seed=1), rnd(2, N, 6, 10, seed=2)
    gx, gw = grads(lambda x, w: F.conv2d(x, w, padding=1), [x, w], dz)
    dx = emulate(T.plan_conv3x3_dgrad(N), nhwc(dz), None, T.pack_conv3x3_dgrad(w), (2, 6, 10, C))
    close(nchw(dx), gx)
    dwp = emulate_wgrad(T.plan_conv3x3(C), nhwc(x), None, nhwc(dz), N)
    close(dwp, T.pack_conv3x3(gw))


def test_linear_dgrad_and_wgrad():
    x, w, dz = rnd(50, 128), rnd(192, 128, seed=1), rnd(50, 192, seed=2)
    gx, gw = grads(lambda x, w: x @ w.t(), [x, w], dz)
    dx = emulate(T.plan_linear(192), dz.reshape(1, 1, 50, 192), None, w.t().contiguous(), (1, 1, 50, 128))
    close(dx.reshape(50, 128), gx)
    close(emulate_wgrad(T.plan_linear(128), x.reshape(1, 1, 50, 128), None, dz.reshape(1, 1, 50, 192), 192), gw)


def test_downsample_backward():
    C, N = 64, 128
    x, y = rnd(2, C, 8, 12), rnd(2, C, 8, 12, seed=5)
    w2, wdc, dz = rnd(N, C, 3, 3, seed=1), rnd(N, 4 * C, 1, 1, seed=3), rnd(2, N, 4, 6, seed=7)
    gy, gx, gw2, gwdc = grads(lambda y, x, w2, wdc: F.conv2d(y, w2, stride=2, padding=1) +
                              F.conv2d(F.pixel_unshuffle(x, 2), wdc), [y, x, w2, wdc], dz)
    dy = emulate(T.plan_downsample_dgrad_main(C, N), nhwc(dz), None, T.pack_conv3x3_dgrad(w2), (2, 8, 12, C))
    close(nchw(dy), gy)
    dx = emulate(T.plan_downsample_dgrad_dc(C, N), nhwc(dz), None, T.pack_downsample_dgrad_dc(wdc), (2, 8, 12, C))
    close(nchw(dx), gx)
    dwp = emulate_wgrad(T.plan_downsample(C), nhwc(y), nhwc(x), nhwc(dz), N)
    close(dwp, T.pack_downsample(gw2, gwdc))


def test_upsample_backward():
    Ci, Co = 128, 64
    x = rnd(2, Ci, 5, 6)
    w1, w2, wdc = rnd(Co, Ci, 3, 3, seed=1), rnd(Co, Co, 3, 3, seed=2), rnd(4 * Co, Ci, 1, 1, seed=3)
    dz1 = rnd(2, Co, 10, 12, seed=4)
    # first conv (nearest2x + 3x3)
    gx, gw1 = grads(lambda x, w: F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1), [x, w1], dz1)
    dx = emulate(T.plan_upsample_conv1_dgrad(Ci, Co), nhwc(dz1), None, T.pack_upsample_conv1_dgrad(w1), (2, 5, 6, Ci))
    close(nchw(dx), gx)
    dwp = emulate_wgrad(T.plan_upsample_conv1(Ci, Co), nhwc(x), None, nhwc(dz1), Co)
    # the packed weight is a linear function of w1 (taps summed): chain rule through the packing
    w1r = w1.clone().requires_grad_(True)
    (T.pack_upsample_conv1(w1r) * dwp).sum().backward()
    close(w1r.grad, gw1)
    # second conv + DC path
    y = rnd(2, Co, 10, 12, seed=6)
    dz2 = rnd(2, Co, 10, 12, seed=8)
    gy, gx2, gw2, gwdc = grads(lambda y, x, w, wdc: F.conv2d(y, w, padding=1) + F.pixel_shuffle(F.conv2d(x, wdc), 2),
                               [y, x, w2, wdc], dz2)
    dy = emulate(T.plan_conv3x3_dgrad(Co), nhwc(dz2), None, T.pack_conv3x3_dgrad(w2), (2, 10, 12, Co))
    close(nchw(dy), gy)
    dx2 = emulate(T.plan_upsample_dc_dgrad(Co), nhwc(dz2), None, T.pack_upsample_dc_dgrad(wdc), (2, 5, 6, Ci))
    close(nchw(dx2), gx2)
    dwp2 = emulate_wgrad(T.plan_upsample_conv2(Co, Ci), nhwc(y), nhwc(x), nhwc(dz2), Co)
    close(dwp2, T.pack_upsample_conv2(gw2, gwdc))
