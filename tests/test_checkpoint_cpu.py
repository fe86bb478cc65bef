"""Checkpoint / optimizer-state interchange with the reference's trainer and the host logic of the training driver
(SURVEY 8f ranks 1-2), CPU only: the schema is what train.py:753-769 / train_2.py:245-260 write
({epoch, global_step, model_state_dict, optimizer_state_dict = torch.optim.AdamW.state_dict(), ...})."""
import importlib.util
import os

import pytest
import torch

from transvae.trainer import FusedAdamW, GradBuckets, Trainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mlp(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(8, 32), torch.nn.Tanh(), torch.nn.Linear(32, 5))


def _driver():
    spec = importlib.util.spec_from_file_location("tvae_train_driver", os.path.join(ROOT, "deepl-project_b200", "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_optimizer_state_loads_into_torch_adamw_and_back():
    m = _mlp()
    opt = FusedAdamW(GradBuckets(m.parameters()), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    opt.m.copy_(torch.randn(opt.m.shape, generator=g))
    opt.v.copy_(torch.rand(opt.v.shape, generator=g))
    opt.step_count = 7
    sd = opt.state_dict()
    # (a) the reference's optimiser accepts it
    ref = torch.optim.AdamW(_mlp().parameters(), lr=1.0, betas=(0.5, 0.5))
    ref.load_state_dict(sd)
    rsd = ref.state_dict()
    assert rsd["param_groups"][0]["lr"] == 3e-4 and tuple(rsd["param_groups"][0]["betas"]) == (0.9, 0.95)
    params = list(m.parameters())
    for i, p in enumerate(params):
        assert tuple(rsd["state"][i]["exp_avg"].shape) == tuple(p.shape)
        assert float(rsd["state"][i]["step"]) == 7.0
    # (b) and its state_dict loads back bit-exactly into a fresh fused optimiser
    opt2 = FusedAdamW(GradBuckets(_mlp(1).parameters()))
    opt2.load_state_dict(rsd)
    for pa, pb in zip(opt.b.params, opt2.b.params):          # (alignment padding between tensors is not state)
        (oa, n), (ob, _) = opt.b._slices[pa], opt2.b._slices[pb]
        assert torch.equal(opt2.m[ob:ob + n], opt.m[oa:oa + n]) and torch.equal(opt2.v[ob:ob + n], opt.v[oa:oa + n])
    assert opt2.step_count == 7
    assert opt2.lr == 3e-4 and opt2.betas == (0.9, 0.95) and opt2.wd == 0.01


def test_reference_trained_optimizer_state_is_read_in_parameter_order():
    """Moments produced by torch.optim.AdamW steps land in the right flat-buffer slices (the flat layout is in reverse
    registration order, the state dict in registration order)."""
    ref_m = _mlp()
    ref = torch.optim.AdamW(ref_m.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(2))
    for _ in range(3):
        ref.zero_grad()
        ref_m(x).pow(2).sum().backward()
        ref.step()
    m = _mlp()
    m.load_state_dict(ref_m.state_dict())
    gb = GradBuckets(m.parameters())
    opt = FusedAdamW(gb)
    opt.load_state_dict(ref.state_dict())
    assert opt.step_count == 3
    for i, p in enumerate(gb.params):
        o, n = gb._slices[p]
        assert torch.equal(opt.m[o:o + n].view(p.shape), ref.state_dict()["state"][i]["exp_avg"])
        assert torch.equal(p.detach(), list(ref_m.parameters())[i].detach())
    bad = ref.state_dict()
    bad["state"][0]["step"] = torch.tensor(5.0)
    with pytest.raises(ValueError):
        opt.load_state_dict(bad)


def test_trainer_checkpoint_schema_round_trip(tmp_path):
    m = _mlp()
    tr = Trainer(m, loss_fn=None, lr=1e-4, warmup_steps=10)
    tr.opt.step_count = 4
    tr.opt.m.fill_(0.25)
    path = str(tmp_path / "ck.pth")
    tr.save(path, epoch=2, args={"variant": "tiny"})
    ck = torch.load(path, weights_only=False)
    assert set(ck) >= {"epoch", "global_step", "model_state_dict", "optimizer_state_dict", "args"}      # train.py:759-765
    assert ck["global_step"] == 4 and ck["epoch"] == 2 and list(ck["model_state_dict"]) == list(m.state_dict())
    m2 = _mlp(3)
    tr2 = Trainer(m2, loss_fn=None, lr=5e-5, warmup_steps=10)
    flat_ptr = tr2.buckets.flat_p.data_ptr()
    tr2.load(path)
    assert tr2.opt.step_count == 4
    for p2 in tr2.buckets.params:
        o, n = tr2.buckets._slices[p2]
        assert bool((tr2.opt.m[o:o + n] == 0.25).all()) and bool((tr2.opt.v[o:o + n] == 0).all())
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    # parameters still live inside the flat buffer after loading
    lo, hi = flat_ptr, flat_ptr + tr2.buckets.flat_p.numel() * 4
    assert all(lo <= p.data_ptr() < hi for p in m2.parameters())
    assert abs(tr2.opt.current_lr() - 1e-4 * 4 / 10) < 1e-12          # base rate comes from the checkpoint, warm-up k / W


def test_checkpoint_resumes_in_the_reference_trainer():
    """train_2.py:480-490 resumes with optimizer.load_state_dict(...) and scheduler.load_state_dict(...) on a
    torch.optim.AdamW + LambdaLR pair: our checkpoint must survive exactly that (LambdaLR.load_state_dict pops
    'lr_lambdas'), and the resumed pair must continue the schedule where we stopped."""
    m = _mlp()
    tr = Trainer(m, loss_fn=None, lr=1e-4, warmup_steps=10)
    tr.opt.step_count = 4
    tr.opt.m.fill_(0.25)
    path = str(tmp_path_factory_dir() / "ck_ref.pth")
    tr.save(path, epoch=1, args={})
    ck = torch.load(path, weights_only=False)
    m2 = _mlp(7)
    m2.load_state_dict(ck["model_state_dict"])
    opt = torch.optim.AdamW(m2.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: float(s) / 10 if s < 10 else 1.0)     # train_2.py:266-274
    opt.load_state_dict(ck["optimizer_state_dict"])
    if ck.get("scheduler_state_dict") is not None:
        sch.load_state_dict(ck["scheduler_state_dict"])
    g0 = opt.param_groups[0]
    assert sch.last_epoch == 4 and g0["initial_lr"] == 1e-4 and abs(g0["lr"] - 4e-5) < 1e-15
    for p in m2.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    sch.step()
    assert abs(opt.param_groups[0]["lr"] - 5e-5) < 1e-15 and float(opt.state_dict()["state"][0]["step"]) == 5.0
    # and back: the reference pair's checkpoint resumes here with the base rate, not the scheduled one
    back = {"epoch": 1, "global_step": 5, "model_state_dict": m2.state_dict(), "optimizer_state_dict": opt.state_dict(),
            "scheduler_state_dict": sch.state_dict(), "args": {}}
    tr3 = Trainer(_mlp(9), loss_fn=None, lr=9.0, warmup_steps=10)
    tr3.load_state_dict(back)
    assert tr3.opt.lr == 1e-4 and tr3.opt.step_count == 5 and abs(tr3.opt.current_lr() - 5e-5) < 1e-15


def tmp_path_factory_dir():
    import pathlib
    import tempfile
    return pathlib.Path(tempfile.mkdtemp(prefix="tvae_ck_"))


def test_driver_schedule_and_args():
    d = _driver()
    assert d.lr_at(0, 1e-4, 1000) == 0.0 and d.lr_at(500, 1e-4, 1000) == 5e-5 and d.lr_at(1000, 1e-4, 1000) == 1e-4
    assert d.lr_at(3, 1e-4, 0) == 1e-4
    # same values as the reference's LambdaLR (train_2.py:266-274)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: s / 8 if s < 8 else 1.0)
    for k in range(12):
        assert abs(opt.param_groups[0]["lr"] - d.lr_at(k, 1e-4, 8)) < 1e-15
        opt.step()
        sch.step()
    a = d.parse_args(["--output_dir", "x", "--variant", "tiny", "--batch_size", "4", "--accumulation_steps", "2"])
    assert a.learning_rate == 1e-4 and a.kl_weight == 1e-8 and a.grad_clip == 1.0 and a.lpips_weight == 0.0
    with pytest.raises(NotImplementedError):
        d.main(["--output_dir", "x", "--lpips_weight", "1.0"])
