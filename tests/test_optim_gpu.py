"""FusedAdamW (grb_adamw_step) against torch.optim.AdamW — the optimizer the reference configures
(configs/model/hstu.yaml, generative_recommenders.py:254-322).  fp32; tolerance: the two differ only
in rounding order (decay folded into one multiply, reciprocal of sqrt(bias_correction2)), 2e-6 rel."""
import pytest
import torch

from mygenerativerecommenders_b200.optim import FusedAdamW

RTOL, ATOL = 2e-6, 1e-8


def _params(device, sizes, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for s in sizes:
        if isinstance(s, tuple) and s[0] == "offset":      # 4-byte aligned only: scalar path
            base = torch.randn(s[1] + 1, generator=g).to(device)
            out.append(torch.nn.Parameter(base[1:]))
        else:
            out.append(torch.nn.Parameter(torch.randn(s, generator=g).to(device)))
    return out


@pytest.mark.gpu
def test_fused_adamw_matches_torch_adamw():
    dev = torch.device("cuda")
    sizes = [(1,), (7,), (8192,), (8193,), (3, 5), (1000, 256), ("offset", 20_001), (0,), (64, 1024)]
    a, b = _params(dev, sizes), _params(dev, sizes)
    kw = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-3)
    ours, ref = FusedAdamW(a, **kw), torch.optim.AdamW(b, foreach=False, fused=False, **kw)
    g = torch.Generator(device="cpu").manual_seed(1)
    for step in range(6):
        for pa, pb in zip(a, b):
            if step == 3 and pa.numel() == 7:
                pa.grad = pb.grad = None       # a parameter without a gradient is skipped, its step too
                continue
            gr = torch.randn(pa.shape, generator=g).to(dev) * (10.0 ** (step - 3))
            pa.grad, pb.grad = gr.clone(), gr.clone()
        if step == 4:
            for grp in ours.param_groups + ref.param_groups:
                grp["lr"] = 5e-4                # schedulers change lr in place
        ours.step()
        ref.step()
        for i, (pa, pb) in enumerate(zip(a, b)):
            torch.testing.assert_close(pa, pb, rtol=RTOL, atol=ATOL, msg=lambda m: f"step {step} param {i}: {m}")
    for pa, pb in zip(a, b):
        if pa.numel() == 0:
            continue
        sa, sb = ours.state[pa], ref.state[pb]
        torch.testing.assert_close(sa["exp_avg"], sb["exp_avg"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=RTOL, atol=1e-12)
        assert float(sa["step"]) == float(sb["step"])


@pytest.mark.gpu
def test_fused_adamw_state_dict_continues_in_torch_adamw():
    dev = torch.device("cuda")
    a, b = _params(dev, [(300, 256), (256,)]), _params(dev, [(300, 256), (256,)])
    kw = dict(lr=1e-3, betas=(0.9, 0.98), weight_decay=1e-3)
    ours, ref = FusedAdamW(a, **kw), torch.optim.AdamW(b, **kw)
    g = torch.Generator(device="cpu").manual_seed(2)

    def grads():
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).to(dev)
            pa.grad, pb.grad = gr.clone(), gr.clone()

    for _ in range(3):
        grads()
        ours.step()
        ref.step()
    cont = torch.optim.AdamW(a, **kw)
    cont.load_state_dict(ours.state_dict())      # same keys as torch.optim.AdamW
    for _ in range(2):
        grads()
        cont.step()
        ref.step()
    for pa, pb in zip(a, b):
        torch.testing.assert_close(pa, pb, rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
def test_fused_adamw_more_than_one_launch_worth_of_tensors():
    dev = torch.device("cuda")
    sizes = [(17 + i,) for i in range(150)]       # > 64 tensors: several launches
    a, b = _params(dev, sizes), _params(dev, sizes)
    ours, ref = FusedAdamW(a, lr=1e-2), torch.optim.AdamW(b, lr=1e-2)
    for pa, pb in zip(a, b):
        pa.grad = torch.ones_like(pa) * 0.5
        pb.grad = pa.grad.clone()
    ours.step()
    ref.step()
    for pa, pb in zip(a, b):
        torch.testing.assert_close(pa, pb, rtol=RTOL, atol=ATOL)


def test_fused_adamw_refuses_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = FusedAdamW([p])
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    with pytest.raises(NotImplementedError):
        FusedAdamW([p], amsgrad=True)
