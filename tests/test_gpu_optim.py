"""GPU: optim.FlatAdam (one kernel over flat buffers) against torch.optim.Adam on the same parameters and gradients."""
import copy

import pytest
import torch

from golden_util import relmax

gpu = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(80, 5, 3, 3), (80,), (2, 1, 5, 5), (2,), (1, 2, 1, 1), (33, 7), (3,)]
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.3).cuda()) for s in shapes]


@gpu
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    from aerial_depth_completion_b200.optim import FlatAdam
    pa = _params(1)
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ref = torch.optim.Adam(pa, lr=3e-3, weight_decay=wd)
    opt = FlatAdam(pb, lr=3e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for it in range(6):
        ref.zero_grad()
        opt.zero_grad()
        for a, b in zip(pa, pb):
            gr = (torch.randn(a.shape, generator=g) * (10.0 ** (it - 3))).cuda()
            a.grad = gr.clone()
            b.grad.copy_(gr) if it % 2 == 0 else setattr(b, "grad", gr.clone())     # in-bucket and assigned-from-outside gradients
        ref.step()
        opt.step()
        for a, b in zip(pa, pb):
            assert relmax(b.detach().cpu(), a.detach().cpu()) < 2e-6, it
    for a, b in zip(pa, pb):
        assert relmax(opt.state[b]["exp_avg"].cpu(), ref.state[a]["exp_avg"].cpu()) < 2e-6
        assert relmax(opt.state[b]["exp_avg_sq"].cpu(), ref.state[a]["exp_avg_sq"].cpu()) < 2e-6
    assert float(opt.step_t) == 6.0
    assert all(b.data_ptr() >= opt.flat_param.data_ptr() for b in pb)


@gpu
def test_flat_adam_state_dict_round_trip_and_version_bump():
    from aerial_depth_completion_b200.optim import FlatAdam
    pa = _params(3)
    opt = FlatAdam(pa, lr=1e-2)
    for p in pa:
        p.grad.fill_(0.5)
    v0 = pa[0]._version
    opt.step()
    assert pa[0]._version > v0
    sd = copy.deepcopy(opt.state_dict())
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    opt2 = FlatAdam(pb, lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 1e-2 and float(opt2.step_t) == 1.0
    for p in pa + pb:
        p.grad.fill_(-0.25)
    opt.step()
    opt2.step()
    for a, b in zip(pa, pb):
        assert torch.equal(a.detach(), b.detach())
    # a torch.optim.Adam state dict loads too
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ta = torch.optim.Adam(pc, lr=5e-3)
    for p in pc:
        p.grad = torch.full_like(p, 0.1)
    ta.step()
    opt3 = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in pc], lr=1.0)
    opt3.load_state_dict(ta.state_dict())
    assert float(opt3.step_t) == 1.0 and opt3.param_groups[0]["lr"] == 5e-3


def test_flat_adam_refuses_cpu_parameters():
    from aerial_depth_completion_b200.optim import FlatAdam
    with pytest.raises(RuntimeError):
        FlatAdam([torch.nn.Parameter(torch.zeros(3))])
