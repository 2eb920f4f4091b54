"""mae_b200.graph.capture_step: a whole MAE.loss + backward (+ clipping + optimizer) replayed as one CUDA graph must
reproduce the eager step (the reference's training loop body, experiments/color_image.py:153-165) on new batches."""
import copy

import pytest
import torch

from tolerance import check_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", ["binary_image_json", "color_resnet"])
def test_captured_step_matches_eager(golden, name, monkeypatch):
    from mae_b200 import graph, synth
    from mae_b200.modules import MAE, GaussianEncoder
    c = golden("models")[name]
    model = synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(DEV)
    model.train()
    x0 = c["x"].to(DEV)
    eps = c["eps_loss"].to(DEV)
    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=eps: e))      # same noise in every step
    args = dict(eta=c["eta"], gamma=c["gamma"], free_bits=c["free_bits"])
    step = graph.capture_step(model, x0, c["ns"], max_norm=None, **args)
    assert step.launches_of_ours >= 2
    g = torch.Generator().manual_seed(5)
    for trial in range(2):
        x = x0 if trial == 0 else (torch.rand(x0.shape, generator=g) < 0.3).float().to(DEV) if x0.size(1) == 1 else \
            (torch.randint(0, 256, x0.shape, generator=g).float() / 127.5 - 1.0).to(DEV)
        out = step(x)
        got = [o.clone() if isinstance(o, torch.Tensor) else o for o in out]
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        ref = model.loss(x, c["ns"], **args)
        ref[0].backward()
        for i, (a, b) in enumerate(zip(got, ref)):
            if isinstance(b, torch.Tensor):
                check_close(a, b.detach().double(), rel=2e-6, what="%s output %d (trial %d)" % (name, i, trial))
            else:
                assert a == b
        gscale = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
        for n, p in model.named_parameters():
            if p.grad is not None:
                # fp32 conv stacks: cuDNN may pick other algorithms / summation orders under capture (measured up to 1.02e-5)
                check_close(grads[n], p.grad.double(), rel=5e-5, floor=1e-6 * gscale, what="%s grad %s (trial %d)" % (name, n, trial))
        model.zero_grad(set_to_none=True)


def test_captured_step_with_clipping_and_sgd(golden, monkeypatch):
    """the optimizer inside the graph: two replays == two eager steps"""
    from mae_b200 import graph, synth
    from mae_b200.modules import MAE, GaussianEncoder
    c = golden("models")["color_resnet"]
    x = c["x"].to(DEV)
    eps = c["eps_loss"].to(DEV)
    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=eps: e))
    args = dict(eta=c["eta"], gamma=c["gamma"], free_bits=c["free_bits"])
    models = [synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(DEV).train() for _ in range(2)]
    opts = [torch.optim.SGD(m.parameters(), lr=1e-3) for m in models]
    step = graph.capture_step(models[0], x, c["ns"], optimizer=opts[0], max_norm=5.0, **args)
    for _ in range(2):
        step(x)
        opts[1].zero_grad()
        models[1].loss(x, c["ns"], **args)[0].backward()
        torch.nn.utils.clip_grad_norm_(models[1].parameters(), 5.0)
        opts[1].step()
    for (n, p), q in zip(models[0].named_parameters(), models[1].parameters()):
        check_close(p, q.detach().double(), rel=5e-5, what="parameter %s after two steps" % n)
