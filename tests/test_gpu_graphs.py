"""CUDA-graph replay of the loss (mamba_clip_b200.enable_cuda_graphs): same kernels, so the results must equal the
eager path bit for bit, for inputs that change from step to step, for losses that are never back-propagated, and when
forward is called again before the pending backward (falls back to eager for that call)."""
import pytest
import torch

from oracle import clip_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def graphs():
    import mamba_clip_b200 as M
    from mamba_clip_b200 import _function
    M.enable_cuda_graphs(True)
    yield _function
    M.enable_cuda_graphs(False)


def _step(crit, img, txt, ls_val, go=1.0):
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(ls_val, device="cuda", requires_grad=True)
    loss = crit(a, b, s, output_dict=False)
    loss.backward(torch.tensor(go, device="cuda"))
    return float(loss.detach()), a.grad.clone(), b.grad.clone(), float(s.grad)


@pytest.mark.parametrize("B,D,dtype", [(1024, 512, torch.bfloat16), (400, 200, torch.bfloat16), (512, 96, torch.float32)])
def test_graph_replay_equals_eager(graphs, B, D, dtype):
    import mamba_clip_b200 as M
    from mamba_clip_b200 import ClipLoss
    crit = ClipLoss()
    data = [O.make_features(B, D, seed=100 + k, correlated=(k % 2 == 0), dtype=dtype) + (10.0 + 3 * k, 1.0 + k) for k in range(6)]
    M.enable_cuda_graphs(False)
    eager = [_step(crit, i, t, ls, go) for i, t, ls, go in data]
    M.enable_cuda_graphs(True)
    graphed = [_step(crit, i, t, ls, go) for i, t, ls, go in data]      # calls 1-2 eager warm-up, 3 captures, 4-6 replay
    key = [k for k in graphs._graph_cache if k[1] == (B, D)]
    assert len(key) == 1 and len(graphs._graph_cache[key[0]].graphs) >= 2      # forward + backward captured
    for (l0, di0, dt0, ds0), (l1, di1, dt1, ds1) in zip(eager, graphed):
        assert l0 == l1 and ds0 == ds1
        assert torch.equal(di0, di1) and torch.equal(dt0, dt1)


def test_graph_buffers_released_when_loss_is_dropped_and_eager_fallback_when_pending(graphs):
    from mamba_clip_b200 import ClipLoss
    crit = ClipLoss()
    img, txt = O.make_features(512, 256, seed=5, dtype=torch.bfloat16)
    ref = O.closed_form(img.float(), txt.float(), 20.0, 1, 0, False, False)
    for _ in range(4):
        _step(crit, img, txt, 20.0)                                    # warm-up + capture + one replay
    a = img.cuda().requires_grad_(True)
    b = txt.cuda().requires_grad_(True)
    s = torch.tensor(20.0, device="cuda", requires_grad=True)
    l1 = crit(a, b, s, output_dict=False)                              # graphed, backward pending
    gl = next(iter(graphs._graph_cache.values()))
    assert gl.pending
    img2, txt2 = O.make_features(512, 256, seed=6, dtype=torch.bfloat16)
    a2 = img2.cuda().requires_grad_(True)
    l2 = crit(a2, txt2.cuda(), torch.tensor(20.0, device="cuda"), output_dict=False)   # must not touch l1's buffers
    l2.backward()
    l1.backward()
    assert abs(float(l1.detach()) - float(ref.loss)) <= 2e-3 * abs(float(ref.loss))
    assert O.rel_err(a.grad.cpu(), ref.d_image) <= 2e-3 and O.rel_err(b.grad.cpu(), ref.d_text) <= 2e-3
    assert not gl.pending
    l3 = crit(a, b, s, output_dict=False)                              # graphed again, then dropped without backward
    assert gl.pending
    del l3
    assert not gl.pending
    with torch.no_grad():
        l4 = crit(img.cuda(), txt.cuda(), torch.tensor(20.0, device="cuda"), output_dict=False)
    assert abs(float(l4) - float(ref.loss)) <= 2e-3 * abs(float(ref.loss))


def test_graph_replay_of_the_shared_recompute_backward(graphs, monkeypatch):
    """mclip_fused_grad forks / joins onto the library's side stream with events: that must capture into the backward graph
    and replay bit-identically.  (By default the sizes that take the shared-recompute backward, B >= 16384, lie above the
    graph work limit; the limit is raised for this test.)"""
    import mamba_clip_b200 as M
    from mamba_clip_b200 import ClipLoss, _cabi
    B, D = 16384, 512
    monkeypatch.setattr(graphs, "_GRAPH_MAX_WORK", 1 << 40)
    crit = ClipLoss()
    assert _cabi.get_backend().fused_supported(torch.empty(B, D, dtype=torch.bfloat16, device="cuda"),
                                               torch.empty(B, D, dtype=torch.bfloat16, device="cuda"))
    data = [O.make_features(B, D, seed=300 + k, dtype=torch.bfloat16) + (14.2857 + k, 1.0 + k) for k in range(5)]
    M.enable_cuda_graphs(False)
    eager = [_step(crit, i, t, ls, go) for i, t, ls, go in data]
    M.enable_cuda_graphs(True)
    graphed = [_step(crit, i, t, ls, go) for i, t, ls, go in data]
    key = [k for k in graphs._graph_cache if k[1] == (B, D)]
    assert len(key) == 1 and len(graphs._graph_cache[key[0]].graphs) >= 2
    for (l0, di0, dt0, ds0), (l1, di1, dt1, ds1) in zip(eager, graphed):
        assert l0 == l1 and ds0 == ds1
        assert torch.equal(di0, di1) and torch.equal(dt0, dt1)
