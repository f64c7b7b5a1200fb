"""PipelinedFitLoop (guassianhand_b200/dist.py): the host-side schedule on CPU, results against the eager step on GPU."""
import numpy as np
import pytest

import util


class _Stub:
    def __init__(self, slots):
        self.slots, self.trace, self.ev = slots, False, []

    def _timed(self, what, fn, *a):
        return fn(*a)

    def upload(self, i, a, c):
        self.ev.append(("U", i))

    def launch(self, i):
        self.ev.append(("R", i))

    def result(self, i):
        self.ev.append(("C", i))
        return i


@pytest.mark.parametrize("slots", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 7])
def test_schedule_uploads_ahead_reads_behind_and_never_reuses_a_busy_slot(slots, n):
    from guassianhand_b200.dist import PipelinedFitLoop
    st = _Stub(slots)
    out = list(PipelinedFitLoop.run(st, ((None, None) for _ in range(n))))
    assert out == list(range(n))
    pos = {e: k for k, e in enumerate(st.ev)}
    assert len(pos) == 3 * n
    for i in range(n):
        assert pos[("U", i)] < pos[("R", i)] < pos[("C", i)]
        if i >= slots:          # the slot's previous step: read by the host before its buffers are reused,
            assert pos[("C", i - slots)] < pos[("R", i)]      # ... its replay enqueued before the new upload
            assert pos[("R", i - slots)] < pos[("U", i)]
        if slots >= 2 and i + 1 < n:
            assert pos[("U", i + 1)] < pos[("R", i)]          # the next step's inputs go up before this one launches
        if i >= slots - 1 and slots >= 2 and i + 1 < n:
            assert pos[("C", i - (slots - 1))] < pos[("R", i + 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [1, 3])
def test_pipelined_loop_matches_eager_steps(cuda_device, slots):
    import torch
    from guassianhand_b200 import scenes
    from guassianhand_b200.dist import PackedGrads, PipelinedFitLoop, fit_step_grads
    dev = cuda_device
    P, H, W, V = 3001, 80, 96, 4
    sc = scenes.two_hand_scene(P, seed=2)
    cams = scenes.fibonacci_cameras(3 * V, H, W, seed=2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
                 colors_precomp=t(sc.colors))
    bg = np.array([0.1, 0.0, 0.2], np.float32)
    groups = [util.gpu_views(cams[k * V:(k + 1) * V], bg, dev) for k in range(3)]
    dL = t((np.random.default_rng(3).normal(size=(V, 3, H, W)) / (H * W)).astype(np.float32))
    ref = PackedGrads(P, 0, device=dev)
    r0 = fit_step_grads(gauss, groups[0], dL, ref)
    cap = int(r0.R * 1.5) + (1 << 14)
    loop = PipelinedFitLoop(gauss, groups[0], dL, cap, overlap=2, slots=slots)
    # 7 steps: the attributes change every step (colours scaled, means shifted) and the cameras rotate
    steps = []
    for k in range(7):
        g = dict(gauss)
        g["colors_precomp"] = gauss["colors_precomp"] * (1.0 - 0.05 * k)
        g["means3D"] = gauss["means3D"] + 0.002 * k
        steps.append((g, groups[k % 3]))
    packed = [(loop.pack_attributes(g), loop.pack_cameras(v)) for g, v in steps]
    got = [(res.grads.clone(), res.loss, res.R) for res in loop.run(packed)]
    assert len(got) == 7
    for (g, v), (grads, loss, R) in zip(steps, got):
        want = PackedGrads(P, 0, device=dev)
        r = fit_step_grads(g, v, dL, want)
        torch.cuda.synchronize()
        assert R == r.R
        assert util.rel_err(grads.numpy(), want.flat.cpu().numpy()) <= 1e-5
        assert abs(loss - float(torch.vdot(r.color.reshape(-1), dL.reshape(-1)))) <= 1e-5 * max(1.0, abs(loss))
    assert loop.h2d_bytes_per_step == (P * 14 + V * (16 + 16 + 3 + 2)) * 4
