"""Multi-start first-order ascent on the device: every restart advances in the same launch.

The update is the reference's Adam rule with box projection and the relative-step stop test
(sbo/lib/stochastic_gradient_descent.py:83-130) applied to all restarts at once through
`bqo_adam_step_batched`; the objective and its gradient come from a caller-supplied closure that
evaluates all restarts in one batched call.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi


def batched_adam_ascent(eng, value_and_grad, start, lower, upper, maxepoch=10, sign=1.0,
                        learning_rate=0.1):
    """Maximise (sign = +1) or minimise (sign = -1) f from every row of `start`.

    :param eng: GPEngine (device, library handle)
    :param value_and_grad: callable(points: torch.Tensor[R, dim]) -> (values[R], grad[R, dim])
    :param start: np.array(R, dim)
    :param lower, upper: np.array(dim) box (use +-inf for free coordinates)
    :return: (points np.array(R, dim), values np.array(R)) at the end points
    """
    pts = eng.to_device(np.asarray(start, dtype=float).copy())
    R, dim = int(pts.shape[0]), int(pts.shape[1])
    m0, v0 = torch.zeros_like(pts), torch.zeros_like(pts)
    active = torch.ones(R, dtype=torch.int32, device=eng.device)
    lo_d = eng.to_device(np.asarray(lower, dtype=float))
    up_d = eng.to_device(np.asarray(upper, dtype=float))
    for t in range(1, int(maxepoch) + 1):
        _, grad = value_and_grad(pts)
        g = (-float(sign)) * grad  # the Adam kernel descends
        g = torch.where(torch.isnan(g), torch.zeros_like(g), g).contiguous()
        _cabi.check(eng.lib.bqo_adam_step_batched(
            _cabi.ptr(pts), _cabi.ptr(g), _cabi.ptr(m0), _cabi.ptr(v0), _cabi.ptr(active),
            _cabi.ptr(lo_d), _cabi.ptr(up_d), R, dim, t, float(learning_rate), 0.9, 0.999, 1e-8,
            torch.cuda.current_stream().cuda_stream), "bqo_adam_step_batched")
        if int(active.sum().item()) == 0:
            break
    values, _ = value_and_grad(pts)
    return pts.cpu().numpy(), values.cpu().numpy()
