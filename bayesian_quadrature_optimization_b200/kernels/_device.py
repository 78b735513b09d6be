"""Shared plumbing of the kernel classes: every covariance, gradient and Hessian is evaluated by
libbqo_sm100.so through a temporary GPEngine on the second argument's points (no CPU path)."""
import numpy as np

from .. import _cabi
from ..engine import GPEngine


def hyper_batch(kind, dim, params, points, n_tasks=0, same_correlation=False, device=None):
    """One-hyper-sample batch whose 'training set' is `points` (rows of dimension `dim`)."""
    eng = GPEngine(kind, dim, n_tasks, same_correlation, device=device)
    points = np.ascontiguousarray(np.asarray(points, dtype=float).reshape(-1, dim))
    eng.set_data(points, np.zeros(points.shape[0]))
    return eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(params, dtype=float).reshape(-1)]))


def cross_cov(kind, dim, params, inputs_1, inputs_2, **kw):
    """k(inputs_1, inputs_2) -> np.array(n1 x n2)."""
    hb = hyper_batch(kind, dim, params, inputs_2, **kw)
    return hb.cross_cov(np.asarray(inputs_1, dtype=float).reshape(-1, dim)).cpu().numpy()[0]


def grad_params(kind, dim, params, inputs, **kw):
    """d K(inputs, inputs) / d params -> np.array(p x n x n)."""
    return hyper_batch(kind, dim, params, inputs, **kw).grad_cov().cpu().numpy()[0]


def grad_point(kind, dim, params, point, inputs, **kw):
    """d k(point, inputs_i) / d point -> np.array(n x dim_m)."""
    hb = hyper_batch(kind, dim, params, inputs, **kw)
    _, dKP = hb.cross_cov(np.asarray(point, dtype=float).reshape(1, dim), grad=True)
    return dKP.cpu().numpy()[0, 0].T


def hessian_point(kind, dim, params, point, inputs, **kw):
    """d2 k(point, inputs_i) / d point d point -> np.array(n x dim_m x dim_m)."""
    hb = hyper_batch(kind, dim, params, inputs, **kw)
    H = hb.cross_cov_hessian(np.asarray(point, dtype=float).reshape(1, dim))
    return H.cpu().numpy()[0, 0].transpose(2, 0, 1)


KIND_MATERN52 = _cabi.KIND_MATERN52
KIND_SCALED_MATERN52 = _cabi.KIND_SCALED_MATERN52
KIND_PRODUCT_TASKS = _cabi.KIND_PRODUCT_TASKS


def simple_dictionary(dictionary, order_keys):
    """sbo/lib/util.py:31-58 convert_dictionary_gradient_to_simple_dictionary: flatten
    {name: {i: nxn} or nxn} to {int: nxn} following `order_keys` = [(name, [(i, None)...] or None)]."""
    result = []
    for key in order_keys:
        if key[1] is None:
            result.append(np.array(dictionary[key[0]]))
        else:
            for sub_key in key[1]:
                result.append(np.array(dictionary[key[0]][sub_key[0]]))
    return {index: element for index, element in enumerate(result)}
