"""The kernels/ API (sbo/kernels/abstract_kernel.py:6-108 on Matern52, ScaledKernel, TasksKernel,
ProductKernels) on the CUDA path: the reference's own closed-form known answers
(tests/kernels/test_matern52.py:60-94, tests/kernels/test_tasks_kernel.py:52-71,128-131,
tests/models/test_gp_fitting_gaussian.py:680-684) and the matrices of the reference fixtures.
The reference asserts `==` on numpy doubles; the device's exp is accurate to 2 ulp, so 1e-13."""
import numpy as np
import pytest

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _kernels():
    from bayesian_quadrature_optimization_b200 import kernels as K
    from bayesian_quadrature_optimization_b200.entities.parameter import ParameterEntity
    return K, ParameterEntity


def test_scaled_matern_known_answers():
    K, PE = _kernels()
    m = K.ScaledKernel(2, K.Matern52(2, PE('scale', np.array([2.0, 3.0]), None)),
                       PE('sigma2', np.array([4.0]), None))
    assert m.dimension_parameters == 3 and list(m.hypers) == ['scale', 'sigma2']
    assert_close(m.cross_cov(np.array([[2.0, 4.0]]), np.array([[3.0, 5.0]])),
                 np.array([[3.0737065834936015]]), rtol=1e-13)
    got = m.cross_cov(np.array([[2.0, 4.0], [3.0, 5.0]]), np.array([[1.5, 9.0], [-3.0, 8.0]]))
    assert_close(got, np.array([[0.87752659905500319, 0.14684671522649542],
                                [1.0880320585678382, 0.084041575076539962]]), rtol=1e-13)
    assert got.shape == (2, 2)
    assert m.cross_cov(np.array([[2.0, 4.0], [3.0, 5.0]]), np.array([[1.5, 9.0]])).shape == (2, 1)
    # sigma2 = 3, ls = [1, 2]: r2 = 1.25 between (1,0) and (0,1)
    m3 = K.ScaledKernel(2, K.Matern52(2, PE('scale', np.array([1.0, 2.0]), None)),
                        PE('sigma2', np.array([3.0]), None))
    r2 = 1.25
    r = np.sqrt(r2)
    want = (1.0 + np.sqrt(5) * r + (5.0 / 3.0) * r2) * np.exp(-np.sqrt(5) * r) * 3.0
    cov = m3.cov(np.array([[1.0, 0.0], [0.0, 1.0]]))
    assert_close(cov, np.array([[3.0, want], [want, 3.0]]), rtol=1e-13)
    # gradient with respect to the parameters: finite differences as in test_matern52.py:104-117
    inputs = np.array([[2.0, 4.0], [3.0, 5.0]])
    base = np.array([2.0, 3.0, 4.0])
    grad = K.ScaledKernel.evaluate_grad_defined_by_params_respect_params(base, inputs, 2,
                                                                         *(['Matern52'],))
    for i in range(3):
        hp = base.copy()
        hm = base.copy()
        hp[i] += 1e-6
        hm[i] -= 1e-6
        fd = (K.ScaledKernel.evaluate_cov_defined_by_params(hp, inputs, 2, *(['Matern52'],)) -
              K.ScaledKernel.evaluate_cov_defined_by_params(hm, inputs, 2, *(['Matern52'],))) / 2e-6
        assert_close(grad[i], fd, rtol=1e-7, atol=1e-9, what="d cov / d param %d" % i)
    g = m.gradient_respect_parameters(inputs)
    assert set(g) == {'scale', 'sigma2'} and set(g['scale']) == {0, 1}
    assert_close(g['sigma2'], m.cov(inputs) / 4.0, rtol=1e-13)
    # gradient and Hessian with respect to the point
    point = np.array([[42.0, 35.0]])
    gp_ = m.grad_respect_point(point, inputs)
    hp_ = m.hessian_respect_point(point, inputs)
    assert gp_.shape == (2, 2) and hp_.shape == (2, 2, 2)
    for l in range(2):
        e = 1e-4
        pp, pm = point.copy(), point.copy()
        pp[0, l] += e
        pm[0, l] -= e
        fd = (m.cross_cov(pp, inputs) - m.cross_cov(pm, inputs))[0] / (2 * e)
        assert_close(gp_[:, l], fd, rtol=1e-6, atol=1e-30, what="d k / d point")
        fdh = (m.grad_respect_point(pp, inputs) - m.grad_respect_point(pm, inputs)) / (2 * e)
        assert_close(hp_[:, l, :], fdh, rtol=1e-6, atol=1e-30, what="d2 k / d point2")


def test_default_kernels_priors_and_bounds():
    K, PE = _kernels()
    k = K.Matern52.define_default_kernel(2, bounds=[[0, 1], [-1, 2]])
    assert_close(k.hypers_values_as_array, [1.0, 1.0], rtol=0)
    assert k.get_bounds_parameters() == [(1e-10, 1000.0), (1e-10, 3000.0)]
    assert k.length_scale.prior.logprob(np.array([0.5, 2999.0])) == 0.0
    assert k.length_scale.prior.logprob(np.array([0.5, 3001.0])) == -np.inf
    s = K.ScaledKernel.define_default_kernel(2, None, np.array([1.0, 2.0, 4.0]), None, *(['Matern52'],))
    assert_close(s.hypers_values_as_array, [1.0, 2.0, 4.0], rtol=0)
    assert s.get_bounds_parameters()[-1] == (1e-10, None)
    assert s.name_parameters_as_list == [('length_scale', [(0, None), (1, None)]), ('sigma2', None)]
    t = K.TasksKernel.define_default_kernel(3)
    assert t.dimension_parameters == 6 and t.lower_triang.prior.logprob(np.zeros(6)) == 0.0
    assert K.Matern52.compare_kernels(k, K.Matern52.define_default_kernel(2, bounds=[[0, 1], [-1, 2]]))
    data = {'points': np.array([[0.0], [1.0], [3.0]]), 'evaluations': np.array([1.0, 2.0, 4.0])}
    # mean |x - y| over the 9 ordered pairs = (2 * (1 + 3 + 2)) / 9
    assert_close(K.Matern52.define_prior_parameters(data, 1)['length_scale'], [12.0 / 9.0 / 0.324],
                 rtol=1e-14)


def test_tasks_kernel_known_answers():
    K, PE = _kernels()
    t1 = K.TasksKernel(1, PE('lower_triang', np.array([3.0]), None))
    t1.compute_cov_matrix()
    assert_close(t1.chol_base_cov_matrix, [[np.exp(3.0)]], rtol=1e-14)
    assert_close(t1.base_cov_matrix, [[np.exp(6.0)]], rtol=1e-14)
    same = K.TasksKernel(2, PE('lw', np.array([0.0, 0.0]), None), same_correlation=True)
    same.compute_cov_matrix()
    assert_close(same.base_cov_matrix, [[2.0, 1.0], [1.0, 2.0]], rtol=1e-15)
    assert_close(same.chol_base_cov_matrix, [[2.0, 1.0], [1.0, 2.0]], rtol=1e-15)
    same1 = K.TasksKernel(1, PE('lw', np.array([0.0]), None), same_correlation=True)
    same1.compute_cov_matrix()
    assert_close(same1.base_cov_matrix, [[1.0]], rtol=1e-15)
    t = K.TasksKernel(1, PE('lower_triang', np.array([1.0]), None))
    assert_close(t.cross_cov(np.array([[0]]), np.array([[0]])), [[np.exp(2.0)]], rtol=1e-14)
    assert t.grad_respect_point(np.array([[0]]), np.array([[0]])).tolist() == [[0.0]]
    grad = K.GradientTasksKernel.gradient_respect_parameters(np.array([[np.exp(1.0)]]), 1)
    assert len(grad) == 1
    assert_close(grad[0], [[2.0 * np.exp(2.0)]], rtol=1e-13)
    # finite differences of the covariance table (test_tasks_kernel.py:96-126)
    inputs = np.array([[0], [1]])
    for params, kw in ((np.array([2.0, 3.0, 4.0]), {}), (np.array([2.0, 3.0]), {'same_correlation': True})):
        g = K.TasksKernel.evaluate_grad_defined_by_params_respect_params(params, inputs, 2, **kw)
        for i in range(len(params)):
            hp, hm = params.copy(), params.copy()
            hp[i] += 1e-6
            hm[i] -= 1e-6
            fd = (K.TasksKernel.evaluate_cov_defined_by_params(hp, inputs, 2, **kw) -
                  K.TasksKernel.evaluate_cov_defined_by_params(hm, inputs, 2, **kw)) / 2e-6
            assert_close(g[i], fd, rtol=1e-7, atol=1e-6, what="d C / d p%d" % i)
    k = K.TasksKernel.define_kernel_from_array(5, np.array([1, 2, 3]))
    assert k.n_tasks == 5 and np.all(k.lower_triang.value == np.array([1, 2, 3]))


def test_product_kernel_known_answer_and_fixtures(golden):
    K, PE = _kernels()
    # tests/models/test_gp_fitting_gaussian.py:680-684
    got = K.ProductKernels.evaluate_cross_cov_defined_by_params_array(
        [np.array([1.0]), np.array([0.0])], np.array([[2.0, 0.0]]), np.array([[1.0, 0.0]]), [1, 1],
        *(['Matern52', 'Tasks_Kernel'],))
    assert_close(got, [[0.52399410883182029]], rtol=1e-13)
    for name in ("product_uniform", "product_weighted"):
        g = golden(name)
        dx, T = int(g["dx"]), int(g["n_tasks"])
        same = bool(int(g["same_correlation"]))
        for s, th in enumerate(g["thetas"]):
            pk = th[2:]
            params = [pk[:dx], pk[dx:]]
            kernel = K.ProductKernels.define_kernel_from_array(
                [dx, T], params, *(['Matern52', 'Tasks_Kernel'],), **{'same_correlation': same})
            assert kernel.dimension == dx + 1 and kernel.dimension_parameters == len(pk)
            assert_close(kernel.cov(g["points"]), g["cov"][s], rtol=1e-13, what="product cov")
            as_dict = kernel.inputs_from_array_to_dict(g["points"])
            assert_close(kernel.cov_dict(as_dict), g["cov"][s], rtol=1e-13, what="product cov_dict")
            grad = K.ProductKernels.evaluate_grad_defined_by_params_respect_params(
                params, as_dict, [dx, T], *(['Matern52', 'Tasks_Kernel'],), **{'same_correlation': same})
            assert_close(np.stack([grad[i] for i in range(len(pk))]), g["grad_cov"][s], rtol=1e-10,
                         atol=1e-15, what="product grad params")
            nested = kernel.gradient_respect_parameters(as_dict)
            assert set(nested) == {'Matern52', 'Tasks_Kernel'}
            q = g["query"][0:1]
            gp_ = kernel.grad_respect_point(q, g["points"])
            assert gp_.shape == (len(g["points"]), dx + 1) and np.all(gp_[:, -1] == 0.0)
            hp_ = kernel.hessian_respect_point(q, g["points"])
            want = golden("hessian_" + name)["hess_k"][s, 0] if name == "product_weighted" else None
            if want is not None:
                assert_close(hp_, want, rtol=1e-9, atol=1e-9 * np.max(np.abs(want)), what="product hessian")


def test_matern_and_scaled_against_fixtures(golden):
    K, PE = _kernels()
    g = golden("matern52_plain")
    for s, th in enumerate(g["thetas"]):
        pk = th[2:]
        assert_close(K.Matern52.evaluate_cov_defined_by_params(pk, g["points"], 3), g["cov"][s], rtol=1e-13)
        grad = K.Matern52.evaluate_grad_defined_by_params_respect_params(pk, g["points"], 3)
        assert_close(np.stack([grad[i] for i in range(3)]), g["grad_cov"][s], rtol=1e-10, atol=1e-15)
        cc = K.Matern52.evaluate_cross_cov_defined_by_params(pk, g["query"], g["points"], 3)
        assert cc.shape == (len(g["query"]), len(g["points"]))
    g = golden("scaled_gamma")
    h = golden("hessian_scaled_gamma")
    for s, th in enumerate(g["thetas"]):
        pk = th[2:]
        args = (['Matern52'],)
        assert_close(K.ScaledKernel.evaluate_cov_defined_by_params(pk, g["points"], 4, *args),
                     g["cov"][s], rtol=1e-13)
        grad = K.ScaledKernel.evaluate_grad_defined_by_params_respect_params(pk, g["points"], 4, *args)
        assert_close(np.stack([grad[i] for i in range(5)]), g["grad_cov"][s], rtol=1e-10, atol=1e-15)
        hess = K.ScaledKernel.evaluate_hessian_respect_point(pk, g["query"][1:2], g["points"], 4, *args)
        assert_close(hess, h["hess_k"][s, 1], rtol=1e-9, atol=1e-9 * np.max(np.abs(h["hess_k"][s, 1])))
