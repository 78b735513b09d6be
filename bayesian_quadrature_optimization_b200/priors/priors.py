"""Prior densities of the GP hyper-parameters (host side, tiny): log-density and sampling of
sbo/priors/{uniform,gaussian,log_normal_square,horseshoe,non_negative,multivariate_normal,
constant}.py.  They feed `GPFittingGaussian.log_prob_parameters`, whose likelihood term is the
GPU log-marginal-likelihood."""
import numpy as np
from scipy.stats import lognorm, multivariate_normal, norm


class UniformPrior(object):
    """sbo/priors/uniform.py:22-42."""

    def __init__(self, dimension, min, max):
        self.dimension = dimension
        self.min = np.array(min, dtype=float)
        self.max = np.array(max, dtype=float)

    def logprob(self, x):
        if np.any(x < self.min) or np.any(x > self.max):
            return -np.inf
        return 0.0

    def sample(self, samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return self.min + np.random.rand(samples, self.dimension) * (self.max - self.min)


class GaussianPrior(object):
    """sbo/priors/gaussian.py."""

    def __init__(self, dimension, mu, sigma):
        self.dimension, self.mu, self.sigma = dimension, mu, sigma

    def logprob(self, x):
        if self.sigma == 0:
            return self.mu == x[0]
        return np.sum(norm.logpdf(x, loc=self.mu, scale=self.sigma))

    def sample(self, samples, random_seed=None):
        if self.sigma == 0:
            return self.mu
        return self.mu + np.random.randn(samples, self.dimension) * self.sigma


class LogNormalSquare(object):
    """sbo/priors/log_normal_square.py:25-53: x = y^2, y ~ LogNormal(s=scale, scale=param)."""

    def __init__(self, dimension, scale, mu):
        self.dimension, self.scale, self.mu = dimension, scale, mu
        self.param = np.sqrt(abs(2.0 * np.log(self.mu)))

    def logprob(self, x):
        x = np.asarray(x, dtype=float)
        if np.any(x <= 0):
            return -np.inf
        y = np.sqrt(x)
        return float(np.sum(lognorm.logpdf(y, s=self.scale, scale=self.param)) - np.sum(np.log(2.0 * y)))

    def sample(self, samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        points = lognorm.rvs(s=self.scale, scale=self.param, size=samples)
        return points.reshape([samples, self.dimension]) ** 2


class HorseShoePrior(object):
    """sbo/priors/horseshoe.py."""

    def __init__(self, dimension, scale):
        self.dimension, self.scale = dimension, scale

    def logprob(self, x):
        x = np.asarray(x, dtype=float)
        if np.any(x == 0.0):
            return np.inf
        return np.sum(np.log(np.log(1 + 3.0 * (self.scale / x) ** 2)))

    def sample(self, samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        lambda_ = np.abs(np.random.standard_cauchy(size=(samples, self.dimension)))
        return np.random.randn(samples, self.dimension) * lambda_ * self.scale


class NonNegativePrior(object):
    """sbo/priors/non_negative.py."""

    def __init__(self, dimension, prior):
        self.dimension, self.prior = dimension, prior

    def logprob(self, x):
        if np.any(np.asarray(x) <= 0):
            return -np.inf
        return self.prior.logprob(x)

    def sample(self, samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return np.abs(self.prior.sample(samples))


class MultivariateNormalPrior(object):
    """sbo/priors/multivariate_normal.py."""

    def __init__(self, dimension, mu, cov):
        self.dimension, self.mu, self.cov = dimension, np.asarray(mu, dtype=float), cov

    def logprob(self, x):
        return np.sum(multivariate_normal.logpdf(x, mean=self.mu, cov=self.cov))

    def sample(self, samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return np.random.multivariate_normal(self.mu, self.cov, size=samples)


class Constant(object):
    """sbo/priors/constant.py: a fixed parameter."""

    def __init__(self, dimension, value):
        self.dimension, self.value = dimension, value

    def logprob(self, x):
        return 0.0 if self.value == x[0] else -np.inf

    def sample(self, samples, random_seed=None):
        return self.value * np.ones((samples, 1))
