"""Slice sampling of the GP hyper-posterior (Neal 2003): the algorithm of
sbo/samplers/slice_sampling.py:49-278 (doubling procedure with the acceptance test, shrinkage,
component-wise or random-direction moves) kept on the host -- it is a sequential Markov chain --
while every `log_prob` probe is a GPU log-marginal-likelihood.  The numpy.random call order
(interval offset, slice height, doubling coin flips, shrinkage proposals) follows the reference so
that a seed reproduces a chain up to the 1e-11 agreement of the two likelihoods.

Batched probes.  One probe costs a Cholesky factorisation whose time on the
GPU hardly depends on how many matrices are factorised together, so when the model offers a batched
log-probability (`log_prob_batch`) the probes whose positions are known in advance go out together:
the two interval ends, the two ends of every halving in the acceptance test, and the next
`speculate` shrinkage proposals -- the k-th proposal only depends on the positions (not the values) of
the rejected ones before it.  The random-number stream is rewound after a speculative draw and
replayed proposal by proposal, so the chain is exactly the sequential one.
"""
import numpy as np
import numpy.random as npr


def combine_vectors(vector_1, vector_2, indexes_1):
    """sbo/lib/util.py combine_vectors: scatter vector_1 at indexes_1, vector_2 elsewhere."""
    n = len(vector_1) + len(vector_2)
    out = np.zeros(n)
    mask = np.zeros(n, dtype=bool)
    mask[list(indexes_1)] = True
    out[mask] = vector_1
    out[~mask] = vector_2
    return out


def separate_vector(vector, indexes):
    """sbo/lib/util.py separate_vector -> [vector[not indexes], vector[indexes]]."""
    vector = np.asarray(vector, dtype=float)
    if indexes is None:
        return [vector, np.array([])]
    mask = np.zeros(len(vector), dtype=bool)
    mask[list(indexes)] = True
    return [vector[~mask], vector[mask]]


class SliceSampling(object):

    def __init__(self, log_prob, indexes, ignore_index=None, **params):
        self.ignore_index = [] if ignore_index is None else ignore_index
        self.log_prob = log_prob
        self.indexes = indexes
        self.sigma = params.get('sigma', 1.0)
        self.step_out = params.get('step_out', True)
        self.max_steps_out = params.get('max_steps_out', 1000)
        self.component_wise = params.get('component_wise', True)
        self.doubling_step = params.get('doubling_step', True)
        self.log_prob_batch = params.get('log_prob_batch')  # optional: (vectors, *args) -> values
        self.speculate = int(params.get('speculate', 4))
        # random stream: the global numpy one (reference behaviour) or a RandomState of its own
        # (independent chains running side by side, GPFittingGaussian.sample_parameters_chains)
        self.rng = params.get('rng', npr)

    def slice_sample(self, point, fixed_parameters, *args_log_prob):
        point = np.asarray(point, dtype=float)
        dimensions = len(point)
        if self.component_wise:
            dims = np.arange(dimensions)
            self.rng.shuffle(dims)
            new_point = point.copy()
            for d in dims:
                if d not in self.ignore_index:
                    direction = np.zeros(dimensions)
                    direction[d] = 1.0
                    new_point = self.direction_slice(direction, new_point, fixed_parameters,
                                                     *args_log_prob)
            return new_point
        direction = self.rng.randn(dimensions)
        for d in self.ignore_index:
            direction[d] = 0.0
        if np.all(direction == 0):
            return point
        direction = direction / np.sqrt(np.sum(direction ** 2))
        return self.direction_slice(direction, point, fixed_parameters, *args_log_prob)

    def directional_log_prob(self, x, direction, point, fixed_parameters=None, *args_log_prob):
        new_point = point + x * direction
        if fixed_parameters is not None:
            new_point = combine_vectors(new_point, fixed_parameters, self.indexes)
        return self.log_prob(new_point, *args_log_prob)

    def directional_log_probs(self, xs, direction, point, fixed_parameters=None, *args_log_prob):
        """log_prob at several positions along the direction: one batched call when available."""
        if self.log_prob_batch is None or len(xs) == 1:
            return [self.directional_log_prob(x, direction, point, fixed_parameters, *args_log_prob)
                    for x in xs]
        pts = []
        for x in xs:
            new_point = point + x * direction
            if fixed_parameters is not None:
                new_point = combine_vectors(new_point, fixed_parameters, self.indexes)
            pts.append(new_point)
        return list(self.log_prob_batch(pts, *args_log_prob))

    def acceptable(self, z, llh, L, U, direction, point, fixed_parameters, *args):
        if not self.doubling_step:
            return True
        while (U - L) > 1.1 * self.sigma:
            old_U, old_L = U, L
            middle = 0.5 * (L + U)
            if z < middle:
                U = middle
            else:
                L = middle
            if U == old_U and L == old_L:
                return False
            D = (middle > 0 and z >= middle) or (middle <= 0 and z < middle)
            lp0, lp1 = self.directional_log_probs([U, L], direction, point, fixed_parameters, *args)
            if D and llh >= lp0 and llh >= lp1:
                return False
        return True

    def find_x_interval(self, llh, lower, upper, direction, point, fixed_parameters, *args):
        l_out = u_out = 0
        if self.doubling_step:
            lp0, lp1 = self.directional_log_probs([lower, upper], direction, point,
                                                  fixed_parameters, *args)
            while (lp0 > llh or lp1 > llh) and (l_out + u_out < self.max_steps_out):
                if self.rng.rand() < 0.5:
                    l_out += 1
                    lower -= (upper - lower)
                    lp0 = self.directional_log_prob(lower, direction, point, fixed_parameters, *args)
                else:
                    u_out += 1
                    upper += (upper - lower)
                    lp1 = self.directional_log_prob(upper, direction, point, fixed_parameters, *args)
        else:
            lp1 = self.directional_log_prob(lower, direction, point, fixed_parameters, *args)
            while lp1 > llh and l_out < self.max_steps_out:
                l_out += 1
                lower -= self.sigma
                lp1 = self.directional_log_prob(lower, direction, point, fixed_parameters, *args)
            lp2 = self.directional_log_prob(upper, direction, point, fixed_parameters, *args)
            while lp2 > llh and u_out < self.max_steps_out:
                u_out += 1
                upper += self.sigma
                lp2 = self.directional_log_prob(upper, direction, point, fixed_parameters, *args)
        return upper, lower

    def find_sample(self, lower, upper, llh, direction, point, fixed_parameters, *args):
        start_upper, start_lower = upper, lower
        ahead = []  # log-probabilities of the coming proposals, evaluated speculatively
        while True:
            if not ahead and self.log_prob_batch is not None and self.speculate > 1:
                # the next proposals if every one before them is rejected: their positions only
                # depend on the positions of the rejected ones.  Draw, evaluate together, rewind.
                state = self.rng.get_state()
                zs, lo, up = [], lower, upper
                for _ in range(self.speculate):
                    z = (up - lo) * self.rng.rand() + lo
                    zs.append(z)
                    if z < 0:
                        lo = z
                    elif z > 0:
                        up = z
                    else:
                        break
                self.rng.set_state(state)
                ahead = self.directional_log_probs(zs, direction, point, fixed_parameters, *args)
            new_z = (upper - lower) * self.rng.rand() + lower
            if ahead:
                new_llh = ahead.pop(0)
            else:
                new_llh = self.directional_log_prob(new_z, direction, point, fixed_parameters,
                                                    *args)
            if np.isnan(new_llh):
                new_llh = -np.inf
            if new_llh > llh and self.acceptable(new_z, llh, start_lower, start_upper, direction,
                                                 point, fixed_parameters, *args):
                return new_z
            elif new_z < 0:
                lower = new_z
            elif new_z > 0:
                upper = new_z
            else:
                raise Exception("Slice sampler shrank to zero!")

    def direction_slice(self, direction, point, fixed_parameters, *args):
        upper = self.sigma * self.rng.rand()
        lower = upper - self.sigma
        llh = np.log(self.rng.rand()) + self.directional_log_prob(0.0, direction, point,
                                                             fixed_parameters, *args)
        if self.step_out:
            upper, lower = self.find_x_interval(llh, lower, upper, direction, point,
                                                fixed_parameters, *args)
        new_z = self.find_sample(lower, upper, llh, direction, point, fixed_parameters, *args)
        return new_z * direction + point
