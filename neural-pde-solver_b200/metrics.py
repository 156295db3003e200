"""utils/metrics.py:5-46 -- iterations-to-threshold bookkeeping of test.py / runtime.py
(SURVEY.md 8f.2).  Consumes the (batch, steps+1) error matrices that
``utils.calculate_errors`` builds from the fused kernel's per-step error rows."""
import numpy as np
import torch


def _to_numpy(a):
    """utils/misc.py:8-17."""
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def first_below(errors, threshold):
    """Per problem: index of the first step whose error is below `threshold`, -1 if none.  (batch,) int64."""
    below = _to_numpy(errors) < threshold
    first = below.argmax(axis=1).astype(np.int64)
    first[~below.any(axis=1)] = -1
    return first


class Metrics(object):
    def __init__(self, scale, error_threshold, verbose=True):
        if verbose:
            print('Runtime scaling: {}'.format(scale))
        self.scale = scale
        self.error_threshold = error_threshold
        self.verbose = verbose
        self.reset()

    def reset(self):
        self.model_steps = []
        self.fd_steps = []
        self.invalid = 0

    def update(self, error_dict):
        """Count, per problem, the first step at which the model and the comparison solver fall below the
        threshold; problems where either never does are 'invalid'.  Model steps are weighted by `scale`
        (the operations ratio of get_iterator)."""
        model = first_below(error_dict['model errors'], self.error_threshold)
        other = first_below(error_dict.get('Jacobi errors', error_dict['model errors']), self.error_threshold)
        for m, o in zip(model, other):
            if m >= 0 and o >= 0:
                self.model_steps.append(m * self.scale)
                self.fd_steps.append(o)
            else:
                self.invalid += 1

    def get_results(self):
        if self.verbose:
            print('Invalid: {}/{}'.format(self.invalid, len(self.fd_steps) + self.invalid))
        self.model_steps = np.array(self.model_steps)
        self.fd_steps = np.array(self.fd_steps)
        with np.errstate(divide='ignore', invalid='ignore'):
            ratios = self.model_steps / self.fd_steps
        return {'Jacobi': self.fd_steps.mean(), 'model': self.model_steps.mean(), 'ratio': ratios.mean()}
