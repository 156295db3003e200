"""The evaluation driver of test.py:10-118 (BASELINE config 1 runs through it): loop a data loader, run
`HeatModel.evaluate` (fused error curves of the model and of its comparison solver) and count iterations to the
error threshold with `Metrics`.  Plotting / tensorboard of the reference are out of scope and left out; the
numbers it logs are returned."""
import os

import numpy as np

from . import heat_utils as utils


class _Printer(object):
    def print(self, msg):
        print(msg)


def evaluate(opt, model, data_loader, logger=None, error_threshold=0.05, limit=None):
    """test.py:10-55 without the image dumps.  Returns the Metrics results dict
    {'Jacobi': mean steps of the comparison solver, 'model': mean steps of the model, 'ratio': mean ratio}."""
    logger = logger or _Printer()
    if model.compare_model is not None:
        logger.print('Comparison: {} ({}), {} ({})'.format(model.iterator.name(), model.iterator.n_operations,
                                                           model.compare_model.name(),
                                                           model.compare_model.n_operations))
    logger.print('Initialization: {}'.format(opt.initialization))
    logger.print('Error threshold: {}'.format(error_threshold))
    metric = utils.Metrics(scale=1, error_threshold=error_threshold, verbose=False)
    for step, batch in enumerate(data_loader):
        bc, gt, x = batch['bc'], batch['final'], batch['x']
        f = batch.get('f')
        if opt.initialization != 'random':  # test time: 'random' keeps the stored start field
            x = utils.initialize(x, bc, opt.initialization)
        results, _ = model.evaluate(x, gt, bc, f, opt.n_evaluation_steps)
        metric.update(results)
        if limit is not None and (step + 1) == limit:
            break
    invalid, total = metric.invalid, metric.invalid + len(metric.fd_steps)
    results = metric.get_results()
    logger.print('Invalid: {}/{}'.format(invalid, total))
    for key in results:
        logger.print('{}: {}'.format(key, results[key]))
    metric.reset()
    return results


def check_eigenvalues(opt, model, logger=None, save_dir=None):
    """test.py:57-94: spectrum of the model's update matrix, of HT + T - H on the periodic grid, and of the
    comparison solver -- all probes batched (utils.construct_matrix).  Returns the three sorted |eigenvalue| arrays."""
    logger = logger or _Printer()
    if opt.iterator == 'cg':
        return None
    if opt.iterator != 'conv':
        raise NotImplementedError
    image_size = 15
    w, v = utils.calculate_eigenvalues(model, image_size)
    if save_dir:
        np.save(os.path.join(save_dir, 'eigenvalues.npy'), w)
        np.save(os.path.join(save_dir, 'eigenvectors.npy'), v)
    w_model = np.sort(np.abs(w))
    logger.print('Absolute eigenvalues:\n{}\n'.format(w_model[-5:]))
    wh, _ = utils.calculate_eigenvalues_wraparound(model, image_size)
    w_h = np.sort(np.abs(wh))
    logger.print('H eigenvalues:\n{}\n'.format(w_h[-5:]))
    w_fd = None
    if model.compare_model is not None:
        keep = model.compare_model.is_bc_mask
        model.compare_model.is_bc_mask = False
        try:
            _, B = utils.construct_matrix(np.zeros((1, 4)), image_size, model.compare_model.iter_step,
                                          device=utils._probe_device(model.iter_step))
        finally:
            model.compare_model.is_bc_mask = keep
        w_fd = np.sort(np.abs(np.linalg.eig(B.numpy())[0]))
        logger.print('Finite difference eigenvalues:\n{}\n'.format(w_fd[-5:]))
    return w_model, w_h, w_fd


def test(opt, model, data_loader, logger=None):
    """test.py:96-123: eigenvalue check for conv models on the square domain, then the 'avg'-initialised
    evaluation with threshold 0.01 (0.1 with a Poisson source)."""
    model.setup(is_train=False)
    spectra = None
    if opt.geometry == 'square' and opt.iterator != 'unet':
        spectra = check_eigenvalues(opt, model, logger)
    opt.initialization = 'avg'
    threshold = 0.1 if opt.poisson else 0.01
    return evaluate(opt, model, data_loader, logger, threshold), spectra
