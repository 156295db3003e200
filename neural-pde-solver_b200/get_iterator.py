"""The drop-in factory: same call and same 4-tuple as models/get_iterator.py:3-48 of the reference.

    iterator, compare_model, operations_ratio, is_train = get_iterator(opt)

Behaviour kept (tests/golden/get_iterator_contract*.npy hold the reference's answers for every name):
  * opt.iterator in {jacobi, multigrid, cg, conv, unet}; anything else raises NotImplementedError;
  * only the learned iterators (conv, unet) keep opt.is_train;
  * conv / multigrid are compared with Jacobi, unet / cg with a multigrid V-cycle, jacobi with nothing;
  * on any geometry but 'square' both get is_bc_mask = True;
  * operations_ratio = n_operations(iterator) / n_operations(compare_model), 1 without a comparison.
"""
import torch

from .iterators import ConjugateGradient, ConvIterator, JacobiIterator, MultigridIterator, UNetIterator


def _vcycle(opt):
    return MultigridIterator(opt.mg_n_layers, opt.mg_pre_smoothing, opt.mg_post_smoothing)


# name -> (constructor, learned?, constructor of the comparison solver or None)
_TABLE = {
    'jacobi': (lambda opt: JacobiIterator(), False, None),
    'multigrid': (_vcycle, False, lambda opt: JacobiIterator()),
    'cg': (lambda opt: ConjugateGradient(opt.cg_n_iters), False, _vcycle),
    'conv': (lambda opt: ConvIterator(opt.activation, opt.conv_n_layers), True, lambda opt: JacobiIterator()),
    'unet': (lambda opt: UNetIterator(opt.activation, opt.mg_n_layers, opt.mg_pre_smoothing, opt.mg_post_smoothing),
             True, _vcycle),
}


def get_iterator(opt):
    if opt.iterator not in _TABLE:
        raise NotImplementedError
    make, learned, make_reference = _TABLE[opt.iterator]
    iterator = make(opt)
    if torch.cuda.is_available():
        iterator = iterator.cuda()
    compare_model = make_reference(opt) if make_reference is not None else None

    masked = opt.geometry != 'square'
    for solver in (iterator, compare_model):
        if solver is not None and masked:
            solver.is_bc_mask = True

    ratio = iterator.n_operations / compare_model.n_operations if compare_model is not None else 1
    return iterator, compare_model, ratio, (opt.is_train if learned else False)
