"""npde-b200: the learned-iterator hot path of ermongroup/Neural-PDE-Solver as
hand-written sm_100a CUDA kernels behind the reference's own Python API.

    import importlib; npde = importlib.import_module("neural-pde-solver_b200")
    it, cmp_model, ratio, is_train = npde.get_iterator(opt)     # models/get_iterator.py
    y = it(x, bc, f)                                             # Iterator.forward
    npde.utils.fd_step(x, bc, f)                                 # utils/heat_utils.py surface
"""
from . import _lib, functional
from . import heat_utils as utils
from .get_iterator import get_iterator
from .iterators import (ConjugateGradient, ConvIterator, Iterator, JacobiIterator, MultigridIterator,
                        MultigridResidualIterator, UNetIterator)
from .heat_model import HeatModel
from . import data, evaluation, generation, geometries, metrics, runtime

__all__ = ["get_iterator", "Iterator", "JacobiIterator", "ConvIterator", "MultigridIterator", "UNetIterator", "ConjugateGradient", "MultigridResidualIterator",
           "HeatModel", "utils", "functional", "data", "evaluation", "generation", "geometries", "metrics", "runtime"]
