"""
`Algorithm` ABC and the `njobs` decorator, mirroring reference nd/algorithm.py:15-105.

Difference from the reference (deliberate, this is the B200 analogue of its only data-parallel
strategy): `njobs` does not fork CPU worker processes over pickled Datasets
(nd/utils.py:343-401); it is the number of GPUs the cube is sharded over, each shard carrying a
halo of `_buffer(dim)` rows along `_parallel_dimension(ds)` (nd/filters.py:424-445).  `njobs=1`
(default) is one GPU, `njobs=-1` all visible GPUs.
"""
import inspect
from abc import ABC, abstractmethod
from collections import OrderedDict
from functools import partial


class Algorithm(ABC):
    @abstractmethod
    def apply(self, ds):
        """Must be implemented by derived classes (reference nd/algorithm.py:18-24)."""
        return

    def _buffer(self, dim):
        """Required halo when sharding over `dim` (reference nd/algorithm.py:26-31)."""
        return 0

    def _parallel_dimension(self, ds):
        """Dimension along which to shard (reference nd/algorithm.py:33-35)."""
        return 'y'


def parallelize(func):
    """Add the `njobs` keyword to `apply` (reference nd/algorithm.py:38-105)."""

    def wrapper(self, ds, *args, njobs=1, **kwargs):
        method = partial(func, self)
        if njobs == -1:
            import torch
            njobs = max(1, torch.cuda.device_count())
        if njobs == 1:
            return method(ds, *args, **kwargs)
        # The shard layer lives below the Dataset marshalling: the filter sees `_njobs`
        # and splits the staged cube over that many GPUs (nd_b200/shard.py).
        if not getattr(self, '_supports_njobs', False):
            import warnings
            warnings.warn('%s has no multi-GPU shard layer: njobs=%d runs on one GPU with the same result (these '
                          'filters are HBM-bound, milliseconds per GB)' % (type(self).__name__, njobs))
        prev = getattr(self, '_njobs', 1)
        self._njobs = int(njobs)
        self._shard_dim = self._parallel_dimension(ds)
        try:
            return method(ds, *args, **kwargs)
        finally:
            self._njobs = prev

    sig_func = inspect.signature(func)
    sig_wrapper = inspect.signature(wrapper)
    parameters = tuple(sig_func.parameters.values()) + (sig_wrapper.parameters['njobs'],)
    parameters = sorted(parameters, key=lambda p: (p.kind, p.default is not inspect._empty))
    new_parameters = []
    for p in parameters:
        if p not in new_parameters:
            new_parameters.append(p)
    wrapper.__signature__ = sig_func.replace(parameters=new_parameters)
    wrapper.__doc__ = (func.__doc__ or '') + (
        "\n        njobs : int, optional\n"
        "            Number of GPUs to shard over (-1: all visible GPUs; default 1).\n")
    wrapper.__name__ = getattr(func, '__name__', 'apply')
    return wrapper


def extract_arguments(fn, args, kwargs):
    """Split call arguments between `fn` and the leftovers (reference nd/utils.py:727-749)."""
    def _(*args, **kwargs):
        pass
    sig = inspect.signature(fn)
    if 'self' in sig.parameters:
        sig = sig.replace(parameters=tuple(sig.parameters.values())[1:])
    parameters = OrderedDict(sig.parameters)
    parameters.update(OrderedDict(inspect.signature(_).parameters))
    parameters = sorted(parameters.values(), key=lambda p: (p.kind, p.default is not inspect._empty))
    bound = sig.replace(parameters=parameters).bind(*args, **kwargs)
    bound.apply_defaults()
    return bound.arguments


def wrap_algorithm(algo, name=None):
    """Function form of an Algorithm class (reference nd/algorithm.py:108-198, without the
    docstring / source-location surgery, which is documentation tooling)."""
    if not issubclass(algo, Algorithm):
        raise ValueError('Class must be an instance of `nd.Algorithm`.')

    def _wrapper(*args, **kwargs):
        apply_kwargs = extract_arguments(algo.apply, args, kwargs)
        init_args = apply_kwargs.pop('args', ())
        init_kwargs = apply_kwargs.pop('kwargs', {})
        return algo(*init_args, **init_kwargs).apply(**apply_kwargs)

    _wrapper.__module__ = algo.__module__
    if name is not None:
        _wrapper.__name__ = name
    sig_init = inspect.signature(algo.__init__)
    sig_apply = inspect.signature(algo.apply)
    parameters = tuple(sig_apply.parameters.values())[1:] + tuple(sig_init.parameters.values())[1:]
    parameters = sorted(parameters, key=lambda p: (p.kind, p.default is not inspect._empty))
    new_parameters = []
    for p in parameters:
        if p not in new_parameters:
            new_parameters.append(p)
    _wrapper.__signature__ = sig_init.replace(parameters=new_parameters)
    _wrapper.__doc__ = "Wrapper for :class:`{}.{}`.\n\n{}".format(algo.__module__, algo.__name__, algo.__doc__ or '')
    return _wrapper
