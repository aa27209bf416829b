"""Mirror of torchreid.models for the hot path (reference torchreid/models/__init__.py:17-41).

Only the model the hot path names is provided: ``init_model('vmgn', ...)``.  Like the reference,
an unknown name raises ``KeyError`` and a ``save_dir`` keyword makes the factory copy the model's
source file there before constructing it (:37-40)."""
import inspect
import os
import shutil

from . import vmgn as vmgn_module
from .vmgn import vmgn, VMGN, pool_clips

_FACTORY = {'vmgn': vmgn}

__all__ = ['init_model', 'get_names', 'vmgn', 'VMGN', 'pool_clips']


def get_names():
    return list(_FACTORY.keys())


def init_model(name, *args, **kwargs):
    if name not in _FACTORY:
        raise KeyError("Unknown model: {}".format(name))
    if 'save_dir' in kwargs:
        src = inspect.getfile(_FACTORY[name])
        shutil.copyfile(src, os.path.join(os.path.abspath(kwargs['save_dir']), os.path.basename(src)))
    return _FACTORY[name](*args, **kwargs)
