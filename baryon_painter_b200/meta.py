"""Reader/writer for the reference ``model_meta`` checkpoint side-car.

The reference writes ``model_meta`` with ``dill.dump`` (reference
``baryon_painter/painter.py:399-417``).  The two transform entries are the
``lambda`` built by ``datasets.compile_transform`` (reference
``baryon_painter/utils/datasets.py:8-13``) pickled *by value*, i.e. as Python-3.7
``CodeType`` objects, which no current interpreter can rebuild
(``TypeError: code expected at least 16 arguments, got 15``).

Everything the paint path needs from those closures is *data*: the mode strings,
``k_values``, ``eps`` and the per-redshift ``stats`` table.  This module walks the
pickle with a restricted unpickler that never builds code objects and never
imports anything outside an exact whitelist of numpy reconstructors (``getattr`` in the stream
only resolves on this module's inert stand-ins), lifts the closure cells out and
re-binds them to :class:`baryon_painter_b200.transforms.CompiledTransform`.

``write_model_meta`` emits a plain-pickle dict with the same keys; the transform
entries are ``CompiledTransform`` instances (picklable by reference to this
package), so files written here load back here without dill.
"""

import collections
import importlib
import io
import pickle

import numpy as np

from . import transforms as _tf

META_KEYS = ("L", "n_grid", "tile_L", "n_tile", "tile_size", "input_field",
             "label_fields", "scale_to_SLICS", "transform", "inverse_transform",
             "model_architecture")


class _Code:
    """Inert stand-in for a pickled ``types.CodeType`` (any Python version)."""

    def __init__(self, *args):
        self.args = args
        strs = [a for a in args if isinstance(a, str)]
        tups = [a for a in args if isinstance(a, tuple)]
        # Python 3.7 layout: (argcount, kwonly, nlocals, stack, flags, code, consts,
        # names, varnames, filename, name, firstlineno, lnotab, freevars, cellvars).
        # 3.8+ adds posonlyargcount in front; 3.11 adds qualname/exceptiontable.
        if len(args) == 15:
            self.name, self.freevars = args[10], args[13]
        elif len(args) == 16:
            self.name, self.freevars = args[11], args[14]
        else:                                   # 3.11+: (..., filename, name, qualname, ...)
            self.name = strs[1] if len(strs) > 1 else "?"
            self.freevars = tups[-2] if len(tups) >= 2 else ()
        self.names = next((t for t in tups if t and all(isinstance(s, str) for s in t)), ())


class _Function:
    """Inert stand-in for a function rebuilt by ``dill._dill._create_function``."""

    def __init__(self, code, globs=None, name=None, defaults=None, closure=None, *rest):
        self.code, self.name, self.defaults, self.closure = code, name, defaults, closure

    def __setstate__(self, state):                 # dill may attach __dict__/__kwdefaults__
        self.state = state

    def cell(self, freevar):
        idx = list(self.code.freevars).index(freevar)
        return self.closure[idx].contents


class _Cell:
    def __init__(self, contents=None):
        self.contents = contents


class _ByReference:
    """A module-level reference function pickled by name (atleast_3d / squeeze)."""

    def __init__(self, module, name):
        self.module, self.name = module, name


def _load_type(name):
    if name != "CodeType":
        raise pickle.UnpicklingError("model_meta: unexpected dill type %r" % (name,))
    return _Code


_NUMPY_MAJOR = int(np.__version__.split(".")[0])
# the only modules a model_meta / transform pickle may name: numpy's scalar and array reconstructors
_NUMPY_MODULES = {"numpy": "numpy", "numpy.core.multiarray": "numpy.core.multiarray",
                  "numpy._core.multiarray": "numpy._core.multiarray", "numpy.core.numeric": "numpy.core.numeric",
                  "numpy._core.numeric": "numpy._core.numeric"}
_NUMPY_NAMES = ("dtype", "scalar", "ndarray", "_reconstruct", "_frombuffer")


class _InertModule:
    """What dill's ``_import_module`` yields here: a NAME, never the module object -- attribute lookups on it go
    through :func:`_safe_getattr`, which resolves only the whitelisted numpy reconstructors."""

    def __init__(self, name):
        self.name = name


def _numpy_module(name):
    if name not in _NUMPY_MODULES:
        raise pickle.UnpicklingError("model_meta: refusing to import %r" % (name,))
    if _NUMPY_MAJOR >= 2:                       # numpy 1.x pickles name numpy.core.*, moved to numpy._core in 2.x
        name = name.replace("numpy.core", "numpy._core")
    return importlib.import_module(name)


def _import_module(name, safe=False):
    if name not in _NUMPY_MODULES and not name.startswith("baryon_painter."):
        raise pickle.UnpicklingError("model_meta: refusing to import %r" % (name,))
    return _InertModule(name)


def _safe_getattr(obj, name, *default):
    """``getattr`` for the pickle stream: only on the inert stand-ins of this module (never on a real module, class
    or builtin, so no chain such as module -> os -> system can be walked)."""
    if isinstance(obj, _InertModule):
        if obj.name.startswith("baryon_painter."):
            return {} if name == "__dict__" else _ByReference(obj.name, name)
        if name in _NUMPY_NAMES:
            return getattr(_numpy_module(obj.name), name)
        raise pickle.UnpicklingError("model_meta: refusing attribute %s.%s" % (obj.name, name))
    if isinstance(obj, (_Function, _ByReference, _Cell, _Code)):
        if name.startswith("__") and name != "__dict__":
            raise pickle.UnpicklingError("model_meta: refusing attribute %r" % (name,))
        return getattr(obj, name, *default) if (default or hasattr(obj, name)) else {}
    raise pickle.UnpicklingError("model_meta: refusing getattr on %s" % (type(obj).__name__,))


class _MetaUnpickler(pickle.Unpickler):
    _DILL = {"_create_function": _Function, "_load_type": _load_type, "_create_cell": _Cell,
             "_get_attr": _safe_getattr, "_import_module": _import_module}

    def find_class(self, module, name):
        if module == "dill._dill" and name in self._DILL:
            return self._DILL[name]
        if module.startswith("baryon_painter."):
            if name == "__dict__":
                return {}
            return _ByReference(module, name)
        if module == "baryon_painter_b200.transforms":
            return getattr(_tf, name)
        if (module, name) == ("collections", "OrderedDict"):
            return collections.OrderedDict
        if module in _NUMPY_MODULES and name in _NUMPY_NAMES:
            return getattr(_numpy_module(module), name)
        if (module, name) == ("_codecs", "encode"):        # protocol-2 encoding of numpy scalar bytes
            import codecs
            return codecs.encode
        if module.endswith(".__dict__") or name == "__dict__":
            return {}
        if (module, name) == ("builtins", "getattr"):
            return _safe_getattr
        if module == "builtins" and name in ("tuple", "list", "dict", "set", "float", "int"):
            return getattr(importlib.import_module("builtins"), name)
        raise pickle.UnpicklingError("model_meta: refusing global %s.%s" % (module, name))


def _find_range_compress(fn):
    """Locate the range-compress closure (reference data_transforms.py:51-110)."""
    if isinstance(fn, _Function):
        if "k_values" in fn.code.freevars:
            return fn
        for c in fn.closure or ():
            hit = _find_range_compress(c.contents)
            if hit is not None:
                return hit
    elif isinstance(fn, (list, tuple)):
        for f in fn:
            hit = _find_range_compress(f)
            if hit is not None:
                return hit
    return None


def _chain_names(fn):
    """Names of the chained steps, in order (reference data_transforms.py:44-49)."""
    if isinstance(fn, _Function) and "transformations" in fn.code.freevars:
        out = []
        for t in fn.cell("transformations"):
            out.append(t.name if isinstance(t, _ByReference) else t.code.name)
        return out
    if isinstance(fn, _Function):
        for c in fn.closure or ():
            hit = _chain_names(c.contents)
            if hit:
                return hit
    return []


def _rebind(fn):
    """dill'd ``compile_transform`` lambda -> CompiledTransform."""
    if isinstance(fn, _tf.CompiledTransform) or fn is None:
        return fn
    if not isinstance(fn, _Function):
        raise ValueError("model_meta: transform entry is not a compiled transform")
    stats = fn.cell("s")
    rc = _find_range_compress(fn)
    if rc is None:
        raise NotImplementedError("model_meta: only range-compress transforms are supported")
    steps = _chain_names(fn)
    inverse = rc.code.name == "inv_transform"
    default_field, default_z = (fn.defaults or (None, None))[:2]
    return _tf.CompiledTransform(
        stats=_tf.normalise_stats(stats), k_values=dict(rc.cell("k_values")),
        modes=dict(rc.cell("modes")), eps=float(rc.cell("eps")),
        sqrt_of_mean=bool(rc.cell("sqrt_of_mean")), inverse=inverse, steps=tuple(steps),
        field=default_field, z=default_z)


def read_model_meta(filename):
    """Load a ``model_meta`` written by the reference (dill, any Python) or by
    :func:`write_model_meta`.  Returns a dict with the keys of reference
    ``painter.py:399-414``; ``transform`` / ``inverse_transform`` are callables
    ``f(x, field=None, z=None)``."""
    with open(filename, "rb") as f:
        d = _MetaUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(d, dict):
        raise ValueError("model_meta: expected a dict")
    for key in ("transform", "inverse_transform"):
        if key in d:
            d[key] = _rebind(d[key])
    return d


def write_model_meta(filename, d):
    """Write the meta dict as a plain pickle (protocol 2, loads without dill)."""
    with open(filename, "wb") as f:
        pickle.dump(dict(d), f, protocol=2)
