"""Registry / config plumbing with the semantics the reference gets from mmcv.

The reference builds its generator with ``build_backbone(cfg.model.generator)`` ->
``build_from_cfg(cfg, BACKBONES)`` (mmedit/models/builder.py:43-66) where ``BACKBONES is MODELS``
is one ``mmcv.utils.Registry`` (mmedit/models/registry.py:5-8), and reads python-file configs with
``mmcv.Config.fromfile`` (``_base_`` inheritance, tools/test.py:67).  mmcv cannot be installed in
this environment, so the same three behaviours are provided here; when the real ``mmedit`` package
is importable the class is *also* registered there (``force=True``) so the reference's own
``tools/test.py`` builds the B200 generator from the unmodified configs.
"""
import copy
import os


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name, force):
        key = name or cls.__name__
        if not force and key in self._module_dict:
            raise KeyError(f"{key} is already registered in {self._name}")
        self._module_dict[key] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls
        return deco


MODELS = Registry("model")
BACKBONES = MODELS
COMPONENTS = MODELS
LOSSES = MODELS


def build_from_cfg(cfg, registry, default_args=None):
    """``type``-keyed construction (mmcv.utils.build_from_cfg semantics incl. its errors)."""
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
    if "type" not in cfg and not (default_args and "type" in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = copy.deepcopy(dict(cfg))
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    elif isinstance(obj_type, type):
        obj_cls = obj_type
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(obj_type)}")
    return obj_cls(**args)


def build_backbone(cfg):
    """mmedit/models/builder.py:60-66"""
    return build_from_cfg(cfg, BACKBONES)


class ConfigDict(dict):
    """dict with attribute access (like mmcv's ConfigDict)."""

    def __getattr__(self, name):
        try:
            v = self[name]
        except KeyError as e:
            raise AttributeError(name) from e
        return v

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_wrap(v) for v in obj)
    return obj


def _merge(base, child):
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
            out[k] = v
    return out


class Config:
    """Python-file configs with ``_base_`` inheritance (the subset of mmcv.Config the path needs)."""

    @staticmethod
    def _load(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError(f'file "{filename}" does not exist')
        scope = {"__file__": filename}
        with open(filename) as f:
            exec(compile(f.read(), filename, "exec"), scope)
        cfg = {k: v for k, v in scope.items()
               if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
        base = cfg.pop("_base_", None)
        if base is not None:
            bases = base if isinstance(base, (list, tuple)) else [base]
            merged = {}
            for b in bases:
                merged = _merge(merged, Config._load(os.path.join(os.path.dirname(filename), b)))
            cfg = _merge(merged, cfg)
        return cfg

    @staticmethod
    def fromfile(filename):
        return _wrap(Config._load(filename))


def register_with_mmedit(cls):
    """Also register in the real mmedit registry when it is importable (reference-side drop-in)."""
    import sys
    if getattr(sys.modules.get("mmcv"), "__pnp_stub__", False):
        return False          # the test-only mmcv stub of oracle/refshim.py, not a real install
    try:
        from mmedit.models.registry import BACKBONES as MM_BACKBONES  # type: ignore
    except Exception:
        return False
    try:
        MM_BACKBONES.register_module(name=cls.__name__, force=True, module=cls)
    except Exception:
        return False
    return True
