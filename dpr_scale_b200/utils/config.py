"""Hydra-style config plumbing without Hydra (absent from this image): ``_target_`` instantiation and a
small YAML-group composer for the ``conf/`` tree, mirroring how /root/reference/dpr_scale/main.py:20-29 and
conf/config.py:9-24 compose ``task`` / ``task/model`` / ``task/transform`` / ``task/optim`` / ``datamodule`` /
``trainer`` groups and apply ``a.b=c`` overrides.  When real Hydra is importable, ``instantiate`` defers to it.
"""
import copy
import importlib
import os
import re

import yaml

CONF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conf")


class AttrDict(dict):
    """dict with attribute access (DictConfig-like enough for the task code)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(x):
    if isinstance(x, dict):
        return AttrDict({k: to_attr(v) for k, v in x.items()})
    if isinstance(x, list):
        return [to_attr(v) for v in x]
    return x


def _locate(path):
    parts = path.split(".")
    for i in range(len(parts) - 1, 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ModuleNotFoundError:
            continue
        for p in parts[i:]:
            obj = getattr(obj, p)
        return obj
    raise ImportError(f"cannot locate {path!r}")


def instantiate(conf, *args, _recursive_=True, **kwargs):
    """``hydra.utils.instantiate`` semantics for the subset the reference uses (positional passthrough,
    kwargs override, optional recursion into nested ``_target_`` nodes)."""
    try:
        import hydra  # noqa
        from omegaconf import DictConfig  # noqa
        if isinstance(conf, DictConfig):
            return hydra.utils.instantiate(conf, *args, _recursive_=_recursive_, **kwargs)
    except Exception:  # noqa
        pass
    if conf is None:
        return None
    conf = dict(conf)
    target = conf.pop("_target_")
    conf.pop("_recursive_", None)
    params = {}
    for k, v in conf.items():
        if _recursive_ and isinstance(v, dict) and "_target_" in v:
            v = instantiate(v)
        params[k] = v
    params.update(kwargs)
    fn = _locate(target) if isinstance(target, str) else target
    return fn(*args, **params)


def _load_yaml(path):
    with open(path) as f:
        return yaml.safe_load(f) or {}


def _set_path(d, dotted, value):
    parts = dotted.split(".")
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = value


def _get_path(d, dotted):
    for p in dotted.split("."):
        d = d[p]
    return d


_INTERP = re.compile(r"\$\{([^}]+)\}")


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node)
        if m:
            return _resolve(_get_path(root, m.group(1)), root)
        return _INTERP.sub(lambda mm: str(_resolve(_get_path(root, mm.group(1)), root)), node)
    return node


DEFAULT_GROUPS = [  # conf/config.py:9-24 of the reference
    ("task", "dpr"), ("task/model", "hf_model"), ("task/transform", "hf_transform"), ("task/optim", "adamw"),
    ("datamodule", "default"), ("trainer", "gpu_1_host"), ("checkpoint_callback", "default"),
]


def compose(config_name=None, overrides=(), conf_dir=CONF_DIR):
    """Compose the default groups (+ an experiment YAML) and apply ``a.b=c`` / ``group=name`` overrides."""
    groups = dict(DEFAULT_GROUPS)
    exp = {}
    if config_name and config_name != "config":
        exp = _load_yaml(os.path.join(conf_dir, config_name + ".yaml"))
        for d in exp.pop("defaults", []):
            if isinstance(d, dict):
                for k, v in d.items():
                    groups[k.replace("override ", "").strip()] = v
    value_overrides = []
    for o in overrides:
        k, _, v = o.partition("=")
        k = k.lstrip("+")
        if k in groups or k.replace(".", "/") in groups:
            groups[k.replace(".", "/")] = v
        else:
            value_overrides.append((k, yaml.safe_load(v)))
    cfg = {"test_only": False}
    for g, name in groups.items():
        path = os.path.join(conf_dir, g, str(name) + ".yaml")
        if os.path.exists(path):
            _set_path(cfg, g.replace("/", "."), _merge(_get_or_empty(cfg, g.replace("/", ".")), _load_yaml(path)))

    cfg = _merge(cfg, exp)
    for k, v in value_overrides:
        _set_path(cfg, k, v)
    return to_attr(_resolve(cfg, cfg))


def _get_or_empty(d, dotted):
    try:
        v = _get_path(d, dotted)
        return v if isinstance(v, dict) else {}
    except KeyError:
        return {}


def _merge(a, b):
    out = copy.deepcopy(a)
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out
