"""Minimal stand-in for gin-config (absent from this image).

Covers what the reference touches: @gin.configurable, @gin.constants_from_enum, gin.parse_config_file with
`import x.y`, `fn.kw = <python literal>` and `%pkg.mod.Enum.NAME` macros (configs/*.gin)."""
import ast
import importlib
import functools

_BINDINGS = {}
_ENUMS = {}


def constants_from_enum(cls=None, module=None):
    def deco(c):
        mod = module or c.__module__
        for member in c:
            _ENUMS[f"{mod}.{c.__name__}.{member.name}"] = member
            _ENUMS[f"{c.__name__}.{member.name}"] = member
        return c
    return deco(cls) if cls is not None else deco


def configurable(fn=None, **_kw):
    def deco(f):
        name = f.__name__
        @functools.wraps(f)
        def wrapper(*a, **k):
            merged = dict(_BINDINGS.get(name, {}))
            merged.update(k)
            return f(*a, **merged)
        return wrapper
    return deco(fn) if callable(fn) else deco


def _resolve_macro(tok: str):
    key = tok[1:]
    if key in _ENUMS:
        return _ENUMS[key]
    tail = ".".join(key.split(".")[-2:])
    if tail in _ENUMS:
        return _ENUMS[tail]
    mod, cls, member = key.rsplit(".", 2)
    return getattr(getattr(importlib.import_module(mod), cls), member)


def parse_config(lines):
    if isinstance(lines, str):
        lines = lines.splitlines()
    for raw in lines:
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("import "):
            try:
                importlib.import_module(line[len("import "):].strip())
            except ImportError:
                pass
            continue
        lhs, rhs = (s.strip() for s in line.split("=", 1))
        fn, kw = lhs.rsplit(".", 1)
        val = _resolve_macro(rhs) if rhs.startswith("%") else ast.literal_eval(rhs)
        _BINDINGS.setdefault(fn.split(".")[-1], {})[kw] = val


def parse_config_file(path, *a, **k):
    with open(path) as f:
        parse_config(f.read())


def clear_config():
    _BINDINGS.clear()
