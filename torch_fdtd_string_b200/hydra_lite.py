"""Hydra-compatible config composition without Hydra / OmegaConf (neither is installable in this image).

Covers exactly what the reference's entry point uses (reference run.py:54 ``@hydra.main(config_path="src/configs",
config_name="config.yaml")``, src/configs/**): defaults lists (``_self_``, relative and absolute ``/group: name`` entries,
``null`` options, ``optional`` / ``override`` keywords, ``group@package`` keys), the ``# @package`` header, command-line
group overrides (``experiment=nsynth-like``), dotted value overrides (``task.num_samples=100``, ``+a.b=1``, ``~a.b``),
``${a.b}`` / ``${now:%fmt}`` / ``${eval:expr}`` / ``${oc.env:VAR}`` interpolation, the ``???`` missing marker and
``hydra.run.dir``.  Semantics follow Hydra 1.1+: a file's own content merges after its defaults unless ``_self_`` says
otherwise; later entries win; dictionaries merge recursively, lists are replaced.

    cfg = compose("/path/to/src/configs", "config.yaml", ["experiment=nsynth-like", "task.num_samples=100"])

``compose`` returns plain nested dicts/lists (``to_namespace`` gives the attribute-and-item access object the
reference's ``get_object`` builds, run.py:15-29).
"""
import copy
import datetime
import os
import re

import yaml

MISSING = "???"
_PKG_RE = re.compile(r"^#\s*@package\s+(\S+)\s*$")
_INTERP_RE = re.compile(r"\$\{([^{}]+)\}")


class ConfigError(Exception):
    pass


# ---------------------------------------------------------------------------------------------------------------
# small dict helpers
# ---------------------------------------------------------------------------------------------------------------
def merge(dst, src):
    """recursive dict merge, ``src`` wins; lists and scalars are replaced (OmegaConf.merge semantics)"""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _wrap(package, content):
    """content placed under the dotted ``package`` ('' = root)"""
    if not package:
        return content
    out = content
    for part in reversed(package.split(".")):
        out = {part: out}
    return out


def select(cfg, dotted, default=KeyError):
    cur = cfg
    for part in dotted.split("."):
        if isinstance(cur, list):
            try:
                cur = cur[int(part)]
                continue
            except (ValueError, IndexError):
                pass
        if not isinstance(cur, dict) or part not in cur:
            if default is KeyError:
                raise ConfigError(f"key '{dotted}' is not in the config")
            return default
        cur = cur[part]
    return cur


def assign(cfg, dotted, value, must_exist):
    parts = dotted.split(".")
    cur = cfg
    for part in parts[:-1]:
        if isinstance(cur, list):
            cur = cur[int(part)]
            continue
        if part not in cur or cur[part] is None:
            if must_exist:
                raise ConfigError(f"could not override '{dotted}': key '{part}' is not in the config (use +{dotted}=... to add it)")
            cur[part] = {}
        cur = cur[part]
    last = parts[-1]
    if isinstance(cur, list):
        cur[int(last)] = value
        return
    if must_exist and last not in cur:
        raise ConfigError(f"could not override '{dotted}': key '{last}' is not in the config (use +{dotted}=... to add it)")
    cur[last] = value


def delete(cfg, dotted):
    parts = dotted.split(".")
    cur = select(cfg, ".".join(parts[:-1])) if len(parts) > 1 else cfg
    if parts[-1] not in cur:
        raise ConfigError(f"could not delete '{dotted}': not in the config")
    del cur[parts[-1]]


# ---------------------------------------------------------------------------------------------------------------
# overrides
# ---------------------------------------------------------------------------------------------------------------
def parse_value(text):
    """command-line value -> python (Hydra's grammar is close to YAML flow syntax: null, true, 1, 1.5, [0,1], {a: 1}, str)"""
    if text == "":
        return ""
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        return text


class Overrides:
    def __init__(self, config_dir, items):
        self.groups = {}        # absolute group path ('experiment', 'model/excitation') -> option name (or None)
        self.added_groups = []  # (+group=name)
        self.values = []        # (op, dotted key, value): op in '=', '+', '++', '~'
        self.raw = list(items)
        for it in items:
            if it.startswith("~"):
                key = it[1:].split("=", 1)[0]
                if os.path.isdir(os.path.join(config_dir, key)):
                    self.groups[key.strip("/")] = None
                else:
                    self.values.append(("~", key, None))
                continue
            if "=" not in it:
                raise ConfigError(f"override '{it}' is not of the form key=value")
            key, val = it.split("=", 1)
            op = "="
            if key.startswith("++"):
                op, key = "++", key[2:]
            elif key.startswith("+"):
                op, key = "+", key[1:]
            gkey = key.split("@", 1)[0].strip("/")
            if os.path.isdir(os.path.join(config_dir, gkey)) and "." not in key:
                name = parse_value(val)
                if op == "=":
                    self.groups[gkey] = name
                else:
                    self.added_groups.append((gkey, name))
            else:
                self.values.append((op, key, parse_value(val)))
        self.used_groups = set()


# ---------------------------------------------------------------------------------------------------------------
# defaults-list composition
# ---------------------------------------------------------------------------------------------------------------
def _read(config_dir, rel):
    path = os.path.join(config_dir, rel)
    if not path.endswith((".yaml", ".yml")):
        path += ".yaml"
    if not os.path.isfile(path):
        return None, None
    with open(path) as f:
        text = f.read()
    header = None
    for line in text.splitlines():
        s = line.strip()
        if not s:
            continue
        m = _PKG_RE.match(s)
        if m:
            header = m.group(1)
        if not s.startswith("#"):
            break
    return (yaml.safe_load(text) or {}), header


def _resolve_package(header, key_pkg, group, parent_pkg):
    """package of a config file: key-level ``group@pkg`` > ``# @package`` header > the group path"""
    default = group.replace("/", ".")
    pkg = key_pkg if key_pkg is not None else (header if header is not None else "_group_")
    pkg = pkg.replace("_group_", default).replace("_name_", "")
    if pkg == "_global_":
        return ""
    if pkg.startswith("_global_."):
        return pkg[len("_global_."):]
    if key_pkg is not None and parent_pkg and not key_pkg.startswith("_global_") and "_group_" not in key_pkg:
        return f"{parent_pkg}.{pkg}"            # a key-level package is relative to the parent's package
    return pkg


def _load(config_dir, group, name, key_pkg, parent_pkg, ov, out, chain):
    rel = f"{group}/{name}" if group else name
    if rel in chain:
        raise ConfigError(f"defaults cycle through '{rel}'")
    content, header = _read(config_dir, rel)
    if content is None:
        raise ConfigError(f"config '{rel}' not found under {config_dir}" +
                          (f" (available: {sorted(os.path.splitext(x)[0] for x in os.listdir(os.path.join(config_dir, group)))})"
                           if os.path.isdir(os.path.join(config_dir, group)) else ""))
    package = _resolve_package(header, key_pkg, group, parent_pkg)
    defaults = content.pop("defaults", None) or []
    if "_self_" not in defaults:
        defaults = list(defaults) + ["_self_"]
    for entry in defaults:
        if entry == "_self_":
            merge(out, _wrap(package, content))
            continue
        if isinstance(entry, str):
            # a config of the same group, same package
            _load(config_dir, group, entry, key_pkg, parent_pkg, ov, out, chain + [rel])
            continue
        if not isinstance(entry, dict) or len(entry) != 1:
            raise ConfigError(f"unsupported defaults entry {entry!r} in '{rel}'")
        (key, option), = entry.items()
        words = key.split()
        optional = "optional" in words[:-1]
        key = words[-1]
        key, _, kp = key.partition("@")
        child_group = key.strip("/") if key.startswith("/") else (f"{group}/{key}" if group else key)
        if child_group in ov.groups:
            option = ov.groups[child_group]
            ov.used_groups.add(child_group)
        if option is None:
            continue
        if option == MISSING:
            raise ConfigError(f"you must specify '{child_group}', e.g. {child_group}=<option>")
        options = option if isinstance(option, list) else [option]
        for opt in options:
            crel = f"{child_group}/{opt}"
            if optional and _read(config_dir, crel)[0] is None:
                continue
            _load(config_dir, child_group, str(opt), kp or None, package, ov, out, chain + [rel])


# ---------------------------------------------------------------------------------------------------------------
# interpolation
# ---------------------------------------------------------------------------------------------------------------
def _resolver(kind, arg, root, now, stack):
    if kind == "now":
        return now.strftime(arg)
    if kind == "eval":
        return eval(arg)                           # the reference registers eval as a resolver (src/utils/config.py:137)
    if kind == "oc.env":
        var, _, dflt = arg.partition(",")
        if var in os.environ:
            return os.environ[var]
        if dflt:
            return parse_value(dflt.strip())
        raise ConfigError(f"environment variable '{var}' is not set")
    raise ConfigError(f"unsupported resolver '{kind}'")


def _resolve_str(s, root, now, stack):
    def one(expr):
        expr = expr.strip()
        if ":" in expr and not expr.startswith("."):
            kind, arg = expr.split(":", 1)
            return _resolver(kind, arg, root, now, stack)
        if expr in stack:
            raise ConfigError(f"interpolation cycle through '{expr}'")
        v = select(root, expr)
        if v == MISSING:
            raise ConfigError(f"interpolation '${{{expr}}}' points at a missing (???) value")
        return _resolve_node(v, root, now, stack + [expr])

    m = _INTERP_RE.fullmatch(s)
    if m:
        return one(m.group(1))                     # keeps the type of the target
    while True:
        m = _INTERP_RE.search(s)
        if not m:
            return s
        s = s[:m.start()] + str(one(m.group(1))) + s[m.end():]


def _resolve_node(node, root, now, stack):
    if isinstance(node, dict):
        return {k: _resolve_node(v, root, now, stack) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve_node(v, root, now, stack) for v in node]
    if isinstance(node, str) and "${" in node:
        return _resolve_str(node, root, now, stack)
    return node


def resolve(cfg, now=None):
    now = now or datetime.datetime.now()
    return _resolve_node(cfg, cfg, now, [])


# ---------------------------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------------------------
def compose(config_dir, config_name="config.yaml", overrides=(), now=None, resolve_interpolations=True):
    """-> (config dict without the ``hydra`` node, hydra node dict)"""
    config_dir = os.path.abspath(config_dir)
    if not os.path.isdir(config_dir):
        raise ConfigError(f"config directory {config_dir} does not exist")
    ov = Overrides(config_dir, list(overrides))
    out = {}
    _load(config_dir, "", os.path.splitext(config_name)[0], None, "", ov, out, [])
    for g, name in ov.added_groups:
        if name is not None:
            _load(config_dir, g, str(name), None, "", ov, out, [])
    unused = set(ov.groups) - ov.used_groups
    if unused:
        raise ConfigError(f"could not override {sorted(unused)}: no match in the defaults list (use +group=option to add one)")
    for op, key, val in ov.values:
        if op == "~":
            delete(out, key)
        else:
            assign(out, key, val, must_exist=(op == "="))
    if resolve_interpolations:
        out = resolve(out, now)
    hydra = out.pop("hydra", {}) or {}
    return out, hydra


def filter_keys(node, fn):
    """reference src/utils/config.py:108-121 (drops keys used only for interpolation, ``__*``)"""
    if isinstance(node, list):
        return [filter_keys(v, fn) for v in node]
    if isinstance(node, dict):
        return {k: filter_keys(v, fn) for k, v in node.items() if fn(k)}
    return node


class ConfigArgument:
    """attribute + item access, like the object reference run.py:15-29 builds from the DictConfig"""

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        return setattr(self, key, value)

    def keys(self):
        return self.__dict__.keys()

    def __repr__(self):
        return f"ConfigArgument({self.__dict__!r})"


def to_namespace(cfg, m=None):
    m = ConfigArgument() if m is None else m
    for key, val in cfg.items():
        if isinstance(val, dict):
            m[key] = to_namespace(val)
        else:
            m[key] = val
    return m
