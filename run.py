#!/usr/bin/env python
"""Entry point with the call shape of the reference's `run.py` (Hydra @main over `configs/`):

    python run.py lightning_datamodule=bwe lightning_module=eben ++trainer.max_steps=200 \
        lightning_module.generator.p=1 lightning_datamodule.batch_size=16

Hydra / OmegaConf / Lightning are not installed in this image, so the composition rules used by this path
are restated here: `defaults` lists with `group@package: choice`, `group=choice` / `a.b.c=value` / `+k=v` /
`++k=v` overrides, `${...}` interpolation of top-level keys, `_target_` / `_partial_` / `_args_`
instantiation.  When Hydra IS available the drop-in classes work unmodified inside the reference's own
run.py through `_target_` overrides."""
from __future__ import annotations

import functools
import importlib
import os
import re
import sys

import yaml

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CONFIGS = os.path.join(ROOT, "configs")


_FLOAT = re.compile(r"[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)")


def _numbers(node):
    """PyYAML (YAML 1.1) reads `3e-4` as a string; OmegaConf reads it as a float - follow OmegaConf."""
    if isinstance(node, dict):
        return {k: _numbers(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_numbers(v) for v in node]
    if isinstance(node, str) and _FLOAT.fullmatch(node):
        return float(node)
    return node


def _load(group: str, choice: str) -> dict:
    path = os.path.join(CONFIGS, group, choice + ".yaml") if group else os.path.join(CONFIGS, choice + ".yaml")
    with open(path) as f:
        return _numbers(yaml.safe_load(f) or {})


def _compose(group: str, choice: str, group_overrides: dict) -> dict:
    cfg = _load(group, choice)
    out = {}
    for item in cfg.pop("defaults", []) or []:
        if item == "_self_":
            continue
        (key, val), = item.items()
        sub, _, package = key.partition("@")
        package = package or sub
        full_group = f"{group}/{sub}" if group else sub
        val = group_overrides.get(package if not group else f"{group}.{package}", group_overrides.get(package, val))
        if val is None:
            raise SystemExit(f"config group '{package}' must be chosen on the command line ({package}=...)")
        out[package] = _compose(full_group, str(val), group_overrides)
    out.update(cfg)
    return out


def _set(cfg: dict, dotted: str, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        cfg = cfg.setdefault(k, {})
    cfg[keys[-1]] = value


def _interp(node, root):
    if isinstance(node, dict):
        return {k: _interp(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_interp(v, root) for v in node]
    if isinstance(node, str):
        def rep(m):
            cur = root
            for k in m.group(1).split("."):
                cur = cur[k]
            return str(cur)
        full = re.fullmatch(r"\$\{([\w\.]+)\}", node)
        if full:
            cur = root
            for k in full.group(1).split("."):
                cur = cur[k]
            return cur
        return re.sub(r"\$\{([\w\.]+)\}", rep, node)
    return node


def instantiate(node, **extra):
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    node = dict(node)
    module, _, name = node.pop("_target_").rpartition(".")
    fn = getattr(importlib.import_module(module), name)
    partial = node.pop("_partial_", False)
    args = [instantiate(a) for a in node.pop("_args_", [])]
    kwargs = {k: instantiate(v) for k, v in node.items()}
    kwargs.update(extra)
    return functools.partial(fn, *args, **kwargs) if partial else fn(*args, **kwargs)


def compose(argv) -> dict:
    group_choice, value_overrides = {}, []
    for arg in argv:
        key, _, val = arg.lstrip("+").partition("=")
        if not _:
            raise SystemExit(f"cannot parse override '{arg}'")
        if "." not in key and os.path.isdir(os.path.join(CONFIGS, key)):
            group_choice[key] = val
        else:
            value_overrides.append((key, _numbers(yaml.safe_load(val))))
    cfg = _compose("", "run", group_choice)
    for k, v in value_overrides:
        _set(cfg, k, v)
    return _interp(cfg, cfg)


def main(argv=None):
    cfg = compose(sys.argv[1:] if argv is None else argv)
    import torch
    torch.manual_seed(42)                                   # run.py:74 seed_everything(42)
    datamodule = instantiate(cfg["lightning_datamodule"])
    module = instantiate(cfg["lightning_module"])
    trainer = instantiate(cfg["trainer"])
    # `ckpt_path=last` (or a file) resumes parameters, both Adam states and the balancing EMA, like
    # `trainer.fit(..., ckpt_path=...)` of Lightning; `trainer.default_root_dir=...` is where last.ckpt is written
    trainer.fit(module, datamodule, ckpt_path=cfg.get("ckpt_path"))
    return module


if __name__ == "__main__":
    main()
