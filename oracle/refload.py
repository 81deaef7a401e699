"""Load the reference's own model classes WITHOUT importing its scripts.  TEST INFRASTRUCTURE ONLY.

The five reference scripts execute dataset I/O and k-fold training at import time
(others/realformer.py:41-42, cmu-mosei/run.py:45-46, Ren-MME/run.py:42, rencecps/run.py:83-84,
robot_demo.py:46), so they cannot be imported.  We parse each file with ``ast`` and ``exec`` only
its top-level ``class`` definitions, the loss ``def``s and the UPPER_CASE constant assignments, in a
namespace pre-seeded with numpy/torch.  Nothing from the reference is copied into this repo; the
source is read from ``$MMEMO_REF`` or ``/root/reference`` at call time and this module reports
``available() == False`` anywhere that tree does not exist (e.g. the GPU box).
"""
from __future__ import annotations

import ast
import math
import os
from types import SimpleNamespace

FILES = {
    "realformer": "others/realformer.py",
    "mosei": "cmu-mosei/run.py",
    "renmme": "Ren-MME/run.py",
    "rencecps": "rencecps/run.py",
    "robot": "robot_demo.py",
}
# loss functions + the pure host-side batch-assembly helpers (oracle/batching_oracle.py is pinned to them)
_LOSS_FUNCS = {"multi_circle_loss", "multi_loss", "masking", "audio_features", "text_features"}


def ref_root() -> str:
    return os.environ.get("MMEMO_REF", "/root/reference")


def available() -> bool:
    return all(os.path.isfile(os.path.join(ref_root(), f)) for f in FILES.values())


def _is_const_assign(node: ast.AST) -> bool:
    if not isinstance(node, ast.Assign) or len(node.targets) != 1:
        return False
    t = node.targets[0]
    if not isinstance(t, ast.Name) or not t.id.isupper():
        return False
    # keep only literal / arithmetic constants (e.g. rencecps ``DIM = 768*3``); skip paths etc.
    try:
        v = eval(compile(ast.Expression(node.value), "<const>", "eval"), {"__builtins__": {}}, {})
    except Exception:
        return False
    return isinstance(v, (int, float))


def load(which: str, device: str = "cpu", functions=(), inject=None, **overrides) -> SimpleNamespace:
    """Return a namespace holding the reference classes/constants of one script.

    ``functions``: extra top-level ``def``s to execute as they are (e.g. the training loops
    ``train`` / ``valid``); ``inject``: names placed in the namespace before anything runs (what the
    script would have imported or built at module level: ``tqdm``, a ``torch`` stand-in on boxes
    without CUDA, ...).

    ``overrides`` replace module-level constants *before* class construction (the classes read
    ``DROP``, ``FFN``, ``L_DIM`` ... from module globals at construction/forward time,
    others/realformer.py:139,159,164; cmu-mosei/run.py:210-212,221)."""
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    path = os.path.join(ref_root(), FILES[which])
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            keep.append(node)
        elif isinstance(node, ast.FunctionDef) and (node.name in _LOSS_FUNCS
                                                    or node.name in functions):
            keep.append(node)
        elif _is_const_assign(node):
            keep.append(node)
    ns = {"np": np, "torch": torch, "nn": nn, "F": F, "math": math,
          "device": torch.device(device)}
    ns.update(inject or {})
    # constants first so that overrides win, then classes
    consts = [n for n in keep if isinstance(n, ast.Assign)]
    others = [n for n in keep if not isinstance(n, ast.Assign)]
    exec(compile(ast.Module(body=consts, type_ignores=[]), path, "exec"), ns)
    ns.update(overrides)
    exec(compile(ast.Module(body=others, type_ignores=[]), path, "exec"), ns)
    return SimpleNamespace(**{k: v for k, v in ns.items() if not k.startswith("__")}, _ns=ns)
