"""L0 oracle: load the reference's own hot-path functions, verbatim.

Only usable where ``/root/reference`` exists (the dev container).  It is used
by ``oracle/make_golden.py`` to produce the committed fixtures and by the
``not gpu`` tests to validate ``oracle.matrix_oracle`` live.  Nothing on the GPU
box may call this (the reference tree does not travel).

The reference modules import three third-party packages that are not installed
here (``matplotlib``, ``pytorch_transformers``, ``pytorch_grad_cam``) plus
``cv2``/``helpers``; none of them is touched by the hot path, so they are
replaced by empty stub modules before import
(``mutual_info_img_txt/main_utils.py:1-25``, ``model.py:14-15``,
``model_utils.py:23-25``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MI_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mutual_info_img_txt", "mi_critics.py"))


class _Anything:
    """Stands in for any attribute of a stubbed third-party module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # behaves as a package

    def _getattr(attr, _name=name):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (_Anything,), {})

    mod.__getattr__ = _getattr  # PEP 562
    return mod


_STUBS = [
    "matplotlib", "matplotlib.pyplot",
    "pytorch_transformers", "pytorch_transformers.optimization",
    "pytorch_transformers.modeling_bert",
    "pytorch_grad_cam", "pytorch_grad_cam.utils",
    "pytorch_grad_cam.utils.model_targets", "pytorch_grad_cam.utils.image",
]


def _install_stubs() -> None:
    for name in _STUBS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _stub(name)
    for name in ("cv2",):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _stub(name)


_cache = {}


def load():
    """Returns a namespace with the reference's own callables:

    ``dv_bound_loss``, ``infonce_bound_loss`` (mi_critics.py:3-23),
    ``create_mi_pairs(X, Y, study_id, device)`` (main_utils.py:80-110, bound to
    an uninitialised ``MultiModalManager`` — the method reads no ``self`` state)
    and ``make_mlp`` (model.py:18-32).
    """
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mi_critics = importlib.import_module("mutual_info_img_txt.mi_critics")
    ns = types.SimpleNamespace(
        dv_bound_loss=mi_critics.dv_bound_loss,
        infonce_bound_loss=mi_critics.infonce_bound_loss,
        create_mi_pairs=None, make_mlp=None,
    )
    try:
        main_utils = importlib.import_module("mutual_info_img_txt.main_utils")
        mgr = main_utils.MultiModalManager.__new__(main_utils.MultiModalManager)
        ns.create_mi_pairs = mgr.create_mi_pairs
        ns.make_mlp = main_utils.make_mlp
    except Exception as exc:  # pragma: no cover - depends on container contents
        ns.import_error = repr(exc)
    _cache["ns"] = ns
    return ns


def load_gdv():
    """The reference's GDV functions (validate.py:16-49: ``z_scored_transform``, ``mean_intra_class_distance``,
    ``mean_inter_class_distance``, ``gdv_calculation``), verbatim.  ``validate.py`` runs its whole validation at import
    time (argument parsing, datasets, checkpoints: validate.py:55-171), so only these four function definitions are
    compiled out of its source — the function bodies are the reference's own, untouched."""
    if "gdv" in _cache:
        return _cache["gdv"]
    import ast
    import math
    import numpy as np
    from sklearn.metrics import pairwise_distances
    from sklearn.preprocessing import StandardScaler
    path = os.path.join(REFERENCE_ROOT, "validate.py")
    if not os.path.isfile(path):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    tree = ast.parse(open(path).read(), filename=path)
    wanted = {"z_scored_transform", "mean_intra_class_distance", "mean_inter_class_distance", "gdv_calculation"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    scope = {"math": math, "np": np, "pairwise_distances": pairwise_distances, "StandardScaler": StandardScaler}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), scope)
    ns = types.SimpleNamespace(**{k: scope[k] for k in wanted})
    _cache["gdv"] = ns
    return ns
