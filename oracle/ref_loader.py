"""Loads the reference's own functions for this path from /root/reference (builder container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that runs
there may call this; it is used by oracle/make_golden.py (fixtures are committed) and by CPU tests
that skip when the tree is absent.

* 32_create_delegate_vector.py imports fine (its prompts are under ``main()``); it needs a
  package named ``qdrant_client`` on sys.path -- the repo's own drop-in provides it.
* 33_run_all_experiments.py prompts at import time, so ``cosine_similarity`` (lines 76-77) is
  pulled out of the AST and compiled on its own.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("RBOD_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "32_create_delegate_vector.py"))


def _import_script(filename: str, modname: str):
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def delegate_module():
    """The reference's 32_create_delegate_vector.py as a module (compute_average, ... :9-31)."""
    return _import_script("32_create_delegate_vector.py", "_ref_delegate")


def embed_module():
    """The reference's 31_clip_embedding_and_save_vector.py (generate_id_from_path :42-43)."""
    return _import_script("31_clip_embedding_and_save_vector.py", "_ref_embed")


def cosine_similarity():
    """The reference's cosine_similarity (33_run_all_experiments.py:76-77), AST-extracted."""
    import numpy as np

    path = os.path.join(REFERENCE_ROOT, "33_run_all_experiments.py")
    tree = ast.parse(open(path, encoding="utf-8").read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "cosine_similarity":
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns["cosine_similarity"]
    raise RuntimeError("cosine_similarity not found in the reference")
