"""CPU: the product path has no CPU / torch fallback and never touches the oracle.

* every reference-facing entry point raises `TicError` when handed CPU tensors (instead of quietly computing with torch);
* no module of the package imports `oracle` (test infrastructure), and `bench.py` imports it only inside its CPU legs."""
import ast
import os
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "socialmedia-textimage-classification-auxlosses_b200")


def test_entry_points_refuse_cpu_tensors():
    from tic_b200 import capi
    from tic_b200 import mm_early
    from tic_b200.eval import compute_metrics
    from tic_b200.mm_late import MMLate_Model
    from tic_b200.utils import clip_loss
    ids = torch.arange(24).view(4, 6)
    with pytest.raises(capi.TicError):
        clip_loss(torch.zeros(4, 4))
    with pytest.raises(capi.TicError):
        mm_early.get_logits_per_text(torch.randn(4, 64), torch.randn(4, 64), torch.tensor(2.6592))
    with pytest.raises(capi.TicError):
        mm_early.itc_loss(torch.randn(4, 64), torch.randn(4, 64), torch.tensor(2.6592))
    with pytest.raises(capi.TicError):
        mm_early.prepare_itm_inputs(ids, torch.ones_like(ids), torch.zeros_like(ids))
    for rng in ("numpy", "device"):
        with pytest.raises(capi.TicError):
            MMLate_Model.prepare_itm_inputs(types.SimpleNamespace(), ids, torch.ones_like(ids), rng=rng)
    with pytest.raises(capi.TicError):
        compute_metrics({"predictions": torch.zeros(3, dtype=torch.int64), "labels": torch.zeros(3, dtype=torch.int64), "loss": 0.0}, 4)


def _imports(path):
    tree = ast.parse(open(path).read())
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            found += [(a.name, node.lineno) for a in node.names]
        elif isinstance(node, ast.ImportFrom):
            found.append((node.module or "", node.lineno))
    return found


def test_package_never_imports_the_oracle():
    for fn in sorted(os.listdir(PKG)):
        if fn.endswith(".py"):
            bad = [(m, ln) for m, ln in _imports(os.path.join(PKG, fn)) if m.split(".")[0] == "oracle"]
            assert not bad, (fn, bad)


def test_bench_imports_the_oracle_only_in_its_cpu_leg():
    path = os.path.join(ROOT, "bench.py")
    tree = ast.parse(open(path).read())
    owners = []
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                owners.append(fn.name)
    top = [m for m, _ in _imports(path) if m.split(".")[0] == "oracle"]
    assert owners == ["cpu_reference_run"] and len(top) == 1, (owners, top)
