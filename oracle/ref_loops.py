"""The reference's OWN loop functions, imported unmodified -- TEST / BASELINE INFRASTRUCTURE ONLY.

    train_link.tgb_train / tgb_test / train / eval      train_link.py:206, 24, 287, 349
    train_sequence.train / eval                         train_sequence.py:16, 68
    negative_sampler.NegativeSampler                    negative_sampler.py:8-48
    TGB Evaluator, NegativeEdgeSampler                  TGB/tgb/linkproppred/{evaluate,negative_sampler}.py

The reference is pure Python; its operators come from un-vendored PyG (oracle/pyg_shim.py restates those).
Nothing here is reference source: `sync_reference_files()` (run by `__graft_entry__.build()` in the build
container, where /root/reference exists) mirrors the handful of files the loops need into the git-ignored
`baseline/_ref/`, which travels to the GPU box with the snapshot the way the built `.so` does; `load()` then
imports them from /root/reference or, failing that, from `baseline/_ref/`.

Only tests/, smoke() and bench.py's baseline legs (`--impl reference`, `cpu_baseline`, the GPU-PyTorch
baseline) may import this module; the product never does (tests/test_cabi_symbols.py checks)."""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MIRROR = os.path.join(ROOT, "baseline", "_ref")
SOURCE = "/root/reference"

# the files the loops above import (everything else of the reference -- experiment drivers, Ray, datasets,
# confs, plotting -- is stubbed in load())
FILES = [
    "train_link.py", "train_sequence.py", "utils.py", "negative_sampler.py",
    "models/__init__.py", "models/ctdg_models.py", "models/gnn_layers.py", "models/memory_layers.py",
    "models/predictors.py", "models/layers.py",
    "TGB/modules/neighbor_loader.py", "TGB/modules/time_enc.py", "TGB/modules/msg_agg.py",
    "TGB/tgb/linkproppred/evaluate.py", "TGB/tgb/linkproppred/negative_sampler.py",
    "LICENSE",
]


def sync_reference_files(verbose: bool = False) -> bool:
    """Mirror FILES from /root/reference into baseline/_ref/ (git-ignored).  No-op when the reference is absent."""
    if not os.path.isdir(os.path.join(SOURCE, "models")):
        return False
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(MIRROR, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copy2(src, dst)
            if verbose:
                print("mirrored", rel)
    return True


def reference_root():
    for r in (os.environ.get("CTAN_REFERENCE_ROOT"), SOURCE, MIRROR):
        if r and os.path.isdir(os.path.join(r, "models")) and os.path.exists(os.path.join(r, "train_link.py")):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


_loaded = None


def load():
    """Import the reference's loop modules (unmodified) with the non-path dependencies stubbed.  Returns a namespace:
    train_link, train_sequence, negative_sampler, utils, Evaluator, NegativeEdgeSampler, plus pyg_shim.install()'s
    classes under `.ns` (reference CTAN / CTAN_tgb / LastNeighborLoader / TemporalData / TemporalDataLoader)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference loops unavailable: neither /root/reference nor baseline/_ref/ is present "
                           "(run `python __graft_entry__.py build` where /root/reference exists)")
    from oracle import pyg_shim
    pyg_shim.REFERENCE_ROOT = root
    ns = pyg_shim.install()

    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            try:
                m = importlib.import_module(name)
            except Exception:
                m = types.ModuleType(name)
                sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    stub("ray", remote=lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f)))
    for name in ("wandb", "clint", "clint.textui", "matplotlib", "matplotlib.pyplot"):
        stub(name)
    stub("tgb")
    stub("tgb.utils")
    stub("tgb.utils.info", DATA_EVAL_METRIC_DICT={"tgbl-wiki": "mrr", "tgbl-review": "mrr", "tgbl-coin": "mrr",
                                                  "tgbl-comment": "mrr", "tgbl-flight": "mrr"}, PROJ_DIR=root)
    stub("tgb.utils.utils", save_pkl=None, load_pkl=None)
    stub("tgb.linkproppred")
    stub("tgb.linkproppred.dataset_pyg", PyGLinkPropPredDataset=object)
    stub("torch_geometric.datasets", JODIEDataset=object)
    # `datasets` / `confs` are the reference's experiment plumbing (dataset download, grids): not on the path
    sys.modules["datasets"] = types.ModuleType("datasets")
    sys.modules["datasets"].get_dataset = None
    stub("confs")
    sys.modules["confs.conf_link_prediction"] = types.ModuleType("confs.conf_link_prediction")
    sys.modules["confs.conf_link_prediction"].edgebank = "EdgeBank"
    ev = pyg_shim._load_file("tgb.linkproppred.evaluate", os.path.join(root, "TGB", "tgb", "linkproppred", "evaluate.py"))
    nes = pyg_shim._load_file("tgb.linkproppred.negative_sampler",
                              os.path.join(root, "TGB", "tgb", "linkproppred", "negative_sampler.py"))
    sys.modules["models"] = ns.models
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("utils", "negative_sampler", "train_link", "train_sequence"):
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(root):
            del sys.modules[name]                 # a same-named module of ours / of a test must not shadow the reference's
    import utils as ref_utils
    import negative_sampler as ref_neg
    import train_link as ref_tl
    import train_sequence as ref_ts
    _loaded = types.SimpleNamespace(ns=ns, root=root, train_link=ref_tl, train_sequence=ref_ts, utils=ref_utils,
                                    negative_sampler=ref_neg, Evaluator=ev.Evaluator,
                                    NegativeEdgeSampler=nes.NegativeEdgeSampler, TemporalData=ns.TemporalData,
                                    TemporalDataLoader=ns.TemporalDataLoader)
    return _loaded


class FixedNegatives:
    """`train_neg_sampler` / `neg_sampler` stand-in that feeds PRE-DRAWN negatives, so that every implementation sees
    the same ones: `sample(src)` returns the next src.numel() entries of `neg` (one per event)."""

    def __init__(self, neg, start: int = 0):
        self.neg, self.pos, self.start = neg, start, start

    def reset(self):
        self.pos = self.start

    def sample(self, src, eval=False, eval_seed=None, *a, **k):
        n = src.shape[0]
        out = self.neg[self.pos:self.pos + n].to(src.device)
        self.pos += n
        return out

    def __repr__(self):
        return "FixedNegatives"


class FixedEvalNegatives:
    """`neg_sampler.query_batch(src, dst, t, split_mode)` stand-in over a dense [n_events, n_neg] table."""

    def __init__(self, neg, start: int = 0):
        self.neg, self.pos = neg, start

    def query_batch(self, pos_src, pos_dst, pos_t, split_mode="val"):
        n = pos_src.shape[0]
        rows = self.neg[self.pos:self.pos + n].tolist()
        self.pos += n
        return rows


class TimedLoader:
    """Wraps a TemporalDataLoader: records a timestamp when batch `skip` is requested and when the iteration ends, so
    the reference's loop function runs UNMODIFIED and we still get "K steps after W warm-up" out of one call."""

    def __init__(self, loader, skip: int, sync=None):
        self.loader, self.skip, self.sync = loader, skip, sync
        self.t0 = self.t1 = None
        self.n_timed = 0

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        n = 0
        for j, b in enumerate(self.loader):
            if j == self.skip:
                if self.sync:
                    self.sync()
                self.t0 = time.perf_counter()
            yield b
            n = j + 1
        if self.sync:
            self.sync()
        self.t1 = time.perf_counter()
        self.n_timed = n - self.skip

    @property
    def seconds(self):
        return self.t1 - self.t0


class legacy_torch_load:
    """The reference targets torch 2.0.1, where `torch.load` un-pickles arbitrary objects; its drivers save numpy scalars
    in their checkpoints (train_link.py:587-603) and re-load them with a bare `torch.load(path, map_location=...)`.
    While one of its drivers runs under torch >= 2.6, restore that default (weights_only=False).  Nothing else changes."""

    def __enter__(self):
        import torch
        self._torch, self._orig = torch, torch.load

        def load(*a, **k):
            k.setdefault("weights_only", False)
            return self._orig(*a, **k)
        torch.load = load
        return self

    def __exit__(self, *exc):
        self._torch.load = self._orig
        return False
