"""Import the UNMODIFIED reference modules from /root/reference on top of oracle/pyg_shim.

TEST INFRASTRUCTURE, build-container only (``/root/reference`` does not exist on the GPU box).
Used by ``oracle/make_golden.py`` and by the container-only tests that compare the
restatement in ``oracle/graphvqa_oracle.py`` with the reference's own code.

What is stubbed, and why:
* ``torch_geometric`` / ``torch_scatter`` / ``torch_sparse``  -> ``oracle/pyg_shim`` (absent wheels);
* ``gqa_dataset_entry``  -> a module object exposing only the class attributes the model files
  read at construction time (``GQATorchDataset.TEXT.vocab``, ``MAX_EXECUTION_STEP``,
  ``GQA_gt_sg_feature_lookup.SG_ENCODING_TEXT.vocab``); the real module needs legacy
  torchtext, spaCy, a GloVe download and a hard-coded ``/home/weixin`` path.
Nothing from the reference is copied: its files are executed where they lie.
"""
import importlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("GVQA_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_shim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TEXT_VOCAB_SIZE = 3657      # questions/GQA_TEXT_obj.pkl vocabulary
SG_VOCAB_SIZE = 2577        # rebuilt scene-graph vocabulary incl. specials


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gat_skip.py"))


class _Vocab:
    def __init__(self, size, dim, seed):
        self._size = size
        self.stoi = {"<unk>": 0, "<pad>": 1, "<start>": 2, "<end>": 3}
        g = torch.Generator().manual_seed(seed)
        self.vectors = torch.randn(size, dim, generator=g) * 0.4   # stand-in for GloVe rows

    def __len__(self):
        return self._size


class _Field:
    pad_token, init_token, eos_token = "<pad>", "<start>", "<end>"

    def __init__(self, size, seed):
        self.vocab = _Vocab(size, 300, seed)


def _dataset_stub():
    mod = types.ModuleType("gqa_dataset_entry")

    class GQATorchDataset:
        TEXT = _Field(TEXT_VOCAB_SIZE, 11)
        MAX_EXECUTION_STEP = 5

    class GQA_gt_sg_feature_lookup:
        SG_ENCODING_TEXT = _Field(SG_VOCAB_SIZE, 12)

    mod.GQATorchDataset = GQATorchDataset
    mod.GQA_gt_sg_feature_lookup = GQA_gt_sg_feature_lookup
    mod.GQATorchDataset_collate_fn = None
    return mod


_loaded = {}
_TT_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torchtext_shim")
_HARD_CODED_ROOT = "/home/weixin/neuralPoolTest/GraphVQA"     # Constants.py:13 + 'GraphVQA' (gqa_dataset_entry.py:28-30)


class _redirect_reference_paths:
    """While active, ``open()`` of a path under the reference's hard-coded checkout location reads the same file
    under REFERENCE_ROOT instead (Constants.py and gqa_dataset_entry.py open their meta_info files at import)."""

    def __enter__(self):
        import builtins
        self._builtins, self._open = builtins, builtins.open

        def redirected(file, *args, **kwargs):
            try:
                path = os.fspath(file)
            except TypeError:
                return self._open(file, *args, **kwargs)
            if isinstance(path, str) and path.startswith(_HARD_CODED_ROOT):
                path = REFERENCE_ROOT + path[len(_HARD_CODED_ROOT):]
            return self._open(path, *args, **kwargs)

        builtins.open = redirected
        return self

    def __exit__(self, *exc):
        self._builtins.open = self._open
        return False


def load_scene_graph_lookup(split="debug"):
    """The reference's own ``GQA_gt_sg_feature_lookup(split)`` (gqa_dataset_entry.py:52-372), built from the
    UNMODIFIED file on top of oracle/torchtext_shim + oracle/pyg_shim: its scene-graph vocabulary and its
    ``convert_one_gqa_scene_graph``.  matplotlib (imported by Constants.py for plotting helpers) is stubbed."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "scene_graph_lookup" in _loaded:
        return _loaded["scene_graph_lookup"]
    for p in (_SHIM, _TT_SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        try:
            importlib.import_module(name)
        except ImportError:
            sys.modules[name] = types.ModuleType(name)
    with _redirect_reference_paths():
        spec = importlib.util.spec_from_file_location("gqa_dataset_entry_reference",
                                                      os.path.join(REFERENCE_ROOT, "gqa_dataset_entry.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)              # imports the reference's Constants.py under the redirect, too
        lookup = mod.GQA_gt_sg_feature_lookup(split)
    _loaded["scene_graph_lookup"] = (mod, lookup)
    return mod, lookup


def load(name):
    """Return a reference module by file stem, e.g. 'gat_skip', 'lcgn', 'pipeline_model_gat',
    'pipeline_model_gine', 'graph_utils.my_graph_layernorm'."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if name in _loaded:
        return _loaded[name]
    for p in (_REPO, _SHIM, REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "baseline_and_test_models")):
        if p not in sys.path:
            sys.path.insert(0, p) if p == _SHIM else sys.path.append(p)
    sys.modules.setdefault("gqa_dataset_entry", _dataset_stub())
    mod = importlib.import_module(name)
    origin = os.path.realpath(getattr(mod, "__file__", ""))
    assert origin.startswith(os.path.realpath(REFERENCE_ROOT)), (name, origin)
    _loaded[name] = mod
    return mod
