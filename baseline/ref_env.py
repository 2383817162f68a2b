"""Run-time environment for executing the UNMODIFIED reference (baseline/_ref, or $TCVOM_REFERENCE) in this image.

Nothing of the reference is edited.  The image lacks three of its imports (matplotlib, yacs, imgaug) and its dataset
(VideoMatting108 is not available offline), so this module provides:

  * empty ``matplotlib`` / ``matplotlib.pyplot`` / ``imgaug`` stubs (imported at module import only:
    models/Index/hldecoder.py:36, dataset/VMD.py:9-15);
  * a minimal ``yacs.config.CfgNode`` (clone / merge_from_file / merge_from_list / freeze / attribute access -- what
    config.py:1-44 and train_ddp.py:364-369 use);
  * ``np.int`` (train_ddp.py:348 uses the alias numpy 2 removed);
  * optionally a synthetic stand-in for ``dataset.VMD.VideoMattingDataset`` with the same constructor and item layout
    (dataset/VMD.py:300-301: fg, bg [S,3,H,W], a [S,1,H,W] float 0..255, idx) built from tcvom_b200.synthetic;
  * for CPU runs, ``torch.cuda.current_device`` -> cpu (VMN_model.py:47,54 hard-code CUDA; SURVEY.md section 8c shim 2).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def ref_dir():
    for d in (os.path.join(HERE, "_ref"), os.environ.get("TCVOM_REFERENCE", "/root/reference")):
        if d and os.path.isdir(os.path.join(d, "models")):
            return d
    return None


def available() -> bool:
    return ref_dir() is not None


class CfgNode(dict):
    """The subset of yacs.config.CfgNode the reference uses."""
    _frozen = False

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        if k == "_frozen":
            object.__setattr__(self, k, v)
            return
        if self._frozen:
            raise AttributeError("CfgNode is frozen")
        self[k] = v

    def clone(self):
        out = CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, CfgNode) else v
        return out

    def _merge(self, d):
        for k, v in d.items():
            if k not in self:
                raise KeyError(f"non-existent config key: {k}")
            if isinstance(self[k], CfgNode):
                self[k]._merge(v)
                continue
            if isinstance(v, str) and not isinstance(self[k], str):     # yacs decodes "(512, 512)" / "1e-4" literals
                import ast
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            if isinstance(self[k], tuple) and isinstance(v, list):
                v = tuple(v)
            if isinstance(self[k], float) and isinstance(v, int):
                v = float(v)
            self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        import ast
        opts = list(opts or [])
        assert len(opts) % 2 == 0, "opts must be KEY VALUE pairs"
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
            node[parts[-1]] = v

    def freeze(self):
        self._frozen = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def __str__(self):
        import yaml

        def plain(n):
            return {k: plain(v) if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v) for k, v in n.items()}
        return yaml.safe_dump(plain(self))


def _synthetic_dataset_module():
    import torch
    from torch.utils import data

    class VideoMattingDataset(data.Dataset):
        """Synthetic stand-in with the constructor and item layout of dataset/VMD.py (train_ddp.py:224-249)."""

        def __init__(self, data_root, image_shape, mode="train", use_subset=False, plus1=False, no_flow=True,
                     sample_length=5, **kw):
            self.image_shape = tuple(image_shape)
            self.sample_length = sample_length
            self.n = int(os.environ.get("TCVOM_STUB_DATASET_LEN", "4"))
            self.samples = [[f"synthetic/{i:04d}_{s}" for s in range(sample_length)] for i in range(self.n)]

        def __len__(self):
            return self.n

        def __getitem__(self, idx):
            from tcvom_b200 import synthetic
            H, W = self.image_shape
            a, fg, bg = synthetic.make_train_batch(1, self.sample_length, H, W, seed=21 + 7 * idx)
            f = lambda t: torch.from_numpy(t[0]).float()
            return f(fg), f(bg), f(a), torch.tensor(idx)

    m = types.ModuleType("dataset.VMD")
    m.VideoMattingDataset = VideoMattingDataset
    return m


def _synthetic_dim_dataset_module():
    import torch
    from torch.utils import data
    from tcvom_b200 import synthetic

    def item(H, W, idx):
        a, fg, bg = synthetic.make_train_batch(1, 3, H, W, seed=51 + 7 * idx)
        f = lambda t: torch.from_numpy(t[0]).float()
        return f(a), f(fg), f(bg)

    class DIMPretrainDataset(data.Dataset):
        """Synthetic stand-in with the constructor and item layout of dataset/DIM.py (pretrain_ddp.py:190-196,52-53):
        3-frame samples (a [3,1,H,W], fg, bg [3,3,H,W])."""

        def __init__(self, data_root, image_shape, min_shape=None, isTrain=True, plus1=False, **kw):
            self.image_shape = tuple(image_shape)
            self.n = int(os.environ.get("TCVOM_STUB_DATASET_LEN", "4"))

        def __len__(self):
            return self.n

        def __getitem__(self, idx):
            return item(*self.image_shape, idx)

    class DIMEvalDataset(data.Dataset):
        """pretrain_ddp.py:207-213,112: items are 5-tuples (a, fg, bg, name, size)."""

        def __init__(self, data_root, min_shape=None, plus1=False, val_mode="origin", **kw):
            self.n = int(os.environ.get("TCVOM_STUB_EVAL_LEN", "1"))

        def __len__(self):
            return self.n

        def __getitem__(self, idx):
            a, fg, bg = item(64, 64, 100 + idx)
            return a, fg, bg, f"synthetic/{idx:04d}", torch.tensor([64, 64])

    m = types.ModuleType("dataset.DIM")
    m.DIMPretrainDataset, m.DIMEvalDataset = DIMPretrainDataset, DIMEvalDataset
    return m


import contextlib


@contextlib.contextmanager
def on_cpu():
    """VMN_model.py:47,54 move TAM's outputs to ``torch.cuda.current_device()``: while the reference runs on the CPU that
    call must name the CPU (SURVEY.md section 8c shim 2).  Restored on exit so that later CUDA work is unaffected."""
    import torch
    orig = torch.cuda.current_device
    torch.cuda.current_device = lambda: torch.device("cpu")
    try:
        yield
    finally:
        torch.cuda.current_device = orig


def activate(cpu: bool = False, synthetic_dataset: bool = False) -> str:
    """Makes ``import models.model`` (and the scripts' other imports) work; returns the reference directory."""
    d = ref_dir()
    if d is None:
        raise RuntimeError("no reference checkout (run baseline/install_ref.py in the build container)")
    for name in ("matplotlib", "matplotlib.pyplot", "imgaug", "imgaug.augmenters", "imgaug.parameters"):
        try:
            __import__(name)
        except ImportError:
            sys.modules.setdefault(name, types.ModuleType(name))
    try:
        import yacs.config  # noqa: F401
    except ImportError:
        y, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")
        yc.CfgNode = CfgNode
        y.config = yc
        sys.modules["yacs"], sys.modules["yacs.config"] = y, yc
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if d not in sys.path:
        sys.path.insert(0, d)
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    if synthetic_dataset:
        pkg = types.ModuleType("dataset")
        pkg.__path__ = []
        sys.modules["dataset"] = pkg
        sys.modules["dataset.VMD"] = pkg.VMD = _synthetic_dataset_module()
        sys.modules["dataset.DIM"] = pkg.DIM = _synthetic_dim_dataset_module()
    if cpu:
        import torch
        torch.cuda.current_device = lambda: torch.device("cpu")
    return d
