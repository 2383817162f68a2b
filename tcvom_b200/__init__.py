"""tcvom_b200 -- B200-native (sm_100a) implementation of TCVOM's GCA+TAM frame-window hot path.

Public surface mirrors the reference's operator/plugin interface for this path:
``get_VMN_models``, ``VMN``, ``FeatureAggregationModule``, ``GuidedCxtAtten``, ``EvalModel``.
"""
from .model import (EvalModel, FeatureAggregationModule, FullModel, FullModel_VMD, GuidedCxtAtten, VMN,  # noqa: F401
                    VMN_FBA, get_VMN_models, trimap_transform)
from .stream import FrameStream  # noqa: F401
from . import metrics  # noqa: F401  (calc_metric.py's metrics on the GPU)

__version__ = "0.1.0"


def install(native_wrapper: bool = True, fba_seam: bool = False, native_tam: bool = False):
    """Hooks this package into an importable reference checkout (``models`` on ``sys.path``).

    * ``models.VMN.get_VMN_models('vmn_gca', ...)`` -> :func:`tcvom_b200.get_VMN_models`
      (the plugin seam ``models/model.py:39-44`` calls at construction time); other archs are
      forwarded to the reference untouched.  ``fba_seam=True`` also routes the inference-only native networks
      (``'vmn_fba'``, ``'vmn_dim'``, ``'vmn_index'``) through the seam -- leave it off when the reference's own training
      wrappers (``FullModel_VMD('vmn_fba')`` etc.) are to keep working;
    * with ``native_wrapper`` also ``models.model.EvalModel`` -> :class:`tcvom_b200.EvalModel`
      when it is built for ``vmn_gca`` / ``vmn_fba`` / ``vmn_dim`` / ``vmn_index`` (fused preprocess / postprocess kernels, CUDA-graph
      replay; for ``vmn_fba`` also the trimap distance transforms on the GPU instead of ``cv2``).

    * ``native_tam=True`` also replaces the reference's ``FeatureAggregationModule`` class (VMN_model.py:9-68, bound by
      name in VMN_DIM.py:4, VMN_Index.py:5, VMN_FBA.py:3, VMN_GCA.py:6) by :class:`tcvom_b200.FeatureAggregationModule`:
      the reference's OWN ``vmn_dim`` / ``vmn_index`` / ``vmn_fba`` networks then run -- and TRAIN, through autograd --
      with the native TAM operator (same parameter names, so checkpoints load unchanged).

    Call it before the reference script imports ``models.model``."""
    import importlib
    import sys
    import types
    for m in ("matplotlib", "matplotlib.pyplot"):          # models/Index/hldecoder.py:36 imports it at module import
        try:
            importlib.import_module(m)
        except ImportError:
            sys.modules.setdefault(m, types.ModuleType(m))
    ref_vmn = importlib.import_module("models.VMN")
    ref_model = importlib.import_module("models.model")
    if native_tam:
        from .model import FeatureAggregationModule as _NativeTAM
        for name in ("VMN_model", "VMN_DIM", "VMN_Index", "VMN_FBA", "VMN_GCA"):
            mod = importlib.import_module("models.VMN." + name)
            cur = getattr(mod, "FeatureAggregationModule", None)
            if cur is not None and cur is not _NativeTAM:
                if not hasattr(_NativeTAM, "_reference"):
                    _NativeTAM._reference = cur
                mod.FeatureAggregationModule = _NativeTAM
    if getattr(ref_vmn.get_VMN_models, "_tcvom_b200", False):
        return
    orig_factory = ref_vmn.get_VMN_models

    def factory(arch, *args, **kwargs):
        if arch == "vmn_gca" or (arch in ("vmn_fba", "vmn_dim", "vmn_index") and fba_seam):
            return get_VMN_models(arch, *args, **kwargs)
        return orig_factory(arch, *args, **kwargs)

    factory._tcvom_b200 = True
    factory._reference = orig_factory
    ref_vmn.get_VMN_models = factory
    if native_wrapper:
        orig_eval = ref_model.EvalModel

        class _EvalModelDispatch:
            """EvalModel('vmn_gca', ...) -> native wrapper; anything else -> reference class."""

            def __new__(cls, model, *args, **kwargs):
                if model in ("vmn_gca", "vmn_fba", "vmn_dim", "vmn_index"):
                    return EvalModel(model, *args, **kwargs)
                return orig_eval(model, *args, **kwargs)

        _EvalModelDispatch._reference = orig_eval
        ref_model.EvalModel = _EvalModelDispatch
        orig_vmd = ref_model.FullModel_VMD

        class _VMDDispatch:
            """FullModel_VMD('vmn_gca', ...) -> native wrapper (inference use, pred_vmn.py); else reference."""
            ARCH_DICT = orig_vmd.ARCH_DICT

            def __new__(cls, model, *args, **kwargs):
                if model == "vmn_gca":
                    return FullModel_VMD(model, *args, **kwargs)
                return orig_vmd(model, *args, **kwargs)

        _VMDDispatch._reference = orig_vmd
        ref_model.FullModel_VMD = _VMDDispatch
