"""Binding the drop-in coders into a running cbench_BaSIC process.

``bind()`` imports the reference's ``NNTrainableModule`` (cbench/nn/base.py:457) and derives, for every drop-in class, a
subclass that also inherits from it, so that inside the reference's harness

* ``isinstance(coder, NNTrainableModule)`` / ``isinstance(coder, NNCacheImpl)`` hold -- ``get_cache`` recursion
  (cbench/nn/base.py:306-321), ``set_optim_state`` / ``set_custom_state`` propagation and the trainer's module walks see
  the coder like any other node coder;
* the profiler is the reference's ``MetricLogger`` (created by ``BaseModule.__init__``), so
  ``collect_profiler_results`` (cbench/modules/base.py:159-169) reports the coder's scopes;
* state_dict keys, constructor keywords and method signatures stay those of ``prior_coder`` / ``z_coder``.

``install()`` additionally replaces the class attributes of the reference's own modules, so that configs which import
``cbench.modules.prior_model.prior_coder.pgm_coder.GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder`` (the ClassBuilder
files under configs/codecs/general/prior_models/prior_coders/) build the CUDA coders without being edited, and aliases
``cbench.ans`` to the CUDA coder module.  Nothing here is needed outside a reference checkout.
"""
import importlib
import sys

from . import ans, prior_coder, z_coder

_bound = None

# drop-in class -> (reference module, attribute) it replaces
TARGETS = {
    "GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder":
        (prior_coder, "cbench.modules.prior_model.prior_coder.pgm_coder"),
    "CombinedNNTrainablePGMPriorCoder": (prior_coder, "cbench.modules.prior_model.prior_coder.pgm_coder"),
    "TopoGroupDynamicMaskConv2dContextModel": (prior_coder, "cbench.nn.layers.masked_conv"),
    "CompressAIEntropyBottleneckPriorCoder": (z_coder, "cbench.modules.prior_model.prior_coder.compressai_coder"),
}


def bind():
    """{class name: subclass of (drop-in, reference NNTrainableModule)}; needs ``cbench`` importable."""
    global _bound
    if _bound is not None:
        return _bound
    base = importlib.import_module("cbench.nn.base").NNTrainableModule
    out = {}
    for name, (mod, _) in TARGETS.items():
        ours = getattr(mod, name)
        out[name] = type(name, (ours, base), {"__module__": __name__, "__doc__": ours.__doc__, "_basic_b200_dropin": True})
    _bound = out
    return out


def install(alias_ans=True):
    """Points the reference's own module attributes at the bound classes (and ``cbench.ans`` at the CUDA coder module).
    Call once in the launcher, before the configs are built."""
    classes = bind()
    for name, (_, ref_mod) in TARGETS.items():
        setattr(importlib.import_module(ref_mod), name, classes[name])
    if alias_ans:
        import cbench
        sys.modules["cbench.ans"] = ans
        cbench.ans = ans
    return classes
