"""The part of the reference's module protocol its callers use on an entropy coder, without importing the reference:

* the cache API of ``NNCacheImpl`` (cbench/nn/base.py:226-420: ``update_cache / get_cache / get_raw_cache / get_all_cache /
  reset_cache / reset_all_cache`` over the dicts ``common, loss_dict, metric_dict, moniter_dict, hist_dict, image_dict,
  text_dict, figure_dict``),
* the profiler of ``BaseModule`` (cbench/modules/base.py:36-47, :159-169; a ``MetricLogger`` whose ``start_time_profile(name)``
  scopes record wall-clock milliseconds under ``"<name> (ms)"``, cbench/utils/logging_utils.py:82-132),
* ``device`` through a non-persistent ``_device_indicator`` buffer, ``set_optim_state`` / ``set_custom_state`` /
  ``load_checkpoint`` / ``post_training_process`` as inert hooks (cbench/nn/base.py:457-520).

``CoderModuleBase.__init__`` is cooperative: when ``reference_integration.bind()`` derives a class from one of the drop-in
coders AND the reference's ``NNTrainableModule`` (so that ``isinstance(coder, NNTrainableModule)`` holds inside the reference's
harness), the reference's own constructor runs first and this one only fills in what is still missing.
"""
import math
import time
from collections import defaultdict, deque

import torch
import torch.nn as nn

CACHE_NAMES = ("common", "loss_dict", "metric_dict", "moniter_dict", "hist_dict", "image_dict", "text_dict", "figure_dict")


class _Meter:
    """SmoothedValue (logging_utils.py:17-78): running total / count plus a short window."""

    def __init__(self, window_size=20):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def reset(self):
        self.deque.clear()
        self.total, self.count = 0.0, 0

    @property
    def global_avg(self):
        return self.total / self.count if self.count > 0 else math.nan

    @property
    def value(self):
        return self.deque[-1] if self.deque else 0


class _TimeScope:
    def __init__(self, meter):
        self.meter = meter

    def __enter__(self):
        self.t0 = time.time()

    def __exit__(self, *exc):
        self.meter.update((time.time() - self.t0) * 1000)


class Profiler:
    """MetricLogger's profiling subset (logging_utils.py:82-132)."""

    def __init__(self):
        self.meters = defaultdict(_Meter)

    def update(self, **kwargs):
        for k, v in kwargs.items():
            self.meters[k].update(v.item() if isinstance(v, torch.Tensor) else v)

    def reset(self):
        for m in self.meters.values():
            m.reset()

    def clear(self):
        self.meters = defaultdict(_Meter)

    def get_global_average(self):
        return {name: m.global_avg for name, m in self.meters.items()}

    def start_time_profile(self, name, profiler_class=None):
        if name not in self.meters:
            name = name + " (ms)"
        scope = _TimeScope if profiler_class is None else profiler_class
        return scope(self.meters[name])

    def record_ms(self, name, ms):
        """A span measured elsewhere (CUDA events inside the library) under the reference's scope name."""
        self.meters[name if name in self.meters else name + " (ms)"].update(ms)


class CoderModuleBase(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)          # nn.Module, or the reference's NNTrainableModule when bound
        for name in CACHE_NAMES:
            if name not in self.__dict__:
                self.__dict__[name] = dict()
        if "_profiler" not in self.__dict__:
            self._profiler = Profiler()
        if "_device_indicator" not in self._buffers:
            self.register_buffer("_device_indicator", torch.zeros(1), persistent=False)
        self.__dict__.setdefault("optim_state", 0)

    # ---- BaseModule
    @property
    def profiler(self):
        return self._profiler

    @profiler.setter
    def profiler(self, profiler):
        self._profiler = profiler

    def start_time_profile(self, name, **kwargs):
        return self.profiler.start_time_profile(name)

    def collect_profiler_results(self, recursive=False, clear=False):
        results = dict(self.profiler.get_global_average())
        if clear:
            self.profiler.clear()
        return results

    @property
    def device(self):
        return self._device_indicator.device

    # ---- NNCacheImpl
    @property
    def cache_names(self):
        return list(CACHE_NAMES)

    def _cache_children(self):
        for name, module in self.named_children():
            if isinstance(module, (nn.ModuleList, nn.ModuleDict)):
                items = enumerate(module) if isinstance(module, nn.ModuleList) else module.items()
                for key, sub in items:
                    if hasattr(sub, "get_cache"):
                        yield f"{name}/{key}", sub
            elif hasattr(module, "get_cache"):
                yield name, module

    def get_cache(self, cache_name="common", prefix=None, recursive=True):
        prefix = cache_name if prefix is None else prefix
        own = getattr(self, cache_name)
        result = {("/".join((prefix, k)) if prefix else k): v for k, v in own.items()}
        if recursive:
            for name, sub in self._cache_children():
                result.update(sub.get_cache(cache_name, prefix="/".join((prefix, name)) if prefix else name, recursive=True))
        return result

    def get_raw_cache(self, cache_name="common"):
        return getattr(self, cache_name)

    def get_all_cache(self, prefix=None, recursive=True):
        return {name: self.get_cache(name, prefix=prefix, recursive=recursive) for name in CACHE_NAMES}

    def update_cache(self, cache_name="common", **kwargs):
        getattr(self, cache_name).update(**kwargs)

    def reset_cache(self, cache_name="common", recursive=True):
        self.__dict__[cache_name] = dict()
        if recursive:
            for _, sub in self._cache_children():
                sub.reset_cache(cache_name, recursive=True)

    def reset_all_cache(self, recursive=True):
        for name in CACHE_NAMES:
            self.reset_cache(name, recursive=recursive)

    # ---- NNTrainableModule hooks a harness may call on every coder; nothing to do on the coding path
    def set_optim_state(self, state=0):
        self.optim_state = state

    def set_custom_state(self, state=None):
        pass

    def load_checkpoint(self, checkpoint_loader=None):
        if isinstance(checkpoint_loader, str):
            self.load_state_dict(torch.load(checkpoint_loader), strict=False)

    def post_training_process(self, *args, **kwargs):
        pass
