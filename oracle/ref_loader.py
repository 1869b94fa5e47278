"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference modules from /root/reference (build container) or from
oracle/_ref/ (the same files staged byte-for-byte by oracle/build_ref.py; this is what the GPU box has).

Used by `oracle/make_golden.py` (golden-vector generation), `tests/test_oracle_vs_reference.py`, the drop-in GPU tests
(tests/test_gpu_dropin.py: our seams installed on the reference's own classes) and `bench.py --impl reference` / its
`cpu_baseline` leg (the reference's CPU forward, kind "reference").  Nothing in `aria_b200/` may import this file.

Harness: the reference package cannot be imported as a package under the installed
transformers 5.5 (`aria/model/__init__.py` pulls in processors whose imports moved), so we
  1. register stub packages `aria`, `aria.model` so `__init__.py` is skipped,
  2. provide `LLAMA_ATTENTION_CLASSES` (removed upstream; reference uses it at moe_lm.py:594),
  3. load the five model files by path.
Two lines of the reference cannot execute on CPU and are shimmed *outside* the reference files:
  - moe_lm.py:264 `torch.histc` on int64 (no CPU kernel)  -> cast to float and back
  - moe_lm.py:483 `torch.cuda.set_device(cpu tensor)`     -> no-op for CPU devices (also on a box that has GPUs)
"""
import importlib.util
import os
import sys
import types

import torch

# the mounted reference when present (build container), else the files staged byte-for-byte by oracle/build_ref.py
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("ARIA_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/aria/model/moe_lm.py") else _STAGED)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "aria", "model", "moe_lm.py"))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference modules: moe_lm, vision_encoder, projector,
    configuration_aria, modeling_aria."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    import transformers.models.llama.modeling_llama as ml

    if not hasattr(ml, "LLAMA_ATTENTION_CLASSES"):
        ml.LLAMA_ATTENTION_CLASSES = {
            k: ml.LlamaAttention for k in ("eager", "sdpa", "flash_attention_2")
        }

    # CPU shims (see module docstring)
    if not getattr(torch.histc, "_aria_shim", False):
        _histc = torch.histc

        def histc(inp, bins=100, min=0, max=0, **kw):
            if not inp.is_floating_point():
                return _histc(inp.float(), bins=bins, min=min, max=max, **kw).to(torch.int64)
            return _histc(inp, bins=bins, min=min, max=max, **kw)

        histc._aria_shim = True
        torch.histc = histc
    if not getattr(torch.cuda.set_device, "_aria_shim", False):
        _set_device = torch.cuda.set_device

        def set_device(device):
            # moe_lm.py:483 calls this with `input.device`; for a CPU tensor there is nothing to select (and torch raises)
            if isinstance(device, torch.device) and device.type != "cuda":
                return
            if isinstance(device, str) and not device.startswith("cuda"):
                return
            if torch.cuda.is_available():
                _set_device(device)

        set_device._aria_shim = True
        torch.cuda.set_device = set_device

    for name in ("aria", "aria.model"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
            sys.modules[name] = m

    for mod in ("moe_lm", "vision_encoder", "projector", "configuration_aria", "modeling_aria"):
        full = f"aria.model.{mod}"
        spec = importlib.util.spec_from_file_location(
            full, os.path.join(REF_ROOT, "aria", "model", mod + ".py")
        )
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        _loaded[mod] = m
    return types.SimpleNamespace(**_loaded)


def load_reference_lora():
    """Loads the unmodified `aria/lora/layers.py` (GroupedGemmLoraLayer).  It subclasses peft's `LoraLayer`, and peft is not
    installed offline, so a minimal stand-in for the two peft symbols it imports is registered first (harness code, outside
    the reference files): just the bookkeeping `LoraLayer.__init__` / `set_adapter` / properties that the reference's
    `__init__`, `update_layer` and `forward` (layers.py:30-152) touch.  The arithmetic under test is the reference's own."""
    ref = load_reference()
    if "lora_layers" in _loaded:
        return _loaded["lora_layers"]
    import math

    from torch import nn

    class LoraLayer:  # stand-in for peft.tuners.lora.LoraLayer (peft 0.13 semantics for the members used)
        def __init__(self, base_layer, **kwargs):
            self.base_layer = base_layer
            self.r, self.lora_alpha, self.scaling, self.use_dora = {}, {}, {}, {}
            self.lora_dropout = nn.ModuleDict({})
            self.lora_A = nn.ModuleDict({})
            self.lora_B = nn.ModuleDict({})
            self._disable_adapters = False
            self.merged_adapters = []
            self.in_features, self.out_features = base_layer.in_features, base_layer.out_features

        @property
        def active_adapters(self):
            a = self._active_adapter
            return [a] if isinstance(a, str) else a

        @property
        def disable_adapters(self):
            return self._disable_adapters

        @property
        def merged(self):
            return bool(self.merged_adapters)

        def get_base_layer(self):
            return self.base_layer

        def set_adapter(self, adapter_names):
            self._active_adapter = adapter_names

        def _move_adapter_to_device_of_base_layer(self, adapter_name):
            pass

        def _check_forward_args(self, x, *args, **kwargs):
            pass

        def reset_lora_parameters(self, adapter_name, init_lora_weights):
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B[adapter_name].weight)

    def check_adapters_to_merge(module, adapter_names=None):
        return adapter_names if adapter_names is not None else module.active_adapters

    for name, attrs in (("peft", {}), ("peft.tuners", {}), ("peft.tuners.lora", {"LoraLayer": LoraLayer}),
                        ("peft.tuners.tuners_utils", {"check_adapters_to_merge": check_adapters_to_merge})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(sys.modules[name], k, v)
    sys.modules["aria.model"].GroupedGEMM = ref.moe_lm.GroupedGEMM   # `from aria.model import GroupedGEMM` (layers.py:26)
    if "aria.lora" not in sys.modules:
        pkg = types.ModuleType("aria.lora")
        pkg.__path__ = [os.path.join(REF_ROOT, "aria", "lora")]
        sys.modules["aria.lora"] = pkg
    spec = importlib.util.spec_from_file_location("aria.lora.layers", os.path.join(REF_ROOT, "aria", "lora", "layers.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["aria.lora.layers"] = m
    spec.loader.exec_module(m)
    _loaded["lora_layers"] = m
    return m
