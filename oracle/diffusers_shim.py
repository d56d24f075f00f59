"""ORACLE (test infrastructure): a minimal fake `diffusers` package so that the reference's UNMODIFIED
/root/reference/i2vgen-xl/pnp_utils.py (imports at :34-36, :138) and its vendored scheduler
/root/reference/consisti2v/ddim_inverse_scheduler.py (imports at :23-26) can be imported and executed on the
oracle modules.  Only usable where /root/reference exists (the build container) — see oracle/make_golden.py.
"""
from __future__ import annotations

import functools
import importlib.util
import inspect
import sys
import types
from types import SimpleNamespace

from . import unet_ref

REF_ROOT = "/root/reference"


def install() -> None:
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_av2v_shim", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    d = mod("diffusers")
    d._av2v_shim = True
    utils = mod("diffusers.utils")
    utils.USE_PEFT_BACKEND = True  # hooks then call plain nn.Linear/Conv2d without `scale` (pnp_utils.py:78,177)

    class BaseOutput(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)
    utils.BaseOutput = BaseOutput
    utils.deprecate = lambda *a, **k: None
    models = mod("diffusers.models")
    up = mod("diffusers.models.upsampling")
    up.Upsample2D = unet_ref.Upsample2D
    down = mod("diffusers.models.downsampling")
    down.Downsample2D = unet_ref.Downsample2D
    ap = mod("diffusers.models.attention_processor")
    ap.AttnProcessor2_0 = unet_ref.AttnProcessor2_0
    cu = mod("diffusers.configuration_utils")

    class ConfigMixin:
        pass

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *a, **k):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **k)
            bound.apply_defaults()
            self.config = SimpleNamespace(**{n: v for n, v in bound.arguments.items() if n != "self"})
            return init(self, *a, **k)
        return wrapper
    cu.ConfigMixin = ConfigMixin
    cu.register_to_config = register_to_config
    sch = mod("diffusers.schedulers")
    su = mod("diffusers.schedulers.scheduling_utils")

    class SchedulerMixin:
        pass
    su.SchedulerMixin = SchedulerMixin
    d.utils, d.models, d.schedulers, d.configuration_utils = utils, models, sch, cu
    models.upsampling, models.downsampling, models.attention_processor = up, down, ap
    sch.scheduling_utils = su

    # pnp_utils.py:10 imports torchvision.io.read_video/write_video (unused on the hook path); torchvision >= 0.22
    # dropped them, so provide inert stand-ins instead of touching the reference file.
    import torchvision.io as tvio
    for name in ("read_video", "write_video"):
        if not hasattr(tvio, name):
            setattr(tvio, name, lambda *a, **k: (_ for _ in ()).throw(RuntimeError("video I/O is out of scope")))


def load_reference_module(relpath: str, name: str):
    """Import a file of the read-only reference tree by path (nothing is copied)."""
    install()
    spec = importlib.util.spec_from_file_location(name, f"{REF_ROOT}/{relpath}")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m
