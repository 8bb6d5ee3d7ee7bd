"""Drop-in module files for the reference's unchanged scripts.

The reference's drivers import the model zoo from their working directory
(``from NIOModules import NIOFP2D, NIOFP2D_FNO, NIOFP2D_FNO_attn`` -- 2d_FPE/train_fno.py:8;
``from NIOModules import NIOFP, NIOFP_FNO`` -- 1d_FPE/train_fno.py:7; NIOModules.py:5-8 in turn imports
``FNOModules``, ``DeepONetModules``, ``Baselines`` and ``debug_tools``).  Two ways to route those imports
to this package without touching the scripts:

  * ``PYTHONPATH=<this dir>/<variant>`` -- each variant directory holds one-line ``NIOModules.py`` /
    ``FNOModules.py`` / ``DeepONetModules.py`` / ``Baselines.py`` / ``debug_tools.py`` shims;
  * ``blindno_b200.dropin.install("2d_FPE")`` -- registers the same modules in ``sys.modules``.

Names on the accelerated path resolve to ``blindno_b200.surface`` classes.  Names the scripts import but
never touch on this path (Transolver / plain U-Net variants, 3-D and ODE leftovers: SURVEY.md section 2,
out of scope) resolve to placeholders that raise on construction, so imports succeed and a wrong
model choice fails loudly instead of silently running something else.
"""
from __future__ import annotations

import sys
import types

VARIANTS = ("1d_FPE", "1d_GPE", "2d_FPE", "2d_Non_conservative_FPE")

_OUT_OF_SCOPE = {
    "NIOModules": ("NIOFP2D_FNO_attn", "NIOFP2D_attn", "NIOFP2D_Trans", "NIOFP3D", "NIOFP_ode", "PermInvUNet",
                   "PermInvUNet_attn", "PermInvUNet_attn1D", "PermInvUNet_attn1D_bag", "PermInvUNet_attn1D_bag_GPE",
                   "PermInvUNet_attn1D_GPE"),
    "FNOModules": ("FNO3d", "SpectralConv3d"),
    "DeepONetModules": ("FeedForwardNN", "FourierFeatures"),
    "Baselines": ("EncoderHelm2", "Encoder_ode", "Encoder3D", "Encoder3D_down", "MLP"),
}


def _placeholder(name: str, module: str):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{module}.{name} is outside the accelerated NIO / NIO-FNO path of blindno_b200 "
                                  "(SURVEY.md section 8f); use the reference implementation for it")
    return type(name, (), {"__init__": __init__, "__doc__": f"placeholder for the reference's {module}.{name}"})


def exports(variant: str, module: str) -> dict:
    """Names the reference's ``module`` (of directory ``variant``) provides on this path."""
    if variant not in VARIANTS:
        raise ValueError(f"unknown variant {variant!r}; expected one of {VARIANTS}")
    from ..surface import deeponet, fno, nio
    from ..surface.baselines import ConvBlock
    models = nio.make_models(variant)
    two_d = variant.startswith("2d")
    if module == "FNOModules":
        out = {"SpectralConv1d": fno.SpectralConv1d, "FNO1d": fno.FNO1d,
               "SpectralConv2d": fno.SpectralConv2d if two_d else fno.SpectralConv2dC64,
               "FNO2d": fno.FNO2d if two_d else fno.FNO2dC64}
    elif module == "NIOModules":
        keep = ("NIOFP2D", "NIOFP2D_FNO") if two_d else (
            ("NIOFP", "NIOFP_FNO") if variant == "1d_FPE" else ("NIOFP_schrodinger", "NIOFP_FNO"))
        out = {k: models[k] for k in keep}
        from ..surface.blindno import make_blindno_models
        out.update(make_blindno_models(variant))      # BlinDNO family: FNO heads on the accelerated path (8f N1)
    elif module == "DeepONetModules":
        out = {"FFN": deeponet.FFN, "DeepOnetNoBiasOrg": deeponet.DeepOnetNoBiasOrg,
               "kaiming_init": deeponet.kaiming_init, "activation": deeponet.activation}
    elif module == "Baselines":
        out = {"ConvBlock": ConvBlock, "Encoder": models["Encoder"], "Encoder2D": models["Encoder2D"]}
    elif module == "debug_tools":
        import torch
        out = {"torch": torch}      # the reference's DeepONetModules/Baselines get `torch` via `from debug_tools import *`
    else:
        raise ValueError(f"unknown module {module!r}")
    for name in _OUT_OF_SCOPE.get(module, ()):
        out.setdefault(name, _placeholder(name, module))
    return out


def _export(variant: str, module: str, namespace: dict) -> None:
    """Used by the one-line shim files: fill ``namespace`` (their globals())."""
    namespace.update(exports(variant, module))
    namespace["__all__"] = sorted(exports(variant, module))


def install(variant: str) -> None:
    """Register NIOModules / FNOModules / DeepONetModules / Baselines / debug_tools of ``variant`` in sys.modules."""
    for module in ("debug_tools", "FNOModules", "DeepONetModules", "Baselines", "NIOModules"):
        mod = types.ModuleType(module)
        mod.__dict__.update(exports(variant, module))
        mod.__file__ = f"<blindno_b200.dropin:{variant}/{module}>"
        sys.modules[module] = mod
