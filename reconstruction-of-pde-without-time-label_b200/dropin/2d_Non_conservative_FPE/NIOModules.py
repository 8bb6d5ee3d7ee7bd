"""NIOModules of the reference directory 2d_Non_conservative_FPE, routed to blindno_b200 (see blindno_b200/dropin/__init__.py)."""
from blindno_b200.dropin import _export

_export("2d_Non_conservative_FPE", "NIOModules", globals())
