"""FNOModules of the reference directory 1d_GPE, routed to blindno_b200 (see blindno_b200/dropin/__init__.py)."""
from blindno_b200.dropin import _export

_export("1d_GPE", "FNOModules", globals())
