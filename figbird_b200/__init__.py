"""figbird_b200 -- B200-native drop-in for Figbird's gap-fill hot path (FillGaps + Figbird workers).

The package is a thin Python face over the C ABI in include/figbird_b200.h:
  figbird_b200.capi.Engine      level-2 engine calls (model / batch upload, fb_em_run)
  figbird_b200.capi.fillgaps    level-1 drop-in (the FillGaps executable as a function)
The compute path is the CUDA library figbird_b200/_build/libfigbird_b200.so; there is no CPU fallback.
"""
from . import capi  # noqa: F401
from .capi import Engine, fillgaps  # noqa: F401
