"""pnpvcve_b200 -- B200-native BAE+CAA enhancement hot path of PnP-VCVE.

Importing the package never needs a GPU; running the generator does (sm_100 only, no fallback).
"""
from .registry import BACKBONES, MODELS, Config, build_backbone, build_from_cfg  # noqa: F401
from .backbone import (BAEGenerator,  # noqa: F401
                       IconVSR_restore_wo_refill_mv_ipb_fast_domain_dynamic_with_par)

__version__ = "0.1.0"
