"""acan_b200 -- B200-native (sm_100a) drop-in for the ACAN hot path: context aggregation module,
attention loss and ordinal / cross-entropy depth head.  See DESIGN.md and INTEGRATION.md."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
