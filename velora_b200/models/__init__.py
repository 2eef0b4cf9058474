"""velora.models' public names (velora/models/__init__.py:1-9)."""
from velora_b200.models.lnn import LiquidNCPNetwork, NCPNetwork
from velora_b200.models.nf import NeuroFlow, NeuroFlowCT

__all__ = ["LiquidNCPNetwork", "NCPNetwork", "NeuroFlowCT", "NeuroFlow"]
