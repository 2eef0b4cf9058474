from velora_b200.models.lnn import LiquidNCPNetwork, NCPNetwork

__all__ = ["LiquidNCPNetwork", "NCPNetwork"]
