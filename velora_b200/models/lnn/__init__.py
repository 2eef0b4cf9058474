from velora_b200.models.lnn.cell import NCPLiquidCell
from velora_b200.models.lnn.ncp import LiquidNCPNetwork, NCPNetwork
from velora_b200.models.lnn.sparse import SparseLinear

__all__ = ["NCPLiquidCell", "LiquidNCPNetwork", "NCPNetwork", "SparseLinear"]
