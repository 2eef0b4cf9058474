"""velora_b200: B200-native Liquid-NN hot path behind the velora public API."""
from velora_b200 import dataparallel
from velora_b200.models import LiquidNCPNetwork, NCPNetwork
from velora_b200.utils.core import set_device, set_seed
from velora_b200.wiring import Wiring

__version__ = "0.1.0"
__all__ = ["LiquidNCPNetwork", "NCPNetwork", "Wiring", "set_seed", "set_device", "dataparallel"]
