from velora_b200.utils.core import set_device, set_seed
from velora_b200.utils.restore import load_state, save_model
from velora_b200.utils.torch import active_parameters, hard_update, soft_update, total_parameters

__all__ = ["set_device", "set_seed", "save_model", "load_state", "soft_update", "hard_update", "total_parameters", "active_parameters"]
