from velora_b200.models.sac.continuous import SACActor, SACCriticNCP
from velora_b200.models.sac.discrete import SACActorDiscrete, SACCriticNCPDiscrete

__all__ = ["SACActor", "SACCriticNCP", "SACActorDiscrete", "SACCriticNCPDiscrete"]
