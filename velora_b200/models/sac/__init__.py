from velora_b200.models.sac.continuous import SACActor, SACCritic, SACCriticNCP
from velora_b200.models.sac.discrete import SACActorDiscrete, SACCriticDiscrete, SACCriticNCPDiscrete

__all__ = ["SACActor", "SACCritic", "SACCriticNCP", "SACActorDiscrete", "SACCriticDiscrete", "SACCriticNCPDiscrete"]
