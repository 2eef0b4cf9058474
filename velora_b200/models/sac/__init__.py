from velora_b200.models.sac.continuous import SACActor, SACCriticNCP

__all__ = ["SACActor", "SACCriticNCP"]
