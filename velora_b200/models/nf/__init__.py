from velora_b200.models.nf.agent import NeuroFlowCTCore
from velora_b200.models.nf.modules import ActorModule, CriticModule, EntropyModule

__all__ = ["NeuroFlowCTCore", "ActorModule", "CriticModule", "EntropyModule"]
