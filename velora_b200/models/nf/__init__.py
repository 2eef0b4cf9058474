from velora_b200.models.nf.agent import NeuroFlowCore, NeuroFlowCTCore
from velora_b200.models.nf.modules import (ActorModule, ActorModuleDiscrete, CriticModule, CriticModuleDiscrete, EntropyModule,
                                           EntropyModuleDiscrete)

__all__ = ["NeuroFlowCore", "NeuroFlowCTCore", "ActorModule", "ActorModuleDiscrete", "CriticModule", "CriticModuleDiscrete",
           "EntropyModule", "EntropyModuleDiscrete"]
