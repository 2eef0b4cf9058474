from velora_b200.models.nf.agent import NeuroFlowCore, NeuroFlowCTCore
from velora_b200.models.nf.env_agent import NeuroFlow, NeuroFlowCT
from velora_b200.models.nf.modules import (ActorModule, ActorModuleDiscrete, CriticModule, CriticModuleDiscrete, EntropyModule,
                                           EntropyModuleDiscrete)

__all__ = ["NeuroFlow", "NeuroFlowCT", "NeuroFlowCore", "NeuroFlowCTCore", "ActorModule", "ActorModuleDiscrete", "CriticModule",
           "CriticModuleDiscrete", "EntropyModule", "EntropyModuleDiscrete"]
