from .mpm import *  # noqa: F401,F403
from .mpm import (MPMStatics, MPMParticleData, MPMConstant, MPMState, MPMModel, MPMModelBuilder,
                  MPMInitData, MPMStateInitializer, MPMStaticsInitializer)
from .interface import MPMSimFunction, MPMSim, MPMDiffSim, MPMCacheDiffSim, MPMForwardSim
