"""rarecoal_b200 — B200-native engine for rarecoal's likelihood hot path (getProb / logl).

The compute path is librarecoal_b200.so (CUDA, sm_100a).  Importing this package without the built
library raises ImportError on first use; there is no CPU fallback."""
from .api import (Engine, Join, JointStateSpace, ModelEvent, ModelSpec, RarecoalCompException,  # noqa: F401
                  RarecoalDeviceError, RarecoalException, RarecoalHistogramException, RarecoalModelException,
                  SetFreeze, SetGrowthRate, SetPopSize, SpectrumEntry, Split, choose, computeAllConfigs,
                  computeLogLikelihoodFromSpec, computeStandardOrder, defaultTimes, desugarGrowth, getRegularizationPenalty,
                  getTimeSteps, makeJointStateSpace, planStats, shardPatterns, validateModel)

__version__ = "0.1.0"
