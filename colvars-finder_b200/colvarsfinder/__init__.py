"""colvarsfinder -- B200-native drop-in for the training step of zwpku/colvars-finder.

Same importable names as the reference for the data-parallel training path:
``colvarsfinder.core.{TrainingTask, EigenFunctionTask, AutoEncoderTask}``,
``colvarsfinder.nn.{create_sequential_nn, EigenFunctions, AutoEncoder}``,
``colvarsfinder.utils.{WeightedTrajectory, Align, FeatureMap, Preprocessing}``.
All arithmetic of the step runs in libcvf_sm100.so (hand-written sm_100a CUDA, include/cvf.h).
"""
__version__ = "0.1.14+b200.1"
