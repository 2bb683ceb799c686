"""B200-native drop-in for the FWI hot path of lidongzh/FwiFlow.jl.

Layout:
  csrc/            hand-written CUDA kernels (sm_100a) + the C ABI of include/fwi_b200.h
  _lib.py          ctypes binding of libfwi_b200.so (built in-tree; no CPU fallback)
  ops.py           fwi_op / fwi_obs_op / fwi_op_grad / Plan   (mirror of src/Core.jl)
  fwi.py           FWI struct, compute_observation, compute_misfit (mirror of src/FWI.jl)
  utils.py         paraGen, surveyGen, sourceGene, velocity_to_moduli (mirror of src/Utils.jl)
  dist.py          shot sharding + one all-reduce per gradient
  synthetic.py     seeded synthetic workloads (BASELINE.json configs)
"""
from .utils import (paraGen, surveyGen, sourceGene, velocity_to_moduli, klauderWave, cs_bounds_cloud,  # noqa: F401
                    resize_bilinear)
from .ops import (fwi_op, fwi_obs_op, fwi_op_grad, fwi_op_and_grad, fwi_op_and_grad_multi, Plan, FwiError,  # noqa: F401
                  release)
from .fwi import (FWI, FWIExample, compute_observation, compute_misfit, compute_misfit_and_gradient,  # noqa: F401
                  compute_misfit_and_gradient_resident,
                  timelapse_misfit_and_gradients, timelapse_misfit_and_gradients_batched)
