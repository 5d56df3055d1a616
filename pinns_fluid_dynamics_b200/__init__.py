"""B200-native PINN loss step with the problem API of giuliamesc/PINNs_Fluid_Dynamics.

    import pinns_fluid_dynamics_b200 as ns     # stands where the scripts `import nisaba as ns`

Host code is Python (PyTorch as tensor/optimizer shell); the arithmetic runs in
``lib/libpinnstep.so`` (hand-written sm_100a CUDA kernels behind the C ABI of include/pinnstep.h).
"""
from . import residuals  # noqa: F401
from .api import (Adam, Loss, LossMeanSquares, OptimizationProblem, TanhMLP, config, minimize,  # noqa: F401
                  optimizers, utils)
from .options import SimulationOptions, read_simulation_options  # noqa: F401
from .residuals import PointSet, ResidualForm  # noqa: F401

__version__ = "0.1.0"
