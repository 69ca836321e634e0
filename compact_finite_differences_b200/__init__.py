"""
compact_finite_differences_b200 -- B200-native fp64 compact (Pade) finite-difference derivatives.

Drop-in for the derivative hot path of ashwinsrnth/compact-finite-differences:

    CompactFiniteDifferenceSolver   code/cuda/compact.py:16-44, code/ocl/compact.py:12-61
    NearToeplitzSolver              code/cuda/solvers/templated/near_toeplitz.py:34-107
    ReducedSolver                   code/cuda/reduced.py:5-18
    ZPartitionedDerivative          the reference's multi-rank dfdx (compact.py:29-44) on a z-partition
    PartitionedDerivative, DA       the same along any direction of a Cartesian process grid (gpuDA.py:7-59,154-180)
    HostGradient                    ndarray in / ndarrays out (code/ocl/compact.py:26-61), copies pipelined

Python is a thin ctypes layer over libcfd_b200.so (hand-written sm_100a kernels); there is no CPU,
PyTorch-eager or Triton path.  PyTorch is used for device memory, streams and torch.distributed only.
"""
from ._lib import CfdError, lib  # noqa: F401
from .compact import CompactFiniteDifferenceSolver, LineDA  # noqa: F401
from .near_toeplitz import NearToeplitzSolver  # noqa: F401
from .reduced import ReducedSolver  # noqa: F401
from .host import HostGradient  # noqa: F401
from .partition import (PartitionedDerivative, ZPartitionedDerivative, exchange_halo_planes,  # noqa: F401
                        exchange_interface_planes, gather_interface_planes)
from .grid import DA, DA_arange, DA_gather_blocks, DA_scatter_blocks  # noqa: F401

__all__ = ["CompactFiniteDifferenceSolver", "LineDA", "NearToeplitzSolver", "ReducedSolver", "ZPartitionedDerivative", "HostGradient",
           "PartitionedDerivative", "DA", "DA_arange", "DA_gather_blocks", "DA_scatter_blocks",
           "exchange_halo_planes", "exchange_interface_planes", "gather_interface_planes", "CfdError", "lib"]
