"""conan-fgw_b200: sm_100a kernels for ConAN's per-conformer message-passing backbone.

The directory name carries a hyphen (it mirrors the reference's repository name), so import it
through the alias module at the repository root: ``import conan_fgw_b200``.

Public surface (same names and signatures as the PyG / ConAN classes it replaces):
``radius_graph, RadiusInteractionGraph, GaussianSmearing, ShiftedSoftplus, CFConv,
InteractionBlock, SchNet, SchNetNoSum`` and, for ViSNet, the classes of ``visnet``.
"""

from . import _lib  # noqa: F401
from .build import build_library, library_is_built  # noqa: F401
from .graph import NeighborList, build_neighbor_list, radius_graph  # noqa: F401
from .nn import (CFConv, GaussianSmearing, InteractionBlock, Linear, RadiusInteractionGraph,  # noqa: F401
                 SchNet, ShiftedSoftplus, SumAggregation)
from .schnet_no_sum import SchNetNoSum, SchNetWithMultipleReturns  # noqa: F401
from . import visnet  # noqa: F401
from .visnet import ViSNet, ViS_MP, ViSNetBlock, TorchGeometricViSNet  # noqa: F401
from . import synthetic  # noqa: F401
from . import aggregation  # noqa: F401
from . import dp, utils  # noqa: F401  (loaded here so that the import alias covers them)
from .aggregation import ConformerAggregationHead, MeanAggregation, create_aggregation_index  # noqa: F401

__all__ = [
    "radius_graph", "build_neighbor_list", "NeighborList", "RadiusInteractionGraph", "GaussianSmearing",
    "ShiftedSoftplus", "CFConv", "InteractionBlock", "SchNet", "SchNetNoSum", "SchNetWithMultipleReturns", "Linear", "SumAggregation",
    "build_library", "library_is_built", "synthetic", "visnet", "ViSNet", "ViS_MP", "ViSNetBlock", "TorchGeometricViSNet",
]
