"""meshflow_b200 -- B200-native data-parallel core of MeshFlow video stabilization.

Drop-in for how4rd/meshflow's ``MeshFlowStabilizer`` (same constructor, same ``stabilize`` entry
point, same returned tuple); the three data-parallel stages run as hand-written sm_100a CUDA kernels
behind the C ABI in ``include/meshflow_b200.h``.  See DESIGN.md.
"""
from .stabilizer import MeshFlowStabilizer
from .pipeline import DeviceCore, MeshSpec, StreamedCore, vertex_xy
from . import _cabi, host_features

__all__ = ["MeshFlowStabilizer", "DeviceCore", "MeshSpec", "StreamedCore", "vertex_xy", "host_features"]
__version__ = "0.1.0"
