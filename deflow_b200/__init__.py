"""deflow_b200 -- B200-native (sm_100a) implementation of DeFlow's data-parallel hot path behind the
reference's PyTorch module surface.  See DESIGN.md; the C ABI is include/deflow_b200.h."""
from .deflow import DeFlow, FastFlow3D, weights_init  # noqa: F401
from .encoder import DynamicEmbedder, DynamicPillarFeatureNet, DynamicVoxelizer, PointPillarsScatter  # noqa: F401
from .decoder import ConvGRU, ConvGRUDecoder, LinearDecoder  # noqa: F401
from .unet import FastFlow3DUNet  # noqa: F401
from .mmcv_ext import DynamicScatter, Voxelization, dynamic_scatter, voxelization  # noqa: F401
from .lossfuncs import deflowLoss, ff3dLoss, seflowLoss, zeroflowLoss, training_step_loss  # noqa: F401
from . import chamfer3D, eval_metric, feed  # noqa: F401,E402
