"""Importing this package registers the B200 modules under the reference's registry names
(what ``plugin=True; plugin_dir='projects/mmdet3d_plugin/'`` does for the reference,
tools/test.py:139-151)."""
from .view_transformer import ViewTransformerLiftSplatShootVoxel  # noqa: F401
from .encoder3d import CustomResNet3D, SECONDFPN3D  # noqa: F401
from .occ_head import OccHead  # noqa: F401
from .image_encoder import CustomEfficientNet, SECONDFPN  # noqa: F401
from .detector import BEVDepthOccupancy  # noqa: F401
from .metrics import SSCMetrics  # noqa: F401
