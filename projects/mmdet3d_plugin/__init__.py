"""Drop-in for the reference's plugin package: ``plugin_dir = "projects/mmdet3d_plugin/"``
(stereoscene.py:7-8) makes the reference's tools import ``projects.mmdet3d_plugin``
(tools/test.py:139-151); importing this package registers the B200 modules under the same
registry names instead."""
from stereoscene_b200.plugin import *  # noqa: F401,F403
from stereoscene_b200.registry import BACKBONES, DETECTORS, HEADS, NECKS, build_model  # noqa: F401
