"""B200-native geometric hot path of SAM-6D's Pose Estimation Model.

Import name: the directory is ``openvino-sam-6d_b200`` (not an identifier); load it with
``__graft_entry__.load_package()`` which registers it as ``openvino_sam_6d_b200``.
"""
from . import _lib  # noqa: F401
from . import data_utils, model_utils, parallel, pointnet2_ext, pointnet2_utils, positional_encoding, stage, synth, transformer  # noqa: F401
from .dropin import install_pointnet2_ext  # noqa: F401
