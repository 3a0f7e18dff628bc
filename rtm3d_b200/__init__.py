"""rtm3d_b200 -- B200-native (sm_100a) replacement of RTM3D's post-head keypoint-heatmap decoder.

Python/PyTorch owns memory and streams; every arithmetic step runs in hand-written CUDA kernels behind the C ABI of
``librtm3d_decode.so`` (include/rtm3d_decode.h).  There is no CPU path and no fallback.
"""
from ._native import LIB_PATH, build  # noqa: F401
from .decoder import (GroupedKeypoints, HeatmapDecoder, HostDecodeSession, KeypointCandidates, PackedDetections,  # noqa: F401
                      decoder_from_config)
from .boxfit import Box3DFit, PackedBoxFit, fit_packed, optim_decode_bbox3d  # noqa: F401
from .train_side import FocalLoss, MainTargets, RTM3DLoss, build_main_targets, gather_l1_loss  # noqa: F401
from .plugin import install, install_boxfit, uninstall, uninstall_boxfit  # noqa: F401

__all__ = ["HeatmapDecoder", "PackedDetections", "KeypointCandidates", "GroupedKeypoints", "HostDecodeSession", "decoder_from_config",
           "install", "uninstall", "install_boxfit", "uninstall_boxfit", "build", "LIB_PATH", "fit_packed", "optim_decode_bbox3d", "PackedBoxFit", "Box3DFit",
           "FocalLoss", "MainTargets", "build_main_targets", "RTM3DLoss", "gather_l1_loss"]
