"""B200-native (sm_100a) SMPL decode -> projection -> visibility mask -> part segmentation / silhouette.

Drop-in for the keras_smpl path of akashsengupta1997/indirect_learning_pose-shape; see layers.py for the mirrored
interface and include/smpl_b200.h for the C ABI underneath.  The directory name contains a hyphen, so import it with

    import importlib; smpl = importlib.import_module("indirect_learning_pose-shape_b200")
"""
from . import smpl_io  # noqa: F401  (numpy only; safe without the CUDA library)
from ._lib import SmplB200Error, launch_count, load as load_library, profile_collect, profile_enable  # noqa: F401
from .layers import (  # noqa: F401
    DeviceModel,
    GraphedDecoderStep,
    PipelinedDecoderSteps,
    PartTable,
    SMPLLayer,
    SmplDecoder,
    categorical_crossentropy,
    categorical_focal_loss,
    compute_mask,
    concat_mean_param,
    get_device_model,
    get_part_table,
    load_mean_set_cam_params,
    orthographic_project,
    projects_to_seg,
    projects_to_seg_focal_loss,
    projects_to_silhouette,
    set_cam_params,
)
from .regressor import Dense, IEFRegressor, PlainRegressor  # noqa: F401
from .renderer import SMPLRenderer  # noqa: F401
from .sharding import all_gather_outputs, shard_bounds, shard_slice  # noqa: F401

__all__ = ["SMPLLayer", "SMPLRenderer", "SmplDecoder", "GraphedDecoderStep", "PipelinedDecoderSteps", "Dense", "IEFRegressor", "PlainRegressor", "orthographic_project", "compute_mask", "projects_to_seg",
           "projects_to_silhouette", "projects_to_seg_focal_loss", "categorical_focal_loss", "categorical_crossentropy", "concat_mean_param", "set_cam_params", "load_mean_set_cam_params",
           "DeviceModel", "PartTable", "get_device_model", "get_part_table", "smpl_io", "SmplB200Error",
           "launch_count", "load_library", "profile_enable", "profile_collect", "shard_bounds", "shard_slice", "all_gather_outputs"]
