"""cpm_r_cnn_b200 -- B200-native (sm_100a) detection-head op layer of CPM R-CNN.

Drop-in for the names `pet/rcnn/**` imports from the reference's op layer:

    from cpm_r_cnn_b200 import ROIAlign, roi_align, nms, ml_nms, boxlist_nms, boxlist_ml_nms, Pooler, LevelMapper

All of them call the C-ABI CUDA library libcpm_ops.so (include/cpm_ops.h); there is no CPU or PyTorch fallback.
"""
from . import _lib
from ._lib import launch_count
from .boxlist_ops import (batched_boxlist_nms, boxlist_ml_nms, boxlist_ml_nms_legacy, boxlist_nms,
                          boxlist_nms_legacy)
from .detect_postprocess import CLSPostProcessor
from .grid_decode import GridPostProcessor, calc_sub_regions, grid_decode
from .grid_targets import GridTargetGenerator, prepare_grid_target
from .matcher import Matcher, boxlist_iou
from .nms import batched_nms, ml_nms, nms
from .poolers import LevelMapper, Pooler
from .roi_align import ROIAlign, roi_align, stage_nhwc
from .rpn import BoxCoder, RPNPostProcessor, rpn_decode
from .structures import BoxList

__all__ = ["ROIAlign", "roi_align", "stage_nhwc", "nms", "ml_nms", "batched_nms", "boxlist_nms", "boxlist_ml_nms",
           "boxlist_nms_legacy", "boxlist_ml_nms_legacy", "batched_boxlist_nms", "Pooler", "LevelMapper",
           "grid_decode", "calc_sub_regions", "GridPostProcessor", "BoxList", "launch_count", "RPNPostProcessor", "BoxCoder",
           "rpn_decode", "CLSPostProcessor", "GridTargetGenerator", "prepare_grid_target", "Matcher", "boxlist_iou"]
