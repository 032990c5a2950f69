"""yolo-u_b200: B200-native (sm_100a) implementation of the YOLO-Seg++ inference hot path of Jhewu/YOLO-U behind the
reference's own Python call signatures.  Import as `yolo_u_b200` (shim at the repo root)."""
from . import _lib  # noqa: F401
from ._lib import YspError, lib  # noqa: F401
from .engine import Engine  # noqa: F401
from .YOLOSegPlusPlus import YOLOSegPlusPlus, DoubleLightConv, ECA  # noqa: F401
from .detector import B200Detector  # noqa: F401
from .nms import non_max_suppression, TorchNMS  # noqa: F401
from .predictor import Predictor, HostPipeline, predict  # noqa: F401
from .metrics import mask_counts, dice_from_counts, SegMetrics  # noqa: F401
from .sharding import shard_volumes, shard_slices, batches  # noqa: F401
from .formats import ingest, objectmap_transform, resize_u8, save_objectmaps, scale_boxes  # noqa: F401
