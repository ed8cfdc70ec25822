"""Drop-in mirrors of scripts/road_segmentation/determine_class.py and final_metrics.py helpers."""
from . import determine_class, final_metrics  # noqa: F401
