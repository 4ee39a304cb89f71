from .coco_results import coco_records, prepare_for_coco_detection, write_coco_json  # noqa: F401
