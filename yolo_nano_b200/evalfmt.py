"""Evaluator-side glue: what the reference's evaluators do with the detections that
`YOLONano.forward` returns — the inverse of the letterbox mapping and the two result formats — so that
accuracy can be re-measured with the stock COCO / VOC tools.  Host code, as in the reference.

  map_to_image      evaluator/cocoapi_evaluator.py:85-87, evaluator/vocapi_evaluator.py:72-74
  coco_result_rows  evaluator/cocoapi_evaluator.py:89-100
  voc_class_dets    evaluator/vocapi_evaluator.py:76-86
  voc_result_lines  evaluator/vocapi_evaluator.py:148-157
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def map_to_image(bboxes: np.ndarray, scale, offset, width: int, height: int) -> np.ndarray:
    """Normalised letterboxed boxes -> pixel boxes of the original image, IN PLACE like the reference
    (`bboxes -= offset; bboxes /= scale; bboxes *= size`)."""
    size = np.array([[width, height, width, height]])
    bboxes -= offset
    bboxes /= scale
    bboxes *= size
    return bboxes


def coco_result_rows(image_id: int, bboxes: np.ndarray, scores: np.ndarray, cls_inds: np.ndarray,
                     class_ids: Sequence[int]) -> List[Dict]:
    """COCO detection-result rows ([x, y, w, h] boxes, category ids through `class_ids`)."""
    rows = []
    for i, box in enumerate(bboxes):
        x1 = float(box[0])
        y1 = float(box[1])
        x2 = float(box[2])
        y2 = float(box[3])
        label = class_ids[int(cls_inds[i])]
        rows.append({"image_id": int(image_id), "category_id": label, "bbox": [x1, y1, x2 - x1, y2 - y1],
                     "score": float(scores[i])})
    return rows


def voc_class_dets(bboxes: np.ndarray, scores: np.ndarray, cls_inds: np.ndarray, num_classes: int) -> List[np.ndarray]:
    """Per-class [n, 5] float32 arrays (x1, y1, x2, y2, score); empty classes give shape (0, 5)."""
    out = []
    for j in range(num_classes):
        inds = np.where(cls_inds == j)[0]
        if len(inds) == 0:
            out.append(np.empty([0, 5], dtype=np.float32))
            continue
        out.append(np.hstack((bboxes[inds], scores[inds][:, np.newaxis])).astype(np.float32, copy=False))
    return out


def voc_result_lines(image_name: str, dets: np.ndarray) -> List[str]:
    """Lines of a VOCdevkit `det_<set>_<class>.txt` file for one image (1-based pixel coordinates)."""
    return ['{:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}\n'.format(image_name, dets[k, -1], dets[k, 0] + 1, dets[k, 1] + 1,
                                                            dets[k, 2] + 1, dets[k, 3] + 1)
            for k in range(dets.shape[0])]
