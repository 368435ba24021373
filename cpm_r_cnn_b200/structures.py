"""Minimal BoxList with the subset of the reference's interface the hot path touches
(pet/utils/data/structures/bounding_box.py): .bbox, .size, .mode, fields, convert, __getitem__, area, __len__.
The ops accept the reference's own BoxList objects just as well (duck typing); this class exists so that the package,
its tests and the benchmark do not need /root/reference."""
import torch


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should have shape (N, 4), got {}".format(tuple(bbox.shape)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size   # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        x1, y1, a, b = self.bbox.unbind(-1)
        if mode == "xyxy":   # from xywh, TO_REMOVE = 1 (bounding_box.py)
            box = torch.stack([x1, y1, x1 + (a - 1).clamp(min=0), y1 + (b - 1).clamp(min=0)], dim=-1)
        else:
            box = torch.stack([x1, y1, a - x1 + 1, b - y1 + 1], dim=-1)
        out = BoxList(box, self.size, mode=mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item] if isinstance(v, torch.Tensor) else v)
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def area(self):
        box = self.bbox
        if self.mode == "xyxy":
            return (box[:, 2] - box[:, 0] + 1) * (box[:, 3] - box[:, 1] + 1)
        return box[:, 2] * box[:, 3]

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            len(self), self.size[0], self.size[1], self.mode)
