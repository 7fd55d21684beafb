"""Minimal BoxList stand-in with the maskrcnn-benchmark surface the on-line modules touch
(`BoxList(bbox, (w, h), mode="xyxy")`, `.bbox`, `.size`, `.mode`, add_field/get_field/fields,
clip_to_image, __len__/__getitem__).  Used only when maskrcnn_benchmark is not installed
(reference import: OnlineRegionClassifier.py:12)."""
import torch


class BoxList:
    def __init__(self, bbox, image_size, mode="xyxy"):
        if not torch.is_tensor(bbox):
            bbox = torch.as_tensor(bbox, dtype=torch.float32)
        if bbox.dim() != 2 or bbox.size(-1) % 4 != 0:
            raise ValueError("bbox should be (n, 4k), got %s" % (tuple(bbox.shape),))
        if mode != "xyxy":
            raise ValueError("only xyxy boxes are supported by this stand-in")
        self.bbox = bbox
        self.size = image_size
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, name, data):
        self.extra_fields[name] = data

    def get_field(self, name):
        return self.extra_fields[name]

    def has_field(self, name):
        return name in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def clip_to_image(self, remove_empty=True):
        w, h = self.size
        self.bbox[:, 0::2].clamp_(min=0, max=w - 1)     # TO_REMOVE = 1
        self.bbox[:, 1::2].clamp_(min=0, max=h - 1)
        if remove_empty:
            keep = (self.bbox[:, 3] > self.bbox[:, 1]) & (self.bbox[:, 2] > self.bbox[:, 0])
            return self[keep]
        return self

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)
