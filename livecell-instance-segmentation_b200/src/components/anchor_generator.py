"""Drop-in for the reference's ``src/components/anchor_generator.py`` (AnchorGenerator, :5-37).

Same constructor, attributes and ``generate_anchors(feature_map_size, stride, device)`` signature;
the [h*w*A, 4] fp32 table is produced by one kernel (lcr_anchors_f32) instead of ~10 ATen launches.
"""
from ... import ops


class AnchorGenerator:
    """Generates anchor boxes at multiple scales and aspect ratios (aspect ratio = w/h, as the
    reference defines it at anchor_generator.py:20-21)."""

    def __init__(self, sizes=(32, 64, 128), aspect_ratios=(0.5, 1.0, 2.0)):
        self.sizes = sizes
        self.aspect_ratios = aspect_ratios
        self.num_anchors_per_location = len(sizes) * len(aspect_ratios)

    def base_anchors(self):
        return ops.base_anchors(self.sizes, self.aspect_ratios)

    def generate_anchors(self, feature_map_size, stride, device):
        h, w = feature_map_size
        return ops.anchors(int(h), int(w), int(stride), self.base_anchors(), device)
