"""psi_6 orientation prologue of ``evaluate.SF2d(rotate=True)`` / ``evaluate.PCF2d``.

Placeholder module: filled in by the order kernels (csrc/order.cu)."""
from __future__ import annotations


def psi6_mean(coords, box_boundary):
    raise NotImplementedError("amep_b200.order.psi6_mean: the psi_6 kernel is not built yet")


def psi6_orientation(coords, box_boundary):
    raise NotImplementedError("amep_b200.order.psi6_orientation: the psi_6 kernel is not built yet")


def rotated_periodic_crop(coords, box_boundary, theta):
    raise NotImplementedError("amep_b200.order.rotated_periodic_crop: not built yet")
