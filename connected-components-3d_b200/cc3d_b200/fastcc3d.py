"""Namespace alias: the reference exposes `runs`, `draw` and `_erase` only as `cc3d.fastcc3d.*`
(cc3d/__init__.py:6-17, 268-275); callers that reach for them there find the B200 versions here."""
from . import (  # noqa: F401
  DimensionError, connected_components, statistics, each, contacts, region_graph, voxel_connectivity_graph,
  color_connectivity_graph, estimate_provisional_labels, runs, draw, erase, _erase,
)
