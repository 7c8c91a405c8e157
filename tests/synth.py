"""Seeded synthetic inputs shared by tests (re-exported from the package)."""
from cerebro_b200.synthetic import band_limited_images, loop_candidate, planted_queries, textured_scenes, unit_rows  # noqa: F401
