"""Seeded synthetic inputs shared by tests (re-exported from the package)."""
from cerebro_b200.synthetic import band_limited_images, loop_candidate, planted_queries, unit_rows  # noqa: F401
