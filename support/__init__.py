"""Bench and test scaffolding (synthetic grids/weather and the reference's text writers).
Not part of the product package ``pyhelp_b200``."""
