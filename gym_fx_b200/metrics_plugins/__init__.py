"""Plugin mirrors for the `metrics.plugins` entry-point group (reference: setup.py:11-35)."""
