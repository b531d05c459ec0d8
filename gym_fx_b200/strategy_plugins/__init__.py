"""Plugin mirrors for the `strategy.plugins` entry-point group (reference: setup.py:11-35)."""
