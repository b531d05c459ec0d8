"""Plugin mirrors for the `broker.plugins` entry-point group (reference: setup.py:11-35)."""
