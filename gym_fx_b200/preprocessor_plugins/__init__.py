"""Plugin mirrors for the `preprocessor.plugins` entry-point group (reference: setup.py:11-35)."""
