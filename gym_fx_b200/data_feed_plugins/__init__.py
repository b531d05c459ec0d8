"""Plugin mirrors for the `data_feed.plugins` entry-point group (reference: setup.py:11-35)."""
