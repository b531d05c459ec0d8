"""Plugin mirrors for the `reward.plugins` entry-point group (reference: setup.py:11-35)."""
