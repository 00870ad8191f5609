"""Placeholder for `matplotlib` (absent offline): parsers/texture_packing.py imports it for a debug plot only."""
