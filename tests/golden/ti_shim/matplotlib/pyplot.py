"""Placeholder (see __init__.py)."""
