"""Minimal `mako` for the reference's kernel templating (see template.py); only used when Mako is not installed."""
