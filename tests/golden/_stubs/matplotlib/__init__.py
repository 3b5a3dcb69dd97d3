"""Inert stand-in for matplotlib, used only by tests/golden/make_golden.py."""
