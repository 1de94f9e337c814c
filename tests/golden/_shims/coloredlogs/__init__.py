"""Test-only stand-in for `coloredlogs` (main.py:3,15) for the golden generator."""


def install(**kw):
    return None
