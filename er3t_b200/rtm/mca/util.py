"""Callers of the hot path that er3t/rtm/mca/util.py provides (IPA reflectance-vs-COT look-up tables)."""

__all__ = []
