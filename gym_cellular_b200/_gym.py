"""Imports `gymnasium`, falling back to the structural stand-in in compat/ when it is absent."""
import os
import sys

try:
    import gymnasium  # noqa: F401
except ImportError:  # the build image has no gymnasium wheel and no network
    sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat"))
    import gymnasium  # noqa: F401

gym = gymnasium
