"""TEST INFRASTRUCTURE ONLY -- `from graph_tool.all import Graph` stand-in."""
from . import Graph  # noqa: F401
