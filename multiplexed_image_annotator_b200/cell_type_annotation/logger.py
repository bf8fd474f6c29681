"""`Logger` under the reference's module path; the implementation is `io.RunLog`."""
from ..io import RunLog as Logger

__all__ = ["Logger"]
