"""Public names of the package (see srgan_b200/__init__.py for the import alias)."""
from . import nets, engine  # noqa: F401
from .nets import describe_module, Net, Layer, Geom  # noqa: F401
from .engine import Engine  # noqa: F401
