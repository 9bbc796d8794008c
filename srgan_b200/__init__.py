"""Import alias: the product package directory is `sr-gan_b200/` (not a valid Python identifier), so
`import srgan_b200` resolves its submodules from there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'sr-gan_b200'))

from ._api import *  # noqa: E402,F401,F403
