"""Public names of the package (see srgan_b200/__init__.py for the import alias)."""
from . import nets, engine  # noqa: F401
from .nets import describe_module, Net, Layer, Geom  # noqa: F401
from .engine import Engine  # noqa: F401
from .experiment import (Settings, Experiment, StepRunner, B200StepMixin, CoefficientMLP, CoefficientGenerator,  # noqa: F401,E402
                         DcganDiscriminator, DcganGenerator, KnnDenseNetCat, abs_mean, abs_mean_neg, abs_plus_one_sqrt_mean_neg,
                         abs_plus_one_log_mean_neg, square_mean, norm_mean)
