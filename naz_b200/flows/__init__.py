from .flow import NormalizingFlow, flow_makers  # noqa: F401
from .mcdpflow import MCDPNormalizingFlow  # noqa: F401
from .bflow import BayesianNormalizingFlow  # noqa: F401
from .transforms import (bounding_transform, inverse_bounding_transform, masked_affine_autoregressive,  # noqa: F401
                         neural_spline_autoregressive)
