"""B200-native cut-selection hot path of rb2309/SDPCutSel-via-NN."""
from . import nn_weights  # noqa: F401
from . import _capi  # noqa: F401
from . import synthetic  # noqa: F401
from . import distributed  # noqa: F401
