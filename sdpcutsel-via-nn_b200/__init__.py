"""B200-native cut-selection hot path of rb2309/SDPCutSel-via-NN."""
from . import nn_weights  # noqa: F401
from . import _capi  # noqa: F401
from . import synthetic  # noqa: F401
from . import distributed  # noqa: F401
from . import cover  # noqa: F401
from . import neartie  # noqa: F401
from . import training_data  # noqa: F401
from .cut_select_qp import B200CutSelection, CutSolver, RankList  # noqa: F401
from .cut_select_qcqp import CutSolverQCQP  # noqa: F401
from .dropin import make_solvers  # noqa: F401
