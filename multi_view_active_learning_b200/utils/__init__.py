"""Drop-in counterparts of the reference's ``utils`` modules on the scoring-and-selection path
(utils/triangulation.py, utils/evaluation.py, utils/coreset.py): same names, arguments and return types,
computed by the sm_100a kernels behind include/mval_b200.h."""
from . import coreset, evaluation, triangulation  # noqa: F401
