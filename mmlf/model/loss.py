from mmlf_b200.model.loss import *  # noqa: F401,F403
from mmlf_b200.model.loss import __doc__  # noqa: F401
