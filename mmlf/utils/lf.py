from mmlf_b200.utils.lf import *  # noqa: F401,F403
from mmlf_b200.utils.lf import __doc__  # noqa: F401
