from mmlf_b200.utils.dl import *  # noqa: F401,F403
from mmlf_b200.utils.dl import __doc__  # noqa: F401
