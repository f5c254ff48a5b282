from mmlf_b200.utils.pfm import *  # noqa: F401,F403
from mmlf_b200.utils.pfm import __doc__, load, save  # noqa: F401
