from mmlf_b200.data.hci4d import *  # noqa: F401,F403
from mmlf_b200.data.hci4d import __doc__  # noqa: F401
