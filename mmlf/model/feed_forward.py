from mmlf_b200.model.feed_forward import *  # noqa: F401,F403
from mmlf_b200.model.feed_forward import __doc__  # noqa: F401
