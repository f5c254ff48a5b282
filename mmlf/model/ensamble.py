from mmlf_b200.model.ensamble import *  # noqa: F401,F403
from mmlf_b200.model.ensamble import __doc__  # noqa: F401
