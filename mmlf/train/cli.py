from mmlf_b200.train.cli import *  # noqa: F401,F403
from mmlf_b200.train.cli import __doc__  # noqa: F401
import sys
from mmlf_b200.train.cli import main

if __name__ == "__main__":
    sys.exit(main())
