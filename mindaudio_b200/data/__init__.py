from . import cmvn, features, spectrum  # noqa: F401
from .cmvn import *  # noqa: F401,F403
from .features import *  # noqa: F401,F403
from .spectrum import *  # noqa: F401,F403
