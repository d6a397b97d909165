"""Drop-in import path of the reference package (`from synchrad.calc import SynchRad`,
`from synchrad.utils import J_in_um`), backed by synchrad_b200."""
from synchrad_b200 import __version__  # noqa: F401
