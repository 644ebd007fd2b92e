"""Drop-in name for `pseudo-lidar/utils/PseudoLiDAR.py` (`from utils.PseudoLiDAR import PseudoLiDAR`,
PseudoLidarPipeline.py:14).  The class lives in `plb200.pseudolidar`; this file only makes the reference's
import path resolve when this directory's `utils/` is the one Python finds.  In the reference's own
`pseudo-lidar/` directory `utils` is a REGULAR package (it has an `__init__.py`) and wins over this namespace
portion whatever the order of `sys.path`: there, call `plb200.dropin.install()`."""
from plb200.pseudolidar import PseudoLiDAR  # noqa: F401
