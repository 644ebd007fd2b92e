"""Drop-in name for `pseudo-lidar/Transform/Transform.py` (`from Transform.Transform import Transform`,
pseudo-lidar/test_pipeline.py:13).  The class lives in `plb200.velodyne`."""
from plb200.velodyne import Transform  # noqa: F401
