"""Import-path compatibility: the metric classes used to live here (parity: ``fl4health/utils/metrics.py``); they are in
``fl4health_b200.metrics`` now.  Importing this module still works and says where to look."""

from logging import WARNING

from fl4health_b200.common.logger import log
from fl4health_b200.metrics import *  # noqa: F401, F403

log(WARNING, "Metrics now reside at fl4health_b200/metrics/metrics.py. This path will be removed in future releases.")
