from rfm_b200.metrics import *  # noqa: F401,F403
from rfm_b200.metrics import metric_candidates  # noqa: F401
